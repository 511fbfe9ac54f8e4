"""`PngCompression` -- quantise the splat attributes onto an image grid and store them as PNG planes; API and file format of
the reference's ``gsplat/compression/png_compression.py`` (:18-163 the class, :166-360 the per-attribute functions), so that
directories written by either implementation decode with the other.

What runs where:
  * the per-attribute quantise / dequantise (normalise per channel, round to k bits, pack into byte planes; the reference
    does it with five torch / numpy passes on the host) is one CUDA kernel each way: gsb200_grid_quantize / _dequantize.
    The stored bytes are identical to the reference's (golden vectors from its numpy arithmetic: tests/golden);
  * PNG encoding / decoding uses Pillow (the reference uses imageio, which is not installed here; same files);
  * sorting (PLAS) and the k-means codebook of shN (TorchPQ) are third-party packages the reference imports lazily and this
    image does not have: `use_sort=True` needs plas installed, and shN falls back to the reference's own generic NPZ
    container (`_compress_npz`) when torchpq is missing.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from .._cabi import call, device_guard, ptr, stream_ptr
from ..utils import inverse_log_transform, log_transform

# attribute -> (bits, container); everything else goes to an NPZ file (png_compression.py:48-63)
_GRID_BITS = {"means": 16, "scales": 8, "quats": 8, "opacities": 8, "sh0": 8}


def quantize_to_grid(params: Tensor, bits: int = 8) -> Tuple[Tensor, Optional[Tensor], Tensor, Tensor]:
    """params [n, ...] (CUDA, float32) -> (lo u8 [n, ch], hi u8 [n, ch] or None, mins [ch], maxs [ch]).
    bits <= 8: lo = round((x - min) / (max - min) * (2^bits - 1)) << (8 - bits)   (_compress_png / _compress_png_kbit)
    bits == 16: the 16-bit value split into a low and a high byte plane            (_compress_png_16bit)"""
    if not params.is_cuda:
        raise RuntimeError("gscodec_studio_b200: quantize_to_grid needs a CUDA tensor (no CPU implementation)")
    x = params.detach().to(torch.float32).reshape(params.shape[0], -1).contiguous()
    ch = x.shape[1]
    mins, maxs = torch.amin(x, dim=0), torch.amax(x, dim=0)
    lo = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    hi = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if bits > 8 else None
    with device_guard(x.device):
        call("gsb200_grid_quantize", x.numel(), ch, ptr(x), ptr(mins), ptr(maxs), int(bits), ptr(lo), ptr(hi), stream_ptr(x.device))
    return lo, hi, mins, maxs


def dequantize_from_grid(lo: Tensor, hi: Optional[Tensor], mins: Tensor, maxs: Tensor, bits: int = 8) -> Tensor:
    """Inverse of quantize_to_grid -> float32 [n, ch] (decoded in float64 like the reference, png_compression.py:205-226)."""
    if not lo.is_cuda:
        raise RuntimeError("gscodec_studio_b200: dequantize_from_grid needs CUDA tensors (no CPU implementation)")
    lo = lo.contiguous()
    hi = None if hi is None else hi.contiguous()
    ch = lo.shape[1]
    mins, maxs = mins.to(lo.device, torch.float32).contiguous(), maxs.to(lo.device, torch.float32).contiguous()
    out = torch.empty(lo.shape, dtype=torch.float32, device=lo.device)
    with device_guard(lo.device):
        call("gsb200_grid_dequantize", lo.numel(), ch, ptr(lo), ptr(hi), ptr(mins), ptr(maxs), int(bits), ptr(out), stream_ptr(lo.device))
    return out


def _write_png(path: str, plane: Tensor, n_sidelen: int) -> None:
    from PIL import Image

    img = plane.reshape(n_sidelen, n_sidelen, -1).cpu().numpy().squeeze()
    Image.fromarray(img).save(path)


def _read_png(path: str) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(path))


def _crop_n_splats(splats: Dict[str, Tensor], n_crop: int) -> Dict[str, Tensor]:
    keep = torch.argsort(splats["opacities"], descending=True)[:-n_crop]
    return {k: v[keep] for k, v in splats.items()}


@dataclass
class PngCompression:
    """Quantise-to-grid PNG container (reference: png_compression.py:18-163).  `splats` hold pre-activation values with the
    fields "means", "scales", "quats", "opacities", "sh0", "shN"; extra fields are stored as NPZ."""

    use_sort: bool = True
    verbose: bool = True

    def compress(self, compress_dir: str, splats: Dict[str, Tensor], entropy_models=None) -> None:
        if entropy_models is not None:
            raise ValueError("PngCompression should not require entropy_models")
        os.makedirs(compress_dir, exist_ok=True)
        splats = dict(splats)
        # outlier filtering: near-transparent splats are dropped (outlier_filter.py:8-33)
        valid = torch.sigmoid(splats["opacities"]) >= 0.005
        splats = {k: v[valid] for k, v in splats.items()}
        splats["means"] = log_transform(splats["means"])
        splats["quats"] = F.normalize(splats["quats"], dim=-1)
        n_gs = len(splats["means"])
        n_sidelen = int(n_gs**0.5)
        n_crop = n_gs - n_sidelen**2
        if n_crop != 0:
            splats = _crop_n_splats(splats, n_crop)
            print(f"Warning: Number of Gaussians was not square. Removed {n_crop} Gaussians.")
        if self.use_sort:
            splats = self._sort(splats)
        meta: Dict[str, Any] = {}
        for name, value in splats.items():
            if name in _GRID_BITS:
                meta[name] = self._compress_grid(compress_dir, name, value, n_sidelen, _GRID_BITS[name])
            elif name == "shN":
                meta[name] = self._compress_shN(compress_dir, name, value)
            else:
                meta[name] = self._compress_npz(compress_dir, name, value)
        with open(os.path.join(compress_dir, "meta.json"), "w") as f:
            json.dump(meta, f)

    def decompress(self, compress_dir: str, device="cuda") -> Dict[str, Tensor]:
        with open(os.path.join(compress_dir, "meta.json")) as f:
            meta = json.load(f)
        splats = {}
        for name, m in meta.items():
            if m.get("container") == "npz" or name not in _GRID_BITS:
                if name == "shN" and "quantization" in m and m.get("container") != "npz":
                    splats[name] = self._decompress_kmeans(compress_dir, name, m)
                else:
                    splats[name] = self._decompress_npz(compress_dir, name, m)
            else:
                splats[name] = self._decompress_grid(compress_dir, name, m, device)
        splats["means"] = inverse_log_transform(splats["means"])
        return splats

    # ---- grid attributes
    @staticmethod
    def _compress_grid(compress_dir, name, params, n_sidelen, bits) -> Dict[str, Any]:
        lo, hi, mins, maxs = quantize_to_grid(params, bits)
        if hi is None:
            _write_png(os.path.join(compress_dir, f"{name}.png"), lo, n_sidelen)
        else:
            _write_png(os.path.join(compress_dir, f"{name}_l.png"), lo, n_sidelen)
            _write_png(os.path.join(compress_dir, f"{name}_u.png"), hi, n_sidelen)
        meta = {"shape": list(params.shape), "dtype": str(params.dtype).split(".")[1], "mins": mins.tolist(), "maxs": maxs.tolist()}
        if name != "means" and name != "opacities":
            meta["quantization"] = bits  # (_compress_png_kbit records it; _compress_png / _png_16bit do not)
        return meta

    @staticmethod
    def _decompress_grid(compress_dir, name, meta, device) -> Tensor:
        if not np.all(meta["shape"]):
            return torch.zeros(meta["shape"], dtype=getattr(torch, meta["dtype"]))
        bits = _GRID_BITS[name] if "quantization" not in meta else int(meta["quantization"])
        if name == "means":
            lo = torch.from_numpy(_read_png(os.path.join(compress_dir, f"{name}_l.png")).copy()).to(device)
            hi = torch.from_numpy(_read_png(os.path.join(compress_dir, f"{name}_u.png")).copy()).to(device)
            bits = 16
        else:
            lo = torch.from_numpy(_read_png(os.path.join(compress_dir, f"{name}.png")).copy()).to(device)
            hi = None
        ch = len(meta["mins"])
        out = dequantize_from_grid(lo.reshape(-1, ch), None if hi is None else hi.reshape(-1, ch), torch.tensor(meta["mins"]),
                                   torch.tensor(meta["maxs"]), bits)
        return out.reshape(meta["shape"]).to(dtype=getattr(torch, meta["dtype"]))

    # ---- the rest
    @staticmethod
    def _compress_npz(compress_dir, name, params) -> Dict[str, Any]:
        np.savez_compressed(os.path.join(compress_dir, f"{name}.npz"), arr=params.detach().cpu().numpy())
        return {"shape": list(params.shape), "dtype": str(params.dtype).split(".")[1], "container": "npz"}

    @staticmethod
    def _decompress_npz(compress_dir, name, meta) -> Tensor:
        arr = np.load(os.path.join(compress_dir, f"{name}.npz"))["arr"]
        return torch.tensor(arr).reshape(meta["shape"]).to(dtype=getattr(torch, meta["dtype"]))

    def _compress_shN(self, compress_dir, name, params) -> Dict[str, Any]:
        try:
            from torchpq.clustering import KMeans  # noqa: F401
        except Exception:
            if self.verbose:
                print("PngCompression: torchpq is not installed; shN is stored losslessly as NPZ instead of a k-means codebook")
            return self._compress_npz(compress_dir, name, params)
        raise NotImplementedError("the k-means codebook of shN (png_compression.py:499-600) needs torchpq; not wired up in this build")

    @staticmethod
    def _decompress_kmeans(compress_dir, name, meta) -> Tensor:
        """Reads a reference-written k-means container (png_compression.py:466-496)."""
        z = np.load(os.path.join(compress_dir, f"{name}.npz"))
        centroids = torch.tensor(z["centroids"] / (2 ** meta["quantization"] - 1))
        mins, maxs = torch.tensor(meta["mins"]), torch.tensor(meta["maxs"])
        params = (centroids * (maxs - mins) + mins)[z["labels"].astype(np.int32)]
        return params.reshape(meta["shape"]).to(dtype=getattr(torch, meta["dtype"]))

    def _sort(self, splats: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """Parallel Linear Assignment Sorting onto the grid (sort.py:7-59); needs the third-party `plas` package."""
        try:
            from plas import sort_with_plas
        except Exception:
            raise ImportError("Please install PLAS with 'pip install git+https://github.com/fraunhoferhhi/PLAS.git' to use "
                              "sorting (or construct PngCompression(use_sort=False))")
        n_gs = len(splats["means"])
        n_sidelen = int(n_gs**0.5)
        keys = [k for k in splats if k != "shN"]
        params = torch.cat([splats[k].reshape(n_gs, -1) for k in keys], dim=-1)
        shuffled = torch.randperm(n_gs, device=params.device)
        grid = params[shuffled].reshape(n_sidelen, n_sidelen, -1)
        _, order = sort_with_plas(grid.permute(2, 0, 1), improvement_break=1e-4, verbose=self.verbose)
        order = shuffled[order.squeeze().flatten()]
        return {k: v[order] for k, v in splats.items()}
