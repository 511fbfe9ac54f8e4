"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Drives the UNMODIFIED reference CUDA kernels (oracle/_ref/gsplat_ref_cuda.so, built by oracle/Makefile from
/root/reference/gsplat/cuda/csrc) through the reference's own op sequence for one training step of the headline
workload:

    fake_quantize_ste (noise)          gsplat/compression_simulation/ops.py:39-54, simulation.py:240-316
    exp / sigmoid / cat(sh0, shN)      examples/simple_trainer.py:766-778
    rasterization(sh_degree=3)         gsplat/rendering.py:296-392, 445-580 (dense, pinhole, classic)
      fully_fused_projection_fwd   ->  spherical_harmonics (+0.5, clamp_min 0)  ->  isect_tiles / isect_offset_encode
      -> rasterize_to_pixels_fwd
    backward                           gsplat/cuda/_wrapper.py:957-1028, 1244-1255, 829-898

The torch glue between the kernels is restated here (the reference's rendering.py / _wrapper.py need the package import,
which needs the un-vendored _gridencoder); the kernels are the reference's, bit for bit.  Used by
tests/test_gpu_parity.py (live parity at the headline size from the RAW parameters) and by bench.py's `ref_cuda`
checker leg (the "reference kernels on the same B200" comparator, timed outside our own timed region)."""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, Optional

import torch
from torch import Tensor

HERE = os.path.dirname(os.path.abspath(__file__))
BOUNDS = {"scales": (-10.0, 2.0), "quats": (-1.0, 1.0), "opacities": (-15.0, 15.0), "sh0": (-2.0, 4.0)}  # simulation.py:52-59


def load_ref():
    """The reference CUDA module, or None when it did not travel with the repo."""
    path = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(path, "gsplat_ref_cuda.so")):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        import gsplat_ref_cuda  # noqa

        return gsplat_ref_cuda
    except Exception:  # pragma: no cover
        return None


def fake_quantize_noise(x: Tensor, lo: float, hi: float, noise: Optional[Tensor], bitwidth: int = 8) -> Tensor:
    """ops.py:44-47 -- clamp(x) + U(-.5, .5) * q_step; the clamp's autograd gives the in-range mask gradient."""
    q_step = (hi - lo) / (2 ** bitwidth - 1)
    if noise is None:
        noise = torch.empty_like(x).uniform_(-0.5, 0.5)
    return torch.clamp(x, lo, hi) + noise * q_step


class RefStep:
    """One M2 step on the reference kernels.  `stage_ms` (filled when `timed=True`) holds CUDA-event times."""

    def __init__(self, ref, width: int, height: int, tile_size: int = 16, sh_degree: int = 3, near=0.01, far=100.0,
                 radius_clip=3.0, eps2d=0.3):
        self.ref, self.W, self.H, self.ts, self.deg = ref, width, height, tile_size, sh_degree
        self.near, self.far, self.clip, self.eps = near, far, radius_clip, eps2d
        self.tw, self.th = math.ceil(width / tile_size), math.ceil(height / tile_size)
        self.stage_ms: Dict[str, float] = {}

    def _timed(self, name, fn, timed):
        if not timed:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self._events.append((name, e0, e1))
        return out

    def run(self, raw: Dict[str, Tensor], viewmats: Tensor, Ks: Tensor, v_render_colors: Optional[Tensor] = None,
            noise: Optional[Dict[str, Tensor]] = None, quantize: bool = True, timed: bool = False, backward: bool = True):
        """raw: leaf tensors means/scales/quats/opacities/sh0/shN (requires_grad for gradients).  Returns a dict with the
        stage outputs and (backward=True) leaves `.grad` on the raw tensors.  The loss is sum(render * v_render_colors)
        (v_render_colors=None: ones)."""
        ref, W, H, ts, tw, th = self.ref, self.W, self.H, self.ts, self.tw, self.th
        self._events = []
        C = viewmats.shape[0]
        cm = ref.CameraModelType.PINHOLE

        def pre():
            q = {}
            for k, (lo, hi) in BOUNDS.items():
                q[k] = fake_quantize_noise(raw[k], lo, hi, None if noise is None else noise[k]) if quantize else raw[k]
            scales, opac = torch.exp(q["scales"]), torch.sigmoid(q["opacities"])
            shs = torch.cat([q["sh0"], raw["shN"]], 1)
            return q["quats"], scales, opac, shs

        quats, scales, opac, shs = self._timed("quant_sim+activations (torch)", pre, timed)
        means = raw["means"]
        md, qd, sd = means.detach().contiguous(), quats.detach().contiguous(), scales.detach().contiguous()
        radii, means2d, depths, conics, _ = self._timed("fully_fused_projection_fwd", lambda: ref.fully_fused_projection_fwd(
            md, None, qd, sd, viewmats, Ks, W, H, self.eps, self.near, self.far, self.clip, False, cm), timed)

        def colour():
            campos = torch.inverse(viewmats)[:, :3, 3]
            dirs = md[None, :, :] - campos[:, None, :]
            masks = radii > 0
            coeffs = shs.detach().expand(C, -1, -1, -1).contiguous()
            rgb = ref.compute_sh_fwd(self.deg, dirs, coeffs, masks)
            return dirs, masks, coeffs, rgb, torch.clamp_min(rgb + 0.5, 0.0)

        dirs, masks, coeffs, rgb, colors = self._timed("spherical_harmonics (+glue)", colour, timed)
        opac_c = opac.detach()[None].repeat(C, 1)

        def isect():
            tpg, ids, fl = ref.isect_tiles(means2d, radii, depths, None, None, C, ts, tw, th, True, True)
            return tpg, ids, fl, ref.isect_offset_encode(ids, C, tw, th)

        tpg, isect_ids, flatten_ids, offsets = self._timed("isect_tiles+offset_encode", isect, timed)
        rc, ra, last = self._timed("rasterize_to_pixels_fwd", lambda: ref.rasterize_to_pixels_fwd(
            means2d, conics, colors, opac_c, None, None, W, H, ts, offsets, flatten_ids), timed)
        out = dict(radii=radii, means2d=means2d, depths=depths, conics=conics, colors=colors, tiles_per_gauss=tpg,
                   isect_ids=isect_ids, flatten_ids=flatten_ids, isect_offsets=offsets, render_colors=rc, render_alphas=ra,
                   last_ids=last)
        if backward:
            if callable(v_render_colors):  # gradient of the caller's loss as a function of the rendered image
                v_rc = self._timed("loss (torch)", lambda: v_render_colors(rc).contiguous(), timed)
            else:
                v_rc = torch.ones_like(rc) if v_render_colors is None else v_render_colors.contiguous()
            v_ra = torch.zeros_like(ra)
            _, v_m2d, v_con, v_col, v_op = self._timed("rasterize_to_pixels_bwd", lambda: ref.rasterize_to_pixels_bwd(
                means2d, conics, colors, opac_c, None, None, W, H, ts, offsets, flatten_ids, ra, last, v_rc, v_ra, False), timed)

            def colour_bwd():
                v_rgb = v_col * (rgb + 0.5 > 0)  # clamp_min backward
                K = coeffs.shape[-2]
                v_coeffs, v_dirs = ref.compute_sh_bwd(K, self.deg, dirs, coeffs, masks, v_rgb.contiguous(), True)
                return v_coeffs.sum(0), v_dirs.sum(0)

            v_shs, v_means_dirs = self._timed("spherical_harmonics bwd (+glue)", colour_bwd, timed)
            v_means, _, v_quats, v_scales, _ = self._timed("fully_fused_projection_bwd", lambda: ref.fully_fused_projection_bwd(
                md, None, qd, sd, viewmats, Ks, W, H, self.eps, cm, radii, conics, None, v_m2d.contiguous(),
                torch.zeros_like(depths), v_con.contiguous(), None, False), timed)

            def post():
                roots, grads = [quats, scales, opac, shs], [v_quats, v_scales, v_op.sum(0), v_shs]
                keep = [(r, g) for r, g in zip(roots, grads) if r.requires_grad]
                if keep:
                    torch.autograd.backward([r for r, _ in keep], [g for _, g in keep])
                if means.requires_grad:
                    means.grad = (v_means + v_means_dirs) if means.grad is None else means.grad + v_means + v_means_dirs

            self._timed("autograd through quant_sim (torch)", post, timed)
            out.update(v_means2d=v_m2d, v_conics=v_con, v_colors=v_col, v_opacities=v_op)
        if timed:
            torch.cuda.synchronize()
            self.stage_ms = {}
            for name, e0, e1 in self._events:
                self.stage_ms[name] = self.stage_ms.get(name, 0.0) + e0.elapsed_time(e1)
        return out
