"""On-disk quantise-to-grid (PngCompression) against golden vectors produced by the REFERENCE's own
gsplat/compression/png_compression.py run on the CPU (oracle/make_golden_png.py -> tests/golden/ref_png_quant.npz):
stored bytes bit-exact, decoded values bit-exact; then a full compress -> decompress round trip through PNG files."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)
sys.path.insert(0, ROOT)

from gscodec_studio_b200.compression import PngCompression, dequantize_from_grid, quantize_to_grid  # noqa: E402

DEV = torch.device("cuda:0")
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_png_quant.npz"))


@pytest.mark.parametrize("name", ["means", "opacities", "scales", "quats", "sh0"])
def test_grid_quantize_matches_reference_bytes(name):
    x = torch.from_numpy(GOLD[f"{name}_in"]).to(DEV)
    bits = int(GOLD[f"{name}_bits"])
    lo, hi, mins, maxs = quantize_to_grid(x, bits)
    assert np.array_equal(mins.double().cpu().numpy(), GOLD[f"{name}_mins"]) and np.array_equal(maxs.double().cpu().numpy(), GOLD[f"{name}_maxs"])
    assert np.array_equal(lo.cpu().numpy().reshape(GOLD[f"{name}_lo"].shape), GOLD[f"{name}_lo"]), "stored bytes differ"
    if bits > 8:
        assert np.array_equal(hi.cpu().numpy().reshape(GOLD[f"{name}_hi"].shape), GOLD[f"{name}_hi"])
    else:
        assert hi is None
    dec = dequantize_from_grid(lo, hi, mins, maxs, bits).reshape(x.shape)
    assert np.array_equal(dec.cpu().numpy(), GOLD[f"{name}_decoded"]), "decoded values differ"
    step = (maxs - mins) / (2**bits - 1)
    assert ((dec - x).reshape(x.shape[0], -1).abs() <= 0.51 * step + 4e-6 * x.abs().max()).all()


def test_png_compression_round_trip(tmp_path):
    g = torch.Generator().manual_seed(1)
    n = 50 * 50 + 17  # not a square: the 17 most transparent splats are dropped
    splats = {
        "means": (torch.randn(n, 3, generator=g) * 3).to(DEV),
        "scales": (torch.randn(n, 3, generator=g) - 4).to(DEV),
        "quats": torch.randn(n, 4, generator=g).to(DEV),
        "opacities": (torch.randn(n, generator=g) * 2 + 1).to(DEV),
        "sh0": torch.randn(n, 1, 3, generator=g).to(DEV),
        "shN": (torch.randn(n, 15, 3, generator=g) * 0.05).to(DEV),
    }
    keep = torch.sigmoid(splats["opacities"]) >= 0.005
    codec = PngCompression(use_sort=False, verbose=False)
    codec.compress(str(tmp_path), {k: v.clone() for k, v in splats.items()})
    for f in ("means_l.png", "means_u.png", "scales.png", "quats.png", "opacities.png", "sh0.png", "shN.npz", "meta.json"):
        assert (tmp_path / f).exists(), f
    out = codec.decompress(str(tmp_path))
    kept = {k: v[keep] for k, v in splats.items()}
    order = torch.argsort(kept["opacities"], descending=True)
    n_side = int(keep.sum().item() ** 0.5)
    order = order[: n_side * n_side]
    assert out["means"].shape == (n_side * n_side, 3)
    ref_means = kept["means"][order]
    assert (out["means"].to(DEV) - ref_means).abs().max() < 2e-3 * ref_means.abs().max()  # 16 bits of log-space
    assert (out["opacities"].to(DEV) - kept["opacities"][order]).abs().max() <= (kept["opacities"].max() - kept["opacities"].min()) / 255 * 0.51
    qn = torch.nn.functional.normalize(kept["quats"][order], dim=-1)
    assert (out["quats"].to(DEV) - qn).abs().max() <= 2.0 / 255 * 0.51 + 1e-6
    assert torch.equal(out["shN"].to(DEV), kept["shN"][order])  # lossless container (no torchpq here)
    with pytest.raises(ImportError, match="PLAS"):
        PngCompression(use_sort=True, verbose=False).compress(str(tmp_path), {k: v.clone() for k, v in splats.items()})
