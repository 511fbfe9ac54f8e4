"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_png_quant.npz by running the REFERENCE's own
gsplat/compression/png_compression.py (_compress_png / _compress_png_kbit / _compress_png_16bit and their inverses,
:166-360) on the CPU of the build container:

    python -m oracle.make_golden_png            (needs /root/reference; the GPU box only reads the fixture)

The reference functions write / read their planes through `imageio.v2`, which is not installed: a stub module captures
the arrays handed to imwrite and serves them back from imread, so the arithmetic that produces and consumes the bytes
is the reference's, line for line.  `_gridencoder` (imported by the package's __init__) is stubbed as in
make_golden_policy.py.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GSPLAT_REF", "/root/reference")


def _reference():
    sys.modules.setdefault("_gridencoder", types.ModuleType("_gridencoder"))
    store = {}
    io = types.ModuleType("imageio.v2")
    io.imwrite = lambda path, img: store.__setitem__(os.path.basename(path), np.array(img, copy=True))
    io.imread = lambda path: store[os.path.basename(path)]
    pkg = types.ModuleType("imageio")
    pkg.v2 = io
    sys.modules["imageio"], sys.modules["imageio.v2"] = pkg, io
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from gsplat.compression import png_compression

    return png_compression, store


def main():
    P, store = _reference()
    g = torch.Generator().manual_seed(0)
    out = {}
    side = 37
    n = side * side
    cases = {
        "means": (torch.randn(n, 3, generator=g) * 4.0, "16bit", {}),
        "opacities": (torch.randn(n, generator=g) * 3.0, "png", {}),
        "scales": (torch.randn(n, 3, generator=g) - 4.0, "kbit", {"quantization": 8}),
        "quats": (torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1), "kbit", {"quantization": 6}),
        "sh0": (torch.randn(n, 1, 3, generator=g), "kbit", {"quantization": 5}),
    }
    fns = {"png": (P._compress_png, P._decompress_png), "kbit": (P._compress_png_kbit, P._decompress_png_kbit),
           "16bit": (P._compress_png_16bit, P._decompress_png_16bit)}
    for name, (x, kind, kw) in cases.items():
        comp, decomp = fns[kind]
        meta = comp("/nowhere", name, x.clone(), n_sidelen=side, **kw)
        out[f"{name}_in"] = x.numpy()
        out[f"{name}_mins"] = np.asarray(meta["mins"], dtype=np.float64)
        out[f"{name}_maxs"] = np.asarray(meta["maxs"], dtype=np.float64)
        out[f"{name}_bits"] = np.int64(16 if kind == "16bit" else kw.get("quantization", 8))
        if kind == "16bit":
            out[f"{name}_lo"], out[f"{name}_hi"] = store[f"{name}_l.png"], store[f"{name}_u.png"]
        else:
            out[f"{name}_lo"] = store[f"{name}.png"]
        out[f"{name}_decoded"] = decomp("/nowhere", name, meta).numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_png_quant.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
