"""Generate tests/golden/ref_cuda_*.npz by running the UNMODIFIED reference CUDA kernels
(oracle/_ref/gsplat_ref_cuda.so, built from /root/reference by oracle/Makefile) on a GPU.
TEST INFRASTRUCTURE.  Run on the GPU box; the vectors pin oracle_cpu.c / oracle.py (CPU tests) and are
an independent check of the product kernels.

    python oracle/make_golden_gpu.py [out_dir]
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import gsplat_ref_cuda as ref  # noqa: E402

from gscodec_studio_b200._helper import load_test_data  # noqa: E402


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    means, quats, scales, opac, colors, viewmats, Ks, w0, h0 = load_test_data(device="cpu", scene_grid=1, seed=42)
    sel = slice(0, None, 6)  # ~18.6k Gaussians
    means, quats, scales, opac, colors = [t[sel].contiguous() for t in (means, quats, scales, opac, colors)]
    scales = scales * 2.0  # keep a healthy number of intersections at the reduced resolution
    C, W, H = 2, 324, 210
    viewmats, Ks = viewmats[:C].clone(), Ks[:C].clone()
    Ks[:, 0, :] *= W / w0
    Ks[:, 1, :] *= H / h0
    d = lambda t: t.to(dev).contiguous()
    cm = ref.CameraModelType.PINHOLE
    radii, m2d, dep, con, comp = ref.fully_fused_projection_fwd(d(means), None, d(quats), d(scales), d(viewmats), d(Ks), W, H,
                                                                0.3, 0.01, 100.0, 0.0, True, cm)
    valid = radii > 0
    zero = lambda t: torch.where(valid.reshape(valid.shape + (1,) * (t.dim() - 2)), t, torch.zeros_like(t))
    m2d, dep, con, comp = zero(m2d), zero(dep), zero(con), zero(comp)  # reference leaves culled rows uninitialised
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, fl = ref.isect_tiles(m2d, radii, dep, None, None, C, 16, tw, th, True, True)
    offs = ref.isect_offset_encode(ids, C, tw, th)
    cols = d(colors)[None].expand(C, -1, -1).contiguous()
    ops = d(opac)[None].expand(C, -1).contiguous()
    g = torch.Generator().manual_seed(17)
    bg = torch.rand(C, 3, generator=g).to(dev)
    v_c = torch.randn(C, H, W, 3, generator=g).to(dev)
    v_a = torch.randn(C, H, W, 1, generator=g).to(dev)
    rc, ra, last = ref.rasterize_to_pixels_fwd(m2d, con, cols, ops, bg, None, W, H, 16, offs, fl)
    v_abs, v_m, v_cn, v_col, v_o = ref.rasterize_to_pixels_bwd(m2d, con, cols, ops, bg, None, W, H, 16, offs, fl, ra, last,
                                                               v_c, v_a, True)
    # projection backward with the blend's gradients as cotangents
    v_dep = torch.zeros_like(dep)
    v_means, _, v_quats, v_scales, v_view = ref.fully_fused_projection_bwd(
        d(means), None, d(quats), d(scales), d(viewmats), d(Ks), W, H, 0.3, cm, radii, con, None, v_m, v_dep, v_cn, None, True)
    torch.cuda.synchronize()
    c = lambda t: t.detach().cpu().numpy()
    np.savez_compressed(
        os.path.join(out_dir, "ref_cuda_pipeline.npz"),
        means=c(means), quats=c(quats), scales=c(scales), opacities=c(opac), colors=c(colors), viewmats=c(viewmats), Ks=c(Ks),
        W=W, H=H, radii=c(radii), means2d=c(m2d), depths=c(dep), conics=c(con), compensations=c(comp),
        tiles_per_gauss=c(tpg), isect_ids=c(ids), flatten_ids=c(fl), isect_offsets=c(offs), backgrounds=c(bg),
        v_render_colors=c(v_c).astype(np.float16), v_render_alphas=c(v_a).astype(np.float16),
        render_colors=c(rc), render_alphas=c(ra), last_ids=c(last),
        v_means2d_abs=c(v_abs), v_means2d=c(v_m), v_conics=c(v_cn), v_colors=c(v_col), v_opacities=c(v_o),
        v_means=c(v_means), v_quats=c(v_quats), v_scales=c(v_scales), v_viewmats=c(v_view))
    print("n_isects", ids.numel(), "visible", int(valid.sum()), "->", os.path.join(out_dir, "ref_cuda_pipeline.npz"))


if __name__ == "__main__":
    main()
