"""Generate tests/golden/ref_cpu_geom_ops.npz from the REFERENCE's own pure-torch code
(gsplat/cuda/_torch_impl.py:41-247: _quat_scale_to_covar_preci, _persp_proj / _ortho_proj / _fisheye_proj,
_world_to_cam), forward values and autograd gradients.  TEST INFRASTRUCTURE; runs only in the build container.

    python oracle/make_golden_geom.py
"""
import os

import numpy as np
import torch

from make_golden import OUT, REF, load_by_path


def main():
    ti = load_by_path("ref_torch_impl", os.path.join(REF, "gsplat/cuda/_torch_impl.py"))
    g = torch.Generator().manual_seed(4321)
    out = {}
    N, C, W, H = 193, 3, 320, 240

    # ---- quat_scale_to_covar_preci, full and triu
    quats = torch.randn(N, 4, generator=g)
    scales = torch.rand(N, 3, generator=g) * 0.5 + 0.1
    out.update(quats=quats, scales=scales)
    for triu in (False, True):
        q, s = quats.clone().requires_grad_(True), scales.clone().requires_grad_(True)
        cov, pre = ti._quat_scale_to_covar_preci(q, s, True, True, triu)
        wc, wp = torch.randn(cov.shape, generator=g), torch.randn(pre.shape, generator=g) * 0.01
        ((cov * wc).sum() + (pre * wp).sum()).backward()
        tag = "triu" if triu else "full"
        out.update({f"qs_{tag}_covars": cov.detach(), f"qs_{tag}_precis": pre.detach(), f"qs_{tag}_wc": wc, f"qs_{tag}_wp": wp,
                    f"qs_{tag}_v_quats": q.grad, f"qs_{tag}_v_scales": s.grad})

    # ---- world_to_cam (slightly non-symmetric covariances exercise the full-matrix formulas)
    means = torch.randn(N, 3, generator=g)
    covars = ti._quat_scale_to_covar_preci(quats, scales, True, False, False)[0] + 0.01 * torch.randn(N, 3, 3, generator=g)
    viewmats = torch.eye(4).repeat(C, 1, 1)
    for c in range(C):
        a = 0.4 * c + 0.1
        viewmats[c, :3, :3] = torch.tensor([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=torch.float32)
        viewmats[c, :3, 3] = torch.tensor([0.2 * c, -0.1 * c, 4.0])
    m, cv, vm = means.clone().requires_grad_(True), covars.clone().requires_grad_(True), viewmats.clone().requires_grad_(True)
    mc, cc = ti._world_to_cam(m, cv, vm)
    wm, wcv = torch.randn(mc.shape, generator=g), torch.randn(cc.shape, generator=g)
    ((mc * wm).sum() + (cc * wcv).sum()).backward()
    out.update(w2c_means=means, w2c_covars=covars, w2c_viewmats=viewmats, w2c_means_c=mc.detach(), w2c_covars_c=cc.detach(),
               w2c_wm=wm, w2c_wc=wcv, w2c_v_means=m.grad, w2c_v_covars=cv.grad, w2c_v_viewmats=vm.grad)

    # ---- proj, three camera models; x/z beyond the field-of-view limits on some points (clamped Jacobian branch)
    Ks = torch.tensor([[260.0, 0, W / 2 + 3], [0, 250.0, H / 2 - 2], [0, 0, 1]]).repeat(C, 1, 1)
    Ks[1, 0, 0] = 300.0
    pm = mc.detach().clone()
    pm[:, ::7, 0] *= 3.0  # far off-axis
    pc = cc.detach().clone()
    out.update(proj_means=pm, proj_covars=pc, proj_Ks=Ks, W=W, H=H)
    fns = {"pinhole": ti._persp_proj, "ortho": ti._ortho_proj, "fisheye": ti._fisheye_proj}
    for name, fn in fns.items():
        a, b = pm.clone().requires_grad_(True), pc.clone().requires_grad_(True)
        m2, c2 = fn(a, b, Ks, W, H)
        w1, w2 = torch.randn(m2.shape, generator=g), torch.randn(c2.shape, generator=g) * 1e-3
        ((m2 * w1).sum() + (c2 * w2).sum()).backward()
        out.update({f"proj_{name}_means2d": m2.detach(), f"proj_{name}_covars2d": c2.detach(), f"proj_{name}_w1": w1,
                    f"proj_{name}_w2": w2, f"proj_{name}_v_means": a.grad, f"proj_{name}_v_covars": b.grad})
    np.savez_compressed(os.path.join(OUT, "ref_cpu_geom_ops.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("wrote ref_cpu_geom_ops.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
