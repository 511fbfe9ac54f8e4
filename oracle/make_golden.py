"""Generate tests/golden/ref_cpu_*.npz from the REFERENCE's own pure-torch code.  TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference).  The reference modules are loaded by
file path because `import gsplat` needs the absent `_gridencoder` extension.
The vectors pin oracle/oracle.py (see tests/test_oracle_golden.py); they travel to the GPU box as
small committed fixtures because /root/reference does not.

    python oracle/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GSPLAT_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_entropy_model():
    # entropy_model.py:311 imports .gaussian_distribution_model (-> _gridencoder): stub that one name.
    pkg = types.ModuleType("refcs")
    pkg.__path__ = [os.path.join(REF, "gsplat", "compression_simulation")]
    sys.modules["refcs"] = pkg
    stub = types.ModuleType("refcs.gaussian_distribution_model")
    stub.hash_based_estimator = object
    sys.modules["refcs.gaussian_distribution_model"] = stub
    return load_by_path("refcs.entropy_model", os.path.join(REF, "gsplat/compression_simulation/entropy_model.py"))


def main():
    os.makedirs(OUT, exist_ok=True)
    ti = load_by_path("ref_torch_impl", os.path.join(REF, "gsplat/cuda/_torch_impl.py"))
    ops = load_by_path("ref_cs_ops", os.path.join(REF, "gsplat/compression_simulation/ops.py"))
    em = load_entropy_model()

    # ---------------- projection (fwd + autograd grads), 3 camera models -----------------
    g = torch.Generator().manual_seed(1234)
    C, N, W, H = 3, 256, 300, 200
    means = torch.randn(N, 3, generator=g) * 1.5
    quats = torch.randn(N, 4, generator=g)
    scales = torch.rand(N, 3, generator=g) * 0.3 + 0.02
    viewmats = torch.eye(4).repeat(C, 1, 1)
    for c in range(C):
        ang = 0.3 * c
        viewmats[c, :3, :3] = torch.tensor(
            [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype=torch.float32)
        viewmats[c, :3, 3] = torch.tensor([0.1 * c, -0.05 * c, 5.0])
    Ks = torch.tensor([[250.0, 0, W / 2], [0, 260.0, H / 2], [0, 0, 1]]).repeat(C, 1, 1)
    out = dict(means=means, quats=quats, scales=scales, viewmats=viewmats, Ks=Ks, W=W, H=H)
    wm = torch.randn(C, N, 2, generator=g)
    wd = torch.randn(C, N, generator=g)
    wc = torch.randn(C, N, 3, generator=g)
    wk = torch.randn(C, N, generator=g)
    out.update(w_means2d=wm, w_depths=wd, w_conics=wc, w_comp=wk)
    for model in ("pinhole", "ortho", "fisheye"):
        m, q, s, v = [t.clone().requires_grad_(True) for t in (means, quats, scales, viewmats)]
        covars, _ = ti._quat_scale_to_covar_preci(q, s, True, False, triu=False)
        radii, m2d, dep, con, comp = ti._fully_fused_projection(
            m, covars, v, Ks, W, H, eps2d=0.3, near_plane=0.01, far_plane=1e10,
            calc_compensations=True, camera_model=model)
        valid = radii > 0
        loss = ((m2d * wm)[valid].sum() + (dep * wd)[valid].sum() + (con * wc)[valid].sum()
                + (comp * wk)[valid].sum())
        loss.backward()
        out.update({
            f"{model}_radii": radii, f"{model}_means2d": m2d.detach(), f"{model}_depths": dep.detach(),
            f"{model}_conics": con.detach(), f"{model}_comp": comp.detach(),
            f"{model}_v_means": m.grad, f"{model}_v_quats": q.grad, f"{model}_v_scales": s.grad,
            f"{model}_v_viewmats": v.grad})
    np.savez_compressed(os.path.join(OUT, "ref_cpu_projection.npz"), **{k: np.asarray(v) for k, v in out.items()})

    # ---------------- spherical harmonics deg 0..4 (fwd + grads) -----------------
    out = {}
    g = torch.Generator().manual_seed(7)
    n = 500
    dirs0 = torch.randn(n, 3, generator=g)
    coeffs0 = torch.randn(n, 25, 3, generator=g)
    wcol = torch.randn(n, 3, generator=g)
    out.update(dirs=dirs0, coeffs=coeffs0, w=wcol)
    for deg in range(5):
        d, c = dirs0.clone().requires_grad_(True), coeffs0.clone().requires_grad_(True)
        col = ti._spherical_harmonics(deg, d, c)
        (col * wcol).sum().backward()
        out.update({f"deg{deg}_colors": col.detach(), f"deg{deg}_v_dirs": d.grad if d.grad is not None else torch.zeros_like(d), f"deg{deg}_v_coeffs": c.grad})
    np.savez_compressed(os.path.join(OUT, "ref_cpu_sh.npz"), **{k: np.asarray(v) for k, v in out.items()})

    # ---------------- tile intersection + offsets (reference Python double loop) -----------------
    g = torch.Generator().manual_seed(42)
    C, N, W, H, ts = 3, 400, 40, 60, 4  # tests/test_basic.py:442-472 uses C=3, N=1000, 40x60
    tw, th = (W + ts - 1) // ts, (H + ts - 1) // ts
    m2d = torch.randn(C, N, 2, generator=g) * W
    radii = torch.randint(0, W, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g)
    depths[:, 10:20] = depths[:, 30:40]  # exact depth ties exercise the stable order
    tpg, ids, fl = ti._isect_tiles(m2d, radii, depths, ts, tw, th, sort=True)
    offs = ti._isect_offset_encode(ids, C, tw, th)
    np.savez_compressed(os.path.join(OUT, "ref_cpu_isect.npz"), means2d=m2d, radii=radii, depths=depths,
                        tile_size=ts, tile_width=tw, tile_height=th, tiles_per_gauss=tpg, isect_ids=ids,
                        flatten_ids=fl, isect_offsets=offs)

    # ---------------- quant-sim ops -----------------
    out = {}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 3, generator=g) * 6 - 4
    out["x"] = x.clone()
    xr = x.clone().requires_grad_(True)
    xr2 = xr * 1.0  # non-leaf so the in-place clamp_ is legal
    res = ops.fake_quantize_ste(xr2, -10.0, 2.0, 8, "round")
    res["output_value"].sum().backward()
    out.update(round_y=res["output_value"].detach(), round_q=res["q_step"], round_x_after=xr2.detach(),
               round_grad=xr.grad)
    torch.manual_seed(99)
    xn = x.clone().requires_grad_(True)
    res = ops.fake_quantize_ste(xn, -10.0, 2.0, 8, "noise")
    (res["output_value"] * 2.0).sum().backward()
    torch.manual_seed(99)
    noise = torch.empty_like(x).uniform_(-0.5, 0.5)
    out.update(noise=noise, noise_y=res["output_value"].detach(), noise_q=res["q_step"], noise_grad=xn.grad)
    np.savez_compressed(os.path.join(OUT, "ref_cpu_quant.npz"), **{k: np.asarray(v) for k, v in out.items()})

    # ---------------- factorized entropy model (the only usable class: *_refactor) -----------------
    out = {}
    for tag, ch, filters, n in (("s3", 3, (3, 3), 1000), ("q4", 4, (3, 3, 3), 1024), ("t3", 3, (3, 3), 37)):
        torch.manual_seed(5)
        np.random.seed(5)
        model = em.Entropy_factorized_optimized_refactor(channel=ch, filters=filters)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.3)
        x = (torch.randn(n, ch) * 2).requires_grad_(True)
        q = 12.0 / 255
        bits = model(x, q)
        w = torch.randn(n, ch)
        (bits * w).sum().backward()
        out.update({f"{tag}_x": x.detach(), f"{tag}_q": q, f"{tag}_bits": bits.detach(), f"{tag}_w": w,
                    f"{tag}_v_x": x.grad, f"{tag}_nlayers": len(filters) + 1})
        for i in range(len(filters) + 1):
            out[f"{tag}_matrix{i}"] = model._matrices[i].detach()
            out[f"{tag}_bias{i}"] = model._bias[i].detach()
            out[f"{tag}_v_matrix{i}"] = model._matrices[i].grad
            out[f"{tag}_v_bias{i}"] = model._bias[i].grad
            if i < len(filters):
                out[f"{tag}_factor{i}"] = model._factor[i].detach()
                out[f"{tag}_v_factor{i}"] = model._factor[i].grad
    np.savez_compressed(os.path.join(OUT, "ref_cpu_entropy.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
