"""Per-stage device timings (CUDA events) of the hot path at the headline size, ours vs the unmodified
reference CUDA kernels (oracle/_ref).  Development aid; bench.py is the contract."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200._helper import load_test_data  # noqa: E402
from gscodec_studio_b200.cuda import _wrapper as W  # noqa: E402

DEV = torch.device("cuda:0")


ONCE = os.environ.get("ONCE") == "1"  # under ncu: one launch per kernel is enough


def timeit(fn, n=20, warm=3):
    if ONCE:
        n, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2] * 1e3  # median, us


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    means, quats, scales, opac, colors, viewmats, Ks, w0, h0 = load_test_data(device=DEV, scene_grid=grid)
    Wd, Hd = 1920, 1080
    viewmats, Ks = viewmats[:1].contiguous(), Ks[:1].clone()
    Ks[:, 0, :] *= Wd / w0
    Ks[:, 1, :] *= Hd / h0
    N = means.shape[0]
    tw, th = math.ceil(Wd / 16), math.ceil(Hd / 16)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import gsplat_ref_cuda as ref
    except Exception as e:  # pragma: no cover
        print("reference CUDA module unavailable:", e)
        ref = None

    radii, m2d, dep, con, _ = W._projection_fwd(means, None, quats, scales, viewmats, Ks, Wd, Hd, 0.3, 0.01, 100.0, 3.0,
                                                False, "pinhole", False)
    tpg, ids, fl, offs = W._isect_tiles_with_offsets(m2d, radii, dep, 16, tw, th)
    cols = colors[None].contiguous()
    ops = opac[None].contiguous()
    I = ids.numel()
    print(f"N={N} visible={(radii > 0).sum().item()} n_isects={I} tiles={tw * th}")
    res = {}
    res["proj_fwd"] = timeit(lambda: W._projection_fwd(means, None, quats, scales, viewmats, Ks, Wd, Hd, 0.3, 0.01, 100.0,
                                                       3.0, False, "pinhole", False))
    res["isect"] = timeit(lambda: W._isect_tiles_with_offsets(m2d, radii, dep, 16, tw, th))
    rc = torch.empty(1, Hd, Wd, 3, device=DEV)
    ra = torch.empty(1, Hd, Wd, 1, device=DEV)
    last = torch.empty(1, Hd, Wd, dtype=torch.int32, device=DEV)
    from gscodec_studio_b200._cabi import call, ptr, stream_ptr

    order = None if os.environ.get("NO_ORDER") == "1" else torch.empty(tw * th, dtype=torch.int32, device=DEV)

    from gscodec_studio_b200._cabi import load_library

    rec = int(load_library().gsb200_blend_stream_record(3, 16))
    bstream = torch.empty(max(I, 1) * rec, device=DEV) if rec else None
    tcnt = torch.empty(tw * th, dtype=torch.int32, device=DEV) if rec else None
    print("blend record floats:", rec)

    def fwd():
        call("gsb200_rasterize_fwd", 1, N, I, 3, ptr(m2d), ptr(con), ptr(cols), ptr(ops), ptr(None), ptr(None), Wd, Hd,
             16, tw, th, ptr(offs), ptr(fl), ptr(rc), ptr(ra), ptr(last), ptr(order), ptr(bstream), ptr(tcnt), stream_ptr(DEV))

    res["blend_fwd"] = timeit(fwd)
    if os.environ.get("TILE_STATS") == "1" and rec:
        # how uneven is the blend's work?  per tile: intersections, records the backward walks, and -- per 8x8 block (= per
        # warp) -- the records that hit it: a warp walks its records serially, so the busiest warp of the longest tile is a
        # lower bound of the kernel time however the other tiles are scheduled
        torch.cuda.synchronize()
        T = tw * th
        ends = torch.cat([offs.view(-1)[1:], torch.tensor([I], device=DEV, dtype=offs.dtype)])
        n_tile = (ends - offs.view(-1)).long()
        cnt = tcnt.long()
        words = bstream[: I * 12].view(-1, 12)[:, 3].contiguous().view(torch.int32)
        tile_of = torch.repeat_interleave(torch.arange(T, device=DEV), n_tile)
        pos = torch.arange(I, device=DEV) - offs.view(-1).long()[tile_of]
        used = pos < cnt[tile_of]
        hits = torch.zeros(T, 4, dtype=torch.long, device=DEV)
        for b in range(4):
            hb = used & (((words >> b) & 1) == 1)
            hits[:, b] = torch.zeros(T, dtype=torch.long, device=DEV).index_add_(0, tile_of[hb], torch.ones(int(hb.sum()), dtype=torch.long, device=DEV))
        busiest = hits.max(1).values
        q = lambda t, p: int(torch.quantile(t.float(), p).item())
        for name, t in (("intersections / tile", n_tile), ("records walked / tile", cnt), ("hits of the busiest warp / tile", busiest)):
            print(f"{name:34s} sum {int(t.sum()):9d} mean {t.float().mean().item():8.1f} p50 {q(t, .5):6d} p90 {q(t, .9):6d} "
                  f"p99 {q(t, .99):6d} max {int(t.max()):6d}")
        total_hits = int(hits.sum())
        print(f"warp iterations (block hits) {total_hits}; busiest warp {int(busiest.max())} = "
              f"{int(busiest.max()) / (total_hits / (148 * 4 * 5)):.2f} x the per-warp-slot average at 5 CTAs/SM")
    v_c = torch.ones(1, Hd, Wd, 3, device=DEV)
    v_a = torch.zeros(1, Hd, Wd, 1, device=DEV)
    scratch = torch.empty(N * 12, device=DEV)
    vm, vc, vcol, vo = torch.empty_like(m2d), torch.empty_like(con), torch.empty_like(cols), torch.empty_like(ops)

    def bwd():
        call("gsb200_rasterize_bwd", 1, N, I, 3, ptr(m2d), ptr(con), ptr(cols), ptr(ops), ptr(None), ptr(None), Wd, Hd,
             16, tw, th, ptr(offs), ptr(fl), ptr(ra), ptr(last), ptr(v_c), ptr(v_a), ptr(order), ptr(bstream), ptr(tcnt),
             ptr(scratch), ptr(None), ptr(vm),
             ptr(vc), ptr(vcol), ptr(vo), stream_ptr(DEV))

    res["blend_bwd"] = timeit(bwd)
    vdep = torch.zeros_like(dep)
    res["proj_bwd"] = timeit(lambda: W._projection_bwd(means, None, quats, scales, viewmats, Ks, Wd, Hd, 0.3, "pinhole",
                                                       radii, con, None, vm, vdep, vc, None, False))
    if ref is not None:
        cm = ref.CameraModelType.PINHOLE
        res["ref_proj_fwd"] = timeit(lambda: ref.fully_fused_projection_fwd(means, None, quats, scales, viewmats, Ks, Wd,
                                                                            Hd, 0.3, 0.01, 100.0, 3.0, False, cm))
        r_radii, r_m2d, r_dep, r_con, _ = ref.fully_fused_projection_fwd(means, None, quats, scales, viewmats, Ks, Wd, Hd,
                                                                         0.3, 0.01, 100.0, 3.0, False, cm)
        print("radii equal frac:", (r_radii == radii).float().mean().item(), "max diff", (r_radii - radii).abs().max().item())
        v = (r_radii > 0) & (radii > 0)
        print("depth bit-equal frac on valid:", (r_dep[v] == dep[v]).float().mean().item(),
              "means2d bit-equal:", (r_m2d[v] == m2d[v]).all(-1).float().mean().item(),
              "conics bit-equal:", (r_con[v] == con[v]).all(-1).float().mean().item())

        def ref_isect():
            t, i, f = ref.isect_tiles(m2d, radii, dep, None, None, 1, 16, tw, th, True, True)
            return ref.isect_offset_encode(i, 1, tw, th)

        res["ref_isect"] = timeit(ref_isect)
        res["ref_blend_fwd"] = timeit(lambda: ref.rasterize_to_pixels_fwd(m2d, con, cols, ops, None, None, Wd, Hd, 16, offs, fl))
        r_rc, r_ra, r_last = ref.rasterize_to_pixels_fwd(m2d, con, cols, ops, None, None, Wd, Hd, 16, offs, fl)
        fwd()
        torch.cuda.synchronize()
        print("blend fwd max abs diff vs ref:", (rc - r_rc).abs().max().item(), "alpha:", (ra - r_ra).abs().max().item(),
              "last_ids equal frac:", (last == r_last).float().mean().item())
        res["ref_blend_bwd"] = timeit(lambda: ref.rasterize_to_pixels_bwd(m2d, con, cols, ops, None, None, Wd, Hd, 16, offs,
                                                                          fl, r_ra, r_last, v_c, v_a, False))
        _, g_m, g_c, g_col, g_o = ref.rasterize_to_pixels_bwd(m2d, con, cols, ops, None, None, Wd, Hd, 16, offs, fl, r_ra,
                                                              r_last, v_c, v_a, False)
        bwd()
        torch.cuda.synchronize()
        for nme, a, b in (("v_means2d", vm, g_m), ("v_conics", vc, g_c), ("v_colors", vcol, g_col), ("v_opac", vo, g_o)):
            print(f"blend bwd {nme}: rel err vs ref {(a - b).abs().max().item() / b.abs().max().item():.3e}")
        res["ref_proj_bwd"] = timeit(lambda: ref.fully_fused_projection_bwd(means, None, quats, scales, viewmats, Ks, Wd, Hd,
                                                                            0.3, cm, radii, con, None, vm, vdep, vc, None,
                                                                            False))
    for k, v in res.items():
        print(f"{k:16s} {v:10.1f} us")
    mine = sum(res[k] for k in ("proj_fwd", "isect", "blend_fwd", "blend_bwd", "proj_bwd"))
    print(f"ours total {mine:.1f} us -> {Wd * Hd / mine:.1f} Mpix/s")
    if ref is not None:
        theirs = sum(res["ref_" + k] for k in ("proj_fwd", "isect", "blend_fwd", "blend_bwd", "proj_bwd"))
        print(f"ref  total {theirs:.1f} us -> {Wd * Hd / theirs:.1f} Mpix/s")


if __name__ == "__main__":
    main()
