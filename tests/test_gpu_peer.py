"""GPU tests of the multi-GPU exchange kernels (csrc/sh.cu, csrc/projection.cu: owner-side projection and colour
exchange, publish / gather / push / reduce of the geometry, the signal-pad barrier) driven from ONE process: the `peer` buffers are separate allocations
on this GPU, so the addressing, the slot layout and the rank-ordered sums are checked bit for bit against torch and the
single-array kernels.  The real two-process run over NVLink is tools/dist_check.py (gpurun --gpus 2)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gscodec_studio_b200._cabi import call, ptr, stream_ptr  # noqa: E402
from gscodec_studio_b200.cuda._wrapper import view_colors_from_sh_split  # noqa: E402

DEV = torch.device("cuda:0")


def table(tensors, float_offsets=None):
    """HOST array of device pointers (+ an offset in floats each)."""
    offs = float_offsets or [0] * len(tensors)
    arr = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() + 4 * o for t, o in zip(tensors, offs)])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def u32(values):
    arr = (ctypes.c_uint32 * len(values))(*values)
    return arr, ctypes.cast(arr, ctypes.c_void_p)


@pytest.mark.parametrize("world,C,deg", [(2, 1, 3), (3, 2, 3), (4, 1, 1), (2, 3, 2), (5, 1, 0)])
def test_owner_colour_exchange_single_process(world, C, deg):
    """masks + camera centres pushed to the owners, owner-side SH forward into the viewers' colour tensors, colour
    gradients pushed back, owner-side SH backward: equal, bit for bit, to ONE local kernel call that treats the job's
    V = world x C views as V cameras over the concatenated rows."""
    g = torch.Generator().manual_seed(5 + 31 * world + C)
    K = 16
    counts = [600 - 13 * r for r in range(world)]
    n_block = (max(counts) + 8 + 3) // 4 * 4
    n_total, V = world * n_block, world * C
    means = [torch.randn(n, 3, generator=g).to(DEV) for n in counts]
    sh0 = [(torch.randn(n, 1, 3, generator=g) * 0.7).to(DEV) for n in counts]
    shN = [(torch.randn(n, K - 1, 3, generator=g) * 0.7).to(DEV) for n in counts]
    campos = [(torch.randn(C, 3, generator=g) * 3).to(DEV) for _ in range(world)]
    radii = [((torch.rand(C, n_total, generator=g) > 0.55).int() * 3).to(DEV) for _ in range(world)]
    for q in range(world):
        for r in range(world):  # padding rows of every block are never visible
            radii[q][:, r * n_block + counts[r]:(r + 1) * n_block] = 0
    weights = [torch.randn(C, n_total, 3, generator=g).to(DEV) for _ in range(world)]

    # reference: one local call, all V views as cameras, rows padded to the block layout
    def padded(ts, shape):
        out = torch.zeros((n_total,) + shape, device=DEV)
        for r, t in enumerate(ts):
            out[r * n_block:r * n_block + counts[r]] = t
        return out

    m_all = padded(means, (3,)).requires_grad_(True)
    s0_all = padded(sh0, (1, 3)).requires_grad_(True)
    sN_all = padded(shN, (K - 1, 3)).requires_grad_(True)
    ref = view_colors_from_sh_split(deg, m_all, torch.cat(campos, 0), s0_all, sN_all, torch.cat(radii, 0))
    (ref * torch.cat(weights, 0)).sum().backward()
    ref = ref.detach()

    # the emulated ranks' symmetric regions
    masks = [torch.full((V, n_block), -7, dtype=torch.int32, device=DEV) for _ in range(world)]
    cams = [torch.full((V, 3), float("nan"), device=DEV) for _ in range(world)]
    vcols = [torch.full((V, n_block, 3), float("nan"), device=DEV) for _ in range(world)]
    colours = [torch.full((C, n_total, 3), float("nan"), device=DEV) for _ in range(world)]
    mask_tab, cam_tab, vcol_tab = table(masks), table(cams), table(vcols)
    s = stream_ptr(DEV)
    for q in range(world):  # viewer q -> every owner
        call("gsb200_shard_push_rows", world, n_block, 1, C, q * C, n_block, ptr(radii[q]), mask_tab[1], ptr(campos[q]), 3,
             cam_tab[1], s)
    for r in range(world):
        assert torch.equal(cams[r], torch.cat(campos, 0))
        for q in range(world):
            assert torch.equal(masks[r][q * C:(q + 1) * C], radii[q][:, r * n_block:(r + 1) * n_block])
    for r in range(world):  # owner r evaluates its rows for all views
        col_tab = table([colours[v // C] for v in range(V)], [((v % C) * n_total + r * n_block) * 3 for v in range(V)])
        call("gsb200_sh_owner_fwd", V, counts[r], n_block, K, deg, ptr(means[r]), ptr(cams[r]), ptr(sh0[r]), ptr(shN[r]),
             ptr(masks[r]), col_tab[1], s)
    for q in range(world):
        vis = radii[q] > 0
        assert torch.equal(colours[q][vis], ref[q * C:(q + 1) * C][vis])
        assert torch.isnan(colours[q][~vis]).all()  # nothing travels for an invisible pair
    for q in range(world):  # backward: viewer q sends its colour gradients home
        call("gsb200_shard_push_rows", world, n_block, 3, C, q * C, n_block * 3, ptr(weights[q]), vcol_tab[1], None, 0, None, s)
    for r in range(world):
        n = counts[r]
        v0 = torch.full((n, 1, 3), float("nan"), device=DEV)
        vN = torch.full((n, K - 1, 3), float("nan"), device=DEV)
        vm = torch.full((n, 3), float("nan"), device=DEV)
        call("gsb200_sh_owner_bwd", V, n, n_block, K, deg, ptr(means[r]), ptr(cams[r]), ptr(sh0[r]), ptr(shN[r]), ptr(masks[r]),
             ptr(vcols[r]), ptr(v0), ptr(vN), ptr(vm), s)
        rows = slice(r * n_block, r * n_block + n)
        assert torch.equal(v0, s0_all.grad[rows]) and torch.equal(vN, sN_all.grad[rows]) and torch.equal(vm, m_all.grad[rows])
    with pytest.raises(RuntimeError):  # more views than GSB200_MAX_VIEWS
        call("gsb200_sh_owner_fwd", 65, 1, n_block, K, deg, ptr(means[0]), ptr(cams[0]), ptr(sh0[0]), ptr(shN[0]), ptr(masks[0]),
             col_tab[1], s)
    with pytest.raises(RuntimeError):  # blocks must be whole 16-byte vectors
        call("gsb200_shard_push_rows", world, n_block + 1, 1, C, 0, n_block + 4, ptr(radii[0]), mask_tab[1], None, 0, None, s)


@pytest.mark.parametrize("world,covars", [(2, False), (3, False), (4, True)])
def test_geometry_exchange_single_process(world, covars):
    """publish (attribute tensors -> part-major block, inert padding rows, sh0 | shN rows side by side) -> gather -> push of
    the gradients -> owner-side reduce, against torch.cat / sums in rank order (bit-exact)."""
    g = torch.Generator().manual_seed(11 + world)
    widths = [3, 6, 1] if covars else [3, 4, 3, 1]
    w = sum(widths)
    counts = [900 - 17 * r for r in range(world)]
    n_block = (max(counts) + 3) // 4 * 4 + 4
    n_total = world * n_block
    pad_row = torch.randn(w, generator=g).to(DEV)
    parts = [[torch.randn(n, wk, generator=g).to(DEV) for wk in widths] for n in counts]
    K = 16
    sh0 = [torch.randn(n, 1, 3, generator=g).to(DEV) for n in counts]
    shN = [torch.randn(n, K - 1, 3, generator=g).to(DEV) for n in counts]
    blocks = [torch.full((n_block * w,), float("nan"), device=DEV) for _ in range(world)]
    coeffs = [torch.full((n_block, K * 3), float("nan"), device=DEV) for _ in range(world)]
    s = stream_ptr(DEV)
    keep_w = u32(widths)
    for r in range(world):
        call("gsb200_shard_publish", counts[r], n_block, len(widths), keep_w[1], table(parts[r])[1], ptr(pad_row), ptr(blocks[r]),
             ptr(sh0[r]), ptr(shN[r]), 3, K * 3, ptr(coeffs[r]), s)
        assert torch.equal(coeffs[r][:counts[r]], torch.cat([sh0[r], shN[r]], 1).reshape(counts[r], K * 3))
        assert torch.isnan(coeffs[r][counts[r]:]).all()
    outs = [torch.full((n_total, wk), float("nan"), device=DEV) for wk in widths]
    call("gsb200_shard_gather_geo", world, n_block, len(widths), keep_w[1], table(blocks)[1], table(outs)[1], s)
    col = 0
    for k, wk in enumerate(widths):
        expect = pad_row[col:col + wk].expand(n_total, wk).clone()
        for r in range(world):
            expect[r * n_block:r * n_block + counts[r]] = parts[r][k]
        assert torch.equal(outs[k], expect), k
        col += wk
    # geometry-only publish (the owner-side colour exchange): no coefficient pointers
    blk = torch.full((n_block * w,), float("nan"), device=DEV)
    call("gsb200_shard_publish", counts[0], n_block, len(widths), keep_w[1], table(parts[0])[1], ptr(pad_row), ptr(blk), None,
         None, 0, 4, None, s)
    assert torch.equal(blk, blocks[0])

    # backward: every source rank p pushes its gradient rows (part 1 has no gradient on odd ranks), the owners reduce
    slot = n_block * w + 8
    stagings = [torch.full((world, slot), float("nan"), device=DEV) for _ in range(world)]
    grads = [[None if (k == 1 and p % 2 == 1) else torch.randn(n_total, wk, generator=g).to(DEV) for k, wk in enumerate(widths)]
             for p in range(world)]
    for p in range(world):
        call("gsb200_shard_push_geo", world, n_block, p, slot, len(widths), keep_w[1], table(grads[p])[1], table(stagings)[1], s)
    for r in range(world):
        res = [torch.full((counts[r], wk), float("nan"), device=DEV) for wk in widths]
        call("gsb200_shard_reduce_geo", world, n_block, counts[r], slot, len(widths), keep_w[1], ptr(stagings[r]), table(res)[1], s)
        for k, wk in enumerate(widths):
            expect = torch.zeros(counts[r], wk, device=DEV)
            for p in range(world):
                if grads[p][k] is not None:
                    expect = expect + grads[p][k][r * n_block:r * n_block + counts[r]]
            assert torch.equal(res[k], expect), (r, k)
    # an unaligned gradient tensor takes the scalar-load path of the push
    odd = torch.randn(n_total * 3 + 1, generator=g).to(DEV)[1:].view(n_total, 3)
    g2 = [odd] + [None] * (len(widths) - 1)
    call("gsb200_shard_push_geo", world, n_block, 0, slot, len(widths), keep_w[1], table(g2)[1], table(stagings)[1], s)
    for r in range(world):
        assert torch.equal(stagings[r][0, :n_block * 3].view(n_block, 3), odd[r * n_block:(r + 1) * n_block])
    with pytest.raises(RuntimeError):
        call("gsb200_shard_gather_geo", world, n_block + 2, len(widths), keep_w[1], table(blocks)[1], table(outs)[1], s)


@pytest.mark.parametrize("world,C,model", [(2, 1, "pinhole"), (3, 2, "pinhole"), (2, 2, "fisheye"), (4, 1, "ortho")])
def test_owner_projection_single_process(world, C, model):
    """Owner-side stage 1 (every emulated rank projects ITS rows for all V = world x C views into the viewers' tensors, then
    runs the projection backward on gradients the viewers pushed) against the single-GPU kernels on the concatenated rows:
    radii equal, visible rows of means2d / depths / conics / opacities bit-equal, invisible rows untouched, the cameras
    delivered by peer_push_small, gradients of means / quats / scales bit-equal."""
    import gscodec_studio_b200 as gs
    from gscodec_studio_b200._helper import load_test_data
    from gscodec_studio_b200.cuda._wrapper import _CAMERA_MODELS

    means, quats, scales, opac, _colors, viewmats, Ks, W, H = load_test_data(device=DEV, scene_grid=1)
    g = torch.Generator().manual_seed(3 + world)
    n_all = 4000
    pick = torch.randperm(means.shape[0], generator=g)[:n_all].to(DEV)
    means, quats, scales, opac = means[pick].contiguous(), quats[pick].contiguous(), scales[pick].contiguous(), opac[pick].contiguous()
    V = world * C
    vm_all = viewmats[torch.arange(V) % viewmats.shape[0]].clone()
    vm_all[:, 0, 3] += torch.linspace(-0.2, 0.2, V, device=DEV)  # V distinct cameras
    ks_all = Ks[torch.arange(V) % Ks.shape[0]].clone().contiguous()
    vm_all = vm_all.contiguous()
    counts = [n_all // world - 5 * r for r in range(world)]
    n_block = (max(counts) + 7) // 4 * 4
    n_total = world * n_block
    starts = [sum(counts[:r]) for r in range(world)]
    shard = lambda t, r: t[starts[r]:starts[r] + counts[r]].contiguous()  # noqa: E731
    s = stream_ptr(DEV)

    # reference: one local projection over the block-padded concatenation (padding rows: NaN mean -> culled)
    def padded(t, fill):
        out = torch.full((n_total,) + tuple(t.shape[1:]), fill, device=DEV)
        for r in range(world):
            out[r * n_block:r * n_block + counts[r]] = shard(t, r)
        return out

    m_ref = padded(means, float("nan")).requires_grad_(True)
    q_ref = padded(quats, 0.0)
    q_ref[:, 0] = torch.where(q_ref.abs().sum(1) == 0, torch.ones_like(q_ref[:, 0]), q_ref[:, 0])
    q_ref.requires_grad_(True)
    s_ref = padded(scales, 0.0).requires_grad_(True)
    radii_r, m2d_r, dep_r, con_r, _ = gs.fully_fused_projection(m_ref, None, q_ref, s_ref, vm_all, ks_all, W, H, camera_model=model)
    w_m2d = torch.randn(V, n_total, 2, generator=g).to(DEV)
    w_dep = torch.randn(V, n_total, generator=g).to(DEV)
    w_con = torch.randn(V, n_total, 3, generator=g).to(DEV)
    ((m2d_r * w_m2d).sum() + (dep_r * w_dep).sum() + (con_r * w_con).sum()).backward()

    # cameras: every rank pushes its C cameras to every peer
    cam_vm = [torch.full((V, 16), float("nan"), device=DEV) for _ in range(world)]
    cam_k = [torch.full((V, 9), float("nan"), device=DEV) for _ in range(world)]
    for q in range(world):
        call("gsb200_peer_push_small", world, q * C, C, 16, ptr(vm_all[q * C:(q + 1) * C].contiguous()), table(cam_vm)[1], s)
        call("gsb200_peer_push_small", world, q * C, C, 9, ptr(ks_all[q * C:(q + 1) * C].contiguous()), table(cam_k)[1], s)
    for r in range(world):
        assert torch.equal(cam_vm[r].view(V, 4, 4), vm_all) and torch.equal(cam_k[r].view(V, 3, 3), ks_all)

    # viewer-side tensors (rank q: its C cameras over the whole gathered set), owner-side copies
    v_radii = [torch.full((C, n_total), -9, dtype=torch.int32, device=DEV) for _ in range(world)]
    v_m2d = [torch.full((C, n_total, 2), float("nan"), device=DEV) for _ in range(world)]
    v_dep = [torch.full((C, n_total), float("nan"), device=DEV) for _ in range(world)]
    v_con = [torch.full((C, n_total, 3), float("nan"), device=DEV) for _ in range(world)]
    v_opa = [torch.full((C, n_total), float("nan"), device=DEV) for _ in range(world)]
    radii_loc, conics_loc = [], []
    for r in range(world):
        tabs = [table([arr[v // C] for v in range(V)], [((v % C) * n_total + r * n_block) * w for v in range(V)])
                for arr, w in ((v_radii, 1), (v_m2d, 2), (v_dep, 1), (v_con, 3), (v_opa, 1))]
        rl = torch.full((V, n_block), -5, dtype=torch.int32, device=DEV)
        cl = torch.full((V, n_block, 3), float("nan"), device=DEV)
        call("gsb200_projection_owner_fwd", V, counts[r], n_block, ptr(shard(means, r)), ptr(shard(quats, r)), ptr(shard(scales, r)),
             ptr(shard(opac, r)), ptr(cam_vm[r]), ptr(cam_k[r]), W, H, 0.3, 0.01, 1e10, 0.0, _CAMERA_MODELS[model], ptr(rl), ptr(cl),
             tabs[0][1], tabs[1][1], tabs[2][1], tabs[3][1], tabs[4][1], s)
        radii_loc.append(rl)
        conics_loc.append(cl)
    opac_full = padded(opac[:, None], 0.0)[:, 0]
    for q in range(world):
        rows = slice(q * C, (q + 1) * C)
        assert torch.equal(v_radii[q], radii_r[rows])
        vis = radii_r[rows] > 0
        assert vis.sum() > 100
        assert torch.equal(v_m2d[q][vis], m2d_r[rows][vis]) and torch.equal(v_dep[q][vis], dep_r[rows][vis])
        assert torch.equal(v_con[q][vis], con_r[rows][vis])
        assert torch.equal(v_opa[q][vis], opac_full[None].expand(C, -1)[vis])
        assert torch.isnan(v_m2d[q][~vis]).all() and torch.isnan(v_con[q][~vis]).all()  # nothing travels for an invisible pair
    for r in range(world):
        assert torch.equal(radii_loc[r], radii_r[:, r * n_block:(r + 1) * n_block])

    # backward: every viewer pushes its gradient rows to the owners, the owners run the projection backward
    st_m2d = [torch.full((V, n_block, 2), float("nan"), device=DEV) for _ in range(world)]
    st_dep = [torch.full((V, n_block), float("nan"), device=DEV) for _ in range(world)]
    st_con = [torch.full((V, n_block, 3), float("nan"), device=DEV) for _ in range(world)]
    for q in range(world):
        rows = slice(q * C, (q + 1) * C)
        for src, w, dst in ((w_m2d, 2, st_m2d), (w_dep, 1, st_dep), (w_con, 3, st_con)):
            call("gsb200_shard_push_rows", world, n_block, w, C, q * C, n_block * w, ptr(src[rows].contiguous()), table(dst)[1],
                 None, 0, None, s)
    for r in range(world):
        n = counts[r]
        vm_, vq_, vs_ = (torch.full((n, k), float("nan"), device=DEV) for k in (3, 4, 3))
        call("gsb200_projection_owner_bwd", V, n, n_block, ptr(shard(means, r)), ptr(shard(quats, r)), ptr(shard(scales, r)),
             ptr(cam_vm[r]), ptr(cam_k[r]), W, H, 0.3, _CAMERA_MODELS[model], ptr(radii_loc[r]), ptr(conics_loc[r]), ptr(st_m2d[r]),
             ptr(st_dep[r]), ptr(st_con[r]), ptr(vm_), ptr(vq_), ptr(vs_), s)
        rows = slice(r * n_block, r * n_block + n)
        assert torch.equal(vm_, m_ref.grad[rows]) and torch.equal(vq_, q_ref.grad[rows]) and torch.equal(vs_, s_ref.grad[rows])


def test_peer_barrier_and_camera_positions():
    """world = 1 barrier (signals and waits on its own pad, epochs only grow); camera centres = -R^T t."""
    pad = torch.zeros(64, dtype=torch.int32, device=DEV)
    tab = table([pad])
    for epoch in (1, 2, 3, 0x7FFFFFFF, 0x80000000):
        call("gsb200_peer_barrier", 1, 0, tab[1], epoch, stream_ptr(DEV))
        torch.cuda.synchronize()
        assert int(pad[0].item()) & 0xFFFFFFFF == epoch
    with pytest.raises(RuntimeError):
        call("gsb200_peer_barrier", 17, 0, tab[1], 1, stream_ptr(DEV))
    g = torch.Generator().manual_seed(3)
    q, _ = torch.linalg.qr(torch.randn(5, 3, 3, generator=g))
    vm = torch.eye(4).repeat(5, 1, 1)
    vm[:, :3, :3] = q
    vm[:, :3, 3] = torch.randn(5, 3, generator=g)
    vm = vm.to(DEV)
    out = torch.empty(5, 3, device=DEV)
    call("gsb200_camera_positions", 5, ptr(vm), ptr(out), stream_ptr(DEV))
    expect = torch.inverse(vm.double())[:, :3, 3]
    assert torch.allclose(out.double(), expect, rtol=1e-5, atol=1e-6)
