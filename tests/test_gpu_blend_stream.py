"""GPU tests of the producer/consumer stream blend (csrc/blend_stream.cu) against the generic per-tile kernels
(csrc/blend.cu) through the C-ABI: the forward must be BIT-identical (same per-pixel arithmetic, same pairs), the
backward equal up to the summation order of the atomics.  The parity of both against the reference CUDA kernels and
the oracle is in test_gpu_parity.py; this file pins the things only the stream path has: the mbarrier ring wrapping
many times, early termination, edge tiles, masks, backgrounds, the record stream handed to the backward."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200 import _cabi  # noqa: E402
from gscodec_studio_b200._cabi import call, ptr, stream_ptr  # noqa: E402
from gscodec_studio_b200._helper import load_test_data  # noqa: E402
from gscodec_studio_b200.cuda._wrapper import _isect_tiles_with_offsets  # noqa: E402

DEV = torch.device("cuda:0")


class blend_mode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.old = _cabi.load_library().gsb200_set_blend_mode(self.mode)

    def __exit__(self, *a):
        _cabi.load_library().gsb200_set_blend_mode(self.old)


def _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a, bg=None, masks=None, absgrad=False):
    """One forward + backward through the raw C-ABI (so that last_ids is visible)."""
    C, th, tw = offs.shape
    D = cols.shape[-1]
    M = opacs.numel()
    I = fl.numel()
    lib = _cabi.load_library()
    rc = torch.zeros(C, H, W, D, device=DEV)
    ra = torch.zeros(C, H, W, 1, device=DEV)
    last = torch.zeros(C, H, W, dtype=torch.int32, device=DEV)
    order = torch.empty(max(C * th * tw, 1), dtype=torch.int32, device=DEV)
    rec = int(lib.gsb200_blend_stream_record(D, 16))
    bstream = torch.empty(max(I, 1) * rec, device=DEV) if rec else None
    tcnt = torch.full((max(C * th * tw, 1),), -7, dtype=torch.int32, device=DEV) if rec else None
    m8 = None if masks is None else masks.to(torch.uint8).contiguous()
    call("gsb200_rasterize_fwd", C, M, I, D, ptr(m2d), ptr(con), ptr(cols), ptr(opacs), ptr(bg), ptr(m8), W, H, 16, tw, th,
         ptr(offs), ptr(fl), ptr(rc), ptr(ra), ptr(last), ptr(order), ptr(bstream), ptr(tcnt), stream_ptr(DEV))
    scratch = torch.empty(max(M, 1) * int(lib.gsb200_grad_record(D)), device=DEV)
    vm, vc, vcol, vo = torch.empty_like(m2d), torch.empty_like(con), torch.empty_like(cols), torch.empty_like(opacs)
    vabs = torch.empty_like(m2d) if absgrad else None
    call("gsb200_rasterize_bwd", C, M, I, D, ptr(m2d), ptr(con), ptr(cols), ptr(opacs), ptr(bg), ptr(m8), W, H, 16, tw, th,
         ptr(offs), ptr(fl), ptr(ra), ptr(last), ptr(v_c), ptr(v_a), ptr(order), ptr(bstream), ptr(tcnt), ptr(scratch),
         ptr(vabs), ptr(vm), ptr(vc), ptr(vcol), ptr(vo), stream_ptr(DEV))
    torch.cuda.synchronize()
    return dict(rc=rc, ra=ra, last=last, vm=vm, vc=vc, vcol=vcol, vo=vo, vabs=vabs, tcnt=tcnt)


def _compare(a, b, pix_mask=None):
    for k in ("rc", "ra", "last"):
        x, y = a[k], b[k]
        if pix_mask is not None and k != "rc":
            x, y = x.reshape(pix_mask.shape)[pix_mask], y.reshape(pix_mask.shape)[pix_mask]
        assert torch.equal(x, y), f"{k}: stream and tile kernels differ ({(x != y).sum().item()} entries)"
    for k in ("vm", "vc", "vcol", "vo", "vabs"):
        if a[k] is None:
            continue
        scale = b[k].abs().max().item()
        err = (a[k] - b[k]).abs().max().item() / max(scale, 1e-20)
        assert err < 2e-5, f"{k}: rel err {err:.2e}"  # summation order only


def _scene_inputs(C, channels, width=None, height=None, scene_grid=1):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = load_test_data(device="cpu", scene_grid=scene_grid)
    viewmats, Ks = viewmats[:C].clone(), Ks[:C].clone()
    if width is not None:
        Ks[:, 0, :] *= width / W
        Ks[:, 1, :] *= height / H
        W, H = width, height
    d = [t.to(DEV) for t in (means, quats, scales, viewmats, Ks)]
    radii, m2d, dep, con, _ = gs.fully_fused_projection(d[0], None, d[1], d[2], d[3], d[4], W, H)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, _, fl, offs = _isect_tiles_with_offsets(m2d, radii, dep, 16, tw, th)
    g = torch.Generator().manual_seed(17)
    N = means.shape[0]
    cols = (colors[None].expand(C, -1, -1) if channels == 3 else torch.rand(C, N, channels, generator=g)).contiguous().to(DEV)
    opacs = opac[None].expand(C, -1).contiguous().to(DEV)
    v_c = torch.randn(C, H, W, channels, generator=g).to(DEV)
    v_a = torch.randn(C, H, W, 1, generator=g).to(DEV)
    return m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a


@pytest.mark.parametrize("channels,size,bg,absgrad", [(3, None, False, False), (3, (645, 419), True, True), (1, (333, 250), False, False),
                                                      (4, None, True, False), (2, (1000, 562), False, True)])
def test_stream_blend_equals_tile_kernels(channels, size, bg, absgrad):
    assert int(_cabi.load_library().gsb200_blend_stream_record(channels, 16)) == 12
    w, h = size if size else (None, None)
    m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a = _scene_inputs(2, channels, w, h)
    g = torch.Generator().manual_seed(3)
    bgt = torch.rand(2, channels, generator=g).to(DEV) if bg else None
    with blend_mode(0):
        a = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a, bg=bgt, absgrad=absgrad)
    with blend_mode(1):
        b = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a, bg=bgt, absgrad=absgrad)
    _compare(a, b)
    # the stream the backward consumed: every tile's record count is within its intersection range
    C, th, tw = offs.shape
    ends = torch.cat([offs.view(-1)[1:], torch.tensor([fl.numel()], dtype=torch.int32, device=DEV)])
    cnt = a["tcnt"][:C * th * tw]
    assert (cnt >= 0).all() and (cnt <= ends - offs.view(-1)).all()


def test_stream_blend_masks():
    m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a = _scene_inputs(1, 3)
    g = torch.Generator().manual_seed(5)
    masks = (torch.rand(offs.shape, generator=g) > 0.4).to(DEV)
    with blend_mode(0):
        a = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a, masks=masks)
    with blend_mode(1):
        b = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a, masks=masks)
    pix = masks.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W]
    _compare(a, b, pix_mask=pix)


@pytest.mark.parametrize("n,opacity", [(6000, 0.004), (6000, 0.05), (70, 0.9), (64, 0.01), (65, 0.01), (1, 0.5)])
def test_stream_blend_long_lists_and_ring_wrap(n, opacity):
    """One 2x2-tile image whose tiles hold up to thousands of faint Gaussians: the 4-stage ring wraps dozens of times in
    both directions; opaque lists terminate early instead."""
    g = torch.Generator().manual_seed(n)
    W = H = 30  # edge tiles: 14 of 16 columns / rows are inside
    m2d = (torch.rand(1, n, 2, generator=g) * 36 - 3).to(DEV)
    s = torch.rand(1, n, generator=g) * 0.02 + 0.002
    con = torch.stack([s, (torch.rand(1, n, generator=g) - 0.5) * s, s * (1 + torch.rand(1, n, generator=g))], -1).to(DEV)
    radii = torch.full((1, n), 40, dtype=torch.int32, device=DEV)
    dep = torch.rand(1, n, generator=g).to(DEV)
    _, _, fl, offs = _isect_tiles_with_offsets(m2d, radii, dep, 16, 2, 2)
    assert fl.numel() == 4 * n
    cols = torch.rand(1, n, 3, generator=g).to(DEV)
    opacs = (torch.rand(1, n, generator=g) * opacity + opacity).to(DEV)
    v_c = torch.randn(1, H, W, 3, generator=g).to(DEV)
    v_a = torch.randn(1, H, W, 1, generator=g).to(DEV)
    with blend_mode(0):
        a = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a)
    with blend_mode(1):
        b = _raw_fwd_bwd(m2d, con, cols, opacs, offs, fl, W, H, v_c, v_a)
    _compare(a, b)


def test_stream_blend_no_intersections_and_inference_only():
    m2d = torch.zeros(1, 4, 2, device=DEV)
    con = torch.ones(1, 4, 3, device=DEV)
    cols = torch.ones(1, 4, 3, device=DEV)
    opacs = torch.ones(1, 4, device=DEV)
    offs = torch.zeros(1, 2, 2, dtype=torch.int32, device=DEV)
    fl = torch.zeros(0, dtype=torch.int32, device=DEV)
    bg = torch.tensor([[0.25, 0.5, 0.75]], device=DEV)
    rc, ra = gs.rasterize_to_pixels(m2d, con, cols, opacs, 20, 20, 16, offs, fl, backgrounds=bg)
    assert torch.equal(rc, bg.view(1, 1, 1, 3).expand(1, 20, 20, 3)) and (ra == 0).all()
    # no gradient requested: the forward runs without a record stream
    m2d, con, cols, opacs, offs, fl, W, H, _, _ = _scene_inputs(1, 3)
    with torch.no_grad():
        with blend_mode(0):
            rc0, ra0 = gs.rasterize_to_pixels(m2d, con, cols, opacs, W, H, 16, offs, fl)
        with blend_mode(1):
            rc1, ra1 = gs.rasterize_to_pixels(m2d, con, cols, opacs, W, H, 16, offs, fl)
    assert torch.equal(rc0, rc1) and torch.equal(ra0, ra1)
