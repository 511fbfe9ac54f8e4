"""Peer-memory (NVLink / NVSwitch) exchange of the SH coefficient shards for view-parallel rendering.

The reference moves whole attribute tensors between ranks with NCCL (gsplat/distributed.py:90-215, called from
rendering.py:282-294 and :397-478).  For the colour stage that is 16x3 floats per Gaussian per direction -- 193 MB
for 1 M splats -- although a view only touches the ~30 % of them it can see.  Here every rank keeps its block of
coefficients in symmetric memory (one allocation, mapped into every peer's address space over NVLink) and

  forward   sh_view_fwd_sharded PULLS the rows of visible Gaussians straight from their owners;
  backward  sh_view_bwd_sharded PUSHES the coefficient-gradient rows of visible Gaussians into the owner's staging
            array (slot [source rank][row]; posted NVLink stores, coalesced per row) plus a 0/1 flag per row, and
            after a cross-rank barrier each owner runs shard_reduce_rows over its LOCAL staging: the flagged rows
            are summed in source-rank order, so the result is reproducible.

No all-gather of coefficients, no dense reduce-scatter of their gradients, and the backward kernel stops writing
zero rows for invisible Gaussians.  torch.distributed._symmetric_memory provides the allocation, the peer pointers
and the stream-ordered barrier; the kernels are ours (csrc/sh.cu).  Geometry (11 floats per Gaussian, needed for
EVERY row by the projection) keeps travelling by NCCL all-gather / reduce-scatter (distributed.py).
"""
from __future__ import annotations

import ctypes
import os
import warnings
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from ._cabi import call, device_guard, ptr, stream_ptr

MAX_SHARDS = 16  # GSB200_MAX_SHARDS in include/gsb200.h

_workspaces: Dict[Tuple[int, int], "PeerWorkspace"] = {}
_unavailable: Optional[str] = None


def _ptr_table(ptrs: Sequence[int], byte_offset: int):
    arr = (ctypes.c_void_p * len(ptrs))(*[p + byte_offset for p in ptrs])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


GEOMETRY_ONLY_ROW = 4  # row width of the workspace that only carries geometry
GEO_WIDTH = 12  # floats reserved per row of the geometry block (means 3 + quats 4 + scales 3 + opacity 1, or covars)


class PeerWorkspace:
    """One symmetric allocation per (device, row width), carved into
         [ coefficient block  cap_block x row_floats ]   this rank's published SH rows
         [ geometry block     cap_block x GEO_WIDTH ]    this rank's published geometry rows (packed [n_block, w])
         [ gradient staging   world x cap_block x row_floats ]   slot [p] is written by rank p (push)
         [ row flags          world x cap_block (int32) ]
         [ geometry-gradient staging   world x cap_block x GEO_WIDTH ]   slot [p] is written by rank p (push)
    Every rank uses the same capacities, so a region sits at the same offset in every peer's buffer."""

    def __init__(self, device: torch.device, row_floats: int, cap_block: int):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.row_floats, self.cap_block = row_floats, cap_block
        n_coeff = cap_block * row_floats
        n_geo = cap_block * GEO_WIDTH
        n_stage = self.world * cap_block * row_floats
        n_flags = self.world * cap_block
        n_gstage = self.world * cap_block * GEO_WIDTH  # geometry-gradient staging: slot [p] is written by rank p
        n_pads = 4 * MAX_SHARDS  # signal pad of the barrier kernel (uint32 words; the first MAX_SHARDS are used)
        self.buf = symm_mem.empty(n_coeff + n_geo + n_stage + n_flags + n_gstage + n_pads, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world and all(p % 16 == 0 for p in ptrs), ptrs
        o_geo, o_stage, o_flags = n_coeff, n_coeff + n_geo, n_coeff + n_geo + n_stage
        self.coeff = self.buf[:o_geo].view(cap_block, row_floats)
        self.geo = self.buf[o_geo:o_stage]
        self.staging = self.buf[o_stage:o_flags]
        o_gstage = o_flags + n_flags
        self.flags = self.buf[o_flags:o_gstage].view(torch.int32)
        o_pads = o_gstage + n_gstage
        self.geo_staging = self.buf[o_gstage:o_pads]
        self.buf[o_pads:].zero_()
        self._pad_tab = _ptr_table(ptrs, 4 * o_pads)
        self.epoch = 0
        self.hdl.barrier(channel=0)  # every pad is zero before the first signal can arrive
        self._gstage_tab = _ptr_table(ptrs, 4 * o_gstage)
        self._coeff_tab = _ptr_table(ptrs, 0)
        self._geo_tab = _ptr_table(ptrs, 4 * o_geo)
        self._stage_tab = _ptr_table(ptrs, 4 * o_stage)
        self._flag_tab = _ptr_table(ptrs, 4 * o_flags)
        self.version = 0      # bumped by every publish()
        self.readers = False  # a pull kernel was launched since the last barrier: peers may still be reading

    def barrier(self) -> None:
        """Stream-ordered barrier over all ranks (a small kernel on the current stream; the host does not block).
        Our own signal-pad kernel: the symmetric-memory handle's barrier costs ~150 us of HOST time per call, four
        calls per step -- enough to let the host fall behind the device."""
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF
        call("gsb200_peer_barrier", self.world, self.rank, self._pad_tab[1], self.epoch, stream_ptr(self.buf.device))
        self.readers = False

    def publish(self, coeff_rows: Tensor, geo_parts: Optional[Sequence[Tensor]] = None, pad_row: Optional[Tensor] = None,
                n_block: int = 0) -> int:
        """Copy this rank's rows into the symmetric buffer and make them visible to every peer (ONE barrier for
        both blocks).  Returns the version number of what is now published."""
        if self.readers:
            self.barrier()
        # one launch writes both blocks: the trainers' (sh0 [n,1,3], shN [n,K-1,3]) land side by side, the [n,K,3]
        # concatenation never exists outside this buffer
        first, rest = coeff_rows if isinstance(coeff_rows, (tuple, list)) else (coeff_rows, None)
        if first is None:  # geometry only (the owner-side colour exchange keeps the coefficients at home)
            n, first_floats = geo_parts[0].shape[0], 0
        else:
            first = first.detach().float().contiguous()
            rest = None if rest is None else rest.detach().float().contiguous()
            n = first.shape[0]
            first_floats = first.numel() // max(n, 1)
            assert first_floats + (0 if rest is None else rest.numel() // max(n, 1)) == self.row_floats
        assert n <= self.cap_block
        n_parts, w_ptr, p_ptr, keep = 0, None, None, None
        if geo_parts is not None:
            parts = [t.detach().float().contiguous() for t in geo_parts]
            widths = [t.shape[1] for t in parts]
            assert sum(widths) <= GEO_WIDTH and n_block <= self.cap_block and all(t.shape[0] == n for t in parts)
            keep = (parts, _u32_array(widths), _ptr_array(parts))
            n_parts, w_ptr, p_ptr = len(parts), keep[1][1], keep[2][1]
            pad_row = None if pad_row is None else pad_row.detach().float().contiguous()
        call("gsb200_shard_publish", n, n_block if geo_parts is not None else n, n_parts, w_ptr, p_ptr,
             ptr(pad_row) if geo_parts is not None else None, ptr(self.geo), ptr(first), ptr(rest), first_floats,
             self.row_floats, ptr(self.coeff), stream_ptr(self.buf.device))
        self.barrier()
        self.version += 1
        return self.version

    coeff_table = property(lambda self: self._coeff_tab[1])
    geo_table = property(lambda self: self._geo_tab[1])
    stage_table = property(lambda self: self._stage_tab[1])
    flag_table = property(lambda self: self._flag_tab[1])
    geo_stage_table = property(lambda self: self._gstage_tab[1])


def workspace(device: torch.device, row_floats: int, n_block: int) -> PeerWorkspace:
    """The workspace for this device and row width, grown (collectively: every rank sees the same n_block) when a
    block no longer fits."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), row_floats)
    ws = _workspaces.get(key)
    if ws is None or ws.cap_block < n_block or ws.world != dist.get_world_size():
        if ws is not None:
            ws.barrier()
            _workspaces.pop(key)
            del ws
        cap = (int(n_block * 1.1) + 64 + 3) // 4 * 4
        _workspaces[key] = ws = PeerWorkspace(device, row_floats, cap)
    return ws


def available(device: torch.device, row_floats: int, n_block: int) -> bool:
    """True if the peer-memory path can be used: NCCL job on CUDA, <= MAX_SHARDS ranks, 16-byte rows, and a
    symmetric allocation that works on this machine.  GSB200_PEER_EXCHANGE=0 forces the NCCL collectives."""
    global _unavailable
    if _unavailable is not None:
        return False
    if os.environ.get("GSB200_PEER_EXCHANGE", "1") == "0":
        _unavailable = "disabled by GSB200_PEER_EXCHANGE=0"
        return False
    if device.type != "cuda" or dist.get_backend() != "nccl":
        return False
    if dist.get_world_size() > MAX_SHARDS or row_floats % 4 != 0:
        return False
    try:
        workspace(device, row_floats, n_block)
    except Exception as e:  # no P2P mapping on this machine: every rank fails alike and falls back to NCCL
        _unavailable = f"{type(e).__name__}: {e}"
        warnings.warn(f"peer-memory exchange unavailable ({_unavailable}); using NCCL collectives")
        return False
    return True


class _ShViewSharded(torch.autograd.Function):
    """coeffs_local: this rank's [n_local,K,3] coefficients, or -- with coeffs_rest given -- the trainers' sh0 [n,1,3]
    next to shN [n,K-1,3]; gradients come back in the same form."""

    @staticmethod
    def forward(ctx, sh_degree, means, campos, coeffs_local, coeffs_rest, radii, n_block, published):
        C, N = campos.shape[0], means.shape[0]
        K = coeffs_local.shape[1] + (coeffs_rest.shape[1] if coeffs_rest is not None else 0)
        ws = workspace(means.device, K * 3, n_block)
        colors = torch.empty((C, N, 3), dtype=torch.float32, device=means.device)
        with device_guard(means.device):
            if published is None or published != ws.version:
                ws.publish(coeffs_local if coeffs_rest is None else (coeffs_local, coeffs_rest))
            # local copy of every pulled row: the backward needs the same rows again (d colour / d direction)
            row_cache = torch.empty((N, K, 3), dtype=torch.float32, device=means.device) if any(ctx.needs_input_grad) else None
            call("gsb200_sh_view_fwd_sharded", C, N, K, sh_degree, ptr(means), ptr(campos), ws.coeff_table, ws.world, n_block,
                 ptr(radii), ptr(colors), ptr(row_cache), stream_ptr(means.device))
        ws.readers = True
        ctx.save_for_backward(means, campos, radii, row_cache)
        ctx.sh_degree, ctx.n_block, ctx.K = sh_degree, n_block, K
        ctx.n_local, ctx.split = coeffs_local.shape[0], coeffs_rest is not None
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, radii, row_cache = ctx.saved_tensors
        n_local, K = ctx.n_local, ctx.K
        C, N = campos.shape[0], means.shape[0]
        n_block = ctx.n_block
        dev = means.device
        ws = workspace(dev, K * 3, n_block)
        v_colors = v_colors.contiguous()
        v_means = torch.empty_like(means) if ctx.needs_input_grad[1] else None
        v_campos = torch.empty_like(campos) if ctx.needs_input_grad[2] else None
        with device_guard(dev):
            call("gsb200_sh_view_bwd_sharded", C, N, K, ctx.sh_degree, ptr(means), ptr(campos), ws.coeff_table, ws.world, n_block,
                 ws.rank, ptr(row_cache), ptr(radii), ptr(v_colors), ws.stage_table, ws.flag_table, ptr(v_means), ptr(v_campos),
                 stream_ptr(dev))
            ws.barrier()  # every rank's pushes have landed
            if ctx.split:
                v_first = torch.empty((n_local, 1, 3), dtype=torch.float32, device=dev)
                v_rest = torch.empty((n_local, K - 1, 3), dtype=torch.float32, device=dev)
                call("gsb200_shard_reduce_rows_split", ws.world, n_block, n_local, K * 3, 3, ptr(ws.staging), ptr(ws.flags),
                     ptr(v_first), ptr(v_rest), stream_ptr(dev))
            else:
                v_first, v_rest = torch.empty((n_local, K, 3), dtype=torch.float32, device=dev), None
                call("gsb200_shard_reduce_rows", ws.world, n_block, n_local, K * 3, ptr(ws.staging), ptr(ws.flags),
                     ptr(v_first), stream_ptr(dev))
            ws.barrier()  # every owner is done reducing: staging and published rows may be overwritten
        return None, v_means, v_campos, v_first, v_rest, None, None, None


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _u32_array(values):
    arr = (ctypes.c_uint32 * len(values))(*values)
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class _GatherGeometry(torch.autograd.Function):
    """The attribute tensors of the gathered set ([world * n_block, w_k] each) straight from the published blocks (one
    pull kernel, no gathered [N, w] intermediate and no column-split copies).  Backward: every rank pushes each owner's
    gradient rows into that owner's staging slot over NVLink, one barrier, every owner sums its slots in rank order."""

    @staticmethod
    def forward(ctx, ws, n_block, *geo_parts):
        widths = [t.shape[1] for t in geo_parts]
        dev = geo_parts[0].device
        outs = [torch.empty((ws.world * n_block, w), dtype=torch.float32, device=dev) for w in widths]
        keep_w, w_ptr = _u32_array(widths)
        keep_o, o_ptr = _ptr_array(outs)
        with device_guard(dev):
            call("gsb200_shard_gather_geo", ws.world, n_block, len(widths), w_ptr, ws.geo_table, o_ptr, stream_ptr(dev))
        ws.readers = True
        ctx.ws, ctx.n_block, ctx.n_local, ctx.widths, ctx.dev = ws, n_block, geo_parts[0].shape[0], widths, dev
        return tuple(outs)

    @staticmethod
    def backward(ctx, *v_outs):
        ws, n_block, widths, dev = ctx.ws, ctx.n_block, ctx.widths, ctx.dev
        grads = [None if g is None else g.contiguous() for g in v_outs]
        outs = [torch.empty((ctx.n_local, w), dtype=torch.float32, device=dev) if ctx.needs_input_grad[2 + k] else None
                for k, w in enumerate(widths)]
        keep_w, w_ptr = _u32_array(widths)
        keep_g, g_ptr = _ptr_array(grads)
        keep_o, o_ptr = _ptr_array(outs)
        with device_guard(dev):
            slot = ws.cap_block * GEO_WIDTH  # floats per staging slot (one dense [n_block * w] block each)
            call("gsb200_shard_push_geo", ws.world, n_block, ws.rank, slot, len(widths), w_ptr, g_ptr, ws.geo_stage_table,
                 stream_ptr(dev))
            ws.barrier()  # every rank's rows have landed in every owner's staging
            call("gsb200_shard_reduce_geo", ws.world, n_block, ctx.n_local, slot, len(widths), w_ptr, ptr(ws.geo_staging),
                 o_ptr, stream_ptr(dev))
            # the staging area is next written in the next step's backward, after that step's publish barrier
        return (None, None) + tuple(outs)


def publish_and_gather_geometry(geo_parts: Sequence[Tensor], pad_row: Tensor, coeffs_local: Optional[Tensor], n_block: int):
    """Start of a distributed render: publish this rank's geometry rows and SH rows (one barrier), pull everybody's
    geometry.  Returns (tuple of the gathered attribute tensors [world * n_block, w_k] -- differentiable w.r.t. geo_parts
    --, version token to hand to view_colors_from_sh_sharded).  n_block must be a multiple of 4 and >= every rank's row
    count."""
    assert n_block % 4 == 0, n_block
    pair = isinstance(coeffs_local, (tuple, list))
    dev = geo_parts[0].device
    if coeffs_local is None:  # coefficients stay with their owner (view_colors_owner): geometry-only workspace
        ws, rows = workspace(dev, GEOMETRY_ONLY_ROW, n_block), None
    else:
        K = coeffs_local[0].shape[1] + coeffs_local[1].shape[1] if pair else coeffs_local.shape[1]
        ws = workspace(dev, K * 3, n_block)
        rows = tuple(t.detach().float() for t in coeffs_local) if pair else coeffs_local.detach().float()
    parts = [t.detach().float() for t in geo_parts]
    with device_guard(dev):
        version = ws.publish(rows, parts, pad_row, n_block)
    return _GatherGeometry.apply(ws, n_block, *[t.float() for t in geo_parts]), version


def view_colors_from_sh_sharded(sh_degree: int, means: Tensor, campos: Tensor, coeffs_local: Tensor, radii: Tensor,
                                n_block: int, published: Optional[int] = None) -> Tensor:
    """Distributed form of cuda._wrapper.view_colors_from_sh: `means` [N,3] and `radii` [C,N] cover the gathered set
    (N = world * n_block, rank r's rows at r * n_block), `coeffs_local` [n_local,K,3] is this rank's shard;
    `published`: version token if the rows were already published by publish_and_gather_geometry."""
    N, C = means.shape[0], campos.shape[0]
    assert N == dist.get_world_size() * n_block, (N, n_block)
    assert radii.shape == (C, N), radii.shape
    first, rest = coeffs_local if isinstance(coeffs_local, (tuple, list)) else (coeffs_local, None)
    assert first.dim() == 3 and first.shape[2] == 3 and first.shape[0] <= n_block, first.shape
    assert (sh_degree + 1) ** 2 <= first.shape[1] + (rest.shape[1] if rest is not None else 0)
    return _ShViewSharded.apply(sh_degree, means.float().contiguous(), campos.float().contiguous(), first.float().contiguous(),
                                None if rest is None else rest.float().contiguous(), radii.contiguous(), n_block, published)


# ---- owner-side colour exchange ---------------------------------------------------------------------------------------------
MAX_VIEWS = 64  # GSB200_MAX_VIEWS in include/gsb200.h
_colour_workspaces: Dict[Tuple[int, int], "ColourExchange"] = {}


class ColourExchange:
    """Symmetric allocation of the owner-side colour exchange for C cameras per rank:
         [ colours   C x world x cap_block x 3 ]   this rank's colour tensor [C, world * n_block, 3]; owner p writes rows
                                                   p * n_block.. of every camera
         [ masks     V x cap_block (int32) ]       view v's radii of THIS rank's rows, pushed by the viewer of v
         [ v_colors  V x cap_block x 3 ]           view v's colour gradients of this rank's rows, pushed in the backward
         [ campos    V x 4 ]                       every view's camera centre
         [ pads ]                                  signal pad of the barrier kernel
    V = world x C views, view v belongs to rank v // C."""

    def __init__(self, device: torch.device, n_cameras: int, cap_block: int):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.C, self.cap_block, self.V = n_cameras, cap_block, self.world * n_cameras
        n_col = n_cameras * self.world * cap_block * 3
        n_mask = self.V * cap_block
        n_vcol = self.V * cap_block * 3
        n_cam = self.V * 4
        n_pads = 4 * MAX_SHARDS
        self.buf = symm_mem.empty(n_col + n_mask + n_vcol + n_cam + n_pads, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.ptrs) == self.world and all(p % 16 == 0 for p in self.ptrs), self.ptrs
        o_mask, o_vcol = n_col, n_col + n_mask
        o_cam, o_pads = o_vcol + n_vcol, o_vcol + n_vcol + n_cam
        self.buf.zero_()  # colours of never-visible pairs are never written: keep them finite
        self.colours = self.buf[:o_mask]
        self.masks = self.buf[o_mask:o_vcol].view(torch.int32)
        self.vcol = self.buf[o_vcol:o_cam]
        self.campos = self.buf[o_cam:o_pads]
        self._mask_tab = _ptr_table(self.ptrs, 4 * o_mask)
        self._vcol_tab = _ptr_table(self.ptrs, 4 * o_vcol)
        self._cam_tab = _ptr_table(self.ptrs, 4 * o_cam)
        self._pad_tab = _ptr_table(self.ptrs, 4 * o_pads)
        self._colour_tabs: Dict[int, tuple] = {}
        self.epoch = 0
        self.hdl.barrier(channel=0)  # every pad is zero before the first signal can arrive

    def barrier(self) -> None:
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF
        call("gsb200_peer_barrier", self.world, self.rank, self._pad_tab[1], self.epoch, stream_ptr(self.buf.device))

    def colour_table(self, n_block: int):
        """view v = (rank q, camera c) -> address of THIS rank's [n_block, 3] rows inside q's colour tensor."""
        tab = self._colour_tabs.get(n_block)
        if tab is None:
            n_total = self.world * n_block
            addrs = [self.ptrs[v // self.C] + 4 * ((v % self.C) * n_total + self.rank * n_block) * 3 for v in range(self.V)]
            arr = (ctypes.c_void_p * self.V)(*addrs)
            self._colour_tabs[n_block] = tab = (arr, ctypes.cast(arr, ctypes.c_void_p))
        return tab[1]


def colour_workspace(device: torch.device, n_cameras: int, n_block: int) -> ColourExchange:
    key = (device.index if device.index is not None else torch.cuda.current_device(), n_cameras)
    ws = _colour_workspaces.get(key)
    if ws is None or ws.cap_block < n_block or ws.world != dist.get_world_size():
        if ws is not None:
            ws.barrier()
            _colour_workspaces.pop(key)
            del ws
        cap = (int(n_block * 1.1) + 64 + 3) // 4 * 4
        _colour_workspaces[key] = ws = ColourExchange(device, n_cameras, cap)
    return ws


def owner_exchange_available(device: torch.device, n_cameras: int, n_block: int) -> bool:
    """The owner-side colour exchange needs what available() needs, at most MAX_VIEWS views in the job, and the same number
    of cameras on every rank (the caller's contract).  GSB200_SH_EXCHANGE=pull keeps the coefficient-pull path."""
    if os.environ.get("GSB200_SH_EXCHANGE", "owner") == "pull":
        return False
    if not available(device, GEOMETRY_ONLY_ROW, n_block) or dist.get_world_size() * n_cameras > MAX_VIEWS:
        return False
    try:
        colour_workspace(device, n_cameras, n_block)
    except Exception as e:  # pragma: no cover - same failure on every rank
        warnings.warn(f"owner-side colour exchange unavailable ({type(e).__name__}: {e}); pulling coefficient rows instead")
        return False
    return True


class _ShOwnerExchange(torch.autograd.Function):
    """View colours [C, world * n_block, 3] of the gathered set, evaluated by the owners of the coefficients.
    forward : push this rank's radii + camera centres to every owner | barrier | evaluate the own rows for all V views,
              storing the colours into the viewers' tensors over NVLink | barrier
    backward: push v_colors rows to their owners | barrier | owner-side SH backward -> gradients of the LOCAL sh0, shN and
              means (no coefficient or coefficient-gradient row ever crosses the link)."""

    @staticmethod
    def forward(ctx, sh_degree, means_local, campos, sh0, shN, radii, n_block):
        dev = means_local.device
        C, n_total = radii.shape
        ws = colour_workspace(dev, C, n_block)
        assert n_total == ws.world * n_block and campos.shape == (C, 3), (radii.shape, campos.shape, n_block)
        K = 1 + shN.shape[1]
        n_local = means_local.shape[0]
        with device_guard(dev):
            call("gsb200_shard_push_rows", ws.world, n_block, 1, C, ws.rank * C, n_block, ptr(radii), ws._mask_tab[1],
                 ptr(campos), 3, ws._cam_tab[1], stream_ptr(dev))
            ws.barrier()  # every viewer's masks and camera centres have landed
            # private copies: the staging is overwritten by the next forward, the backward of THIS one may come later
            masks = ws.masks[: ws.V * n_block].clone()
            campos_all = ws.campos[: ws.V * 3].clone()
            call("gsb200_sh_owner_fwd", ws.V, n_local, n_block, K, sh_degree, ptr(means_local), ptr(campos_all), ptr(sh0),
                 ptr(shN), ptr(masks), ws.colour_table(n_block), stream_ptr(dev))
            ws.barrier()  # every owner's colours have landed
        ctx.save_for_backward(means_local, sh0, shN, masks, campos_all)
        ctx.cfg = (sh_degree, n_block, C, K)
        colours = ws.colours[: C * n_total * 3].view(C, n_total, 3)
        return colours

    @staticmethod
    def backward(ctx, v_colors):
        means_local, sh0, shN, masks, campos_all = ctx.saved_tensors
        sh_degree, n_block, C, K = ctx.cfg
        dev = means_local.device
        ws = colour_workspace(dev, C, n_block)
        n_local = means_local.shape[0]
        v_colors = v_colors.contiguous()
        v_sh0 = torch.empty((n_local, 1, 3), dtype=torch.float32, device=dev)
        v_shN = torch.empty((n_local, K - 1, 3), dtype=torch.float32, device=dev)
        v_means = torch.empty_like(means_local) if ctx.needs_input_grad[1] else None
        with device_guard(dev):
            call("gsb200_shard_push_rows", ws.world, n_block, 3, C, ws.rank * C, n_block * 3, ptr(v_colors), ws._vcol_tab[1],
                 None, 0, None, stream_ptr(dev))
            ws.barrier()  # every viewer's colour gradients have landed
            call("gsb200_sh_owner_bwd", ws.V, n_local, n_block, K, sh_degree, ptr(means_local), ptr(campos_all), ptr(sh0), ptr(shN),
                 ptr(masks), ptr(ws.vcol), ptr(v_sh0), ptr(v_shN), ptr(v_means), stream_ptr(dev))
        return None, v_means, None, v_sh0, v_shN, None, None


def view_colors_owner(sh_degree: int, means_local: Tensor, campos: Tensor, sh_pair, radii: Tensor, n_block: int) -> Tensor:
    """Distributed colour stage, owner-side form: `means_local` [n_local,3], `sh_pair` = (sh0 [n_local,1,3],
    shN [n_local,K-1,3]) are THIS rank's shard, `campos` [C,3] and `radii` [C, world * n_block] this rank's views of the
    gathered set.  Returns the colours [C, world * n_block, 3] (rows of invisible pairs hold stale finite values; the blend
    never reads them).  Differentiable w.r.t. means_local, sh0, shN; not w.r.t. campos (use the pull path for that)."""
    sh0, shN = sh_pair
    assert sh0.dim() == 3 and sh0.shape[1:] == (1, 3) and shN.dim() == 3 and shN.shape[2] == 3, (sh0.shape, shN.shape)
    assert (sh_degree + 1) ** 2 <= 1 + shN.shape[1]
    return _ShOwnerExchange.apply(sh_degree, means_local.float().contiguous(), campos.detach().float().contiguous(),
                                  sh0.float().contiguous(), shN.float().contiguous(), radii.contiguous(), n_block)


# ---- owner-side projection + colour exchange ("owner render") ---------------------------------------------------------------
_render_workspaces: Dict[Tuple[int, int], "RenderExchange"] = {}
_VIEW_ARRAYS = (("radii", 1), ("means2d", 2), ("depths", 1), ("conics", 3), ("opacities", 1), ("colours", 3))
_STAGE_ARRAYS = (("v_means2d", 2), ("v_conics", 3), ("v_opacities", 1), ("v_colours", 3), ("v_depths", 1))


class RenderExchange:
    """Symmetric allocation of the owner-side render exchange for C cameras per rank (V = world x C views):
         viewer side, written by the owners : radii | means2d | depths | conics | opacities | colours, each
                                              [C, world * n_block, w] -- the tensors this rank's tile intersection and
                                              blend read; owner p writes rows p * n_block.. of every camera
         owner side, written by the viewers : v_means2d | v_conics | v_opacities | v_colours | v_depths, each
                                              [V, n_block, w] -- gradients of THIS rank's rows in every view
         cameras                            : viewmats [V, 16], Ks [V, 9] of every view
         pads                               : signal pad of the barrier kernel"""

    def __init__(self, device: torch.device, n_cameras: int, cap_block: int):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.C, self.cap_block, self.V = n_cameras, cap_block, self.world * n_cameras
        off, self.off = 0, {}
        for name, w in _VIEW_ARRAYS:
            self.off[name] = off
            off += n_cameras * self.world * cap_block * w
        for name, w in _STAGE_ARRAYS:
            self.off[name] = off
            off += self.V * cap_block * w
        for name, w in (("viewmats", 16), ("Ks", 9)):
            self.off[name] = off
            off += (self.V * w + 3) // 4 * 4
        self.off["pads"] = off
        off += 4 * MAX_SHARDS
        self.buf = symm_mem.empty(off, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.ptrs) == self.world and all(p % 16 == 0 for p in self.ptrs), self.ptrs
        self.buf.zero_()  # rows of never-visible pairs are never written: keep them finite
        self.tabs = {name: _ptr_table(self.ptrs, 4 * self.off[name]) for name in self.off}
        self._view_tabs: Dict[int, dict] = {}
        self.epoch = 0
        self.hdl.barrier(channel=0)  # every pad is zero before the first signal can arrive

    def barrier(self) -> None:
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF
        call("gsb200_peer_barrier", self.world, self.rank, self.tabs["pads"][1], self.epoch, stream_ptr(self.buf.device))

    def region(self, name: str, n_floats: int) -> Tensor:
        return self.buf[self.off[name]: self.off[name] + n_floats]

    def view_tables(self, n_block: int) -> dict:
        """array name -> HOST table: view v = (rank q, camera c) -> address of THIS rank's n_block rows in q's array."""
        tabs = self._view_tabs.get(n_block)
        if tabs is None:
            n_total, tabs = self.world * n_block, {}
            for name, w in _VIEW_ARRAYS:
                addrs = [self.ptrs[v // self.C] + 4 * (self.off[name] + ((v % self.C) * n_total + self.rank * n_block) * w)
                         for v in range(self.V)]
                arr = (ctypes.c_void_p * self.V)(*addrs)
                tabs[name] = (arr, ctypes.cast(arr, ctypes.c_void_p))
            self._view_tabs[n_block] = tabs
        return tabs


def render_workspace(device: torch.device, n_cameras: int, n_block: int) -> RenderExchange:
    key = (device.index if device.index is not None else torch.cuda.current_device(), n_cameras)
    ws = _render_workspaces.get(key)
    if ws is None or ws.cap_block < n_block or ws.world != dist.get_world_size():
        if ws is not None:
            ws.barrier()
            _render_workspaces.pop(key)
            del ws
        cap = (int(n_block * 1.1) + 64 + 3) // 4 * 4
        _render_workspaces[key] = ws = RenderExchange(device, n_cameras, cap)
    return ws


def owner_render_available(device: torch.device, n_cameras: int, n_block: int) -> bool:
    """Owner-side projection + colours: NCCL job on CUDA with peer mappings, <= MAX_SHARDS ranks, <= MAX_VIEWS views, the
    same number of cameras on every rank (the caller's contract).  GSB200_DIST_RENDER=gather keeps the gather path."""
    global _unavailable
    if _unavailable is not None or os.environ.get("GSB200_DIST_RENDER", "owner") != "owner":
        return False
    if os.environ.get("GSB200_PEER_EXCHANGE", "1") == "0":
        return False
    if device.type != "cuda" or dist.get_backend() != "nccl":
        return False
    if dist.get_world_size() > MAX_SHARDS or dist.get_world_size() * n_cameras > MAX_VIEWS:
        return False
    try:
        render_workspace(device, n_cameras, n_block)
    except Exception as e:  # no P2P mapping on this machine: every rank fails alike
        _unavailable = f"{type(e).__name__}: {e}"
        warnings.warn(f"peer-memory exchange unavailable ({_unavailable}); using NCCL collectives")
        return False
    return True


class _OwnerRenderInputs(torch.autograd.Function):
    """Stage 1 + colour stage of a distributed render, computed by the OWNERS of the Gaussians:
    forward : exchange the cameras | barrier | every rank projects its rows and evaluates their SH colours for all V views,
              storing radius / 2-D mean / depth / conic / opacity / colour of the visible pairs into the viewers' tensors
              over NVLink | barrier
    backward: the viewers push the blend's gradients of those tensors to the owners | barrier | owner-side projection and
              SH backward -> gradients of the LOCAL means, quats, scales, opacities, sh0, shN.
    No attribute, coefficient or attribute-gradient row crosses the link; what crosses is ~12 bytes per (view, Gaussian)
    forward and 36 bytes backward, instead of 44 + 44 bytes per Gaussian plus the SH rows."""

    @staticmethod
    def forward(ctx, cfg, means, quats, scales, opacities, sh0, shN, viewmats, Ks):
        (sh_degree, n_block, W, H, eps2d, near, far, radius_clip, model) = cfg
        dev = means.device
        C = viewmats.shape[0]
        ws = render_workspace(dev, C, n_block)
        V, n_local, n_total = ws.V, means.shape[0], ws.world * n_block
        K = 1 + shN.shape[1]
        s = stream_ptr(dev)
        with device_guard(dev):
            call("gsb200_peer_push_small", ws.world, ws.rank * C, C, 16, ptr(viewmats), ws.tabs["viewmats"][1], s)
            call("gsb200_peer_push_small", ws.world, ws.rank * C, C, 9, ptr(Ks), ws.tabs["Ks"][1], s)
            ws.barrier()  # every rank's cameras have landed
            vm_all = ws.region("viewmats", V * 16).clone()  # private copies: the backward of THIS step may come after the
            ks_all = ws.region("Ks", V * 9).clone()         # next forward has overwritten the staging
            radii_loc = torch.empty((V, n_block), dtype=torch.int32, device=dev)
            conics_loc = torch.empty((V, n_block, 3), dtype=torch.float32, device=dev)
            tabs = ws.view_tables(n_block)
            call("gsb200_projection_owner_fwd", V, n_local, n_block, ptr(means), ptr(quats), ptr(scales), ptr(opacities),
                 ptr(vm_all), ptr(ks_all), W, H, eps2d, near, far, radius_clip, model, ptr(radii_loc), ptr(conics_loc),
                 tabs["radii"][1], tabs["means2d"][1], tabs["depths"][1], tabs["conics"][1], tabs["opacities"][1], s)
            campos_all = torch.empty((V, 3), dtype=torch.float32, device=dev)
            call("gsb200_camera_positions", V, ptr(vm_all), ptr(campos_all), s)
            call("gsb200_sh_owner_fwd", V, n_local, n_block, K, sh_degree, ptr(means), ptr(campos_all), ptr(sh0), ptr(shN),
                 ptr(radii_loc), tabs["colours"][1], s)
            ws.barrier()  # every owner's rows have landed
        ctx.save_for_backward(means, quats, scales, sh0, shN, vm_all, ks_all, campos_all, radii_loc, conics_loc)
        ctx.cfg = cfg
        ctx.C = C
        radii = ws.region("radii", C * n_total).view(torch.int32).view(C, n_total)
        ctx.mark_non_differentiable(radii)
        return (radii, ws.region("means2d", C * n_total * 2).view(C, n_total, 2), ws.region("depths", C * n_total).view(C, n_total),
                ws.region("conics", C * n_total * 3).view(C, n_total, 3), ws.region("opacities", C * n_total).view(C, n_total),
                ws.region("colours", C * n_total * 3).view(C, n_total, 3))

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_opacities, v_colours):
        means, quats, scales, sh0, shN, vm_all, ks_all, campos_all, radii_loc, conics_loc = ctx.saved_tensors
        (sh_degree, n_block, W, H, eps2d, near, far, radius_clip, model) = ctx.cfg
        dev, C = means.device, ctx.C
        ws = render_workspace(dev, C, n_block)
        V, n_local, n_total = ws.V, means.shape[0], ws.world * n_block
        K = 1 + shN.shape[1]
        s = stream_ptr(dev)

        def dense(t, w):
            return torch.zeros((C, n_total, w), dtype=torch.float32, device=dev) if t is None else t.contiguous()

        v_sh0 = torch.empty((n_local, 1, 3), dtype=torch.float32, device=dev)
        v_shN = torch.empty((n_local, K - 1, 3), dtype=torch.float32, device=dev)
        v_means = torch.empty_like(means)
        v_means_sh = torch.empty_like(means)
        v_quats, v_scales = torch.empty_like(quats), torch.empty_like(scales)
        with device_guard(dev):
            for name, w, t in (("v_means2d", 2, v_means2d), ("v_conics", 3, v_conics), ("v_opacities", 1, v_opacities),
                               ("v_colours", 3, v_colours), ("v_depths", 1, v_depths)):
                if name == "v_depths" and t is None:
                    continue
                call("gsb200_shard_push_rows", ws.world, n_block, w, C, ws.rank * C, n_block * w, ptr(dense(t, w)),
                     ws.tabs[name][1], None, 0, None, s)
            ws.barrier()  # every viewer's gradients have landed
            call("gsb200_projection_owner_bwd", V, n_local, n_block, ptr(means), ptr(quats), ptr(scales), ptr(vm_all), ptr(ks_all),
                 W, H, eps2d, model, ptr(radii_loc), ptr(conics_loc), ptr(ws.region("v_means2d", V * n_block * 2)),
                 None if v_depths is None else ptr(ws.region("v_depths", V * n_block)),
                 ptr(ws.region("v_conics", V * n_block * 3)), ptr(v_means), ptr(v_quats), ptr(v_scales), s)
            call("gsb200_sh_owner_bwd", V, n_local, n_block, K, sh_degree, ptr(means), ptr(campos_all), ptr(sh0), ptr(shN),
                 ptr(radii_loc), ptr(ws.region("v_colours", V * n_block * 3)), ptr(v_sh0), ptr(v_shN), ptr(v_means_sh), s)
            v_means += v_means_sh
            # opacity: the viewers' rows are zero for invisible pairs, so the plain sum over the views is the gradient
            v_opac = ws.region("v_opacities", V * n_block).view(V, n_block)[:, :n_local].sum(0)
        return None, v_means, v_quats, v_scales, v_opac, v_sh0, v_shN, None, None


def owner_render_inputs(sh_degree: int, means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, sh_pair, viewmats: Tensor,
                        Ks: Tensor, width: int, height: int, n_block: int, eps2d: float, near_plane: float, far_plane: float,
                        radius_clip: float, camera_model: int):
    """-> (radii [C,N], means2d [C,N,2], depths [C,N], conics [C,N,3], opacities [C,N], colours [C,N,3]) of the gathered set
    (N = world * n_block, rank r's rows at r * n_block) for THIS rank's cameras, computed by the owners of the rows.  Inputs
    are this rank's shard ([n_local, ...]) and cameras.  Rows of invisible pairs hold radius 0 and stale finite values
    elsewhere (the tile intersection and the blend never read them).  Differentiable w.r.t. the shard, not the cameras."""
    sh0, shN = sh_pair
    assert sh0.dim() == 3 and sh0.shape[1:] == (1, 3) and shN.dim() == 3 and shN.shape[2] == 3, (sh0.shape, shN.shape)
    assert (sh_degree + 1) ** 2 <= 1 + shN.shape[1]
    cfg = (int(sh_degree), int(n_block), int(width), int(height), float(eps2d), float(near_plane), float(far_plane),
           float(radius_clip), int(camera_model))

    def f(t):
        return t.float().contiguous()

    return _OwnerRenderInputs.apply(cfg, f(means), f(quats), f(scales), f(opacities), f(sh0), f(shN), f(viewmats.detach()),
                                    f(Ks.detach()))
