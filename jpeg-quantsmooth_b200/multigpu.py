"""MCU-row sharding of one image across the GPUs of a box.

One process per GPU (torch.distributed, NCCL over NVLink).  Every rank owns a contiguous
range of MCU rows of every component.  Within an iteration blocks only need the
previous-pass samples of the 1-pixel ring around them (reference quantsmooth.h:1396-1401,
Jacobi structure), so the only exchange is ONE pixel row per component per slab boundary
between the IDCT pass and the smoothing pass: a neighbour exchange (send/recv), not a
collective.  The IDCT is integer-deterministic, so the result is bit-identical to the
single-GPU run for any shard count (tests/test_multigpu_gloo.py, tests/test_gpu_slabs.py).

The pass implementation is pluggable: the product uses `CudaPasses` (the C ABI's pass-level
entry points); the CPU tests drive the same sharding/exchange logic with oracle passes.
JOINT_YUV / UPSAMPLE_UV (q>=5) need luma->chroma plane hand-over across slabs and are run
single-GPU for now.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

PASS_DEQUANT, PASS_CLAMP = 1, 2
PLANE_PAD = 16


def plane_stride(wblk: int) -> int:
    return wblk * 8 + 2 * PLANE_PAD


def split_mcu_rows(total_mcu_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [m0, m1) MCU-row ranges, sizes differing by at most one."""
    base, rem = divmod(total_mcu_rows, world)
    out, m = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((m, m + n))
        m += n
    return out


def comp_block_rows(mcu_range: Tuple[int, int], v_samp: int, hblk: int) -> Tuple[int, int]:
    """Block rows of a component covered by an MCU-row range (libjpeg geometry: the last
    MCU row may be only partly present, image.blocks_for)."""
    return min(mcu_range[0] * v_samp, hblk), min(mcu_range[1] * v_samp, hblk)


@dataclass
class SlabComp:
    coef: object            # tensor int16 [rows, wblk, 64] on the pass backend's device
    plane: object           # tensor uint8 [rows*8+2, stride]
    wblk: int
    rows: int               # block rows in this slab
    quant: np.ndarray       # raw quantval [64]
    luma: bool


class CudaPasses:
    """Pass backend over the C ABI (jpegqs_cuda_pass_idct / jpegqs_cuda_pass_smooth)."""

    def __init__(self, ctx, stream: int = 0):
        # stream: a cudaStream_t handle.  0 is mapped to cudaStreamLegacy (0x1) so that the
        # passes are ordered with torch work on the default stream; the C ABI reserves NULL
        # for "the context's own stream".
        self.ctx, self.stream = ctx, (stream or 1)

    def _jobs(self, comps: Sequence[SlabComp], top_edge: bool, bottom_edge: bool):
        return [self.ctx.make_job(c.coef.data_ptr(), c.plane.data_ptr(), None, c.wblk, c.rows,
                                  c.quant, c.luma, top_edge, bottom_edge) for c in comps if c.rows]

    def idct(self, comps, mode, top_edge, bottom_edge, want_bad):
        return self.ctx.pass_idct(self._jobs(comps, top_edge, bottom_edge), mode, want_bad, self.stream)

    def smooth(self, comps, flags, clamp_out, top_edge, bottom_edge):
        self.ctx.pass_smooth(self._jobs(comps, top_edge, bottom_edge), flags, clamp_out, self.stream)

    def clamp(self, comps):
        for c in comps:
            c.coef.clamp_(-1023, 1023)


def make_flag_allreduce(dist, device):
    """Returns f(bitmask) -> bitwise OR of the 32-bit mask over all ranks.  NCCL has no BOR,
    so the mask travels as 32 0/1 lanes reduced with MAX."""
    import torch
    lanes = torch.zeros(32, dtype=torch.int32, device=device)
    weights = [1 << i for i in range(32)]

    def f(mask: int) -> int:
        mask &= 0xFFFFFFFF
        lanes.copy_(torch.tensor([(mask >> i) & 1 for i in range(32)], dtype=torch.int32))
        dist.all_reduce(lanes, op=dist.ReduceOp.MAX)
        got = lanes.cpu().tolist()
        return sum(w for w, g in zip(weights, got) if g)
    return f


def exchange_halos(comps: Sequence[SlabComp], rank: int, world: int, dist, group=None):
    """Fill row 0 / row h+1 of every plane with the neighbour slab's adjacent pixel row
    (side border bytes included).  Image top/bottom rows are replicated by the IDCT pass."""
    if world == 1:
        return
    ops = []
    for c in comps:
        if not c.rows:
            continue
        h = c.rows * 8
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, c.plane[1], rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, c.plane[0], rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, c.plane[h], rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, c.plane[h + 1], rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def run_slab(passes, comps: Sequence[SlabComp], flags: int, niter: int, rank: int, world: int,
             dist=None, allreduce_flag=None) -> int:
    """The iteration loop of do_quantsmooth (reference quantsmooth.h:2580-2689) on one slab.
    Returns the reference's `stop`.  Components are independent for flags without
    JOINT_YUV/UPSAMPLE_UV, so they share the passes."""
    if flags & (2 | 4):
        raise NotImplementedError("multi-GPU sharding does not cover JOINT_YUV / UPSAMPLE_UV yet (DESIGN.md 5)")
    niter = max(0, min(int(niter), 100))
    if niter == 0:
        return 0
    top, bottom = rank == 0, rank == world - 1
    stop = 0
    work = []
    for c in comps:                                    # per-component prelude, 2497-2566
        val = int(np.bitwise_or.reduce(np.asarray(c.quant, dtype=np.int64)))
        if val >= 0x800:
            stop = 1
        if val <= 1:
            continue                                   # niter2 == 0 and no extra refresh
        if stop:
            q = c.quant
            import torch
            qt = torch.as_tensor(np.asarray(q, dtype=np.int32), device=c.coef.device)
            c.coef.copy_((c.coef.to(torch.int32) * qt).to(torch.int16))    # dequantize only
            continue
        work.append(c)
    for it in range(niter):
        if not work:
            break
        bad = passes.idct(work, PASS_DEQUANT if it == 0 else 0, top, bottom, it == 0)
        if it == 0:
            if allreduce_flag is not None:
                bad = allreduce_flag(bad)              # OR across ranks (quantsmooth.h:2602-2610)
            if bad:
                # bit i = work[i] out of range.  The reference runs components one after
                # another: those before the first bad one finish normally, the bad one is
                # only clamped, the later ones are only de-quantized (which the pass did).
                stop = 1
                b = (bad & -bad).bit_length() - 1
                passes.clamp(work[b:b + 1])
                work = work[:b]
                if not work:
                    break
        if dist is not None:
            exchange_halos(work, rank, world, dist)
        passes.smooth(work, flags, it == niter - 1, top, bottom)
    return stop
