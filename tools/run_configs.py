"""Measures BASELINE.json configs 2-5 on device-resident data (generated on the GPU with the
torch generator).  Under torchrun, config 5 (and --check) shard by MCU rows over the ranks.

    python tools/run_configs.py --configs 2,3,4           # one GPU
    torchrun --nproc-per-node N tools/run_configs.py --configs 5 [--size 32768] [--check]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpegqs_b200 as qs                      # noqa: E402
from jpegqs_b200 import multigpu as mg        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="2,3,4")
ap.add_argument("--size", type=int, default=32768)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true", help="config 5: verify the sharded result against a single-GPU run of a smaller image")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
ctx = qs.cuda.QsContext(local)
tstream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(tstream)
stream = tstream.cuda_stream


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


def timed(fn, reps):
    fn()                                   # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def single_image(cfg, w, h, flags, niter):
    im = qs.synth.make_image_torch(w, h, "420", device=dev)
    src = [c.coef for c in im.comps]
    work = [torch.empty_like(t) for t in src]
    y = im.comps[0]
    ups = [None] + [torch.empty((y.hblk, y.wblk, 64), dtype=torch.int16, device=dev).data_ptr() for _ in range(2)] \
        if flags & 4 else []

    def fn():
        for a, b in zip(work, src):
            a.copy_(b)
        ctx.run_device(im, [t.data_ptr() for t in work], ups, flags, niter, stream)

    def copy_only():
        for a, b in zip(work, src):
            a.copy_(b)
    ms = timed(fn, args.reps) - timed(copy_only, args.reps)
    emit({"config": cfg, "shape": f"{w}x{h} 4:2:0", "flags": flags, "niter": niter, "ms": round(ms, 3),
          "mpix_s": round(w * h / 1e6 / (ms / 1e3), 1), "launches": ctx.last_launches})


def batch(cfg, n, flags, niter):
    ims = [qs.synth.make_image_torch(1920, 1080, "420", seed=1000 + i, device=dev) for i in range(min(n, 8))]
    ims = [ims[i % len(ims)] for i in range(n)]                 # 8 distinct images, cycled
    src = [[c.coef for c in im.comps] for im in ims]
    work = [[torch.empty_like(t) for t in s] for s in src]

    def copy_only():
        for ws, ss in zip(work, src):
            for a, b in zip(ws, ss):
                a.copy_(b)

    def fn():
        copy_only()
        ctx.run_batch_device(ims, [[t.data_ptr() for t in ws] for ws in work], [], flags, niter, stream)
    ms = timed(fn, args.reps) - timed(copy_only, args.reps)
    emit({"config": cfg, "shape": f"{n} x 1920x1080 4:2:0", "flags": flags, "niter": niter, "ms": round(ms, 3),
          "images_s": round(n / (ms / 1e3), 1), "mpix_s": round(n * 1920 * 1080 / 1e6 / (ms / 1e3), 1),
          "launches": ctx.last_launches})


def sharded(cfg, size, flags, niter, check):
    total = (size + 15) // 16
    rng = mg.split_mcu_rows(total, world)[rank]
    slab = qs.synth.make_image_torch(size, size, "420", mcu_rows=rng, device=dev)
    geom = mg.SlabGeom(True, size, size)
    fullh = [qs.blocks_for(size, c.v_samp, 2) for c in slab.comps]
    src = [c.coef for c in slab.comps]
    work = [torch.empty_like(t) for t in src]
    planes = [torch.empty((c.hblk * 8 + 2, mg.plane_stride(c.wblk)), dtype=torch.uint8, device=dev) for c in slab.comps]
    passes = mg.CudaPasses(ctx, stream)
    allreduce_flag = mg.make_flag_allreduce(dist, dev) if dist is not None else None

    state = {}

    def copy_only():
        for a, b in zip(work, src):
            a.copy_(b)

    def fn():
        copy_only()
        comps = [mg.SlabComp(work[k], planes[k], c.wblk, c.hblk, c.quant, k == 0, c.h_samp, c.v_samp,
                             mg.comp_block_rows(rng, c.v_samp, fullh[k])[0], fullh[k]) for k, c in enumerate(slab.comps)]
        state["ups"] = mg.run_slab(passes, comps, flags, niter, rank, world, dist, allreduce_flag, geom)[1]
        state["comps"] = comps
    ms = timed(fn, args.reps) - timed(copy_only, args.reps)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res = {"config": cfg, "shape": f"{size}x{size} 4:2:0", "flags": flags, "niter": niter, "n_gpus": world,
           "ms": round(ms, 3), "mpix_s": round(size * size / 1e6 / (ms / 1e3), 1)}
    if check:
        # the sharded result of every rank's slab must equal the same rows of a single-GPU run
        fn()                                  # `work` now holds the sharded result again
        torch.cuda.synchronize()
        full = qs.synth.make_image_torch(size, size, "420", device=dev)
        fw = [c.coef.clone() for c in full.comps]
        y = full.comps[0]
        fup = [torch.empty((y.hblk, y.wblk, 64), dtype=torch.int16, device=dev) for _ in range(2)]
        _, ups1 = ctx.run_device(full, [t.data_ptr() for t in fw], [None] + [t.data_ptr() for t in fup], flags, niter, stream)
        torch.cuda.synchronize()
        bad = int(ups1 != state["ups"])
        for k, c in enumerate(full.comps):
            if ups1 and k in (1, 2):
                r0, r1 = mg.comp_block_rows(rng, 2, y.hblk)
                bad += int((fup[k - 1][r0:r1] != state["comps"][k].coef_up).sum().item())
            else:
                r0, r1 = mg.comp_block_rows(rng, c.v_samp, c.hblk)
                bad += int((fw[k][r0:r1] != work[k]).sum().item())
        b = torch.tensor([bad], dtype=torch.int64, device=dev)
        if dist is not None:
            dist.all_reduce(b)
        res["mismatches_vs_single_gpu"] = int(b.item())
    emit(res)


for cfg in [int(x) for x in args.configs.split(",")]:
    if cfg == 2:
        single_image(2, 3840, 2160, 0, 3)
    elif cfg == 3:
        single_image(3, 7680, 4320, 7, 3)
        single_image("3b (q4)", 7680, 4320, 1, 3)
    elif cfg == 4:
        for n in (1, 4, 16, 64, 256):
            batch(4, n, 0, 3)
    elif cfg == 5:
        sharded(5, args.size, 1, 5, args.check)
    elif cfg == 6:                      # q6 sharded (JOINT_YUV + UPSAMPLE_UV across slabs)
        sharded("6 (q6 sharded)", args.size, 7, 3, args.check)
if dist is not None:
    dist.destroy_process_group()
