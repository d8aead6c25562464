"""Where the host-buffer call's time goes (8K 4:2:0 by default): raw pinned copy times, then
jpegqs_cuda_run_host with the slab pipeline off / on / with other wave sizes.
    python tools/e2e_probe.py [--flags 0] [--niter 3] [--waves 0,37888,151552]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpegqs_b200 as qs
from jpegqs_b200.image import CoefImage, Component

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--niter", type=int, default=3)
ap.add_argument("--width", type=int, default=7680)
ap.add_argument("--height", type=int, default=4320)
ap.add_argument("--waves", default="0,37888,151552")
ap.add_argument("--reps", type=int, default=6)
args = ap.parse_args()

ctx = qs.cuda.QsContext(0)
im = qs.synth.make_image(args.width, args.height, "420")
nbytes = sum(c.coef.nbytes for c in im.comps)
pins = [qs.cuda.PinnedArray(c.coef.shape) for c in im.comps]
dev = [torch.empty(c.coef.shape, dtype=torch.int16, device="cuda") for c in im.comps]


def best(fn, reps=args.reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


def h2d():
    for p, d in zip(pins, dev):
        d.copy_(torch.from_numpy(p.array), non_blocking=True)


def d2h():
    for p, d in zip(pins, dev):
        torch.from_numpy(p.array).copy_(d, non_blocking=True)


for p, c in zip(pins, im.comps):
    p.array[...] = c.coef
a, b = best(h2d)
print(f"H2D {nbytes / 1e6:.1f} MB: best {a:.3f} ms ({nbytes / a / 1e6:.1f} GB/s), median {b:.3f} ms")
a, b = best(d2h)
print(f"D2H {nbytes / 1e6:.1f} MB: best {a:.3f} ms ({nbytes / a / 1e6:.1f} GB/s), median {b:.3f} ms")


def run():
    for p, c in zip(pins, im.comps):
        p.array[...] = c.coef
    w = CoefImage(im.width, im.height, im.colorspace,
                  [Component(p.array, c.quant.copy(), c.h_samp, c.v_samp, c.quant_tbl_no) for p, c in zip(pins, im.comps)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.do_quantsmooth(w, args.flags, args.niter, inplace=True)
    return (time.perf_counter() - t0) * 1e3


def report(label):
    ts = [run() for _ in range(args.reps)]
    a, b, c, d = ctx.kernel_stats()
    print(f"{label}: best {min(ts):.3f} ms, median {sorted(ts)[len(ts) // 2]:.3f} ms, device {ctx.last_device_ms:.3f} ms, "
          f"launches {ctx.last_launches}; kernels: idct {a:.3f} ms / {b}, smooth {c:.3f} ms / {d}", flush=True)


# ---- device-resident run, alone and with copy traffic on a side stream ----
ctx.set_profiling(True)
side = torch.cuda.Stream()
big_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
big_d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for traffic in (0, 1, 2):
    ts = []
    for _ in range(4):
        for p, d in zip(pins, dev):
            d.copy_(torch.from_numpy(p.array))
        torch.cuda.synchronize()
        if traffic:
            with torch.cuda.stream(side):
                for _ in range(8):
                    if traffic == 1:
                        big_d.copy_(big_h, non_blocking=True)
                    else:
                        big_h.copy_(big_d, non_blocking=True)
        ctx.run_device(im, [d.data_ptr() for d in dev], [], args.flags, args.niter, 1)
        torch.cuda.synchronize()
        ts.append(ctx.last_device_ms)
    a, b, c, d = ctx.kernel_stats()
    print(f"run_device, side traffic {['none', 'H2D', 'D2H'][traffic]}: device best {min(ts):.3f} ms; "
          f"kernels: idct {a:.3f} ms / {b}, smooth {c:.3f} ms / {d}", flush=True)


ctx.set_tuning(6, 0)
report("run_host, slabs off")
ctx.set_tuning(6, 1)
for wv in args.waves.split(","):
    ctx.set_tuning(7, int(wv))
    report(f"run_host, slabs on, wave={wv}")
