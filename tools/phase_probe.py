#!/usr/bin/env python
"""Where do the smoothing kernel's cycles go?  Needs the measurement build
(`make -C jpeg-quantsmooth_b200/csrc phase`), selected with JPEGQS_B200_LIB:

    JPEGQS_B200_LIB=jpeg-quantsmooth_b200/csrc/libjpegqs_b200_phase.so python tools/phase_probe.py [--flags 0]

Every warp of qs_smooth_kernel accumulates SM clock cycles per phase (QS_PHASE_CLOCKS in
csrc/qs_kernels.cu); the four warps of a sub-partition run in lock step, so the per-phase share
of the summed warp cycles is the share of sub-partition time."""
import argparse, ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpegqs_b200 as qs

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--niter", type=int, default=3)
ap.add_argument("--width", type=int, default=7680)
ap.add_argument("--height", type=int, default=4320)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--merge", type=int, default=-1, help="tuning key 8 (mixed chunks), -1 = library default")
a = ap.parse_args()
NAMES = ["tile fetch (atomic + barriers)", "tile prologue (job, coef/pixel loads, mixed check)",
         "chunk header + chunk barrier", "refresh IDCT (+ barrier)", "chunk set-up (Rs, table ptrs)",
         "section h", "section border", "section v", "section diag", "coefficient update (div, clamp)",
         "chunk loop exit", "rebalance", "write-back", "-", "-", "-"]
ctx = qs.cuda.QsContext(0)
if a.merge >= 0:
    ctx.set_tuning(8, a.merge)
lib = qs.cuda.load()
if not hasattr(lib, "qs_read_phase_clocks"):
    raise SystemExit("not a phase build: set JPEGQS_B200_LIB to libjpegqs_b200_phase.so")
lib.qs_read_phase_clocks.argtypes = [C.c_void_p, C.c_int]
im = qs.synth.make_image(a.width, a.height, "420")
dev = torch.device("cuda", 0)
host = [torch.from_numpy(np.ascontiguousarray(c.coef)) for c in im.comps]
stream = torch.cuda.current_stream().cuda_stream or 1
buf = (C.c_ulonglong * 16)()
ctx.set_profiling(True)
for i in range(a.steps + 1):
    bufs = [h.to(dev) for h in host]
    ups = []
    if a.flags & 4:
        y = im.comps[0]
        ups = [None] + [torch.empty((y.hblk, y.wblk, 64), dtype=torch.int16, device=dev).data_ptr() for _ in range(2)]
    torch.cuda.synchronize()
    if i == 1:
        lib.qs_read_phase_clocks(buf, 1)       # reset after the warm-up
    ctx.run_device(im, [t.data_ptr() for t in bufs], ups, a.flags, a.niter, stream)
torch.cuda.synchronize()
_, _, sm_ms, sm_n = ctx.kernel_stats()
lib.qs_read_phase_clocks(buf, 0)
v = np.array(list(buf), dtype=np.float64)
tot = v.sum()
ntile = im.num_blocks / 32.0 * a.niter * a.steps
print(f"flags={a.flags} {a.width}x{a.height}: smoothing {sm_ms / sm_n:.3f} ms/launch (instrumented build), "
      f"{tot / ntile:.0f} cycles per warp tile")
for k, n in enumerate(NAMES):
    if v[k]:
        print(f"  {100 * v[k] / tot:6.2f} %  {v[k] / ntile:9.0f} cyc/tile  {n}")
