#!/usr/bin/env python
"""Wall time of the drop-in call itself: the library's exported
`do_quantsmooth(j_decompress_ptr, jvirt_barray_ptr*, jpegqs_control_t*)` (include/libjpegqs.h,
reference libjpegqs.h:47-48) driven through the fake libjpeg boundary with pageable,
separately allocated block rows (what libjpeg's memory manager hands over).  Only the call is
timed; loading the input into the arrays happens outside (in real life that is libjpeg's own
decoding).  Also prints jpegqs_cuda_run_host on caller-pinned flat buffers for comparison.
    python tools/dropin_probe.py [--flags 0] [--niter 3] [--reps 8] [--threads 0]"""
import argparse, ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jpegqs_b200 as qs
import oracle_lib as ol

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--niter", type=int, default=3)
ap.add_argument("--width", type=int, default=7680)
ap.add_argument("--height", type=int, default=4320)
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--contiguous", action="store_true")
ap.add_argument("--gpus", type=int, default=0, help="JPEGQS_GPUS for do_quantsmooth; the image is stacked that many times (weak scaling)")
a = ap.parse_args()
if a.gpus > 1:
    os.environ["JPEGQS_GPUS"] = str(a.gpus)
    a.height *= a.gpus
ol.ensure_built()
lib = qs.cuda.load()
im = qs.synth.make_image(a.width, a.height, "420")
ctx = qs.cuda.QsContext(0)
ret0, want = ctx.do_quantsmooth(im, a.flags, a.niter)          # flat host-buffer path, for equality
fn = C.cast(lib.do_quantsmooth, C.c_void_p)
s = ol.BoundarySession(im, scatter_rows=not a.contiguous)
ts = []
for r in range(a.reps + 2):
    s.load()
    t0 = time.perf_counter()
    ret = s.run(fn, a.flags | 64, a.niter, threads=a.threads)    # JPEGQS_TRANSCODE
    ts.append((time.perf_counter() - t0) * 1e3)
    if r == 0:
        out = s.result()
        same = ol.images_equal(out, want)
        if a.flags & 4:
            break                                               # geometry changed: one run only
s.close()
ts = ts[2:] or ts
mp = a.width * a.height / 1e6
print(f"do_quantsmooth via fake libjpeg ({'contiguous' if a.contiguous else 'scattered pageable'} rows), flags={a.flags} "
      f"niter={a.niter}: ret={ret} same_as_run_host={same}  min {min(ts):.2f} ms  median {sorted(ts)[len(ts)//2]:.2f} ms "
      f"({mp / (sorted(ts)[len(ts)//2] / 1e3):.0f} Mpix/s)")
