#!/usr/bin/env python
"""How often could the smoothing kernel skip the in-register IDCT refresh?

The reference re-renders a block at the start of an anti-diagonal only if a coefficient of
the previous diagonals changed (need_refresh, quantsmooth.h:1408, 1563).  The CUDA kernel
refreshes unconditionally; a warp (32 consecutive blocks in row-major order) could skip a
refresh only when NO lane changed a coefficient in the previous diagonal.  This tool uses the
oracle's change probe (oracle/qs_oracle.c::qso_set_change_probe) on the bench image (luma of
the synthetic 4:2:0 image, grayscale run) and prints, per iteration, the per-block and
per-warp skip rates for the 13 refresh points s = 13..1 (refresh at diagonal s happens iff
diagonal s+1 changed something)."""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jpegqs_b200 as qs
import oracle_lib as ol

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=3840); ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--quality", type=int, default=50); ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--comp", type=int, default=0)
a = ap.parse_args()
ol.ensure_built()
im = qs.synth.make_image(a.width, a.height, "420", quality=a.quality, seed=12345)
from jpegqs_b200.image import CoefImage
c = im.comps[a.comp]
gray = CoefImage(c.wblk * 8, c.hblk * 8, 1, [c])          # JCS_GRAYSCALE, one component
lib = ol.oraclelib()
lib.qso_set_change_probe.argtypes = [C.c_void_p, C.c_void_p]
nb = c.wblk * c.hblk
for it in range(3):
    work = gray.clone()
    arr = np.ascontiguousarray(work.comps[0].coef, dtype=np.int16)
    work.comps[0].coef = arr
    probe = np.zeros(nb, dtype=np.uint16)
    # run_oracle clones again, so probe addressing goes through a direct call
    qi = ol.QsoImage(); qi.ncomp = 1; qi.is_ycbcr = 0
    qi.image_width, qi.image_height = work.width, work.height
    qc = qi.comp[0]; qc.coef = arr.ctypes.data; qc.wblk, qc.hblk, qc.h_samp, qc.v_samp = c.wblk, c.hblk, 1, 1
    qc.has_qtbl = 1
    for k in range(64): qc.quant[k] = int(c.quant[k])
    lib.qso_set_threads(0)
    lib.qso_set_change_probe(arr.ctypes.data, probe.ctypes.data)
    lib.qso_run.argtypes = [C.POINTER(ol.QsoImage), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.qso_run(C.byref(qi), a.flags, it + 1, 0, None, None)
    lib.qso_set_change_probe(None, None)
    # warps: 32 consecutive blocks (tile mapping of qs_smooth_kernel)
    pad = (-nb) % 32
    pw = np.concatenate([probe, np.zeros(pad, np.uint16)]).reshape(-1, 32)
    wor = np.bitwise_or.reduce(pw, axis=1)
    blk, wrp = [], []
    for s in range(14, 1, -1):            # change in diagonal s -> refresh needed at diagonal s-1
        blk.append(1.0 - np.count_nonzero(probe & (1 << (s - 1))) / nb)
        wrp.append(1.0 - np.count_nonzero(wor & (1 << (s - 1))) / len(wor))
    print(f"iteration {it}: refresh skippable  per block: mean {np.mean(blk):.3f}  per warp: mean {np.mean(wrp):.3f}")
    print("   diag changed  " + " ".join(f"{s:5d}" for s in range(14, 1, -1)))
    print("   block skip    " + " ".join(f"{v:5.2f}" for v in blk))
    print("   warp  skip    " + " ".join(f"{v:5.2f}" for v in wrp))
