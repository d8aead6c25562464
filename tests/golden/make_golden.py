"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqsref_scalar.so,
built from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each file holds one input (quantized coefficients, raw quant tables, geometry) and the
reference's output for it, so the parity tests can pin both the C restatement (CPU) and the
CUDA path (GPU) without /root/reference being present.  The reference has no golden vectors
of its own; these are outputs of the reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jpegqs_b200 as qs          # noqa: E402
import oracle_lib as ol           # noqa: E402
from golden_io import save_case, adversarial_image   # noqa: E402

CASES = [
    # name, (w, h, subsampling, quality, seed), flags, niter
    ("gray_q3", (40, 32, "gray", 50, 1), 0, 3),
    ("420_q3", (64, 48, "420", 50, 2), 0, 3),
    ("420_q3_hi", (56, 40, "420", 92, 3), 0, 2),
    ("420_q4", (64, 48, "420", 50, 4), 1, 2),
    ("420_q5", (64, 48, "420", 50, 5), 3, 2),
    ("420_q6", (64, 48, "420", 60, 6), 7, 2),
    ("444_q6", (48, 32, "444", 50, 7), 7, 2),
    ("422_q6", (64, 32, "422", 50, 8), 7, 1),
    ("420_norebal", (48, 32, "420", 50, 9), 16, 2),
    ("420_norebal_uv", (48, 32, "420", 50, 10), 33, 2),
    ("420_lowq0", (64, 48, "420", 25, 11), 8 | 1, 3),
    ("420_lowq2", (64, 48, "420", 30, 12), 8 | 7, 2),
    ("gray_lowq", (40, 32, "gray", 15, 13), 8, 3),
]


def main():
    ol.ensure_built()
    assert ol.have_ref("scalar"), "needs oracle/_ref (build container only)"
    for name, (w, h, ss, quality, seed), flags, niter in CASES:
        im = qs.synth.make_image(w, h, ss, quality=quality, seed=seed)
        ret, out = ol.run_reference(im, flags, niter)
        save_case(os.path.join(HERE, name + ".npz"), im, out, flags, niter, ret)
        print(name, "ret", ret, [c.coef.shape for c in out.comps])
    # adversarial blocks: NaN (a3 == 0) case, out-of-range coefficient, huge quant value,
    # zero quant values, all-zero AC
    for name, kind, flags, niter in [("adv_nan", "nan", 0, 1), ("adv_badcoef", "badcoef", 0, 2),
                                      ("adv_bigquant", "bigquant", 0, 2), ("adv_zeroquant", "zeroquant", 0, 2),
                                      ("adv_flat", "flat", 1, 2), ("adv_q1", "q1", 0, 3)]:
        im = adversarial_image(kind)
        ret, out = ol.run_reference(im, flags, niter)
        save_case(os.path.join(HERE, name + ".npz"), im, out, flags, niter, ret)
        print(name, "ret", ret)


if __name__ == "__main__":
    main()
