"""Pins the C restatement (oracle/qs_oracle.c) against the UNMODIFIED reference
(oracle/_ref/libqsref_scalar.so, the -DNO_SIMD build of /root/reference/quantsmooth.h).
Skipped where oracle/_ref is absent; the committed golden vectors (test_golden.py) cover
that case."""
import ctypes as C

import numpy as np
import pytest

import jpegqs_b200 as qs
import oracle_lib as ol
from golden_io import adversarial_image

pytestmark = pytest.mark.skipif(not ol.have_ref("scalar"), reason="oracle/_ref not built (needs /root/reference)")


def test_natural_order_and_refresh_points():
    # reference idct.h:24-33 / quantsmooth.h:313-322: the refresh flag is set exactly at
    # the first-visited coefficient of each anti-diagonal in reverse zig-zag order
    zz = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
          13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52,
          45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    assert sorted(zz) == list(range(64))
    refresh = {2, 4, 6, 8, 15, 24, 31, 40, 47, 56, 58, 60, 62, 63}      # the table's 1s, minus DC
    got = set()
    for k in range(63, 0, -1):
        i = zz[k]
        if k == 63 or (zz[k + 1] // 8 + zz[k + 1] % 8) != (i // 8 + i % 8):
            got.add(i)
    assert got == refresh


def test_idct_islow_exact():
    ref, orc = ol.reflib(), ol.oraclelib()
    rng = np.random.RandomState(1)
    for scale in (8, 200, 2047):
        for _ in range(200):
            c = rng.randint(-scale, scale + 1, size=64).astype(np.int16)
            if rng.rand() < 0.3:
                c[rng.rand(64) < 0.8] = 0
            a = np.zeros((8, 24), dtype=np.uint8); b = np.zeros((8, 24), dtype=np.uint8)
            ref.qsref_idct_islow(c.ctypes.data, a.ctypes.data, 24)
            orc.qso_idct_islow(c.ctypes.data, b.ctypes.data, 24)
            assert np.array_equal(a, b)


def test_float_dcts_bit_exact():
    ref, orc = ol.reflib(), ol.oraclelib()
    rng = np.random.RandomState(2)
    for _ in range(300):
        x = (rng.rand(64).astype(np.float32) - 0.5) * np.float32(rng.choice([1, 255, 4096]))
        for fr, fo in ((ref.qsref_fdct_float, orc.qso_fdct_float), (ref.qsref_idct_float, orc.qso_idct_float)):
            a = np.zeros(64, dtype=np.float32); b = np.zeros(64, dtype=np.float32)
            xa, xb = x.copy(), x.copy()
            fr(xa.ctypes.data, a.ctypes.data); fo(xb.ctypes.data, b.ctypes.data)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("flags", [0, 1])
def test_tables_bit_exact(flags):
    size = 272 if flags else 160
    a = np.zeros((64, size), dtype=np.float32); b = np.zeros((64, size), dtype=np.float32)
    assert ol.reflib().qsref_tables(flags, a.ctypes.data) == size
    ol.oraclelib().qso_tables(flags, b.ctypes.data)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_orig_coef_exhaustive_vs_reference_formula():
    # plain form (quantsmooth.h:338-341) over the whole valid range, vectorised
    orc = ol.oraclelib()
    c = np.arange(-0x4000, 0x4000, dtype=np.int64)
    for q in list(range(1, 300)) + [511, 512, 513, 1023, 1024, 2047]:
        h = q >> 1
        n = c + np.where(c < 0, -h, h)
        want = np.sign(n) * (np.abs(n) // q) * q
        for cc in (-0x4000, -q, -h - 1, -h, -1, 0, 1, h, h + 1, q, 0x3fff):
            assert orc.qso_orig_coef(int(cc), q) == int(want[cc + 0x4000])


CASES = [
    (512, 512, "gray", 0, 3, 50),            # BASELINE config 1 (CPU scalar plumbing)
    (256, 128, "420", 0, 3, 50), (250, 130, "420", 0, 2, 92), (256, 128, "420", 1, 2, 50),
    (256, 128, "420", 3, 2, 50), (256, 128, "420", 7, 2, 50), (200, 120, "420", 7, 3, 75),
    (256, 128, "444", 7, 2, 50), (256, 128, "422", 7, 2, 50), (248, 136, "440", 7, 2, 75),
    (128, 64, "420", 16, 2, 50), (128, 64, "420", 33, 2, 50), (128, 64, "420", 6, 0, 50),
    (128, 64, "420", 0, 0, 50), (8, 8, "gray", 0, 3, 50), (128, 64, "420", 5, 1, 30),
    (256, 128, "420", 9, 3, 30), (250, 130, "420", 11, 2, 20), (256, 128, "420", 15, 2, 40),
    (128, 64, "gray", 8, 3, 10), (256, 128, "444", 15, 2, 50), (64, 48, "420", 24, 2, 30),
]


@pytest.mark.parametrize("w,h,ss,flags,niter,quality", CASES)
def test_driver_matches_reference(w, h, ss, flags, niter, quality):
    im = qs.synth.make_image(w, h, ss, quality=quality)
    r1, o1 = ol.run_reference(im, flags, niter)
    r2, o2 = ol.run_oracle(im, flags, niter)
    assert r1 == r2
    assert ol.diff_count(o1, o2) == 0 and ol.images_equal(o1, o2)


@pytest.mark.parametrize("kind,flags,niter", [("nan", 0, 1), ("badcoef", 0, 2), ("bigquant", 0, 2),
                                              ("zeroquant", 0, 2), ("flat", 1, 2), ("q1", 0, 3),
                                              ("badcoef", 7, 2), ("bigquant", 7, 1)])
def test_adversarial_matches_reference(kind, flags, niter):
    im = adversarial_image(kind)
    r1, o1 = ol.run_reference(im, flags, niter)
    r2, o2 = ol.run_oracle(im, flags, niter)
    assert r1 == r2
    assert ol.images_equal(o1, o2)


def test_nan_case_moves_coefficients():
    # with a3 == 0 the x86 conversion yields INT_MIN and coefficients
    # move to their clamp edge; a conversion that yields 0 would leave them unchanged
    im = adversarial_image("nan")
    _, out = ol.run_oracle(im, 0, 1)
    deq = im.comps[0].coef.astype(np.int32) * im.comps[0].quant.astype(np.int32)
    assert np.count_nonzero(out.comps[0].coef != deq) > 0


def test_non_contiguous_rows_and_threads():
    im = qs.synth.make_image(96, 80, "420")
    _, a = ol.run_reference(im, 7, 2, scatter_rows=True, threads=3)
    _, b = ol.run_reference(im, 7, 2, threads=1)
    assert ol.images_equal(a, b)


def test_progress_sequence_and_cancel():
    im = qs.synth.make_image(96, 80, "420")
    seq_ref, seq_orc = [], []
    ol.run_reference(im, 0, 3, progprec=7, progress=lambda d, cur, mx: seq_ref.append((cur, mx)) or 0)
    ol.run_oracle(im, 0, 3, progprec=7, progress=lambda d, cur, mx: seq_orc.append((cur, mx)) or 0)
    assert seq_ref == seq_orc and len(seq_ref) > 2
    # cancel at the second callback
    def mk(seq):
        return lambda d, cur, mx: (seq.append(cur), 5 if len(seq) >= 2 else 0)[1]
    s1, s2 = [], []
    r1, o1 = ol.run_reference(im, 0, 3, progprec=7, progress=mk(s1))
    r2, o2 = ol.run_oracle(im, 0, 3, progprec=7, progress=mk(s2))
    assert r1 == r2 == 5 and s1 == s2
    assert ol.images_equal(o1, o2)


def test_randomized_sweep_matches_reference():
    """The same seeded sweep the GPU parity test runs (tests/test_gpu_parity.py), here pinning
    the checker itself against the unmodified reference."""
    rng = np.random.RandomState(20260923)
    modes = ["gray", "444", "422", "420", "440"]
    for case in range(48):
        ss = modes[rng.randint(len(modes))]
        w, h = int(rng.randint(8, 180)), int(rng.randint(8, 140))
        quality = int(rng.choice([3, 10, 25, 50, 75, 90, 97, 100]))
        flags = int(rng.randint(0, 64))
        if (flags & 4) and ss in ("420", "422", "440"):
            w = max(16, w // 16 * 16)
        niter = int(rng.randint(1, 4))
        im = qs.synth.make_image(w, h, ss, quality=quality, seed=1000 + case, noise=int(rng.randint(0, 12)))
        r1, o1 = ol.run_reference(im, flags, niter)
        r2, o2 = ol.run_oracle(im, flags, niter)
        assert r1 == r2 and ol.images_equal(o1, o2), (case, w, h, ss, quality, flags, niter)
