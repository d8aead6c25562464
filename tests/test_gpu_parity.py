"""GPU parity tests proper: the CUDA path, called through the C ABI (host buffers), must be
bit-identical to the oracle on the same seeded inputs."""
import numpy as np
import pytest

import jpegqs_b200 as qs
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = qs.cuda.QsContext(0)
    yield c
    c.close()


def _check(ctx, im, flags, niter, **kw):
    ret, out = ctx.do_quantsmooth(im, flags, niter, **kw)
    oret, oout = ol.run_oracle(im, flags, niter)
    assert ret == oret
    nd = ol.diff_count(out, oout)
    assert nd == 0, f"{nd} coefficients differ from the oracle"
    assert ol.images_equal(out, oout)
    return out


CASES = [
    # (w, h, subsampling, flags, niter, quality)
    (64, 64, "gray", 0, 1, 50),
    (512, 512, "gray", 0, 3, 50),          # BASELINE config 1 shape
    (256, 128, "420", 0, 3, 50),
    (250, 130, "420", 0, 2, 90),           # ragged geometry, small q
    (256, 128, "420", 1, 2, 50),           # DIAGONALS (q4)
    (256, 128, "420", 3, 2, 50),           # + JOINT_YUV (q5)
    (256, 128, "420", 7, 2, 50),           # + UPSAMPLE_UV (q6)
    (200, 120, "420", 7, 3, 75),
    (256, 128, "444", 7, 2, 50),
    (256, 128, "422", 7, 2, 50),
    (248, 136, "440", 7, 2, 75),
    (256, 128, "420", 16, 2, 50),          # NO_REBALANCE
    (256, 128, "420", 32 | 1, 2, 50),      # NO_REBALANCE_UV
    (8, 8, "gray", 0, 3, 50),              # a single block
    (1920, 1080, "420", 0, 3, 50),         # BASELINE config 4 single image
]


@pytest.mark.parametrize("w,h,ss,flags,niter,quality", CASES)
def test_parity(ctx, w, h, ss, flags, niter, quality):
    im = qs.synth.make_image(w, h, ss, quality=quality)
    _check(ctx, im, flags, niter)


def test_randomized_configurations(ctx):
    """Seeded sweep over geometry, sub-sampling, quality (incl. extreme tables), flag combinations
    and iteration counts - every case bit-exact against the oracle."""
    rng = np.random.RandomState(20260923)
    modes = ["gray", "444", "422", "420", "440"]
    for case in range(48):
        ss = modes[rng.randint(len(modes))]
        w, h = int(rng.randint(8, 180)), int(rng.randint(8, 140))
        quality = int(rng.choice([3, 10, 25, 50, 75, 90, 97, 100]))
        flags = int(rng.randint(0, 64))          # all 6 behaviour bits
        if (flags & 4) and ss in ("420", "422", "440"):
            # UPSAMPLE_UV on a sub-sampled image: keep the geometry where the reference is
            # well defined (DESIGN.md "reference quirks"): width a multiple of the MCU
            w = max(16, w // 16 * 16)
        niter = int(rng.randint(1, 4))
        im = qs.synth.make_image(w, h, ss, quality=quality, seed=1000 + case, noise=int(rng.randint(0, 12)))
        ret, out = ctx.do_quantsmooth(im, flags, niter)
        oret, want = ol.run_oracle(im, flags, niter)
        assert ret == oret, (case, w, h, ss, quality, flags, niter)
        assert ol.images_equal(out, want), (case, w, h, ss, quality, flags, niter, ol.diff_count(out, want))
