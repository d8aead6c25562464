"""The host entry point's slab pipeline (upload | iteration 0 and last iteration | download
overlap, qs_cuda.cu::run_images): forced onto small images with a tiny wave size so that every
flag combination, the stop paths and ragged slab heights are checked against the oracle."""
import numpy as np
import pytest

import jpegqs_b200 as qs
import oracle_lib as ol
from golden_io import adversarial_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = qs.cuda.QsContext(0)
    yield c
    c.close()


def _same(ctx, im, flags, niter, wave):
    ctx.set_tuning(7, wave)
    try:
        ret, out = ctx.do_quantsmooth(im, flags, niter, pinned=True)
    finally:
        ctx.set_tuning(7, 0)
    oret, oout = ol.run_oracle(im, flags, niter)
    assert ret == oret
    assert ol.images_equal(out, oout), f"{ol.diff_count(out, oout)} coefficients differ"
    return out


@pytest.mark.parametrize("w,h,ss,flags,niter,wave", [
    (200, 136, "420", 0, 3, 64),        # q3: slabs in and out, ragged heights
    (200, 136, "420", 0, 1, 64),        # one iteration: first pass is also the last
    (200, 136, "420", 0, 2, 40),
    (256, 200, "420", 1, 3, 96),        # q4 DIAGONALS (corner pixels across slab seams)
    (256, 200, "420", 3, 3, 96),        # q5 JOINT_YUV: chroma slabs read the luma plane2 at an offset
    (256, 200, "420", 7, 3, 96),        # q6 UPSAMPLE_UV: extra render pass, no slab download
    (250, 130, "444", 7, 2, 50),
    (250, 130, "444", 0, 3, 50),
    (160, 400, "gray", 1, 3, 70),
    (136, 120, "422", 3, 2, 30),
    (136, 120, "440", 0, 3, 30),
    (64, 1000, "420", 0, 3, 16),        # more wave-slabs than QS_MAX_SLABS: several waves per slab
])
def test_slab_pipeline_matches_oracle(ctx, w, h, ss, flags, niter, wave):
    im = qs.synth.make_image(w, h, ss)
    _same(ctx, im, flags, niter, wave)
    n_on = ctx.last_launches
    ctx.do_quantsmooth(im, flags, niter, pinned=True)      # default wave: far too small to be cut
    assert n_on > ctx.last_launches


@pytest.mark.parametrize("flags,niter", [(0, 3), (7, 2), (1, 1)])
def test_slab_pipeline_bad_coefficient_falls_back(ctx, flags, niter):
    """A coefficient out of range in the LAST slab, found after earlier slabs were already
    smoothed: the group is fetched again and stops like the reference (quantsmooth.h:2602-2610)."""
    im = qs.synth.make_image(160, 240, "420")
    y = im.comps[0]
    y.coef[y.hblk - 1, 3, 5] = 3000
    _same(ctx, im, flags, niter, 40)
    im = qs.synth.make_image(160, 240, "420")
    c = im.comps[2]
    c.coef[c.hblk - 1, 2, 1] = -3000
    _same(ctx, im, flags, niter, 40)


@pytest.mark.parametrize("kind", ["badcoef", "bigquant", "zeroquant", "nan", "q1"])
def test_slab_pipeline_adversarial(ctx, kind):
    _same(ctx, adversarial_image(kind), 0, 2, 8)
    _same(ctx, adversarial_image(kind), 7, 2, 8)


def test_slab_pipeline_on_off_identical_at_full_size(ctx):
    """8K 4:2:0 q3: the default plan (one wave per slab) against the un-pipelined path."""
    im = qs.synth.make_image(7680, 4320, "420")
    ret1, a = ctx.do_quantsmooth(im, 0, 3, pinned=True)
    n_on = ctx.last_launches
    ctx.set_tuning(6, 0)
    try:
        ret0, b = ctx.do_quantsmooth(im, 0, 3, pinned=True)
        n_off = ctx.last_launches
    finally:
        ctx.set_tuning(6, 1)
    assert ret0 == ret1 == 0 and ol.images_equal(a, b)
    assert n_on > n_off            # the pipelined run really was cut into slabs
