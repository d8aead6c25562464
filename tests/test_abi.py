"""The C-ABI shared library loads and exports every symbol declared in include/*.h; the
host-side pieces that need no GPU (weight tables, exact-division helper, error paths) are
checked against the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import jpegqs_b200 as qs
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jpegqs_cuda_\w+|do_quantsmooth)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(qs.cuda.lib_path())
    names = [n for n in _declared("jpegqs_cuda.h") if n != "jpegqs_cuda_progress_fn"]
    assert len(names) >= 20
    for n in names + ["do_quantsmooth"]:
        assert hasattr(lib, n), f"{n} is declared in include/ but not exported"


def test_libjpegqs_header_keeps_the_reference_surface(tmp_path):
    """Compile a C program against include/libjpegqs.h and check every constant the reference
    header defines (libjpegqs.h:14-32), the struct layout (41-45) and the prototypes (47-55)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include <jpeglib.h>
#include "libjpegqs.h"
static int cb(void *d, int cur, int max) { (void)d; return cur > max; }
int main(void) {
	jpegqs_control_t c = { 1, 2, 3, 4, NULL, cb };
	int (*f)(j_decompress_ptr, jvirt_barray_ptr*, jpegqs_control_t*) = do_quantsmooth;
	printf("%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", JPEGQS_ITER_MAX, JPEGQS_DIAGONALS,
		JPEGQS_JOINT_YUV, JPEGQS_UPSAMPLE_UV, JPEGQS_LOW_QUALITY, JPEGQS_NO_REBALANCE, JPEGQS_NO_REBALANCE_UV,
		JPEGQS_TRANSCODE, JPEGQS_FLAGS_MASK, JPEGQS_CPU_SHIFT, JPEGQS_CPU_MASK, JPEGQS_INFO_SHIFT,
		JPEGQS_INFO_COMP1 >> 16, JPEGQS_INFO_QUANT >> 16, JPEGQS_INFO_COMP2 >> 16, JPEGQS_INFO_TIME >> 16,
		JPEGQS_INFO_CPU >> 16);
	printf("%zu %zu %zu %zu %zu %zu %d\n", offsetof(jpegqs_control_t, flags), offsetof(jpegqs_control_t, niter),
		offsetof(jpegqs_control_t, threads), offsetof(jpegqs_control_t, progprec),
		offsetof(jpegqs_control_t, userdata), offsetof(jpegqs_control_t, progress), c.progress(NULL, 1, 0) + (f != NULL));
	return 0;
}
''')
    exe = tmp_path / "t"
    lib = os.path.dirname(qs.cuda.lib_path())
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include", "libjpeg62"), "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", lib, "-ljpegqs_b200", f"-Wl,-rpath,{lib}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0].split() == [str(v) for v in (100, 1, 2, 4, 8, 16, 32, 64, 0x7f, 12, 15, 16, 1, 2, 4, 8, 16)]
    assert out[1].split() == ["0", "4", "8", "12", "16", "24", "2"]
    for q, f in [(0, 8 | 1), (2, 8 | 7), (3, 0), (4, 1), (5, 3), (6, 7)]:     # quantsmooth.c:380-393
        assert qs.quality_to_flags(q) == f


@pytest.mark.parametrize("flags", [0, 1])
def test_device_tables_are_the_reference_bits(flags):
    t = qs.cuda.tables(flags)
    o = np.zeros_like(t)
    ol.oraclelib().qso_tables(flags, o.ctypes.data)
    assert np.array_equal(t.view(np.uint32), o.view(np.uint32))
    nz = np.abs(t[t != 0])
    # exact power-of-two rescaling needs every non-zero weight well above 2^-39
    assert nz.min() >= 2.0 ** -30 and nz.max() < 8.0


def test_exact_division_helper_matches_plain_form():
    lib = qs.cuda.load()
    orc = ol.oraclelib()
    rng = np.random.RandomState(5)
    qs_ = list(range(1, 260)) + [511, 512, 513, 1000, 1023, 1024, 1025, 2046, 2047]
    for q in qs_:
        cs = np.concatenate([np.arange(-3 * q - 2, 3 * q + 3), rng.randint(-0x4000, 0x4000, 40),
                             [-0x4000, 0x3fff, -2048, 2047]])
        for c in cs:
            assert lib.jpegqs_cuda_orig_coef(int(c), q) == orc.qso_orig_coef(int(c), q), (c, q)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = qs.cuda.load()
    h = C.c_void_p()
    rc = lib.jpegqs_cuda_create(-1, C.byref(h))
    assert rc < 0 and not h.value
    assert len(lib.jpegqs_cuda_last_error(None)) > 0
    with pytest.raises(qs.cuda.QsError):
        qs.cuda.QsContext(0)
    # the libjpeg-facing entry point reports the failure and leaves the coefficients alone
    im = qs.synth.make_image(32, 32, "gray")
    ret, out = ol.run_libjpeg_boundary(lib.do_quantsmooth, im, 64, 3)
    assert ret < 0
    assert np.array_equal(out.comps[0].coef, im.comps[0].coef)
    assert np.array_equal(out.comps[0].quant, im.comps[0].quant)


def test_plane_geometry_helpers():
    lib = qs.cuda.load()
    from jpegqs_b200 import multigpu as mg
    assert lib.jpegqs_cuda_plane_pad() == mg.PLANE_PAD
    for w in (1, 7, 240, 960):
        assert lib.jpegqs_cuda_plane_stride(w) == mg.plane_stride(w)
        assert lib.jpegqs_cuda_plane_bytes(w, 5) == mg.plane_stride(w) * (5 * 8 + 2)
