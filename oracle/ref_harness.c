/*
 * TEST INFRASTRUCTURE - not part of the product.
 *
 * Compiles the UNMODIFIED reference hot path (reference quantsmooth.h + idct.h,
 * included from where they lie under $(REF), never copied) into a shared
 * object under oracle/_ref/, against include/libjpeg62/jpeglib.h (libjpeg API 6.2).
 * Built twice by oracle/Makefile:
 *   libqsref_scalar.so  gcc -O2 -DNO_SIMD -ffp-contract=off       (parity oracle)
 *   libqsref_avx512.so  gcc -O2 -mavx512f -mavx512dq -mavx512bw -mfma -fopenmp
 *                                                                  (timed CPU baseline)
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load these objects.
 *
 * Exports: qsref_do_quantsmooth (= the reference's do_quantsmooth, renamed via
 * QS_NAME, reference quantsmooth.h:2401-2404) and thin wrappers that expose the
 * reference's `static` building blocks for unit-level parity tests.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "jpeglib.h"

static char qsref_log[4096];
static int qsref_log_len;
#define logfmt(...) do { \
	int n_ = snprintf(qsref_log + qsref_log_len, sizeof(qsref_log) - qsref_log_len, __VA_ARGS__); \
	if (n_ > 0 && qsref_log_len + n_ < (int)sizeof(qsref_log)) qsref_log_len += n_; } while (0)

#define WITH_LOG
#define TRANSCODE_ONLY
#define QS_NAME qsref_do_quantsmooth
#define JPEGQS_ATTR
#include "quantsmooth.h"

const char *qsref_take_log(void) { qsref_log_len = 0; return qsref_log; }

int qsref_num_procs(void) {
#ifdef _OPENMP
	return omp_get_num_procs();
#else
	return 1;
#endif
}

const char *qsref_variant(void) {
#if defined(NO_SIMD)
	return "scalar";
#elif defined(USE_AVX512)
	return "avx512";
#elif defined(USE_AVX2)
	return "avx2";
#elif defined(USE_SSE2)
	return "sse2";
#else
	return "simulated";
#endif
}

void qsref_idct_islow(JCOEFPTR coef, JSAMPROW out, unsigned stride) { idct_islow(coef, out, stride); }
void qsref_idct_float(float *in, float *out) { idct_float(in, out); }
void qsref_fdct_float(float *in, float *out) { fdct_float(in, out); }
void qsref_fdct_clamp(float *buf, JCOEFPTR coef, UINT16 *quantval) { fdct_clamp(buf, coef, quantval); }

/* copies the 64 per-coefficient tables (natural index) into out[64][size] */
int qsref_tables(int flags, float *out) {
	int i, size = flags & JPEGQS_DIAGONALS ? 272 : 160;
	float **t = quantsmooth_init(flags);
	if (!t) return -1;
	for (i = 0; i < 64; i++) memcpy(out + i * size, t[i], size * sizeof(float));
	free(t);
	return size;
}

void qsref_quantsmooth_block(JCOEFPTR coef, UINT16 *quantval, JSAMPLE *image,
		JSAMPLE *image2, int stride, int flags, int luma) {
	float **t = NULL;
	if (!(flags & JPEGQS_LOW_QUALITY)) t = quantsmooth_init(flags);
	quantsmooth_block(coef, quantval, image, image2, stride, flags, t, luma);
	free(t);
}

void qsref_upsample_row(int w1, int y0, int y1, JSAMPLE *image, JSAMPLE *image2,
		int stride, JSAMPLE *image1, int stride1, JSAMPLE *mem, int st,
		int ww, int ws, int hs) {
	upsample_row(w1, y0, y1, image, image2, stride, image1, stride1, mem, st, ww, ws, hs);
}
