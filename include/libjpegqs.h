/*
 * libjpegqs.h - public API of the B200-native quantsmooth library.
 *
 * Source-compatible with the reference's libjpegqs.h (reference
 * libjpegqs.h:14-55): same flag values, same jpegqs_control_t layout and the
 * same three entry points, so a caller written against the reference
 * (quantsmooth.c:550, example.c:96, irfanview/plugin.c:103) links against
 * libjpegqs_b200.so unchanged.  Include <jpeglib.h> before this header.
 *
 * Differences in behaviour (all documented in INTEGRATION.md):
 *   - the smoothing runs on a CUDA device (sm_100a); there is no CPU
 *     fallback - if no device/extension is usable the call aborts through
 *     cinfo->err->error_exit when available, else returns JPEGQS_ERR_CUDA;
 *   - opts->threads is ignored (reference: OpenMP thread count,
 *     quantsmooth.h:2467-2472);
 *   - (flags >> JPEGQS_CPU_SHIFT) & JPEGQS_CPU_MASK selects the CUDA device
 *     ordinal + 1 (0 = current device) instead of a SIMD tier
 *     (reference libjpegqs.c:123-129).
 */
#ifndef JPEGQS_H
#define JPEGQS_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
	JPEGQS_ITER_MAX = 100,          /* niter is clamped to [0, 100]          */
	JPEGQS_DIAGONALS = 1,           /* q>=4: add the 98 diagonal pair terms  */
	JPEGQS_JOINT_YUV = 2,           /* q>=5: chroma predicted from luma      */
	JPEGQS_UPSAMPLE_UV = 4,         /* q>=6: chroma re-sampled at luma size  */
	JPEGQS_LOW_QUALITY = 8,         /* q<=2: one-shot 8-neighbour filter      */
	JPEGQS_NO_REBALANCE = 16,
	JPEGQS_NO_REBALANCE_UV = 32,
	JPEGQS_TRANSCODE = 64,
	JPEGQS_FLAGS_MASK = 0x7f,
	JPEGQS_CPU_SHIFT = 12,
	JPEGQS_CPU_MASK = 15,
	JPEGQS_INFO_SHIFT = 16,
	JPEGQS_INFO_COMP1 = 1 << JPEGQS_INFO_SHIFT,
	JPEGQS_INFO_QUANT = 2 << JPEGQS_INFO_SHIFT,
	JPEGQS_INFO_COMP2 = 4 << JPEGQS_INFO_SHIFT,
	JPEGQS_INFO_TIME = 8 << JPEGQS_INFO_SHIFT,
	JPEGQS_INFO_CPU = 16 << JPEGQS_INFO_SHIFT
};

#ifndef JPEGQS_ATTR
#define JPEGQS_ATTR
#endif

#define JPEGQS_VERSION "1.20230818-b200"

typedef struct {
	int flags, niter, threads, progprec;
	void *userdata;
	int (*progress)(void *data, int cur, int max);
} jpegqs_control_t;

/* replaces reference libjpegqs.h:47-48 / quantsmooth.h:2404 */
JPEGQS_ATTR
int do_quantsmooth(j_decompress_ptr srcinfo, jvirt_barray_ptr *coef_arrays,
		jpegqs_control_t *opts);

#ifndef TRANSCODE_ONLY
/* replace reference libjpegqs.h:50-55 / quantsmooth.h:2880-2904 */
JPEGQS_ATTR
boolean jpegqs_start_decompress(j_decompress_ptr cinfo, jpegqs_control_t *opts);

JPEGQS_ATTR
boolean jpegqs_finish_decompress(j_decompress_ptr cinfo);
#endif

#ifdef __cplusplus
}
#endif
#endif
