/*
 * do_quantsmooth.c - the libjpeg-facing entry points of libjpegqs (include/libjpegqs.h),
 * kept in C like the reference's host code.  They do what the reference driver does on
 * the libjpeg side of the boundary (reference quantsmooth.h:2404-2453 argument handling,
 * 2836-2876 result plumbing, 2880-2904 decode helpers) and hand the arithmetic to the CUDA
 * back end through the C ABI of include/jpegqs_cuda.h: each component's block rows are
 * gathered from libjpeg's virtual arrays (access_virt_barray, one row contiguous, rows not
 * necessarily adjacent - SURVEY.md 8b) into one pinned array, smoothed on the device, and
 * scattered back in place.
 *
 * There is no CPU implementation here: if the CUDA back end cannot run, the call reports
 * the error on stderr and returns a negative code with the coefficients untouched.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <jpeglib.h>
#ifdef JPEGQS_COMPAT_JPEGLIB
#ifndef TRANSCODE_ONLY
#define TRANSCODE_ONLY          /* the stand-in header has no decoder internals */
#endif
#endif
#include "libjpegqs.h"
#include "jpegqs_cuda.h"

#ifndef JPEGQS_NO_LOG
#define WITH_LOG
#define logfmt(...) fprintf(stderr, __VA_ARGS__)
#include <sys/time.h>
static int64_t now_usec(void) {
	struct timeval tv; gettimeofday(&tv, NULL);
	return tv.tv_sec * (int64_t)1000000 + tv.tv_usec;
}
#endif

#if !defined(TRANSCODE_ONLY) && !defined(JPEG_INTERNALS)
/* decoder re-initialisation hooks from jpegint.h, used when the caller goes on to
 * jpeg_read_scanlines (reference quantsmooth.h:33-61, 2861-2876) */
#define DSTATE_SCANNING 205
#define DSTATE_RAW_OK 206
EXTERN(void) jinit_d_main_controller(j_decompress_ptr, boolean);
EXTERN(void) jinit_inverse_dct(j_decompress_ptr);
EXTERN(void) jinit_upsampler(j_decompress_ptr);
EXTERN(void) jinit_color_deconverter(j_decompress_ptr);
#endif

/* one lazily created context per device ordinal; callers are serialised like the
 * reference's plugin does with its own lock (irfanview/plugin.c:148-162) */
#define QS_MAX_DEVICES 16
static jpegqs_cuda_ctx *g_ctx[QS_MAX_DEVICES + 1];

static jpegqs_cuda_ctx *get_ctx(int ordinal_plus1) {
	int slot = ordinal_plus1 < 0 || ordinal_plus1 > QS_MAX_DEVICES ? 0 : ordinal_plus1;
	if (!g_ctx[slot]) {
		int rc = jpegqs_cuda_create(slot ? slot - 1 : -1, &g_ctx[slot]);
		if (rc) {
			fprintf(stderr, "jpegqs: CUDA back end unavailable: %s\n", jpegqs_cuda_last_error(NULL));
			g_ctx[slot] = NULL;
		}
	}
	return g_ctx[slot];
}

JPEGQS_ATTR
int do_quantsmooth(j_decompress_ptr srcinfo, jvirt_barray_ptr *coef_arrays, jpegqs_control_t *opts) {
	jpegqs_cuda_image img; jpegqs_cuda_ctx *ctx;
	jpeg_component_info *comp = srcinfo->comp_info;
	int ci, i, ret, ncomp = srcinfo->num_components, niter = opts->niter;
	int need_downsample = 0, flags = opts->flags;
	int16_t *host[JPEGQS_CUDA_MAX_COMP], *host_up[2] = { NULL, NULL };
	JDIMENSION y;
#ifdef WITH_LOG
	int64_t t0 = 0;
	if (flags & JPEGQS_INFO_COMP1)
		for (ci = 0; ci < ncomp; ci++)
			logfmt("component[%i] : table %i, samp %ix%i\n", ci, comp[ci].quant_tbl_no,
					comp[ci].h_samp_factor, comp[ci].v_samp_factor);
	if (flags & JPEGQS_INFO_QUANT)
		for (i = 0; i < NUM_QUANT_TBLS; i++) {
			JQUANT_TBL *t = srcinfo->quant_tbl_ptrs[i]; int k;
			if (!t) continue;
			logfmt("quant[%i]:\n", i);
			for (k = 0; k < DCTSIZE2; k++) {
				logfmt("%04x ", t->quantval[k]);
				if ((k & 7) == 7) logfmt("\n");
			}
		}
	if (flags & JPEGQS_INFO_TIME) t0 = now_usec();
#endif
#define LOG_COMP2 if (flags & JPEGQS_INFO_COMP2) for (ci = 0; ci < ncomp; ci++) if (comp[ci].quant_table) \
	logfmt("component[%i] : size %ix%i\n", ci, comp[ci].width_in_blocks, comp[ci].height_in_blocks);
	if (ncomp < 1 || ncomp > JPEGQS_CUDA_MAX_COMP) return 0;

	/* same early exit as the reference (quantsmooth.h:2447-2458): nothing is touched */
	if (flags & (JPEGQS_JOINT_YUV | JPEGQS_UPSAMPLE_UV) && srcinfo->jpeg_color_space == JCS_YCbCr &&
			ncomp >= 3 && comp[1].h_samp_factor == 1 && comp[1].v_samp_factor == 1 &&
			comp[2].h_samp_factor == 1 && comp[2].v_samp_factor == 1) need_downsample = 1;
	if (niter > JPEGQS_ITER_MAX) niter = JPEGQS_ITER_MAX;
	if (niter <= 0 && !(flags & JPEGQS_UPSAMPLE_UV && need_downsample)) return 0;

	ctx = get_ctx((flags >> JPEGQS_CPU_SHIFT) & JPEGQS_CPU_MASK);
	if (!ctx) return JPEGQS_ERR_CUDA;
#ifdef WITH_LOG
	if (flags & JPEGQS_INFO_CPU) logfmt("SIMD type: CUDA sm_100a (%s)\n", jpegqs_cuda_device_name(ctx));
	LOG_COMP2                                          /* quantsmooth.h:2569-2572 */
#endif

	memset(&img, 0, sizeof(img));
	memset(host, 0, sizeof(host));
	img.ncomp = ncomp;
	img.is_ycbcr = srcinfo->jpeg_color_space == JCS_YCbCr;
	img.image_width = srcinfo->image_width; img.image_height = srcinfo->image_height;
	ret = JPEGQS_ERR_CUDA;
	for (ci = 0; ci < ncomp; ci++) {
		jpegqs_cuda_comp *c = &img.comp[ci];
		size_t rowb = (size_t)comp[ci].width_in_blocks * sizeof(JBLOCK);
		c->wblk = comp[ci].width_in_blocks; c->hblk = comp[ci].height_in_blocks;
		c->h_samp = comp[ci].h_samp_factor; c->v_samp = comp[ci].v_samp_factor;
		c->has_qtbl = comp[ci].quant_table != NULL;
		if (c->has_qtbl) memcpy(c->quant, comp[ci].quant_table->quantval, sizeof(c->quant));
		host[ci] = (int16_t*)jpegqs_cuda_host_alloc(rowb * c->hblk);
		if (!host[ci]) goto done;
		c->coef = host[ci];
		for (y = 0; y < c->hblk; y++) {
			JBLOCKARRAY rows = (*srcinfo->mem->access_virt_barray)
					((j_common_ptr)srcinfo, coef_arrays[ci], y, 1, TRUE);
			memcpy((char*)host[ci] + y * rowb, rows[0], rowb);
		}
	}
	if (need_downsample && flags & JPEGQS_UPSAMPLE_UV &&
			(comp[0].h_samp_factor != 1 || comp[0].v_samp_factor != 1))
		for (i = 0; i < 2; i++) {
			host_up[i] = (int16_t*)jpegqs_cuda_host_alloc((size_t)img.comp[0].wblk * img.comp[0].hblk * sizeof(JBLOCK));
			if (!host_up[i]) goto done;
			img.comp[1 + i].coef_up = host_up[i];
		}

	ret = jpegqs_cuda_run_host(ctx, &img, flags & JPEGQS_FLAGS_MASK, opts->niter, opts->progprec,
			opts->progress, opts->userdata);
	if (ret < 0) {
		fprintf(stderr, "jpegqs: CUDA back end failed (%d): %s\n", ret, jpegqs_cuda_last_error(ctx));
		goto done;
	}

	for (ci = 0; ci < ncomp; ci++) {
		jpegqs_cuda_comp *c = &img.comp[ci];
		size_t rowb = (size_t)c->wblk * sizeof(JBLOCK);
		if (img.upsampled && (ci == 1 || ci == 2)) continue;
		for (y = 0; y < c->hblk; y++) {
			JBLOCKARRAY rows = (*srcinfo->mem->access_virt_barray)
					((j_common_ptr)srcinfo, coef_arrays[ci], y, 1, TRUE);
			memcpy(rows[0], (char*)host[ci] + y * rowb, rowb);
		}
	}
	if (img.upsampled) {                               /* quantsmooth.h:2700-2702, 2836-2849 */
		JDIMENSION W0 = comp[0].width_in_blocks, H0 = comp[0].height_in_blocks;
		size_t rowb = (size_t)W0 * sizeof(JBLOCK);
		jvirt_barray_ptr up[2];
		for (i = 0; i < 2; i++)
			up[i] = (*srcinfo->mem->request_virt_barray)((j_common_ptr)srcinfo, JPOOL_IMAGE, FALSE, W0, H0, 1);
		(*srcinfo->mem->realize_virt_arrays)((j_common_ptr)srcinfo);
		for (i = 0; i < 2; i++) {
			for (y = 0; y < H0; y++) {
				JBLOCKARRAY rows = (*srcinfo->mem->access_virt_barray)((j_common_ptr)srcinfo, up[i], y, 1, TRUE);
				memcpy(rows[0], (char*)host_up[i] + y * rowb, rowb);
			}
			coef_arrays[1 + i] = up[i];
			comp[1 + i].width_in_blocks = W0; comp[1 + i].height_in_blocks = H0;
		}
		srcinfo->max_h_samp_factor = 1; srcinfo->max_v_samp_factor = 1;
		comp[0].h_samp_factor = 1; comp[0].v_samp_factor = 1;
	}
	for (i = 0; i < NUM_QUANT_TBLS; i++) {             /* quantsmooth.h:2851-2859 */
		JQUANT_TBL *t = srcinfo->quant_tbl_ptrs[i]; int k;
		if (t) for (k = 0; k < DCTSIZE2; k++) t->quantval[k] = 1;
	}
	for (ci = 0; ci < ncomp; ci++) {
		JQUANT_TBL *t = comp[ci].quant_table; int k;
		if (t) for (k = 0; k < DCTSIZE2; k++) t->quantval[k] = 1;
	}
#ifndef TRANSCODE_ONLY
	if (!(flags & JPEGQS_TRANSCODE)) {                 /* quantsmooth.h:2861-2876 */
		if (img.upsampled) {
#ifdef LIBJPEG_TURBO_VERSION
			srcinfo->master->last_MCU_col[1] = srcinfo->master->last_MCU_col[0];
			srcinfo->master->last_MCU_col[2] = srcinfo->master->last_MCU_col[0];
#endif
			jinit_color_deconverter(srcinfo);
			jinit_upsampler(srcinfo);
			jinit_d_main_controller(srcinfo, FALSE);
			srcinfo->input_iMCU_row = (srcinfo->output_height + DCTSIZE - 1) / DCTSIZE;
		}
		jinit_inverse_dct(srcinfo);
	}
#endif
#ifdef WITH_LOG
	if (!ret && flags & JPEGQS_INFO_TIME) {
		logfmt("quantsmooth: %.3fms\n", (now_usec() - t0) * 0.001);
		logfmt("quantsmooth (device kernels): %.3fms\n", jpegqs_cuda_last_device_ms(ctx));
	}
#endif
done:
	for (ci = 0; ci < ncomp; ci++) jpegqs_cuda_host_free(host[ci]);
	jpegqs_cuda_host_free(host_up[0]); jpegqs_cuda_host_free(host_up[1]);
	return ret;
}

#ifndef TRANSCODE_ONLY
/* decode helpers, reference quantsmooth.h:2880-2904: read every scan in buffered-image
 * mode, smooth the coefficient arrays, then let the caller pull scanlines as usual */
JPEGQS_ATTR
boolean jpegqs_start_decompress(j_decompress_ptr cinfo, jpegqs_control_t *opts) {
	boolean ok;
	int active = opts->niter > 0 || (opts->flags & JPEGQS_UPSAMPLE_UV);
	if (active) cinfo->buffered_image = TRUE;
	ok = jpeg_start_decompress(cinfo);
	if (!active) return ok;
	while (!jpeg_input_complete(cinfo)) {
		jpeg_start_output(cinfo, cinfo->input_scan_number);
		jpeg_finish_output(cinfo);
	}
	do_quantsmooth(cinfo, jpeg_read_coefficients(cinfo), opts);
	jpeg_start_output(cinfo, cinfo->input_scan_number);
	return ok;
}

JPEGQS_ATTR
boolean jpegqs_finish_decompress(j_decompress_ptr cinfo) {
	if (cinfo->buffered_image &&
			(cinfo->global_state == DSTATE_SCANNING || cinfo->global_state == DSTATE_RAW_OK))
		jpeg_finish_output(cinfo);
	return jpeg_finish_decompress(cinfo);
}
#endif
