/*
 * TEST INFRASTRUCTURE - not part of the product.
 *
 * A fake libjpeg front end: builds a jpeg_decompress_struct (the real libjpeg 6.2
 * layout, include/libjpeg62/jpeglib.h) plus an in-memory jpeg_memory_mgr around flat
 * coefficient arrays, calls a do_quantsmooth-shaped function (the reference's,
 * the C restatement's, or the product's) and copies the results back out.
 * This is how the parity tests drive every implementation through the SAME
 * libjpeg-facing boundary (reference libjpegqs.h:47-48).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "jpeglib.h"
#include "libjpegqs.h"

struct jvirt_barray_control {
	JDIMENSION w, h;
	JBLOCKROW *rows;
	JBLOCK *data;
	int owned;
	struct jvirt_barray_control *next;
};

typedef struct {
	struct jpeg_memory_mgr pub;
	struct jvirt_barray_control *head;
	int scatter;   /* non-zero: block rows are separately malloc'd (non-contiguous) */
} fake_mem;

static jvirt_barray_ptr fake_request(j_common_ptr cinfo, int pool, boolean pre_zero,
		JDIMENSION w, JDIMENSION h, JDIMENSION maxaccess) {
	fake_mem *m = (fake_mem*)cinfo->mem; JDIMENSION y;
	struct jvirt_barray_control *a = calloc(1, sizeof(*a));
	(void)pool; (void)pre_zero; (void)maxaccess;
	a->w = w; a->h = h; a->owned = 1;
	a->rows = malloc(sizeof(JBLOCKROW) * (h ? h : 1));
	if (m->scatter) {
		a->data = NULL;
		for (y = 0; y < h; y++) a->rows[y] = calloc(w ? w : 1, sizeof(JBLOCK));
	} else {
		a->data = calloc(((size_t)w * h) != 0 ? (size_t)w * h : 1, sizeof(JBLOCK));
		for (y = 0; y < h; y++) a->rows[y] = a->data + (size_t)y * w;
	}
	a->next = m->head; m->head = a;
	return a;
}
static void fake_realize(j_common_ptr cinfo) { (void)cinfo; }
static JBLOCKARRAY fake_access(j_common_ptr cinfo, jvirt_barray_ptr a,
		JDIMENSION start, JDIMENSION n, boolean writable) {
	(void)cinfo; (void)n; (void)writable;
	return a->rows + start;
}

typedef struct {
	int num_components, color_space;
	unsigned image_width, image_height;
	int h_samp[4], v_samp[4], quant_tbl_no[4];
	unsigned width_in_blocks[4], height_in_blocks[4];
	uint16_t quant[4][64];        /* NUM_QUANT_TBLS slots; slot_present mask below */
	int slot_present;
	int16_t *coef[4];             /* in: [h][w][64] quantized; out: result (same dims) */
	int16_t *coef_up[2];          /* out: buffers of luma dims for UPSAMPLE_UV, or NULL */
	int upsampled;                /* out: 1 if coef_arrays[1,2] were replaced */
	int max_h_samp, max_v_samp;   /* out */
	int scatter_rows;
} fake_image;

typedef int (*qs_fn)(j_decompress_ptr, jvirt_barray_ptr*, jpegqs_control_t*);

/* A session keeps the fake decompress object alive between calls, so that a benchmark can time
 * fn(&ci, arrays, opts) alone - exactly what a libjpeg application spends inside do_quantsmooth
 * (reference quantsmooth.c:550) - with the set-up (libjpeg's own decoding work in real life)
 * outside the timed region:
 *   s = fakejpeg_open(im); { fakejpeg_load(s, im); ret = fakejpeg_run(s, fn, opts); } ...
 *   fakejpeg_store(s, im); fakejpeg_close(s);                                              */
typedef struct {
	struct jpeg_decompress_struct ci; fake_mem mem;
	JQUANT_TBL tbl[4]; jpeg_component_info comp[4]; jvirt_barray_ptr arrays[4];
} fake_session;

static void fake_setup(fake_session *s, const fake_image *im, int reuse_arrays) {
	int i, c; JDIMENSION y;
	jvirt_barray_ptr keep[4]; struct jvirt_barray_control *head = NULL;
	if (reuse_arrays) { memcpy(keep, s->arrays, sizeof(keep)); head = s->mem.head; }
	memset(s, 0, sizeof(*s));
	if (reuse_arrays) { memcpy(s->arrays, keep, sizeof(keep)); s->mem.head = head; }
	s->mem.pub.request_virt_barray = fake_request;
	s->mem.pub.realize_virt_arrays = fake_realize;
	s->mem.pub.access_virt_barray = fake_access;
	s->mem.scatter = im->scatter_rows;
	s->ci.mem = &s->mem.pub;
	s->ci.image_width = im->image_width; s->ci.image_height = im->image_height;
	s->ci.num_components = im->num_components;
	s->ci.jpeg_color_space = (J_COLOR_SPACE)im->color_space;
	s->ci.comp_info = s->comp;
	for (i = 0; i < 4; i++) {
		memcpy(s->tbl[i].quantval, im->quant[i], sizeof(s->tbl[i].quantval));
		s->ci.quant_tbl_ptrs[i] = (im->slot_present >> i) & 1 ? &s->tbl[i] : NULL;
	}
	for (c = 0; c < im->num_components; c++) {
		size_t rowb = (size_t)im->width_in_blocks[c] * sizeof(JBLOCK);
		jpeg_component_info *cp = &s->comp[c];
		cp->component_index = c; cp->component_id = c + 1;
		cp->h_samp_factor = im->h_samp[c]; cp->v_samp_factor = im->v_samp[c];
		cp->quant_tbl_no = im->quant_tbl_no[c];
		cp->width_in_blocks = im->width_in_blocks[c];
		cp->height_in_blocks = im->height_in_blocks[c];
		cp->quant_table = s->ci.quant_tbl_ptrs[im->quant_tbl_no[c] & 3];
		if (cp->h_samp_factor > s->ci.max_h_samp_factor) s->ci.max_h_samp_factor = cp->h_samp_factor;
		if (cp->v_samp_factor > s->ci.max_v_samp_factor) s->ci.max_v_samp_factor = cp->v_samp_factor;
		if (!s->arrays[c]) s->arrays[c] = fake_request((j_common_ptr)&s->ci, JPOOL_IMAGE, FALSE,
				im->width_in_blocks[c], im->height_in_blocks[c], 1);
		for (y = 0; y < im->height_in_blocks[c]; y++)
			memcpy(s->arrays[c]->rows[y], (char*)im->coef[c] + y * rowb, rowb);
	}
}

static void fake_store(fake_session *s, fake_image *im) {
	int i, c; JDIMENSION y;
	im->upsampled = 0;
	for (c = 0; c < im->num_components; c++) {
		jvirt_barray_ptr a = s->arrays[c];
		int replaced = a->w != im->width_in_blocks[c] || a->h != im->height_in_blocks[c] ||
				s->comp[c].width_in_blocks != im->width_in_blocks[c] ||
				s->comp[c].height_in_blocks != im->height_in_blocks[c];
		int16_t *dst = im->coef[c]; size_t rowb;
		if (replaced) {
			im->upsampled = 1;
			dst = (c >= 1 && c <= 2) ? im->coef_up[c - 1] : NULL;
		}
		im->width_in_blocks[c] = s->comp[c].width_in_blocks;
		im->height_in_blocks[c] = s->comp[c].height_in_blocks;
		im->h_samp[c] = s->comp[c].h_samp_factor; im->v_samp[c] = s->comp[c].v_samp_factor;
		rowb = (size_t)a->w * sizeof(JBLOCK);
		if (dst) for (y = 0; y < a->h; y++)
			memcpy((char*)dst + y * rowb, a->rows[y], rowb);
	}
	im->max_h_samp = s->ci.max_h_samp_factor; im->max_v_samp = s->ci.max_v_samp_factor;
	for (i = 0; i < 4; i++) memcpy(im->quant[i], s->tbl[i].quantval, sizeof(s->tbl[i].quantval));
}

static void fake_free(fake_session *s) {
	JDIMENSION y;
	while (s->mem.head) {
		struct jvirt_barray_control *a = s->mem.head; s->mem.head = a->next;
		if (a->data) free(a->data);
		else for (y = 0; y < a->h; y++) free(a->rows[y]);
		free(a->rows); free(a);
	}
}

fake_session *fakejpeg_open(const fake_image *im) {
	fake_session *s = calloc(1, sizeof(*s));
	if (s) fake_setup(s, im, 0);
	return s;
}
/* (re)load the input coefficients and tables of `im` into the session's arrays; only valid
 * while the geometry is the original one (i.e. not after an UPSAMPLE_UV run) */
void fakejpeg_load(fake_session *s, const fake_image *im) { fake_setup(s, im, 1); }
int fakejpeg_run(fake_session *s, qs_fn fn, jpegqs_control_t *opts) { return fn(&s->ci, s->arrays, opts); }
void fakejpeg_store(fake_session *s, fake_image *im) { fake_store(s, im); }
void fakejpeg_close(fake_session *s) { if (s) { fake_free(s); free(s); } }

int fakejpeg_call(qs_fn fn, fake_image *im, jpegqs_control_t *opts) {
	fake_session s; int ret;
	fake_setup(&s, im, 0);
	ret = fn(&s.ci, s.arrays, opts);
	fake_store(&s, im);
	fake_free(&s);
	return ret;
}
