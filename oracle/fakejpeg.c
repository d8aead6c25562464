/*
 * TEST INFRASTRUCTURE - not part of the product.
 *
 * A fake libjpeg front end: builds a jpeg_decompress_struct (compat layout,
 * include/compat/jpeglib.h) plus an in-memory jpeg_memory_mgr around flat
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

int fakejpeg_call(qs_fn fn, fake_image *im, jpegqs_control_t *opts) {
	struct jpeg_decompress_struct ci; fake_mem mem;
	JQUANT_TBL tbl[4]; jpeg_component_info comp[4]; jvirt_barray_ptr arrays[4];
	int i, c, ret; JDIMENSION y;
	memset(&ci, 0, sizeof(ci)); memset(&mem, 0, sizeof(mem));
	memset(tbl, 0, sizeof(tbl)); memset(comp, 0, sizeof(comp));
	mem.pub.request_virt_barray = fake_request;
	mem.pub.realize_virt_arrays = fake_realize;
	mem.pub.access_virt_barray = fake_access;
	mem.scatter = im->scatter_rows;
	ci.mem = &mem.pub;
	ci.image_width = im->image_width; ci.image_height = im->image_height;
	ci.num_components = im->num_components;
	ci.jpeg_color_space = (J_COLOR_SPACE)im->color_space;
	ci.comp_info = comp;
	for (i = 0; i < 4; i++) {
		memcpy(tbl[i].quantval, im->quant[i], sizeof(tbl[i].quantval));
		ci.quant_tbl_ptrs[i] = (im->slot_present >> i) & 1 ? &tbl[i] : NULL;
	}
	for (c = 0; c < im->num_components; c++) {
		size_t rowb = (size_t)im->width_in_blocks[c] * sizeof(JBLOCK);
		comp[c].component_index = c; comp[c].component_id = c + 1;
		comp[c].h_samp_factor = im->h_samp[c]; comp[c].v_samp_factor = im->v_samp[c];
		comp[c].quant_tbl_no = im->quant_tbl_no[c];
		comp[c].width_in_blocks = im->width_in_blocks[c];
		comp[c].height_in_blocks = im->height_in_blocks[c];
		comp[c].quant_table = ci.quant_tbl_ptrs[im->quant_tbl_no[c] & 3];
		if (comp[c].h_samp_factor > ci.max_h_samp_factor) ci.max_h_samp_factor = comp[c].h_samp_factor;
		if (comp[c].v_samp_factor > ci.max_v_samp_factor) ci.max_v_samp_factor = comp[c].v_samp_factor;
		arrays[c] = fake_request((j_common_ptr)&ci, JPOOL_IMAGE, FALSE,
				im->width_in_blocks[c], im->height_in_blocks[c], 1);
		for (y = 0; y < im->height_in_blocks[c]; y++)
			memcpy(arrays[c]->rows[y], (char*)im->coef[c] + y * rowb, rowb);
	}

	ret = fn(&ci, arrays, opts);

	im->upsampled = 0;
	for (c = 0; c < im->num_components; c++) {
		int replaced = arrays[c]->w != im->width_in_blocks[c] ||
				arrays[c]->h != im->height_in_blocks[c] ||
				comp[c].width_in_blocks != im->width_in_blocks[c] ||
				comp[c].height_in_blocks != im->height_in_blocks[c];
		int16_t *dst = im->coef[c]; size_t rowb;
		if (replaced) {
			im->upsampled = 1;
			dst = (c >= 1 && c <= 2) ? im->coef_up[c - 1] : NULL;
		}
		im->width_in_blocks[c] = comp[c].width_in_blocks;
		im->height_in_blocks[c] = comp[c].height_in_blocks;
		im->h_samp[c] = comp[c].h_samp_factor; im->v_samp[c] = comp[c].v_samp_factor;
		rowb = (size_t)arrays[c]->w * sizeof(JBLOCK);
		if (dst) for (y = 0; y < arrays[c]->h; y++)
			memcpy((char*)dst + y * rowb, arrays[c]->rows[y], rowb);
	}
	im->max_h_samp = ci.max_h_samp_factor; im->max_v_samp = ci.max_v_samp_factor;
	for (i = 0; i < 4; i++) memcpy(im->quant[i], tbl[i].quantval, sizeof(tbl[i].quantval));

	while (mem.head) {
		struct jvirt_barray_control *a = mem.head; mem.head = a->next;
		if (a->data) free(a->data);
		else for (y = 0; y < a->h; y++) free(a->rows[y]);
		free(a->rows); free(a);
	}
	return ret;
}
