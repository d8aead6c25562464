/*
 * Minimal stand-in for libjpeg's <jpeglib.h>.
 *
 * The real libjpeg headers are not installed in the build image, so this file
 * declares ONLY the public types and fields that the coefficient-smoothing
 * path touches (the list in SURVEY.md section 8c; the uses are at
 * reference quantsmooth.h:2404-2878).  Field names and meanings follow the
 * public libjpeg API; the struct layouts are NOT ABI-compatible with a real
 * libjpeg - a production build must compile csrc/do_quantsmooth.c against the
 * system <jpeglib.h> instead.
 */
#ifndef JPEGQS_COMPAT_JPEGLIB_H
#define JPEGQS_COMPAT_JPEGLIB_H

#include <stddef.h>

#define JPEGQS_COMPAT_JPEGLIB 1

#define DCTSIZE 8
#define DCTSIZE2 64
#define NUM_QUANT_TBLS 4
#define MAX_COMPONENTS 10
#define BITS_IN_JSAMPLE 8
#define MAXJSAMPLE 255
#define CENTERJSAMPLE 128
#define JPOOL_PERMANENT 0
#define JPOOL_IMAGE 1

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define EXTERN(type) extern type

typedef int boolean;
typedef unsigned char JSAMPLE;
typedef JSAMPLE *JSAMPROW;
typedef JSAMPROW *JSAMPARRAY;
typedef short JCOEF;
typedef JCOEF JBLOCK[DCTSIZE2];
typedef JBLOCK *JBLOCKROW;
typedef JBLOCKROW *JBLOCKARRAY;
typedef JCOEF *JCOEFPTR;
typedef unsigned short UINT16;
typedef unsigned int JDIMENSION;

typedef enum {
	JCS_UNKNOWN, JCS_GRAYSCALE, JCS_RGB, JCS_YCbCr, JCS_CMYK, JCS_YCCK
} J_COLOR_SPACE;

typedef struct { UINT16 quantval[DCTSIZE2]; boolean sent_table; } JQUANT_TBL;

typedef struct {
	int component_id, component_index;
	int h_samp_factor, v_samp_factor;
	int quant_tbl_no;
	JDIMENSION width_in_blocks, height_in_blocks;
	JQUANT_TBL *quant_table;
} jpeg_component_info;

typedef struct jvirt_barray_control *jvirt_barray_ptr;
typedef struct jpeg_common_struct *j_common_ptr;
typedef struct jpeg_decompress_struct *j_decompress_ptr;

struct jpeg_memory_mgr {
	jvirt_barray_ptr (*request_virt_barray)(j_common_ptr cinfo, int pool_id,
			boolean pre_zero, JDIMENSION blocksperrow, JDIMENSION numrows,
			JDIMENSION maxaccess);
	void (*realize_virt_arrays)(j_common_ptr cinfo);
	JBLOCKARRAY (*access_virt_barray)(j_common_ptr cinfo, jvirt_barray_ptr ptr,
			JDIMENSION start_row, JDIMENSION num_rows, boolean writable);
	long max_memory_to_use;
};

struct jpeg_common_struct { struct jpeg_memory_mgr *mem; void *client_data; };

struct jpeg_decompress_struct {
	struct jpeg_memory_mgr *mem;
	void *client_data;
	JDIMENSION image_width, image_height;
	int num_components;
	J_COLOR_SPACE jpeg_color_space;
	JQUANT_TBL *quant_tbl_ptrs[NUM_QUANT_TBLS];
	jpeg_component_info *comp_info;
	int max_h_samp_factor, max_v_samp_factor;
};

#endif
