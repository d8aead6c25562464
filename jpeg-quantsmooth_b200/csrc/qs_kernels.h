/* launch wrappers implemented in qs_kernels.cu (internal, C++ linkage) */
#ifndef QS_KERNELS_H
#define QS_KERNELS_H
#include <cuda_runtime.h>
#include "qs_common.h"

size_t qs_smooth_smem_bytes(int diag, int wpg);
cudaError_t qs_smooth_configure(void);
cudaError_t qs_set_chunks(const QsChunk *chunks, int n);
#ifdef QS_EXPERIMENTS
cudaError_t qs_set_chunks2(const QsChunk2 *chunks, int n, int nslots);
size_t qs_smooth_smem_bytes_x2(int diag, int nslots);
cudaError_t qs_launch_smooth_x2(const QsJob *jobs_dev, int njobs, int total_tiles, const float *tables_dev,
		int nslots, int *tile_counter, int flags, int clamp_out, int num_sms, int sync, cudaStream_t st);
#endif
cudaError_t qs_launch_idct_pass(const QsJob *jobs_dev, int njobs, int total_tiles, int mode,
		int *bad_flags, cudaStream_t st);
cudaError_t qs_launch_smooth(const QsJob *jobs_dev, int njobs, int total_tiles, const float *tables_dev,
		int *tile_counter, int flags, int clamp_out, int num_sms, int sync, int wpg, const int *bad, cudaStream_t st);
cudaError_t qs_launch_lowq(const QsJob *jobs_dev, int njobs, int total_tiles, int flags, int clamp_out,
		const int *bad, cudaStream_t st);
cudaError_t qs_launch_stop_fixup(const QsJob *jobs_dev, int njobs, const int *bad, cudaStream_t st);
cudaError_t qs_launch_xchg(const QsXchgPush *push, const QsXchgPull *pull, cudaStream_t st);
cudaError_t qs_launch_scale_clamp(int16_t *coef, size_t n, const QsQuantDev *qd, int dequant, int clamp,
		cudaStream_t st);
cudaError_t qs_launch_downsample(const uint8_t *src, int sstride, int w, int h, uint8_t *dst, int dstride,
		int w2, int h2, int ws, int hs, int src_row0, int dst_row_first, int dst_rows, int h1_total,
		cudaStream_t st);
cudaError_t qs_launch_upsample(const uint8_t *C, const uint8_t *Yd, int cstride, const uint8_t *Yf, int ystride,
		uint8_t *out, int ostride, int w1, int h1, int ws, int hs, int ww, int hh, int oy0, cudaStream_t st);
cudaError_t qs_launch_fdct_plane(const uint8_t *px, int pstride, int16_t *coef, int W, int H, cudaStream_t st);
cudaError_t qs_launch_render_rgb(const uint8_t *const *planes, const int *strides, const int *cw, const int *ch,
		const int *hs, const int *vs, int ncomp, int width, int height, int ycc, uint8_t *rgb, cudaStream_t st);
extern "C" int qs_host_orig_coef(int c, int q);
/* control data without the copy engines (see qs_kernels.cu) */
#define QS_JOB_PACK 48
cudaError_t qs_store_jobs(QsJob *dst, const QsJob *jobs_host, int n, cudaStream_t st);
cudaError_t qs_copy_flags(const int *src_dev, int *dst_mapped, int n, cudaStream_t st);

#endif
