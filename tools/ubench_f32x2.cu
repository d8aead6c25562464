// Micro-benchmark: issue / pipe throughput of packed FP32x2 (FMUL2/FADD2) vs scalar FMUL/FADD on
// sm_100a, alone and mixed with ALU-pipe integer work.  Build: nvcc -gencode
// arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float mul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE> __global__ void k(float *out, int iters, float seed) {
	float a[8]; u64 p[8]; unsigned q[4];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
#pragma unroll
	for (int i = 0; i < 4; i++) q[i] = threadIdx.x * 7 + i;
	float c = 1.0000001f; u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			if (MODE == 0) {        // 16 scalar FP ops
#pragma unroll
				for (int i = 0; i < 8; i++) { a[i] = mul1(a[i], c); a[i] = add1(a[i], c); }
			} else if (MODE == 1) { // 8 packed ops = 16 FP ops
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = (i & 1) ? add2(p[i], c2) : mul2(p[i], c2); }
			} else if (MODE == 2) { // 16 packed ops = 32 FP ops
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = mul2(p[i], c2); p[i] = add2(p[i], c2); }
			} else if (MODE == 3) { // 8 packed + 8 ALU (LOP3/SHF) ops
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = (i & 1) ? add2(p[i], c2) : mul2(p[i], c2); q[i & 3] = (q[i & 3] >> 3) ^ (q[(i + 1) & 3] + 0x9e3779b9u); }
			} else if (MODE == 5) { // 16 packed multiplies only
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = mul2(p[i], c2); p[i] = mul2(p[i], c2); }
			} else if (MODE == 6) { // 16 packed adds only
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = add2(p[i], c2); p[i] = add2(p[i], c2); }
			} else if (MODE == 7) { // 16 packed FMAs
#pragma unroll
				for (int i = 0; i < 8; i++) { p[i] = fma2(p[i], c2, c2); p[i] = fma2(p[i], c2, c2); }
			} else if (MODE == 8) { // the smoothing term mix: 2 scalar adds + 5 mul2 + 2 add2, x2 chains
#pragma unroll
				for (int i = 0; i < 8; i += 2) {
					float ta = add1(a[i], c), tb = add1(a[i + 1], c);
					u64 t = ((u64)__float_as_uint(tb) << 32) | __float_as_uint(ta);
					t = mul2(t, t); u64 a0 = mul2(c2, t), a1 = mul2(p[i], t);
					p[i] = add2(p[i], mul2(a0, a1)); p[i + 1] = add2(p[i + 1], mul2(a1, a1));
				}
			} else if (MODE == 4) { // 16 scalar + 8 ALU
#pragma unroll
				for (int i = 0; i < 8; i++) { a[i] = mul1(a[i], c); a[i] = add1(a[i], c); q[i & 3] = (q[i & 3] >> 3) ^ (q[(i + 1) & 3] + 0x9e3779b9u); }
			}
		}
	}
	float s = 0; for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + q[0] + q[1] + q[2] + q[3];
}

template <int MODE> void run(const char *name, double fp_per_iter, double inst_per_iter) {
	float *out; cudaMalloc(&out, 148 * 8 * 512 * 4);
	int iters = 20000;
	k<MODE><<<148, 512>>>(out, 100, 1.0f);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0); k<MODE><<<148, 512>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double warps = 148.0 * 16, cyc = ms * 1e-3 * 1.965e9;
	printf("%-34s %8.3f ms  FP32 ops/clk/SM %7.1f   warp-instr/clk/SMSP %5.2f\n", name, ms,
		fp_per_iter * iters * 4 * warps * 32 / cyc / 148, inst_per_iter * iters * 4 * warps / cyc / 592);
	cudaFree(out);
}
int main() {
	run<0>("16 scalar FMUL/FADD", 16, 16);
	run<1>("8 FMUL2/FADD2 (=16 FP ops)", 16, 8);
	run<2>("16 FMUL2/FADD2 (=32 FP ops)", 32, 16);
	run<5>("16 FMUL2 only (=32 FP ops)", 32, 16);
	run<6>("16 FADD2 only (=32 FP ops)", 32, 16);
	run<7>("16 FFMA2 only (=32 FMA)", 32, 16);
	run<8>("term mix 4x(2 FADD+5 FMUL2+2 FADD2)", 4 * 16, 4 * 9);
	run<3>("8 FMUL2/FADD2 + 8x2 ALU ops", 16, 8 + 16);
	run<4>("16 scalar FP + 8x2 ALU ops", 16, 16 + 16);
	return 0;
}
