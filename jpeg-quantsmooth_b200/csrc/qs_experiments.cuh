/*
 * qs_experiments.cuh - kernel experiments that are NOT part of the shipped library.  Included by
 * qs_kernels.cu only with -DQS_EXPERIMENTS (make experiments -> libjpegqs_b200_exp.so).
 *
 *  - packed FP32x2 pair path (FMUL2 / FFMA2): bit exact, measured slower than the scalar path
 *    (2.95 vs 2.73 ms per 8K launch, profiles/README.md round 1), kept for reference together
 *    with tools/ubench_f32x2.cu;
 *  - the lock-step variants other than "barrier per chunk, 4 warps per sub-partition".
 */
#ifndef QS_EXPERIMENTS_CUH
#define QS_EXPERIMENTS_CUH

__constant__ QsChunk2 c_chunks2[QS_MAX_CHUNKS];
__constant__ int c_nchunks2, c_nslots2;
__constant__ unsigned long long c_one2;      /* {1.0f, 1.0f}, deliberately opaque to ptxas (see qs_add2) */

/* ------------------------------------------------------------------------------------------
 * Packed FP32x2 path.  Blackwell's FMUL2 / FADD2 (PTX mul/add.rn.f32x2) perform two IEEE-RN
 * FP32 operations per lane per issue slot at full rate (tools/ubench_f32x2.cu: 253 FP32
 * ops/clk/SM vs 114 scalar).  Two coefficients of an anti-diagonal advance through the same
 * pixel-difference terms in the two halves of 64-bit register pairs: per (term, pair)
 * 2 FADD.SAT + 5 FMUL2 + 2 FADD2 = 9 issue slots instead of 16, each lane still being the
 * reference's sequential, separately rounded sum (.rn forbids contraction into FFMA2).
 * ------------------------------------------------------------------------------------------ */
typedef unsigned long long qs_u64;
__device__ __forceinline__ qs_u64 qs_pk(float lo, float hi) {
	qs_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void qs_unpk(qs_u64 v, float &lo, float &hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ qs_u64 qs_mul2(qs_u64 a, qs_u64 b) {
	qs_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
/* a + b, per lane, IEEE RN.  NOT written as add.rn.f32x2: ptxas 12.9 contracts a mul.rn.f32x2
 * feeding an add.rn.f32x2 into one FFMA2 even with -fmad=false (it keeps the scalar mul.rn/add.rn
 * pair apart), which would drop the rounding of the product.  fma(b, 1.0, a) with the 1.0 pair
 * in a register ptxas cannot see through is exact (b*1 is exact, one rounding of the sum) and
 * cannot be merged with the producer of b. */
__device__ __forceinline__ qs_u64 qs_add2(qs_u64 a, qs_u64 b, qs_u64 one) {
	qs_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(one), "l"(a)); return r;
}

/* CORE (quantsmooth.h:1519-1520) for the two lanes of a pair; dd = {ds, ds}, w = {w_a, w_b} */
__device__ __forceinline__ void qs_term2(qs_u64 dd, float nad, qs_u64 w, float Rsa, float Rsb,
		qs_u64 &a2, qs_u64 &a3, qs_u64 one) {
	float ta, tb;
	asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(ta) : "f"(Rsa), "f"(nad));
	asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(tb) : "f"(Rsb), "f"(nad));
	qs_u64 t = qs_pk(ta, tb);
	t = qs_mul2(t, t);
	qs_u64 a0 = qs_mul2(dd, t), a1 = qs_mul2(w, t);
	a2 = qs_add2(a2, qs_mul2(a0, a1), one);
	a3 = qs_add2(a3, qs_mul2(a1, a1), one);
}

/* all terms of one row step for NP pairs; tab[c] = pair table (float2 per term) */
template <int NP, int NT, class Prep>
__device__ __forceinline__ void qs_terms_row2(const float *d, const float *const *tab, int off,
		const float *Rs, qs_u64 *a2, qs_u64 *a3, qs_u64 one, Prep prep) {
	qs_u64 dd[8]; float nad[8];
#pragma unroll
	for (int x = 0; x < NT; x++) { dd[x] = qs_pk(d[x], d[x]); nad[x] = -fabsf(d[x]); }
#pragma unroll
	for (int c = 0; c < NP; c++) {
		const ulonglong2 *t = (const ulonglong2 *)(tab[c] + off * 2);
		qs_u64 w[8];
#pragma unroll
		for (int k = 0; k < 4; k++) { ulonglong2 v = t[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
#pragma unroll
		for (int x = 0; x < NT; x++) qs_term2(dd[x], nad[x], w[x], Rs[2 * c], Rs[2 * c + 1], a2[c], a3[c], one);
		prep(c);
	}
}

template <int NP>
__device__ __forceinline__ void qs_prep_slice2(uint2 w, float *f, int c) {
#pragma unroll
	for (int j = 0; j < 8; j++) if (j * NP / 8 == c) f[j] = qs_px8(w, j);
}

template <int NP, bool DIAG>
__device__ __forceinline__ void qs_pair_sections(const uint2 *pw, const float *const *tab, const float *Rs,
		qs_u64 *a2, qs_u64 *a3) {
	float d[8];
	const qs_u64 one = c_one2;
	{                                                   /* horizontal, quantsmooth.h:1527 */
		float f[8];
		qs_unpack8(pw[0], f);
#pragma unroll
		for (int x = 0; x < 7; x++) d[x] = FS(f[x], f[x + 1]);
#pragma unroll 1
		for (int y = 0; y < 8; y++) {
			uint2 wn = pw[((y + 1) & 7) * 32];
			qs_terms_row2<NP, 7>(d, tab, y * 8, Rs, a2, a3, one, [&](int c) { qs_prep_slice2<NP>(wn, f, c); });
#pragma unroll
			for (int x = 0; x < 7; x++) d[x] = FS(f[x], f[x + 1]);
		}
	}
	{                                                   /* border, quantsmooth.h:1529-1530 */
		float fa[8], fb[8];
		qs_unpack8(pw[0], fa); qs_unpack8(pw[10 * 32], fb);
#pragma unroll
		for (int x = 0; x < 8; x++) d[x] = FS(fa[x], fb[x]);
#pragma unroll 1
		for (int s = 0; s < 4; s++) {
			int sn = (s + 1) & 3;
			int wi = sn == 0 ? 0 : sn == 1 ? 7 : 6 + sn;
			uint2 wa = pw[wi * 32], wb = pw[(10 + sn) * 32];
			qs_terms_row2<NP, 8>(d, tab, 64 + s * 8, Rs, a2, a3, one,
					[&](int c) { qs_prep_slice2<NP>(wa, fa, c); qs_prep_slice2<NP>(wb, fb, c); });
#pragma unroll
			for (int x = 0; x < 8; x++) d[x] = FS(fa[x], fb[x]);
		}
	}
	{                                                   /* vertical, quantsmooth.h:1531 */
		float fp[8], fn[8];
		qs_unpack8(pw[0], fn); qs_unpack8(pw[32], fp);
#pragma unroll
		for (int x = 0; x < 8; x++) d[x] = FS(fn[x], fp[x]);
#pragma unroll 1
		for (int y = 0; y < 7; y++) {
			uint2 wn = pw[min(y + 2, 7) * 32];
			qs_terms_row2<NP, 8>(d, tab, 96 + y * 8, Rs, a2, a3, one, [&](int c) { qs_prep_slice2<NP>(wn, fn, c); });
#pragma unroll
			for (int x = 0; x < 8; x++) { d[x] = FS(fp[x], fn[x]); fp[x] = fn[x]; }
		}
	}
	if (DIAG) {                                         /* diagonals, quantsmooth.h:1533-1540 */
		float fp[8], fn[8], d1[8], d2[8];
		qs_unpack8(pw[0], fn); qs_unpack8(pw[32], fp);
#pragma unroll
		for (int x = 0; x < 7; x++) { d1[x] = FS(fn[x], fp[x + 1]); d2[x] = FS(fn[x + 1], fp[x]); }
#pragma unroll 1
		for (int y = 0; y < 7; y++) {
			uint2 wn = pw[min(y + 2, 7) * 32];
			qs_u64 e1[8], e2[8];
#pragma unroll
			for (int x = 0; x < 7; x++) { e1[x] = qs_pk(d1[x], d1[x]); e2[x] = qs_pk(d2[x], d2[x]); }
#pragma unroll
			for (int c = 0; c < NP; c++) {
				const ulonglong2 *t = (const ulonglong2 *)(tab[c] + (160 + y * 16) * 2);
				qs_u64 w[16];
#pragma unroll
				for (int k = 0; k < 8; k++) { ulonglong2 v = t[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
#pragma unroll
				for (int x = 0; x < 7; x++) {
					qs_term2(e1[x], -fabsf(d1[x]), w[x], Rs[2 * c], Rs[2 * c + 1], a2[c], a3[c], one);
					qs_term2(e2[x], -fabsf(d2[x]), w[8 + x], Rs[2 * c], Rs[2 * c + 1], a2[c], a3[c], one);
				}
				qs_prep_slice2<NP>(wn, fn, c);
			}
#pragma unroll
			for (int x = 0; x < 7; x++) { d1[x] = FS(fp[x], fn[x + 1]); d2[x] = FS(fp[x + 1], fn[x]); }
#pragma unroll
			for (int x = 0; x < 8; x++) fp[x] = fn[x];
		}
	}
}

template <int NP, bool DIAG>
__device__ __forceinline__ void qs_chunk_pairs(const QsChunk2 &ch, const float *tabs, const uint2 *pw,
		const QsQuantDev *__restrict__ qd, uint16_t *cs, long long *msum) {
	const float *tab[NP]; float Rs[2 * NP]; qs_u64 a2[NP], a3[NP];
	const int TS = DIAG ? QS_TAB_DIAG : QS_TAB_PLAIN;
#pragma unroll
	for (int c = 0; c < NP; c++) {
		tab[c] = tabs + (int)ch.slot[c] * TS * 2;
		a2[c] = 0ull; a3[c] = 0ull;
#pragma unroll
		for (int h = 0; h < 2; h++) {
			int i = ch.idx[2 * c + h];
			Rs[2 * c + h] = i < 64 ? __ldg(&qd->Rs[i]) : 0.0f;
		}
	}
	qs_pair_sections<NP, DIAG>(pw, tab, Rs, a2, a3);
#pragma unroll
	for (int c = 0; c < NP; c++) {
		float x2a, x2b, x3a, x3b;
		qs_unpk(a2[c], x2a, x2b); qs_unpk(a3[c], x3a, x3b);
		if (ch.idx[2 * c] < 64) qs_coef_update<1>(&x2a, &x3a, &ch.idx[2 * c], qd, cs, msum);
		if (ch.idx[2 * c + 1] < 64) qs_coef_update<1>(&x2b, &x3b, &ch.idx[2 * c + 1], qd, cs, msum);
	}
}

#endif
