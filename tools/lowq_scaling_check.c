/* Checks on the CPU that the rescaled arithmetic of the LOW_QUALITY kernel (qs_kernels.cu, qs_lowq_kernel:
 * pixels as 1 + p * 2^-15, range * 2^-15, one saturating add for the clamp) gives the plain form's results
 * (reference quantsmooth.h:1162-1178, scalar branch): every (range, difference, weight) term bit for bit,
 * and 2e7 random pixels through the sums, the division and the truncation.
 *   gcc -O2 -ffp-contract=off -o /tmp/lowq_check tools/lowq_scaling_check.c -lm && /tmp/lowq_check */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
static float satadd(float a, float b) { float r = a + b; return r < 0.0f ? 0.0f : r > 1.0f ? 1.0f : r; }
static float px(int p) { uint32_t u = 0x3F800000u | ((uint32_t)p << 8); float f; memcpy(&f, &u, 4); return f; }
static int cvtt(float x) { return fabsf(x) < 2147483648.0f ? (int)x : (int)0x80000000; }
int main(void) {
	const float c0 = 2.0f, c1 = 2.0f * 0.70710678118654752440f;
	long bad = 0, n = 0;
	/* per term, exhaustive */
	for (int range = 0; range <= 128; range++) for (int d = -255; d <= 255; d++) for (int ci = 0; ci < 2; ci++) {
		float c = ci ? c1 : c0;
		float t0 = (float)d, t = (float)range - fabsf(t0); t = t < 0 ? 0 : t; t = t * t;
		float aw = c * t, p = (t0 * t) * aw, q = aw * aw;
		/* scaled */
		float ds = px(d > 0 ? d : 0) - px(d > 0 ? 0 : -d);           /* = d * 2^-15 */
		float rs = (float)range * 3.0517578125e-05f;
		float ts = satadd(rs, -fabsf(ds)); ts = ts * ts;
		float aws = c * ts, ps = (ds * ts) * aws, qs = aws * aws;
		if (ps * 0x1p75f != p || qs * 0x1p60f != q) bad++;
		n++;
	}
	printf("per-term: %ld cases, %ld mismatches\n", n, bad);
	/* whole pixel: 8 neighbours, random, incl. the division and the final truncation */
	srand(1); bad = 0;
	for (long it = 0; it < 20000000; it++) {
		int range = rand() % 129, a = rand() & 255, v[8]; float cc[8] = { c1, c0, c1, c0, c0, c1, c0, c1 };
		int spread = 1 + rand() % 255;
		for (int k = 0; k < 8; k++) { v[k] = a + rand() % (2 * spread + 1) - spread; v[k] = v[k] < 0 ? 0 : v[k] > 255 ? 255 : v[k]; }
		float a0 = 0, an = 0, a0s = 0, ans = 0, rs = (float)range * 3.0517578125e-05f;
		for (int k = 0; k < 8; k++) {
			float t0 = (float)(a - v[k]), t = (float)range - fabsf(t0); t = t < 0 ? 0 : t; t = t * t;
			float aw = cc[k] * t; a0 = a0 + (t0 * t) * aw; an = an + aw * aw;
			float ds = px(a) - px(v[k]), ts = satadd(rs, -fabsf(ds)); ts = ts * ts;
			float aws = cc[k] * ts; a0s = a0s + (ds * ts) * aws; ans = ans + aws * aws;
		}
		int r1 = a, r2 = a;
		if (an > 0.0f) r1 = cvtt((float)a - a0 / an);
		if (ans > 0.0f) r2 = cvtt((float)a - (a0s / ans) * 32768.0f);
		if (r1 != r2 || (an > 0.0f) != (ans > 0.0f)) { if (bad < 5) printf("mismatch: %d vs %d (an %g ans %g)\n", r1, r2, an, ans); bad++; }
	}
	printf("whole pixel: %ld mismatches\n", bad);
	return 0;
}
