/* Test-only: exhaustive check of the divider-free x / 6 of gcm_math.cuh (div6).
 * For every float bit pattern with biased exponent E (finite), compare
 *     q = fma(fma(-6, x*c, x), c, x*c),  c = RN(1/6)
 * with the IEEE quotient x / 6.0f.  Prints one line per exponent that has mismatches and the total;
 * the device code uses the FMA form only for exponents 2..254, where this program finds none.
 * usage: div6_check [stride]   (stride > 1 samples every stride-th mantissa; 1 = all 2^32 patterns)
 * build: gcc -O2 -ffp-contract=off -fopenmp -mfma div6_check.c -o div6_check -lm                  */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv)
{
	const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
	const volatile float one = 1.0f, six = 6.0f;
	const float c = one / six;
	unsigned long long bad_total = 0, bad_in_range = 0;
	#pragma omp parallel for reduction(+:bad_total,bad_in_range) schedule(dynamic, 1)
	for(int e = 0; e < 255; e++) {
		unsigned long long bad = 0;
		for(uint32_t s = 0; s < 2; s++)
			for(uint32_t m = (uint32_t)e % stride; m < (1u << 23); m += stride) {
				const float x = asf((s << 31) | ((uint32_t)e << 23) | m);
				const float qt = x / 6.0f;
				const float q0 = x * c;
				const float r = fmaf(-6.0f, q0, x);
				const float q1 = fmaf(r, c, q0);
				if(asu(qt) != asu(q1)) bad++;
			}
		if(bad) {
			#pragma omp critical
			printf("exponent %d: %llu mismatches\n", e, bad);
		}
		bad_total += bad;
		if(e >= 2) bad_in_range += bad;
	}
	printf("c = %.30g (0x%08x)\n", (double)c, asu(c));
	printf("total mismatches %llu, in exponents 2..254: %llu\n", bad_total, bad_in_range);
	return bad_in_range ? 1 : 0;
}
