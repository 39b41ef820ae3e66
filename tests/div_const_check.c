/* Exhaustive check behind div_const_rn (sage_b200/csrc/device_common.cuh): for c in {1e6, 100, 3} and EVERY float x with 1e-20 <= x <= 1e30,
 * fma(fma(-q0, c, x), rc, q0) with q0 = x * rc, rc = rn(1 / c) equals the IEEE division x / c bit for bit (negative x follows by symmetry of
 * round-to-nearest). Prints "tested <n> bad <m>" per constant; exit status 0 iff every m is 0.  Build: gcc -O2 -fopenmp -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static unsigned long long check(float c) {
    const float rc = 1.0f / c;
    unsigned long long bad = 0, tested = 0;
#pragma omp parallel for reduction(+ : bad, tested) schedule(static)
    for (int64_t b = 0; b < (1ll << 31); b++) {
        const uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        if (!(x >= 1e-20f && x <= 1e30f)) continue;
        const float ref = x / c;
        const float q0 = x * rc;
        const float q = fmaf(fmaf(-q0, c, x), rc, q0);
        tested++;
        bad += memcmp(&q, &ref, 4) != 0;
    }
    printf("c=%g tested %llu bad %llu\n", (double)c, tested, bad);
    return bad;
}

/* 3: fragment / charge for triply charged fragments in k_score's straight-line task body (scoring.rs:707) */
int main(void) { return (check(1000000.0f) | check(100.0f) | check(3.0f)) != 0; }
