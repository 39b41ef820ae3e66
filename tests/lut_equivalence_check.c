/* CPU restatement (same IEEE-754 binary32 operations, no contraction) of two device-side search accelerators, checked against the plain
 * searches they replace on random inputs:
 *   1. lut_build_walk (sage_b200/csrc/kernels.cuh): start[c] = #{arr[i] < edge(c)} built by 128 "threads", each owning a run of consecutive cells
 *      (binary search for the first, sentinel-terminated forward walk for the others), must give the same table as one binary search per
 *      cell (lut_build), for ascending arrays with duplicates and clustered values. (lut_new is the O(cells + n) builder of round 1, kept as a
 *      second independent construction.)
 *   2. pep_partition: the precursor-mass LUT bracket [lut[c-1], lut[c+2]] must always contain the partition point, i.e. the bracketed binary
 *      search returns the same index as the full one, for query values inside, outside and exactly on array elements.
 * Exit status 0 iff no mismatch.  Build: gcc -O2 -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }
static float frand(void) { return (float)(rnd() >> 8) / 16777216.0f; }
static int cmpf(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return x < y ? -1 : x > y; }

static void lut_ref(const float* arr, uint32_t n, float base, float inv_w, uint16_t* start, uint32_t cells) {   /* lut_build */
    for (uint32_t c = 0; c < cells; c++) {
        uint32_t lo = 0;
        if (c > 0 && inv_w > 0.0f) {
            const float e = base + (float)c * (1.0f / inv_w);
            uint32_t hi = n;
            while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (arr[m] < e) lo = m + 1; else hi = m; }
        }
        start[c] = (uint16_t)lo;
    }
}
static int cell_of(float m, float base, float inv_w, float w, uint32_t cells) {
    const float t = (m - base) * inv_w;
    int c = t > 0.0f ? (int)fminf(t, (float)(cells - 1)) : 0;
    while (c + 1 < (int)cells && base + (float)(c + 1) * w <= m) c++;
    while (c > 0 && base + (float)c * w > m) c--;
    return c;
}
static void lut_new(const float* arr, uint32_t n, float base, float inv_w, uint16_t* start, uint32_t cells) {   /* lut_build_sorted */
    if (!(inv_w > 0.0f)) { for (uint32_t c = 0; c < cells; c++) start[c] = 0; return; }
    const float w = 1.0f / inv_w;
    for (uint32_t j = 0; j <= n; j++) {
        const int c_prev = j == 0 ? -1 : cell_of(arr[j - 1], base, inv_w, w, cells);
        const int c_here = j == n ? (int)cells - 1 : cell_of(arr[j], base, inv_w, w, cells);
        for (int c = c_prev + 1; c <= c_here; c++) start[c] = (uint16_t)j;
    }
}

static void lut_walk(const float* arr /* arr[n] = +inf */, uint32_t n, float base, float inv_w, uint16_t* start, uint32_t cells) {   /* lut_build_walk */
    const uint32_t stride = 128, per = (cells + stride - 1) / stride;
    for (uint32_t t0 = 0; t0 < stride; t0++) {
        const uint32_t c0 = t0 * per, c1 = c0 + per < cells ? c0 + per : cells;
        if (!(inv_w > 0.0f)) { for (uint32_t c = c0; c < c1; c++) start[c] = 0; continue; }
        if (c0 >= c1) continue;
        uint32_t i = 0;
        if (c0 > 0) {
            const float e = base + (float)c0 * (1.0f / inv_w);
            uint32_t hi = n;
            while (i < hi) { const uint32_t m = (i + hi) >> 1; if (arr[m] < e) i = m + 1; else hi = m; }
        }
        start[c0] = (uint16_t)i;
        for (uint32_t c = c0 + 1; c < c1; c++) {
            const float e = base + (float)c * (1.0f / inv_w);
            while (arr[i] < e) i++;
            start[c] = (uint16_t)i;
        }
    }
}

int main(void) {
    unsigned long long bad = 0, cases = 0;
    enum { CELLS = 1024, NMAX = 400 };
    static float arr[NMAX + 1];
    static uint16_t a[CELLS], b[CELLS], w[CELLS];
    for (int it = 0; it < 20000; it++) {
        const uint32_t n = 1 + rnd() % NMAX;
        const int mode = it % 4;
        for (uint32_t i = 0; i < n; i++) {
            float x = 100.0f + 1900.0f * frand();
            if (mode == 1) x = 500.0f + 3.0f * frand();                 /* clustered */
            if (mode == 2) x = (float)(100 + rnd() % 40) * 1.5f;         /* many duplicates */
            if (mode == 3 && i > n / 2) x = 1.0e6f * frand() + 2000.0f;  /* long sparse tail */
            arr[i] = x;
        }
        qsort(arr, n, sizeof(float), cmpf);
        const float base = arr[0], w0 = (arr[n - 1] - arr[0]) / (float)CELLS;     /* lut_params */
        const float inv_w = (w0 > 0.0f && w0 < 3.0e38f) ? 1.0f / w0 : 0.0f;
        lut_ref(arr, n, base, inv_w, a, CELLS);
        lut_new(arr, n, base, inv_w, b, CELLS);
        arr[n] = INFINITY;
        lut_walk(arr, n, base, inv_w, w, CELLS);
        for (int c = 0; c < CELLS; c++) bad += (a[c] != b[c]) + (a[c] != w[c]);
        cases++;
    }
    printf("lut_build_walk / lut_build_sorted: %llu arrays, %llu mismatching cells\n", cases, bad);

    /* ---- pep_partition bracket */
    enum { PCELLS = 65536, NP = 300000 };
    static float mono[NP];
    static uint32_t plut[PCELLS + 1];
    unsigned long long pbad = 0, q = 0;
    for (int rep = 0; rep < 3; rep++) {
        for (uint32_t i = 0; i < NP; i++) mono[i] = rep == 2 ? 500.0f + (float)(rnd() % 2000) * 0.25f : 500.0f + 4500.0f * frand() * frand();
        qsort(mono, NP, sizeof(float), cmpf);
        const float base = mono[0], pw = (mono[NP - 1] - mono[0]) / (float)PCELLS, inv = 1.0f / pw;
        for (uint32_t c = 0; c <= PCELLS; c++) {   /* k_build_pep_lut */
            uint32_t lo = c == PCELLS ? NP : 0;
            if (c > 0 && c < PCELLS) { const float e = base + (float)c * (1.0f / inv); uint32_t hi = NP; while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (mono[m] < e) lo = m + 1; else hi = m; } }
            plut[c] = lo;
        }
        for (int k = 0; k < 400000; k++) {
            float x;
            switch (k % 5) {
                case 0: x = mono[rnd() % NP]; break;                                  /* exactly an element */
                case 1: x = nextafterf(mono[rnd() % NP], k & 8 ? 1e9f : -1e9f); break; /* one ulp off an element */
                case 2: x = base - 100.0f * frand(); break;                          /* below the table */
                case 3: x = mono[NP - 1] + 100.0f * frand(); break;                  /* above the table */
                default: x = 400.0f + 4800.0f * frand(); break;
            }
            for (int le = 0; le < 2; le++) {
                uint32_t flo = 0, fhi = NP;                                            /* full search (positive values: float order == total_cmp) */
                while (flo < fhi) { uint32_t m = (flo + fhi) >> 1; if (le ? mono[m] <= x : mono[m] < x) flo = m + 1; else fhi = m; }
                const float t = (x - base) * inv;
                int c = (int)fminf(fmaxf(floorf(t), -2.0f), (float)PCELLS + 2.0f);
                int c0 = c - 1, c1 = c + 2;
                if (c0 < 0) c0 = 0; if (c0 > PCELLS) c0 = PCELLS; if (c1 < 0) c1 = 0; if (c1 > PCELLS) c1 = PCELLS;
                uint32_t lo = plut[c0], hi = plut[c1];
                while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (le ? mono[m] <= x : mono[m] < x) lo = m + 1; else hi = m; }
                pbad += lo != flo;
                q++;
            }
        }
    }
    printf("pep_partition: %llu queries, %llu mismatches\n", q, pbad);
    return (bad | pbad) != 0;
}
