// CPU check of sage_b200/csrc/glibc_log.cuh (compiled by tests/test_glibc_log.py with g++): evaluates both variants of glibc's log()
// exactly as the device does and counts the inputs on which each differs from this host's libm log().
#include "glibc_log.cuh"

#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char** argv) {
    std::mt19937_64 rng(12345);
    const long n = argc > 1 ? atol(argv[1]) : 1000000;
    long bad[2] = {0, 0};
    auto chk = [&](double x) {
        volatile double vx = x;
        const double ref = std::log(vx);
        const double a = sb::glog::glibc_log<true>(x), b = sb::glog::glibc_log<false>(x);
        if (memcmp(&a, &ref, 8) && !(a != a && ref != ref)) bad[0]++;
        if (memcmp(&b, &ref, 8) && !(b != b && ref != ref)) bad[1]++;
    };
    for (long i = 0; i < n; i++) {
        const uint64_t u = rng();
        double x;
        switch (i % 7) {
            case 0: { uint64_t b = u & 0x7fffffffffffffffull; memcpy(&x, &b, 8); break; }                   // any non-negative bit pattern
            case 1: x = 0.95 + (double)(u >> 11) * 0x1p-53 * 0.1; break;                                    // near 1
            case 2: { float a = (float)(u & 0xffffff) * 0.37f + 1.0f, b = (float)((u >> 24) & 0xffffff) * 1.91f + 1.0f; x = (double)a * (double)b; break; }
            case 3: x = (double)(u >> 11) * 0x1p-53 * 100.0; break;                                         // lambda-like
            case 4: x = std::exp(((double)(u >> 11) * 0x1p-53 - 0.5) * 1400.0); break;
            case 5: x = 0.93 + (double)(u >> 11) * 0x1p-53 * 0.15; break;                                   // edges of the near-1 interval
            default: { uint64_t b = (u & 0x000fffffffffffffull); memcpy(&x, &b, 8); break; }                // subnormals
        }
        chk(x);
    }
    const double sp[] = {0.0, -0.0, -1.0, 1.0, INFINITY, NAN, 0x1p-1074, 0x1.fffffffffffffp1023, 0.9375, 1.064697265625, 0x1.dffffffffffffp-1, 0x1.109p0};
    for (double x : sp) chk(x);
    printf("n=%ld variant0=%ld variant1=%ld\n", n, bad[0], bad[1]);
    return 0;
}
