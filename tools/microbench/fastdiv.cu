// checks: q = a/b (IEEE) == Markstein division with r = RN(1/b): q0 = a*r; 2 x { e = fma(-q,b,a); q = fma(e,r,q) }
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_div(double a, double b, double rb)
{
    double q = a * rb;
    double e = fma(-q, b, a);
    q = fma(e, rb, q);
    e = fma(-q, b, a);
    q = fma(e, rb, q);
    return q;
}
__device__ uint64_t sm64(uint64_t &s) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__global__ void k(unsigned long long *bad, unsigned long long *bad1, int iters, int mode)
{
    uint64_t s = (blockIdx.x * 1024ull + threadIdx.x) * 7919ull + mode * 104729ull;
    unsigned long long nb = 0, nb1 = 0;
    for (int i = 0; i < iters; i++) {
        uint64_t ra = sm64(s), rb = sm64(s);
        double a, b;
        if (mode == 0) { // random mantissas, exponents within +-60
            a = __longlong_as_double((ra & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 60 + (ra >> 52) % 121) << 52));
            b = __longlong_as_double((rb & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 60 + (rb >> 52) % 121) << 52));
        } else if (mode == 1) { // mantissas with long runs of ones / zeros (hard cases)
            uint64_t ma = (ra & 0xFFFFFFFFFFFFFull), mb = (rb & 0xFFFFFFFFFFFFFull);
            int sh = (ra >> 56) % 52; ma = (ra & 1) ? (ma >> sh) : ~(ma >> sh) & 0xFFFFFFFFFFFFFull;
            sh = (rb >> 56) % 52; mb = (rb & 1) ? (mb >> sh) : ~(mb >> sh) & 0xFFFFFFFFFFFFFull;
            a = __longlong_as_double(ma | (1023ull << 52)); b = __longlong_as_double(mb | (1022ull << 52));
        } else { // small integers / simple ratios
            a = (double)(int)(ra % 2000) - 1000.0; b = (double)(int)(rb % 2000) - 999.5;
        }
        const double r = 1.0 / b;
        const double q = a / b, f = fast_div(a, b, r);
        if (__double_as_longlong(q) != __double_as_longlong(f) && !(q == 0.0 && f == 0.0)) nb++;
        double q1 = a * r; double e = fma(-q1, b, a); q1 = fma(e, r, q1);
        if (__double_as_longlong(q) != __double_as_longlong(q1) && !(q == 0.0 && q1 == 0.0)) nb1++;
    }
    atomicAdd(bad, nb); atomicAdd(bad1, nb1);
}
int main()
{
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(d, 0, 16);
        k<<<148 * 8, 256>>>(d, d + 1, 20000, mode);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d: %.2e pairs: mismatches two-step %llu, one-step %llu\n", mode, 148.0 * 8 * 256 * 20000, h[0], h[1]);
    }
    return 0;
}
