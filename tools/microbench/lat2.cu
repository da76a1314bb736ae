// dependent-chain latencies with the loop overhead amortised: 16 dependent ops per iteration (compile with -fmad=false)
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void k(double *out, double a, double b, long long *cyc)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    unsigned long long u = (unsigned long long)threadIdx.x + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (OP == 0) x = x + y;
            if (OP == 1) x = x * y;
            if (OP == 2) x = fma(x, y, y);
            if (OP == 3) { x = x * y; x = x - y; }                       // 2 ops
            if (OP == 4) u = u * 0x9E3779B97F4A7C15ull + 1;              // integer multiply-add chain
            if (OP == 5) x = (x > 1.5) ? x - y : x + y;                  // DSETP + select + add
            if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + (double)u;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "DMUL,DADD (2 ops)", "IMAD.WIDE chain (64-bit, 1 op)", "DSETP+FSEL+DADD", "SHFL (64-bit)"};
#define RUN(OP) k<OP><<<1, 32>>>(o, 1.000001, 0.999999, c); k<OP><<<1, 32>>>(o, 1.000001, 0.999999, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-36s %8.2f cycles per unrolled step\n", names[OP], (double)h / N / 16);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
