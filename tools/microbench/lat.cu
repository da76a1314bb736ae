// single-warp dependent-chain latencies on the target GPU (cycles per op), no FMA contraction (-fmad=false)
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(double *out, double a, double b, long long *cyc)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = x + y;
        if (OP == 1) x = x * y;
        if (OP == 2) x = x - x * y;               // mul then sub (no fma)
        if (OP == 3) x = x / y;
        if (OP == 4) x = sqrt(x) + y;
        if (OP == 5) x = __shfl_xor_sync(0xffffffffu, x, 1) + y;
        if (OP == 6) { sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1] + y; __syncwarp(); }
        if (OP == 7) x = fma(x, y, y);
        if (OP == 8) { __syncwarp(); x = x + y; }
        if (OP == 9) { if (threadIdx.x == (i & 31)) x = x / y; x = __shfl_sync(0xffffffffu, x, i & 31); }  // one-lane div + broadcast
        if (OP == 10) { float xf = (float)x; xf = xf * 1.0001f + 0.5f; x = xf; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 8);
    const char *names[] = {"DADD", "DMUL", "DMUL+DADD", "DDIV", "DSQRT+DADD", "SHFL(double)+DADD", "STS+syncwarp+LDS+DADD+syncwarp", "DFMA", "syncwarp+DADD", "1-lane DDIV + shfl bcast", "cvt+FFMA+cvt"};
#define RUN(OP) k<OP><<<1, 32>>>(o, 1.000001, 0.999999, c); k<OP><<<1, 32>>>(o, 1.000001, 0.999999, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-36s %8.1f cycles/iter\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    return 0;
}
