// fp64_lat2.cu -- dependent-chain latencies (cycles, one warp) of the FP64 operations on the serial spine of the BA kernels'
// Cholesky: rsqrt, reciprocal / division, sqrt, DFMA, DMMA (m8n8k4), double shuffle, float-seeded Newton rsqrt.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rsqrt_newton(double d)
{
    double r = (double)rsqrtf((float)d);
    r = r * (1.5 - 0.5 * d * r * r);
    r = r * (1.5 - 0.5 * d * r * r);
    return r;
}
template <int OP>
__global__ void k(double *out, long long *cyc, double x0, int iters)
{
    double x = x0 + threadIdx.x * 1e-3, y = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) x = rsqrt(x) + 1.5;
        if (OP == 1) x = 1.0 / x + 1.5;
        if (OP == 2) x = sqrt(x) + 1.5;
        if (OP == 3) x = fma(x, 0.999, 0.001);
        if (OP == 4) dmma884(x, y, x0, 1e-3);
        if (OP == 5) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 6) x = rsqrt_newton(x) + 1.5;
        if (OP == 7) x = x + 1.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 64); cudaMalloc(&c, 8);
    const char *names[8] = {"rsqrt(x)+1.5", "1/x+1.5", "sqrt(x)+1.5", "fma", "DMMA m8n8k4 (dependent accumulator)", "shfl double", "float-seeded Newton rsqrt +1.5", "add"};
    const int iters = 2048;
#define RUN(OP) k<OP><<<1, 32>>>(o, c, 2.0, iters); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(o, c, 2.0, iters); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %7.1f cycles per dependent step\n", names[OP], (double)h / iters);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    return 0;
}
