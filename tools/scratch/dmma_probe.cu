// probe for round 2: FP64 tensor-core MMA (mma.sync m8n8k4) on sm_100a -- compiles? which SASS?  Timing harness to be run
// on the GPU box: rank-8 update C(8x8) -= A(8x8) B(8x8)^T per warp, the shape of the Cholesky trailing update in k_ba_solve.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double *out, const double *in, long long *cyc, int iters)
{
    const int lane = threadIdx.x & 31;
    double c0 = 0, c1 = 0, a = in[lane], b = in[32 + lane];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma884(c0, c1, a, b); a += 1e-9; }
    long long t1 = clock64();
    out[threadIdx.x * 2] = c0; out[threadIdx.x * 2 + 1] = c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double *o, *in; long long *c, h;
    cudaMalloc(&o, 8 * 2048); cudaMalloc(&in, 8 * 64); cudaMalloc(&c, 8);
    cudaMemset(in, 0, 8 * 64);
    for (int threads : {32, 512}) {
        k<<<1, threads>>>(o, in, c, 1024); cudaDeviceSynchronize();
        k<<<1, threads>>>(o, in, c, 1024); cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d: 1024 dependent DMMA m8n8k4 in %lld cycles (%.1f per MMA; 256 FMA each)\n", threads, h, h / 1024.0);
    }
    return 0;
}
