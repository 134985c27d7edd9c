// fp64_tput.cu -- FP64 throughput of one B200 SM: vector DFMA vs tensor-core DMMA (mma.sync m8n8k4), the two ways the
// dense contractions of the BA kernels (Cholesky trailing update, Schur product) can be issued.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double *out, const double *in, long long *cyc, int iters)
{
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = q;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) c[q] = fma(a, c[q], b);            // 8 independent DFMA chains per thread
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) dmma884(c[2 * q], c[2 * q + 1], a, b);   // 4 independent DMMA accumulators per warp
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += c[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double *o, *in; long long *c;
    cudaMalloc(&o, 8 * 1024 * 1024); cudaMalloc(&in, 8 * 64); cudaMalloc(&c, 8 * 1024);
    cudaMemset(in, 0, 8 * 64);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {32, 128, 512, 1024}) {
            for (int grid : {1, 148}) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                if (mode == 0) k<0><<<grid, threads>>>(o, in, c, iters); else k<1><<<grid, threads>>>(o, in, c, iters);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, threads>>>(o, in, c, iters); else k<1><<<grid, threads>>>(o, in, c, iters);
                cudaEventRecord(e1); cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                const double warps = threads / 32.0;
                const double fma_per_sm = mode == 0 ? warps * 32.0 * 8 * iters : warps * 4.0 * 256 * iters;
                printf("%s threads %4d grid %3d: %9lld cycles -> %.2f FMA/clk/SM (%.2f warp-instr/clk/SM), whole GPU %.2f TFLOP/s\n",
                       mode == 0 ? "DFMA" : "DMMA", threads, grid, h, fma_per_sm / h, fma_per_sm / h / (mode == 0 ? 32.0 : 256.0),
                       2.0 * fma_per_sm * grid / (ms * 1e-3) / 1e12);
            }
        }
    return 0;
}
