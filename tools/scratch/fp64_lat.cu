// micro-benchmark: FP64 dependent-issue latency / throughput, rsqrt(double), shared-memory FP64 atomicAdd on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int mode)
{
    __shared__ double sm[64];
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
    if (threadIdx.x < 64) sm[threadIdx.x] = 0;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) { for (int i = 0; i < 1024; ++i) a = fma(a, b, c); }
    else if (mode == 1) { double a2 = a + 1, a3 = a + 2, a4 = a + 3; for (int i = 0; i < 256; ++i) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); } a += a2 + a3 + a4; }
    else if (mode == 2) { for (int i = 0; i < 256; ++i) a = rsqrt(a) + 1.5; }
    else if (mode == 3) { for (int i = 0; i < 256; ++i) a = 1.0 / a + 1.5; }
    else if (mode == 4) { for (int i = 0; i < 256; ++i) a = sqrt(a) + 1.5; }
    else if (mode == 5) { for (int i = 0; i < 256; ++i) atomicAdd(&sm[threadIdx.x & 63], a); }
    else if (mode == 6) { for (int i = 0; i < 256; ++i) atomicAdd(&sm[0], a); }
    else if (mode == 7) { for (int i = 0; i < 256; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0; }
    else if (mode == 8) { volatile double *v = sm; for (int i = 0; i < 256; ++i) { v[threadIdx.x & 63] = a; a = v[(threadIdx.x + 1) & 63] + 1.0; } }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + sm[threadIdx.x & 63];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double *o; long long *c, h[1];
    cudaMalloc(&o, 8 * 1024 * 148); cudaMalloc(&c, 8 * 148);
    const char *names[] = {"dfma chain x1024", "4 indep dfma chains x256", "rsqrt chain x256", "1/x chain x256", "sqrt chain x256",
                           "smem atomicAdd f64, 64 addrs, x256", "smem atomicAdd f64, 1 addr, x256", "shfl+dadd chain x256", "sts+lds chain x256"};
    for (int threads : {32, 512}) for (int m = 0; m < 9; ++m) {
        k<<<1, threads>>>(o, c, m); cudaDeviceSynchronize();
        k<<<1, threads>>>(o, c, m); cudaDeviceSynchronize();
        cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d  %-40s %8lld cycles\n", threads, names[m], h[0]);
    }
    return 0;
}
