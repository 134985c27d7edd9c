// Bisecting TMA tile loads on B200: the CUDA programming guide's example generalised over element type (int32 / u8),
// rank (2 / 3) and box origin (aligned / unaligned / out of bounds).   usage: tma_guide <u8:0|1> <rank> <x> <y>
#include <cuda.h>
#include <cuda/barrier>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define GW 1024
#define GH 1024
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int rank, int x, int y, int bytes, unsigned *out)
{
    __shared__ alignas(128) unsigned smem_buffer[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (rank == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(&smem_buffer, &tensor_map, x, y, 0, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    out[threadIdx.x] = smem_buffer[threadIdx.x];
}
int main(int argc, char **argv)
{
    const int u8 = atoi(argv[1]), rank = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]);
    const int es_ = u8 ? 1 : 4;
    unsigned char *t; cudaMalloc(&t, (size_t)GW * GH * es_ * 2);
    unsigned char *h = (unsigned char *)malloc((size_t)GW * GH * es_ * 2);
    for (size_t i = 0; i < (size_t)GW * GH * es_ * 2; ++i) h[i] = (unsigned char)(i * 7 + (i >> 10));
    cudaMemcpy(t, h, (size_t)GW * GH * es_ * 2, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    auto fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    CUtensorMap m{};
    uint64_t size[3] = {GW, GH, 2}; uint64_t stride[2] = {(uint64_t)GW * es_, (uint64_t)GW * GH * es_};
    uint32_t box[3] = {32, u8 ? 32u : 8u, 1}; uint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&m, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, rank, t, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    unsigned *out; cudaMalloc(&out, 256 * 4);
    kernel<<<1, 256>>>(m, rank, x, y, 1024, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned ho[256]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    // expected first word of the box
    unsigned want = 0;
    for (int b = 0; b < 4; ++b) {
        long xx = (long)x * es_ + b, yy = y;
        unsigned char v = (xx >= 0 && xx < (long)GW * es_ && yy >= 0 && yy < GH) ? h[(size_t)yy * GW * es_ + xx] : 0;
        want |= (unsigned)v << (8 * b);
    }
    printf("u8=%d rank=%d origin=(%d,%d): encode %d sync '%s' word0=%08x want %08x\n", u8, rank, x, y, (int)r, cudaGetErrorString(e), ho[0], want);
    return 0;
}
