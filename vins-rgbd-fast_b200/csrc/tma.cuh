// tma.cuh -- thin inline-PTX wrappers for the Blackwell/Hopper async-copy machinery used by the
// image kernels: mbarrier (transaction barrier), cp.async.bulk.tensor (TMA tiled loads into shared
// memory, SASS: UTMALDG), proxy fence, and the host-side tensor-map encoder obtained through
// cudaGetDriverEntryPoint (libvrf.so links only cudart).
#pragma once
#include <cuda.h>            // CUtensorMap + enums only; no libcuda symbol is referenced
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrf {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async proxy (TMA completes on the barrier)
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// generic-proxy accesses of shared memory (LDS / STS) before -> async-proxy accesses (TMA writes) after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Bounded wait: a TMA that never completes (bad tensor map) traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();
}

// TMA tiled load of one box of a 3-D u8 tensor (x = column, y = row, z = sequence) into shared memory; out-of-bounds
// elements are filled with zeros.  Completion (box bytes) is signalled on `bar`.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// 16 x 8 -> 32 dot products (SASS IDP.2A / IDP.4A): signed 16-bit weights x unsigned pixel bytes.
//   dp2a_lo(w, p, c) = c + s16(w.lo) * u8(p.b0) + s16(w.hi) * u8(p.b1)
__device__ __forceinline__ int dp2a_lo_su(int w, unsigned p, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(p), "r"(c));
    return d;
}
//   dp4a_us(p, w, c) = c + sum_k u8(p.bk) * s8(w.bk)
__device__ __forceinline__ int dp4a_us(unsigned p, int w, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(p), "r"(w), "r"(c));
    return d;
}

// ---- host side ------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// u8 tensor [nz][ny][nx] with byte strides (pitch_y, pitch_z), box (bx, by, 1), no swizzle, zero OOB fill.
inline int tma_encode_u8_3d(CUtensorMap *out, void *base, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t pitch_y,
                            uint64_t pitch_z, uint32_t bx, uint32_t by)
{
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p ||
            qr != cudaDriverEntryPointSuccess)
            return -1;
        fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    const cuuint64_t dims[3] = {nx, ny, nz};
    const cuuint64_t strides[2] = {pitch_y, pitch_z};
    const cuuint32_t box[3] = {bx, by, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace vrf
