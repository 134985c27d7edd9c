// tma_probe.cu -- isolates the TMA tile load used by k_lk: one warp per CTA loads a 32x32 u8 box from a
// [nz][rows][cols] tensor at arbitrary (x, y) incl. out-of-bounds origins, checks the bytes, and times it.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcudart
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../vins-rgbd-fast_b200/csrc/tma.cuh"
using namespace vrf;
struct Maps { CUtensorMap m[2][4]; };

template <int V>
__global__ void k(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap single, const CUtensorMap *gmap, int which, int lvl,
                  const int *xy, int n, uint8_t *out, long long *cyc, int *err, const uint8_t *gsrc)
{
    extern __shared__ unsigned char raw[];
    unsigned char *sm = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    uint8_t *tile = sm + 128;
    const int lane = threadIdx.x;
    if (lane == 0) mbar_init(bar, 1);
    if (V & 8) fence_proxy_async(); else fence_mbar_init();
    __syncwarp();
    const CUtensorMap *mp = (V & 3) == 0 ? &maps.m[which][lvl] : (V & 3) == 1 ? &single : gmap;
    unsigned ph = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        fence_proxy_async();
        __syncwarp();
        long long t0 = clock64();
        if (lane == 0) {
            if (V & 4) {            // no TMA at all: plain arrive completes the phase
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
            } else if (V & 32) {    // 1-D bulk copy (no tensor map): 1024 bytes from a 16-byte aligned global address
                mbar_arrive_expect_tx(bar, 1024);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tile)), "l"(gsrc + (size_t)(i % 64) * 1024), "r"(1024), "r"(smem_u32(bar)) : "memory");
            } else if (V & 16) {    // rank-2 map, 2d instruction
                mbar_arrive_expect_tx(bar, 1024);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(tile)), "l"(gmap), "r"(xy[2 * i]), "r"(xy[2 * i + 1]), "r"(smem_u32(bar)) : "memory");
            } else {
                mbar_arrive_expect_tx(bar, 1024);
                tma_load_3d(tile, mp, bar, xy[2 * i], xy[2 * i + 1], 0);
            }
        }
        unsigned spin = 0;
        while (!mbar_try_wait(bar, ph)) { if (++spin > (1u << 20)) { if (lane == 0) atomicAdd(err, 1); return; } }
        ph ^= 1u;
        long long t1 = clock64();
        for (int r = 0; r < 32; ++r) out[(size_t)i * 1024 + r * 32 + lane] = tile[r * 32 + lane];
        if (lane == 0) cyc[i] = t1 - t0;
        __syncwarp();
    }
}

template <int V>
void launch(const Maps &maps, const CUtensorMap *gmap, int *d_xy, int n, uint8_t *d_out, long long *d_cyc, int *d_err, const uint8_t *gsrc)
{
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096);
    k<V><<<296, 32, 2048>>>(maps, maps.m[1][2], gmap, 1, 2, d_xy, n, d_out, d_cyc, d_err, gsrc);
}

int main(int argc, char **argv)
{
    const int V = argc > 1 ? atoi(argv[1]) : 0;
    { cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
      printf("device %s cc %d.%d driver %d runtime %d\n", p.name, p.major, p.minor, drv, rt); }
    const int cols = 640, rows = 480, pitch = 640, nz = 3;
    std::vector<uint8_t> img((size_t)nz * rows * pitch);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d_img; cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    Maps maps; memset(&maps, 0, sizeof(maps));
    for (int a = 0; a < 2; ++a) for (int l = 0; l < 4; ++l) {
        int rc = tma_encode_u8_3d(&maps.m[a][l], d_img, cols, rows, nz, pitch, (uint64_t)rows * pitch, 32, 32);
        if (rc) { printf("encode failed %d\n", rc); return 1; }
    }
    {
        const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&maps.m[1][2]);
        printf("descriptor words:");
        for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]);
        printf("\n");
    }
    const int n = 4096;
    std::vector<int> xy(2 * n);
    srand(1);
    for (int i = 0; i < n; ++i) { xy[2 * i] = rand() % (cols + 40) - 30; xy[2 * i + 1] = rand() % (rows + 40) - 30; }
    int *d_xy; cudaMalloc(&d_xy, xy.size() * 4); cudaMemcpy(d_xy, xy.data(), xy.size() * 4, cudaMemcpyHostToDevice);
    uint8_t *d_out; cudaMalloc(&d_out, (size_t)n * 1024);
    long long *d_cyc; cudaMalloc(&d_cyc, n * 8);
    int *d_err; cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &maps.m[1][2], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (V & 16) {       // rank-2 u8 map through the by-version entry point, L2 promotion none (the programming guide's example shape)
        void *p = nullptr; cudaDriverEntryPointQueryResult qr;
        cudaError_t ee = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qr);
        printf("by-version entry point: %s qr=%d p=%p\n", cudaGetErrorString(ee), (int)qr, p);
        PFN_tmapEncodeTiled fn = (PFN_tmapEncodeTiled)p;
        CUtensorMap m2; memset(&m2, 0, sizeof(m2));
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = fn(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank-2 encode: %d\n", (int)r);
        const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&m2);
        printf("rank-2 descriptor words:");
        for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]);
        printf("\n");
        cudaMemcpy(gmap, &m2, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    }
    printf("variant %d: map source %s, %s, init fence %s\n", V, (V & 3) == 0 ? "grid_constant struct array" : (V & 3) == 1 ? "grid_constant single" : "global memory",
           (V & 4) ? "NO TMA (mbarrier only)" : "TMA", (V & 8) ? "fence.proxy.async" : "fence.mbarrier_init");
    switch (V) {
    case 0: launch<0>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 1: launch<1>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 2: launch<2>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 4: launch<4>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 9: launch<9>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 12: launch<12>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 32: launch<32>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    case 18: launch<18>(maps, gmap, d_xy, n, d_out, d_cyc, d_err, d_img); break;
    default: printf("unknown variant\n"); return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    int err = 0; cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
    std::vector<uint8_t> out((size_t)n * 1024); cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    std::vector<long long> cyc(n); cudaMemcpy(cyc.data(), d_cyc, n * 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < n; ++i) for (int r = 0; r < 32; ++r) for (int c = 0; c < 32; ++c) {
        int x = xy[2 * i] + c, y = xy[2 * i + 1] + r;
        uint8_t want = (x >= 0 && x < cols && y >= 0 && y < rows) ? img[(size_t)y * pitch + x] : 0;
        if (V & 32) want = img[(size_t)(i % 64) * 1024 + r * 32 + c];
        if (out[(size_t)i * 1024 + r * 32 + c] != want) ++bad;
    }
    double s = 0; for (int i = 0; i < n; ++i) s += cyc[i];
    printf("timeouts %d  mismatching bytes %ld  mean TMA round trip %.0f cycles\n", err, bad, s / n);
    return 0;
}
