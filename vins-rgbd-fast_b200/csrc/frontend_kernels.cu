// frontend_kernels.cu -- hand-written sm_100a kernels for the per-frame front end
// (FeatureTracker::readImage, reference vins_estimator/src/feature_tracker/
// feature_tracker.cpp:263-439).  Compiled with -fmad=false so that float/double
// expressions round exactly like the reference's x86 (no-FMA) build.
//
// Kernel sequence for one batched readImage call (all sequences of the batch in
// each launch; no host round trip in between):
//   k_pyr<frame>    frame (GRAY8 | RGB8) -> pyramid levels 0 and 1 of the "forw" buffer: ingest fused with the first
//                   cv::pyrDown, the frame is read once (+ work-list prefix for k_lk)   [HBM-bound]
//   k_pyr<level>    cv::pyrDown level l -> l+1 for the deeper levels                    [HBM/L2-bound]
//                   (EQUALIZE: k_ingest -> k_clahe_lut -> k_clahe_apply -> k_pyr<level> from level 0)
//   k_lk            predictPtsInNextFrame + cv::calcOpticalFlowPyrLK, one warp / feature   (lk_kernels.cu)
//   k_post_a        status fix-up, inBorder, reduceVector x5, track_cnt++
//   k_ransac        rejectWithF (cv::findFundamentalMat RANSAC)   (ransac_kernels.cu)
//   k_post_b        reduceVector by inliers, setMask (std::sort + greedy circles),
//                   grid occupancy + cell selection
//   k_fast          gridDetect: FAST-9/16 + NMS + mask + top-K slots, one CTA / cell
//   k_finish        addPoints, undistortedPoints (+velocity), updateID, outputs
#include <stdlib.h>

#include "common.cuh"
#include "handle.h"
#include "introsort.h"

namespace vrf {

// exclusive prefix of the LK work items (n_pts of every sequence of the batch), by one CTA
__device__ void lk_work_prefix(const SeqCall *calls, int ncalls, const FrontDev &d, int *s_part)
{
    const int t = threadIdx.x;
    const int per = (ncalls + 255) / 256;
    const int b = t * per, e = min(ncalls, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += calls[i].first ? 0 : d.n_pts[calls[i].seq];
    s_part[t] = s;
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int i = 0; i < 256; ++i) { int v = s_part[i]; s_part[i] = acc; acc += v; }
        d.work_prefix[ncalls] = acc;
    }
    __syncthreads();
    int acc = s_part[t];
    for (int i = b; i < e; ++i) {
        d.work_prefix[i] = acc;
        acc += calls[i].first ? 0 : d.n_pts[calls[i].seq];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// k_ingest: copy / convert the input frame into pyramid level 0 (EQUALIZE path and single-level pyramids only; otherwise
// the ingest is fused into k_pyr below).
// GRAY8: 16 px per thread (uint4 load/store).  RGB8: 16 px per thread =
// 3 x uint4 coalesced loads, fixed-point cv::cvtColor RGB2GRAY
// ((R*9798 + G*19235 + B*3735 + 16384) >> 15), one uint4 store.
// Block (0,0) additionally builds the LK work-list prefix over the batch.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned gray4(unsigned r0, unsigned g0, unsigned b0)
{
    return (r0 * 9798u + g0 * 19235u + b0 * 3735u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256)
k_ingest(FrontCfg c, const SeqCall *calls, int ncalls, FrontDev d, const uint8_t *frames,
         size_t frame_bytes, int fmt)
{
    const int ci = blockIdx.y;
    if (blockIdx.x == 0 && ci == 0) {
        __shared__ int s_part[256];
        lk_work_prefix(calls, ncalls, d, s_part);
    }
    const SeqCall call = calls[ci];
    uint8_t *dst = d.pyr[call.buf_cur] + (size_t)call.seq * c.pyr_bytes;   // level 0 at offset 0
    const uint8_t *src = frames + (size_t)ci * frame_bytes;
    const int groups_per_row = c.cols >> 4;                 // cols % 16 == 0 enforced at create
    const int total = groups_per_row * c.rows;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        int y = g / groups_per_row, xg = g - y * groups_per_row;
        uint4 out;
        if (fmt == VRF_FMT_GRAY8) {
            out = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)y * c.cols) + xg);
        } else {
            const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t)y * c.cols * 3) + xg * 3;
            uint4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
            unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
            unsigned o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // 4 pixels = 12 bytes = words w[3q..3q+2]
                unsigned w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
                unsigned p0 = gray4(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
                unsigned p1 = gray4(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
                unsigned p2 = gray4((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
                unsigned p3 = gray4((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
                o[q] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
            }
            out = make_uint4(o[0], o[1], o[2], o[3]);
        }
        reinterpret_cast<uint4 *>(dst + (size_t)y * c.lp[0])[xg] = out;
    }
}

// ---------------------------------------------------------------------------
// EQUALIZE: cv::createCLAHE(3.0, Size(8, 8))->apply(img) (feature_tracker.cpp:269-275; OpenCV imgproc/clahe.cpp,
// restated bit-exactly in oracle/frontend_spec.py::clahe and pinned against cv2 there), incl. OpenCV's border
// extension for frame sizes that are not multiples of the 8x8 tile grid.
//   k_clahe_lut   one CTA per (tile, frame): 256-bin histogram in shared memory, clip at
//                 int(3.0 * tile_pixels / 256), uniform redistribution + the strided residual, prefix sum,
//                 lut[i] = cvRound(float(sum_i) * (255.f / tile_pixels)).
//   k_clahe_apply per pixel: bilinear blend of the four neighbouring tiles' LUT entries in OpenCV's float32
//                 operation order, in place on pyramid level 0 (reads and writes the same byte only).
// Both are pure image scans: WH bytes read (+ WH written by the second) per frame.
// ---------------------------------------------------------------------------
// tile size of the 8x8 grid.  Frames whose size is not a multiple of 8 are extended like OpenCV does:
// copyMakeBorder(src, ext, 0, 8 - rows % 8, 0, 8 - cols % 8, BORDER_REFLECT_101) -- note the full 8 extra columns (rows)
// when only the other dimension needs padding -- and the tiles are cut from the extended frame.
__device__ __forceinline__ void clahe_tiles(const FrontCfg &c, int &th, int &tw, bool &padded)
{
    padded = (c.rows & 7) || (c.cols & 7);
    const int eh = padded ? c.rows + (8 - (c.rows & 7)) : c.rows, ew = padded ? c.cols + (8 - (c.cols & 7)) : c.cols;
    th = eh >> 3; tw = ew >> 3;
}

__global__ void __launch_bounds__(256)
k_clahe_lut(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    __shared__ int s_hist[256];
    __shared__ int s_scan[256];
    __shared__ int s_warp[8];
    const int t = threadIdx.x;
    const SeqCall call = calls[blockIdx.y];
    const uint8_t *img = d.pyr[call.buf_cur] + (size_t)call.seq * c.pyr_bytes;
    int th, tw;
    bool padded;
    clahe_tiles(c, th, tw, padded);
    const int pitch = c.lp[0];
    const int ty = blockIdx.x >> 3, tx = blockIdx.x & 7;
    const uint8_t *tile = img + (size_t)(ty * th) * pitch + tx * tw;
    s_hist[t] = 0;
    __syncthreads();
    if (padded) {
        for (int i = t; i < th * tw; i += 256) {
            const int y = i / tw, x = i - y * tw;
            const int sy = reflect101(ty * th + y, c.rows), sx = reflect101(tx * tw + x, c.cols);
            atomicAdd(&s_hist[__ldg(img + (size_t)sy * pitch + sx)], 1);
        }
    } else if ((tw & 3) == 0) {
        const int wpr = tw >> 2;
        for (int i = t; i < th * wpr; i += 256) {
            const int y = i / wpr, xw = i - y * wpr;
            const unsigned v = __ldg(reinterpret_cast<const unsigned *>(tile + (size_t)y * pitch) + xw);
            atomicAdd(&s_hist[v & 255u], 1); atomicAdd(&s_hist[(v >> 8) & 255u], 1);
            atomicAdd(&s_hist[(v >> 16) & 255u], 1); atomicAdd(&s_hist[v >> 24], 1);
        }
    } else {
        for (int i = t; i < th * tw; i += 256) {
            const int y = i / tw, x = i - y * tw;
            atomicAdd(&s_hist[__ldg(tile + (size_t)y * pitch + x)], 1);
        }
    }
    __syncthreads();
    const int total = th * tw;
    int clip = (int)(3.0 * total / 256);
    if (clip < 1) clip = 1;
    int v = s_hist[t];
    int excess = v > clip ? v - clip : 0;
    v = v > clip ? clip : v;
    // clipped = sum of the excess over the 256 bins
    for (int o = 16; o > 0; o >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, o);
    if ((t & 31) == 0) s_warp[t >> 5] = excess;
    __syncthreads();
    int clipped = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) clipped += s_warp[i];
    const int batch = clipped / 256, resid = clipped - batch * 256;
    v += batch;
    if (resid != 0) {
        const int step = 256 / resid > 1 ? 256 / resid : 1;
        if (t % step == 0 && t / step < resid) ++v;
    }
    // inclusive prefix sum over the bins
    s_scan[t] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int add = t >= o ? s_scan[t - o] : 0;
        __syncthreads();
        s_scan[t] += add;
        __syncthreads();
    }
    const float lut_scale = 255.0f / (float)total;
    int q = __float2int_rn((float)s_scan[t] * lut_scale);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    d.clahe_lut[((size_t)call.seq * 64 + blockIdx.x) * 256 + t] = (uint8_t)q;
}

__global__ void __launch_bounds__(256)
k_clahe_apply(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    const SeqCall call = calls[blockIdx.y];
    uint8_t *img = d.pyr[call.buf_cur] + (size_t)call.seq * c.pyr_bytes;
    const uint8_t *lut = d.clahe_lut + (size_t)call.seq * 64 * 256;
    int th, tw;
    bool padded;
    clahe_tiles(c, th, tw, padded);
    const int pitch = c.lp[0];
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    const int wpr = c.cols >> 2;                            // 4 pixels per thread
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.rows * wpr; i += gridDim.x * blockDim.x) {
        const int y = i / wpr, x0 = (i - y * wpr) << 2;
        const float tyf = (float)y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf);
        const float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
        int ty2 = ty1 + 1;
        ty1 = ty1 < 0 ? 0 : ty1; ty2 = ty2 > 7 ? 7 : ty2;
        const uint8_t *l1 = lut + ty1 * 8 * 256, *l2 = lut + ty2 * 8 * 256;
        unsigned *p = reinterpret_cast<unsigned *>(img + (size_t)y * pitch + x0);
        const unsigned w = *p;
        unsigned out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sv = (w >> (8 * k)) & 255u;
            const float txf = (float)(x0 + k) * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf);
            const float xa = txf - (float)tx1, xa1 = 1.0f - xa;
            int tx2 = tx1 + 1;
            tx1 = tx1 < 0 ? 0 : tx1; tx2 = tx2 > 7 ? 7 : tx2;
            const float a = (float)l1[tx1 * 256 + sv], b = (float)l1[tx2 * 256 + sv];
            const float cc = (float)l2[tx1 * 256 + sv], dd = (float)l2[tx2 * 256 + sv];
            const float res = (a * xa1 + b * xa) * ya1 + (cc * xa1 + dd * xa) * ya;
            int q = __float2int_rn(res);
            q = q < 0 ? 0 : (q > 255 ? 255 : q);
            out |= (unsigned)q << (8 * k);
        }
        *p = out;
    }
}

// ---------------------------------------------------------------------------
// k_pyr: one pyramid step, cv::pyrDown (5x5 separable [1 4 6 4 1], (sum+128)>>8, REFLECT_101), fused with the frame
// ingest when the source is the incoming frame (cv::buildOpticalFlowPyramid's first level, feature_tracker.cpp:302-310):
//   SRC_RGB8 / SRC_GRAY8 : frame -> level 0 (stored) and level 1, the frame is read from HBM exactly once
//   SRC_PYR              : level l -> level l + 1 (the deeper levels; level 0 -> 1 on the EQUALIZE path, where CLAHE
//                          needs the whole level-0 image first)
// A CTA owns a full-width strip of R1 destination rows.  Three phases over shared memory:
//   (1) the 2 R1 + 3 source rows are loaded 16 px per thread (RGB8: 3 x uint4, fixed-point RGB2GRAY with IDP.2A:
//       two dot-product instructions per pixel on doubled 16-bit weights, the gray value lands in byte 2), written to
//       the level-0 image (own rows only) and parked in shared memory; the two reflected border columns are patched in;
//   (2) horizontal pass: 8 outputs per thread from one 16-byte shared-memory word + two halo words, two IDP.4A per
//       output (the five taps 1 4 6 4 1 split over the 4-byte word boundary), stored as packed u16 pairs.  The
//       rounding constant rides along: +8 per row sum = +128 after the vertical weights (sum 16);
//   (3) vertical pass on the packed pairs (<= 16 * (16 * 255 + 8) < 2^16), 8 output bytes per thread.
// All of it exact integer arithmetic => bit-identical to cv::cvtColor / cv::pyrDown (parity: tests/test_frontend_gpu.py,
// tests/test_golden.py).  ncu: 15 executed instructions per source pixel for ingest + first pyrDown (r1: 30 for the two kernels),
// 5.2 TB/s of DRAM traffic for the frame launch: HBM-bound.
// ---------------------------------------------------------------------------
enum { SRC_RGB8 = 0, SRC_GRAY8 = 1, SRC_PYR = 2 };

__device__ __forceinline__ unsigned dp2a_lo_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp2a_hi_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 4 RGB pixels (12 bytes, words w0 w1 w2) -> 4 gray bytes.  (R*9798 + G*19235 + B*3735 + 16384) >> 15 evaluated as
// (R*19596 + G*38470 + B*7470 + 32768) >> 16: same quotient, the doubled weights still fit 16 bits and the result is byte 2.
__device__ __forceinline__ unsigned rgb4_to_gray(unsigned w0, unsigned w1, unsigned w2)
{
    const unsigned WRG = 19596u | (38470u << 16), WB0 = 7470u, W0R = 19596u << 16, WGB = 38470u | (7470u << 16);
    const unsigned v0 = dp2a_hi_uu(WB0, w0, dp2a_lo_uu(WRG, w0, 32768u));   // R G B = w0.b0 b1 b2
    const unsigned v1 = dp2a_lo_uu(WGB, w1, dp2a_hi_uu(W0R, w0, 32768u));   // w0.b3 w1.b0 b1
    const unsigned v2 = dp2a_lo_uu(WB0, w2, dp2a_hi_uu(WRG, w1, 32768u));   // w1.b2 b3 w2.b0
    const unsigned v3 = dp2a_hi_uu(WGB, w2, dp2a_lo_uu(W0R, w2, 32768u));   // w2.b1 b2 b3
    return __byte_perm(__byte_perm(v0, v1, 0x0062), __byte_perm(v2, v3, 0x0062), 0x5410);
}

template <int SRC>
__global__ void __launch_bounds__(256)
k_pyr(FrontCfg c, const SeqCall *calls, int ncalls, FrontDev d, const uint8_t *frames, size_t frame_bytes, int level, int R1)
{
    extern __shared__ __align__(16) uint8_t s_pyr[];
    const int tid = threadIdx.x;
    const int ci = blockIdx.y;
    if (SRC != SRC_PYR && blockIdx.x == 0 && ci == 0) lk_work_prefix(calls, ncalls, d, reinterpret_cast<int *>(s_pyr));
    const SeqCall call = calls[ci];
    uint8_t *pyr = d.pyr[call.buf_cur] + (size_t)call.seq * c.pyr_bytes;
    const int sw = c.lw[level], sh = c.lh[level], sp = c.lp[level];
    const int dw = c.lw[level + 1], dh = c.lh[level + 1], dp = c.lp[level + 1];
    const uint8_t *src = SRC == SRC_PYR ? pyr + c.loff[level] : frames + (size_t)ci * frame_bytes;
    uint8_t *dst0 = pyr + c.loff[level];          // level-l image (written by the ingest variants)
    uint8_t *dst1 = pyr + c.loff[level + 1];
    const int gpr = sp >> 4;                      // 16-px groups per source row (pitches are multiples of 16)
    const int hpr = (dw + 7) >> 3;                // 8-output groups per destination row
    const int NR = 2 * R1 + 3;
    const int gpitch = sp + 32;                   // shared gray rows: 16 B margin | row | 16 B margin
    const int hpitch = hpr * 16;
    uint8_t *gs = s_pyr;
    uint8_t *hs = s_pyr + (size_t)NR * gpitch;
    const int y1_0 = blockIdx.x * R1;
    const int r0 = 2 * y1_0 - 2;                  // source row of shared row 0

    // ---- (1) load / convert / store level l, park in shared memory; two items per thread in flight ----
    constexpr int NLD = SRC == SRC_RGB8 ? 3 : 1;
    // items (row r, 16-px group g) are walked without divisions: 256 items further = dq rows and dm groups further
    const int dq1 = 256 / gpr, dm1 = 256 - dq1 * gpr;
    auto item_row = [&](int it, int r) -> int {               // source row of item `it` (>= -2), or -1000 when nothing needs it
        const int sy = r0 + r;
        return (it < NR * gpr && sy <= sh + 1) ? sy : -1000;  // rows beyond sh + 1: only destination rows >= dh would use them
    };
    auto item_load = [&](int sy, int g, uint4 (&v)[NLD]) {
        const int syr = reflect101(sy, sh);
        if (SRC == SRC_RGB8) {
            const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t)syr * sw * 3) + g * 3;
            v[0] = __ldg(p); v[NLD > 1 ? 1 : 0] = __ldg(p + 1); v[NLD > 2 ? 2 : 0] = __ldg(p + 2);
        } else if (SRC == SRC_GRAY8) v[0] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)syr * sw) + g);
        else v[0] = *(reinterpret_cast<const uint4 *>(src + (size_t)syr * sp) + g);
    };
    auto item_store = [&](int sy, int r, int g, const uint4 (&v)[NLD]) {
        uint4 out = v[0];
        if (SRC == SRC_RGB8) {
            const uint4 a = v[0], b = v[NLD > 1 ? 1 : 0], cc = v[NLD > 2 ? 2 : 0];
            out = make_uint4(rgb4_to_gray(a.x, a.y, a.z), rgb4_to_gray(a.w, b.x, b.y), rgb4_to_gray(b.z, b.w, cc.x),
                             rgb4_to_gray(cc.y, cc.z, cc.w));
        }
        *reinterpret_cast<uint4 *>(gs + (size_t)r * gpitch + 16 + 16 * g) = out;
        if (SRC != SRC_PYR && r >= 2 && r < 2 + 2 * R1 && sy < sh)
            *reinterpret_cast<uint4 *>(dst0 + (size_t)sy * sp + 16 * g) = out;
    };
    {
        int rA = tid / gpr, gA = tid - rA * gpr;
        for (int it = tid; it < NR * gpr; it += 512) {
            int rB = rA + dq1, gB = gA + dm1;
            if (gB >= gpr) { gB -= gpr; ++rB; }
            const int syA = item_row(it, rA), syB = item_row(it + 256, rB);
            uint4 vA[NLD], vB[NLD];
            if (syA >= -2) item_load(syA, gA, vA);
            if (syB >= -2) item_load(syB, gB, vB);
            if (syA >= -2) item_store(syA, rA, gA, vA);
            if (syB >= -2) item_store(syB, rB, gB, vB);
            rA = rB + dq1; gA = gB + dm1;
            if (gA >= gpr) { gA -= gpr; ++rA; }
        }
    }
    __syncthreads();
    for (int r = tid; r < NR; r += 256) {         // REFLECT_101 columns -2, -1, sw, sw + 1
        uint8_t *row = gs + (size_t)r * gpitch + 16;
        row[-1] = row[1]; row[-2] = row[2];
        row[sw] = row[sw - 2]; row[sw + 1] = row[sw - 3];
    }
    __syncthreads();

    // ---- (2) horizontal pass: h[x] = p[2x-2] + 4 p[2x-1] + 6 p[2x] + 4 p[2x+1] + p[2x+2] + 8, packed (h[2m], h[2m+1]) ----
    const int dq2 = 256 / hpr, dm2 = 256 - dq2 * hpr;
    int r = tid / hpr, q = tid - r * hpr;
    for (int it = tid; it < NR * hpr; it += 256) {
        const unsigned *rw = reinterpret_cast<const unsigned *>(gs + (size_t)r * gpitch + 16) + 4 * q;
        const uint4 w = *reinterpret_cast<const uint4 *>(rw);
        const unsigned wv[6] = {rw[-1], w.x, w.y, w.z, w.w, rw[4]};
        unsigned o[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            // outputs 2m (taps: bytes 2,3 of word m-1, bytes 0..2 of word m) and 2m+1 (word m, byte 0 of word m+1)
            const unsigned lo = __dp4a(wv[m], 0x04010000u, __dp4a(wv[m + 1], 0x00010406u, 8u));
            const unsigned hi = __dp4a(wv[m + 1], 0x04060401u, __dp4a(wv[m + 2], 0x00000001u, 8u));
            o[m] = __byte_perm(lo, hi, 0x5410);
        }
        *reinterpret_cast<uint4 *>(hs + (size_t)r * hpitch + 16 * q) = make_uint4(o[0], o[1], o[2], o[3]);
        r += dq2; q += dm2;
        if (q >= hpr) { q -= hpr; ++r; }
    }
    __syncthreads();

    // ---- (3) vertical pass ----
    int y = tid / hpr;
    q = tid - y * hpr;
    for (int it = tid; it < R1 * hpr; it += 256, y += dq2, q += dm2) {
        if (q >= hpr) { q -= hpr; ++y; }
        const int dy = y1_0 + y, dx0 = 8 * q;
        if (dy >= dh) break;
        const uint8_t *hp = hs + (size_t)(2 * y) * hpitch + 16 * q;
        const uint4 a0 = *reinterpret_cast<const uint4 *>(hp), a1 = *reinterpret_cast<const uint4 *>(hp + hpitch),
                    a2 = *reinterpret_cast<const uint4 *>(hp + 2 * hpitch), a3 = *reinterpret_cast<const uint4 *>(hp + 3 * hpitch),
                    a4 = *reinterpret_cast<const uint4 *>(hp + 4 * hpitch);
        const unsigned msk = 0x00FF00FFu;
        const unsigned v0 = ((a0.x + a4.x + 4u * (a1.x + a3.x) + 6u * a2.x) >> 8) & msk;
        const unsigned v1 = ((a0.y + a4.y + 4u * (a1.y + a3.y) + 6u * a2.y) >> 8) & msk;
        const unsigned v2 = ((a0.z + a4.z + 4u * (a1.z + a3.z) + 6u * a2.z) >> 8) & msk;
        const unsigned v3 = ((a0.w + a4.w + 4u * (a1.w + a3.w) + 6u * a2.w) >> 8) & msk;
        const unsigned lo = __byte_perm(v0, v1, 0x6420), hi = __byte_perm(v2, v3, 0x6420);
        uint8_t *qd = dst1 + (size_t)dy * dp + dx0;
        if (dx0 + 8 <= dw) *reinterpret_cast<uint2 *>(qd) = make_uint2(lo, hi);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (dx0 + k < dw) qd[k] = (uint8_t)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xFFu);
        }
    }
}

// ---------------------------------------------------------------------------
// Block-wide exclusive scan helper (blockDim.x == 256), returns total in *tot.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int block_excl_scan_256(int v, int *s_warp, int *tot)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = (lane < 8) ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += n;
        }
        if (lane < 8) s_warp[lane] = xi - x;
        if (lane == 7) s_warp[8] = xi;
    }
    __syncthreads();
    int res = inc - v + s_warp[w];
    *tot = s_warp[8];
    __syncthreads();
    return res;
}

__device__ __forceinline__ bool in_border(const FrontCfg &c, float2 p)
{
    // FeatureTracker::inBorder (feature_tracker.cpp:96-103): cvRound = round half to even
    int x = __float2int_rn(p.x), y = __float2int_rn(p.y);
    return 1 <= x && x < c.cols - 1 && 1 <= y && y < c.rows - 1;
}

// ---------------------------------------------------------------------------
// k_post_a: feature_tracker.cpp:313-349 -- status fix-up with inBorder,
// unstable_pts, reduceVector of the five arrays (order preserving), track_cnt++.
// One CTA (256 threads) per batch item; 4 consecutive elements per thread.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_post_a(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    __shared__ int s_warp[9];
    const SeqCall call = calls[blockIdx.x];
    const int seq = call.seq;
    const size_t base = (size_t)seq * VRF_CAP;
    const int n = call.first ? 0 : d.n_pts[seq];
    const int t4 = threadIdx.x * 4;
    int keep[4], uns[4];
    int nk = 0, nu = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int i = t4 + e;
        keep[e] = 0; uns[e] = 0;
        if (i < n) {
            int st = d.lk_status[base + i];
            float2 f = d.lk_pts[base + i];
            bool ib = in_border(c, f);
            if (!st && ib) uns[e] = 1;
            else if (st && !ib) st = 0;
            keep[e] = st;
        }
        nk += keep[e]; nu += uns[e];
    }
    int totk, totu;
    int ok = block_excl_scan_256(nk, s_warp, &totk);
    int ou = block_excl_scan_256(nu, s_warp, &totu);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int i = t4 + e;
        if (keep[e]) {
            d.t_prev[base + ok] = d.cur_pts[base + i];
            d.t_forw[base + ok] = d.lk_pts[base + i];
            d.t_ids[base + ok] = d.ids[base + i];
            d.t_cnt[base + ok] = d.cnt[base + i] + 1;      // for (auto &n : track_cnt) n++;
            d.t_prevun[base + ok] = d.prev_un[base + i];
            d.t_keep[base + ok] = 1;
            ok++;
        }
        if (uns[e]) { d.unstable[base + ou] = d.lk_pts[base + i]; ou++; }
    }
    if (threadIdx.x == 0) {
        d.t_n[seq] = totk;
        d.n_unstable[seq] = totu;
        d.n_lk[seq] = n;
    }
}

// ---------------------------------------------------------------------------
// k_post_b: reduceVector by the RANSAC inlier mask (feature_tracker.cpp:463-468),
// setMask (:173-208: std::sort by track_cnt desc, greedy accept against the
// virtual circle mask, circles for unstable_pts), grid occupancy and cell
// selection (:361-395).  One CTA per batch item.
// The H x W mask image of the reference is *virtual* here: cv::circle(mask,p,r,0,-1)
// zeroes exactly {q : |q-p|^2 <= r^2} (pinned in tests), so "mask.at(q)==255"
// <=> no stored centre within r of q.  Centres are kept in d.maskpts.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_post_b(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    extern __shared__ unsigned char smem_raw[];
    float2 *s_forw = reinterpret_cast<float2 *>(smem_raw);                 // CAP
    float2 *s_prev = s_forw + VRF_CAP;                                      // CAP
    float2 *s_pun = s_prev + VRF_CAP;                                       // CAP
    int *s_ids = reinterpret_cast<int *>(s_pun + VRF_CAP);                  // CAP
    int *s_cnt = s_ids + VRF_CAP;                                           // CAP
    SortItem *s_sort = reinterpret_cast<SortItem *>(s_cnt + VRF_CAP);       // CAP
    int2 *s_acc = reinterpret_cast<int2 *>(s_sort + VRF_CAP);               // CAP accepted centres
    int *s_order = reinterpret_cast<int *>(s_acc + VRF_CAP);                // CAP accepted source index
    __shared__ int s_warp[9];
    __shared__ int s_n, s_nacc;

    const SeqCall call = calls[blockIdx.x];
    const int seq = call.seq;
    const size_t base = (size_t)seq * VRF_CAP;
    const int n0 = d.t_n[seq];
    // 1. compaction by RANSAC mask
    {
        const int t4 = threadIdx.x * 4;
        int keep[4], nk = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int i = t4 + e;
            keep[e] = (i < n0) ? (int)d.t_keep[base + i] : 0;
            nk += keep[e];
        }
        int tot;
        int o = block_excl_scan_256(nk, s_warp, &tot);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int i = t4 + e;
            if (keep[e]) {
                s_forw[o] = d.t_forw[base + i];
                s_prev[o] = d.t_prev[base + i];
                s_pun[o] = d.t_prevun[base + i];
                s_ids[o] = d.t_ids[base + i];
                s_cnt[o] = d.t_cnt[base + i];
                o++;
            }
        }
        if (threadIdx.x == 0) s_n = tot;
    }
    __syncthreads();
    int n = s_n;
    const int nuns = d.n_unstable[seq];
    if (call.pub) {
        // 2. setMask: std::sort (libstdc++ introsort restatement, single thread)
        for (int i = threadIdx.x; i < n; i += 256) { s_sort[i].key = s_cnt[i]; s_sort[i].val = i; }
        __syncthreads();
        if (threadIdx.x == 0) std_sort_desc(s_sort, n);
        __syncthreads();
        // greedy accept in sorted order (warp 0; lanes test the accepted list in parallel)
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int r2 = c.min_dist * c.min_dist;
            int nacc = 0;
            for (int k = 0; k < n; ++k) {
                int src = s_sort[k].val;
                float2 p = s_forw[src];
                int px = __float2int_rn(p.x), py = __float2int_rn(p.y);
                // FISHEYE: the mask starts as fisheye_mask.clone() (feature_tracker.cpp:175-178); the test is == 255
                bool hit = d.fisheye && d.fisheye[(size_t)py * c.cols + px] != 255;
                for (int l = lane; l < nacc; l += 32) {
                    int dx = s_acc[l].x - px, dy = s_acc[l].y - py;
                    hit |= (dx * dx + dy * dy <= r2);
                }
                if (!__any_sync(0xffffffffu, hit)) {
                    if (lane == 0) { s_acc[nacc] = make_int2(px, py); s_order[nacc] = src; }
                    nacc++;
                    __syncwarp();
                }
            }
            if (lane == 0) s_nacc = nacc;
        }
        __syncthreads();
        const int nacc = s_nacc;
        // write back in accepted (sorted) order + the mask centres (accepted + unstable)
        for (int i = threadIdx.x; i < nacc; i += 256) {
            int src = s_order[i];
            d.t_forw[base + i] = s_forw[src];
            d.t_prev[base + i] = s_prev[src];
            d.t_prevun[base + i] = s_pun[src];
            d.t_ids[base + i] = s_ids[src];
            d.t_cnt[base + i] = s_cnt[src];
            d.maskpts[(size_t)seq * 2 * VRF_CAP + i] = s_acc[i];
        }
        for (int i = threadIdx.x; i < nuns; i += 256) {
            float2 u = d.unstable[base + i];
            d.maskpts[(size_t)seq * 2 * VRF_CAP + nacc + i] = make_int2(__float2int_rn(u.x), __float2int_rn(u.y));
        }
        n = nacc;
        // 3. grid occupancy + cell selection (only when n_max_cnt > 0)
        int *gcnt = d.grid_cnt + (size_t)seq * VRF_MAX_CELLS;
        uint8_t *tex = d.tex_status + (size_t)seq * VRF_MAX_CELLS;
        int *ck = d.cell_k + (size_t)seq * VRF_MAX_CELLS;
        const bool detect = (c.max_cnt - n) > 0;
        if (detect) {
            for (int i = threadIdx.x; i < c.ncells; i += 256) gcnt[i] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < nacc; i += 256) {
                float2 p = s_forw[s_order[i]];
                int col = (int)p.x / c.gw, row = (int)p.y / c.gh;
                if (col == c.gcols) --col;
                if (row == c.grows) --row;
                atomicAdd(&gcnt[col + c.gcols * row], 1);
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < c.ncells; i += 256) {
            int k = 0;
            if (detect) {
                if (gcnt[i] < c.thr && tex[i]) k = c.thr - gcnt[i] + 2;
                else tex[i] = 1;
            }
            ck[i] = k;
            d.ncand[(size_t)seq * VRF_MAX_CELLS + i] = 0;
        }
        if (threadIdx.x == 0) { d.t_n[seq] = n; d.n_maskpts[seq] = nacc + nuns; }
    } else {
        for (int i = threadIdx.x; i < n; i += 256) {
            d.t_forw[base + i] = s_forw[i];
            d.t_prev[base + i] = s_prev[i];
            d.t_prevun[base + i] = s_pun[i];
            d.t_ids[base + i] = s_ids[i];
            d.t_cnt[base + i] = s_cnt[i];
        }
        for (int i = threadIdx.x; i < c.ncells; i += 256) d.cell_k[(size_t)seq * VRF_MAX_CELLS + i] = 0;
        if (threadIdx.x == 0) { d.t_n[seq] = n; d.n_maskpts[seq] = 0; }
    }
}

// ---------------------------------------------------------------------------
// k_fast: FeatureTracker::gridDetect (feature_tracker.cpp:105-171) for one grid
// cell: cv::FastFeatureDetector (threshold 10, NMS, TYPE_9_16) on forw_img(rect)
// with the snapshot mask, then the reference's top-K slot selection.
// One CTA (256 threads) per (cell, batch item).  Dynamic shared memory:
//   u8 img[h][w] | u8 score[h][w] | u32 kp[(w*h)/4+64] | int2 mpts[...]
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool has_run9(unsigned m)
{
    // 16-bit cyclic mask: any 9 contiguous set bits?
    unsigned x = m | (m << 16);
    x &= x >> 1;  // runs of 2
    x &= x >> 2;  // 4
    x &= x >> 4;  // 8
    x &= (m | (m << 16)) >> 8;  // 9
    return (x & 0xFFFFu) != 0;
}

__global__ void __launch_bounds__(256)
k_fast(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ int s_warp[9];
    __shared__ int s_nm, s_nkp;
    const SeqCall call = calls[blockIdx.y];
    const int seq = call.seq, cell = blockIdx.x;
    if (!call.pub) return;
    const int K = d.cell_k[(size_t)seq * VRF_MAX_CELLS + cell];
    if (K == 0) return;
    // rect (feature_tracker.cpp:44-86): +3 px overlap on interior edges
    const int ci = cell / c.gcols, cj = cell - ci * c.gcols;
    const int grw = c.cols - (c.gcols - 1) * c.gw, grh = c.rows - (c.grows - 1) * c.gh;
    int rx, ry, rw, rh;
    if (cj == 0) { rx = 0; rw = c.gw + 3; }
    else if (cj < c.gcols - 1) { rx = cj * c.gw - 3; rw = c.gw + 6; }
    else { rx = cj * c.gw - 3; rw = grw + 3; }
    if (ci == 0) { ry = 0; rh = c.gh + 3; }
    else if (ci < c.grows - 1) { ry = ci * c.gh - 3; rh = c.gh + 6; }
    else { ry = ci * c.gh - 3; rh = grh + 3; }
    const int npx = rw * rh;
    uint8_t *s_img = smem_raw;
    uint8_t *s_sc = s_img + ((npx + 15) & ~15);
    unsigned *s_kp = reinterpret_cast<unsigned *>(s_sc + ((npx + 15) & ~15));
    const int kp_cap = (npx / 4 + 64) & ~1;     // even: keeps s_m 8-byte aligned
    int2 *s_m = reinterpret_cast<int2 *>(s_kp + kp_cap);

    const uint8_t *img = d.pyr[call.buf_cur] + (size_t)seq * c.pyr_bytes;   // level 0
    for (int i = threadIdx.x; i < npx; i += 256) {
        int y = i / rw, x = i - y * rw;
        s_img[i] = __ldg(img + (size_t)(ry + y) * c.lp[0] + rx + x);
        s_sc[i] = 0;
    }
    if (threadIdx.x == 0) { s_nm = 0; s_nkp = 0; }
    __syncthreads();
    // mask centres whose circle can reach this rect
    {
        const int nm = d.n_maskpts[seq];
        const int2 *mp = d.maskpts + (size_t)seq * 2 * VRF_CAP;
        const int r = c.min_dist;
        for (int i = threadIdx.x; i < nm; i += 256) {
            int2 p = mp[i];
            if (p.x + r >= rx && p.x - r < rx + rw && p.y + r >= ry && p.y - r < ry + rh) {
                int o = atomicAdd(&s_nm, 1);
                s_m[o] = p;
            }
        }
    }
    // scores on the rect interior
    const int iw = rw - 6, ih = rh - 6;
    const int off[16] = {3 * rw, 3 * rw + 1, 2 * rw + 2, rw + 3, 3, -rw + 3, -2 * rw + 2, -3 * rw + 1,
                         -3 * rw, -3 * rw - 1, -2 * rw - 2, -rw - 3, -3, rw - 3, 2 * rw - 2, 3 * rw - 1};
    for (int i = threadIdx.x; i < iw * ih; i += 256) {
        int y = i / iw + 3, x = i - (y - 3) * iw + 3;
        const uint8_t *p = s_img + y * rw + x;
        int v = p[0];
        int dd[16];
        unsigned mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            dd[k] = v - (int)p[off[k]];
            mb |= (dd[k] >= 11 ? 1u : 0u) << k;     // centre brighter than ring by > threshold
            md |= (dd[k] <= -11 ? 1u : 0u) << k;
        }
        if (has_run9(mb) || has_run9(md)) {
            // exact cornerScore: max over the 16 arcs of 9 of max(min d, -max d), minus 1
            // NOTE (toolchain): ptxas 12.9 for sm_100a folds `max(max(mn, -mx), best)` into a
            // VIMNMX3 and silently drops the negation (seen in SASS, confirmed on a B200:
            // scores came out as max|d|-1).  The bright / dark arcs are therefore tracked in
            // two separate chains and combined once at the end without an integer negate
            // feeding a min/max: -max(d) == min(255 - d) - 255.
            int best_b = -100000;      // max over arcs of min(d)
            int best_e = -100000;      // max over arcs of min(255 - d)
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                int mn = dd[s], me = 255 - dd[s];
#pragma unroll
                for (int j = 1; j < 9; ++j) {
                    int e = dd[(s + j) & 15];
                    mn = min(mn, e); me = min(me, 255 - e);
                }
                best_b = max(best_b, mn);
                best_e = max(best_e, me);
            }
            int best_d = best_e - 255;
            int best = best_b > best_d ? best_b : best_d;
            int sc = best - 1;
            s_sc[y * rw + x] = (uint8_t)(sc >= 10 ? sc : 0);
        }
    }
    __syncthreads();
    // NMS (strict maximum over the 8 neighbours) + mask, ordered (row-major) compaction
    const int r2 = c.min_dist * c.min_dist;
    const int nmask = s_nm;
    const int nchunks = (iw * ih + 255) / 256;
    int basecount = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        int i = ch * 256 + threadIdx.x;
        int flag = 0, x = 0, y = 0, sc = 0;
        if (i < iw * ih) {
            y = i / iw + 3; x = i - (y - 3) * iw + 3;
            const uint8_t *q = s_sc + y * rw + x;
            sc = q[0];
            if (sc > 0 && sc > q[-1] && sc > q[1] && sc > q[-rw - 1] && sc > q[-rw] && sc > q[-rw + 1] &&
                sc > q[rw - 1] && sc > q[rw] && sc > q[rw + 1]) {
                flag = 1;
                int gx = rx + x, gy = ry + y;
                if (d.fisheye && d.fisheye[(size_t)gy * c.cols + gx] == 0) flag = 0;      // detect(img, kps, mask): mask != 0
                for (int m = 0; m < nmask; ++m) {
                    int dx = s_m[m].x - gx, dy = s_m[m].y - gy;
                    if (dx * dx + dy * dy <= r2) { flag = 0; break; }
                }
            }
        }
        int tot;
        int o = block_excl_scan_256(flag, s_warp, &tot);
        if (flag && basecount + o < kp_cap) s_kp[basecount + o] = (unsigned)x | ((unsigned)y << 12) | ((unsigned)sc << 24);
        basecount += tot;
    }
    __syncthreads();
    // top-K slot selection (feature_tracker.cpp:119-167), sequential exactly as the reference
    if (threadIdx.x == 0) {
        const int nk = min(basecount, kp_cap);
        float *cand = d.cand + ((size_t)seq * VRF_MAX_CELLS + cell) * c.kmax * 3;
        int nout = 0;
        if (nk == 0) {
            d.tex_status[(size_t)seq * VRF_MAX_CELLS + cell] = 0;
        } else if (nk <= K) {
            for (int j = 0; j < nk; ++j) {
                unsigned e = s_kp[j];
                cand[3 * j] = (float)((int)(e & 0xFFF) + rx);
                cand[3 * j + 1] = (float)((int)((e >> 12) & 0xFFF) + ry);
                cand[3 * j + 2] = (float)(e >> 24);
            }
            nout = nk;
        } else {
            int minid = 0;
            for (int j = 0; j < nk; ++j) {
                unsigned e = s_kp[j];
                float resp = (float)(e >> 24);
                if (j < K) {
                    cand[3 * j] = (float)((int)(e & 0xFFF) + rx);
                    cand[3 * j + 1] = (float)((int)((e >> 12) & 0xFFF) + ry);
                    cand[3 * j + 2] = resp;
                    if (resp < cand[3 * minid + 2]) minid = j;
                } else if (resp > cand[3 * minid + 2]) {
                    cand[3 * minid] = (float)((int)(e & 0xFFF) + rx);
                    cand[3 * minid + 1] = (float)((int)((e >> 12) & 0xFFF) + ry);
                    cand[3 * minid + 2] = resp;
                    for (int k = 0; k < K; ++k)
                        if (cand[3 * k + 2] < cand[3 * minid + 2]) minid = k;
                }
            }
            nout = K;
        }
        d.ncand[(size_t)seq * VRF_MAX_CELLS + cell] = nout;
    }
}

// ---------------------------------------------------------------------------
// k_finish: addPoints(vector<KeyPoint>&) in cell order (feature_tracker.cpp:405-409,
// :220-233), cur <- forw, undistortedPoints (:542-593) incl. velocity, the
// nodelet's updateID loop (estimator_nodelet.cpp:324-330), outputs.
// One CTA (256 threads) per batch item.
// ---------------------------------------------------------------------------
// Depth: the nodelet's depth decode (estimator_nodelet.cpp:512-533) and the lookup
// depth_img.at<unsigned short>((int)v, (int)u) + DEPTH_MIN_DIST test of FeatureManager::addFeatureCheckParallax
// (feature_manager.cpp:71-80) are fused here: the 32FC1 -> 16UC1 convertTo(…, 1000) is elementwise, so only the
// looked-up pixels are converted (float multiply, round-half-even, x86 out-of-range -> INT_MIN, saturate).
__device__ __forceinline__ unsigned depth_mm_at(const uint8_t *plane, int fmt, int cols, int x, int y)
{
    if (fmt == VRF_DEPTH_16UC1) return __ldg(reinterpret_cast<const unsigned short *>(plane) + (size_t)y * cols + x);
    const float t = __ldg(reinterpret_cast<const float *>(plane) + (size_t)y * cols + x) * 1000.0f;
    const int r = (t >= -2147483648.0f && t < 2147483648.0f) ? __float2int_rn(t) : (int)0x80000000;
    return (unsigned)min(max(r, 0), 65535);
}

__global__ void __launch_bounds__(256)
k_finish(FrontCfg c, const SeqCall *calls, FrontDev d, const uint8_t *depth, size_t depth_frame_bytes, int depth_fmt)
{
    __shared__ int2 s_acc[VRF_CAP];      // newly accepted centres
    __shared__ float2 s_new[VRF_CAP];
    __shared__ int s_warp[9];
    __shared__ int s_nnew;
    const SeqCall call = calls[blockIdx.x];
    const int seq = call.seq;
    const size_t base = (size_t)seq * VRF_CAP;
    const int n = d.t_n[seq];
    const uint8_t *dplane = (depth && depth_fmt != VRF_DEPTH_NONE && call.dslot >= 0 && call.pub)
                                ? depth + (size_t)call.dslot * depth_frame_bytes : nullptr;
    if (threadIdx.x == 0) s_nnew = 0;
    __syncthreads();
    if (call.pub && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int r2 = c.min_dist * c.min_dist;
        const int2 *mp = d.maskpts + (size_t)seq * 2 * VRF_CAP;
        const int nm = d.n_maskpts[seq];
        int nnew = 0;
        for (int cell = 0; cell < c.ncells; ++cell) {
            if (d.cell_k[(size_t)seq * VRF_MAX_CELLS + cell] == 0) continue;
            const int nc = d.ncand[(size_t)seq * VRF_MAX_CELLS + cell];
            const float *cand = d.cand + ((size_t)seq * VRF_MAX_CELLS + cell) * c.kmax * 3;
            for (int j = 0; j < nc; ++j) {
                float px = cand[3 * j], py = cand[3 * j + 1];
                int ix = __float2int_rn(px), iy = __float2int_rn(py);
                bool hit = d.fisheye && d.fisheye[(size_t)iy * c.cols + ix] != 255;
                for (int l = lane; l < nm; l += 32) {
                    int dx = mp[l].x - ix, dy = mp[l].y - iy;
                    hit |= (dx * dx + dy * dy <= r2);
                }
                for (int l = lane; l < nnew; l += 32) {
                    int dx = s_acc[l].x - ix, dy = s_acc[l].y - iy;
                    hit |= (dx * dx + dy * dy <= r2);
                }
                if (!__any_sync(0xffffffffu, hit)) {
                    if (lane == 0 && n + nnew < VRF_CAP) { s_acc[nnew] = make_int2(ix, iy); s_new[nnew] = make_float2(px, py); }
                    if (n + nnew < VRF_CAP) nnew++;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) s_nnew = nnew;
    }
    __syncthreads();
    const int nnew = s_nnew;
    const int ntot = n + nnew;
    // per point: final arrays, undistort, velocity, id assignment
    int n_id0 = d.n_id[seq];
    const int t4 = threadIdx.x * 4;
    int isnew[4], cntn = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int i = t4 + e;
        isnew[e] = (i < ntot) && ((i >= n) || d.t_ids[base + i] == -1);
        cntn += isnew[e];
    }
    int totnew;
    int o = block_excl_scan_256(cntn, s_warp, &totnew);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int i = t4 + e;
        if (i < ntot) {
            float2 p; int id, cnt; float2 pun;
            if (i < n) { p = d.t_forw[base + i]; id = d.t_ids[base + i]; cnt = d.t_cnt[base + i]; pun = d.t_prevun[base + i]; }
            else { p = s_new[i - n]; id = -1; cnt = 1; pun = make_float2(0.f, 0.f); }
            double mx, my;
            cam_lift(c, (double)p.x, (double)p.y, mx, my);
            float2 un = make_float2((float)mx, (float)my);   // b.x()/b.z(), z == 1
            // velocity: only points that already carried an id when the previous
            // undistortedPoints ran are found in prev_un_pts_map: cnt >= 3
            float2 v = make_float2(0.f, 0.f);
            if (!call.first && id != -1 && cnt >= 3) {
                v.x = (float)((double)(un.x - pun.x) / call.dt);
                v.y = (float)((double)(un.y - pun.y) / call.dt);
            }
            if (isnew[e]) { id = n_id0 + o; o++; }
            d.cur_pts[base + i] = p;
            d.ids[base + i] = id;
            d.cnt[base + i] = cnt;
            d.prev_un[base + i] = un;
            if (i >= d.out_pitch) continue;        // reported through the status word below
            const size_t ob = (size_t)blockIdx.x * d.out_pitch + i;
            d.o_pts[ob] = p;
            d.o_un[ob] = un;
            d.o_vel[ob] = v;
            d.o_ids[ob] = id;
            d.o_cnt[ob] = cnt;
            unsigned mm = 0;
            if (dplane) mm = depth_mm_at(dplane, depth_fmt, c.cols, min(max((int)p.x, 0), c.cols - 1), min(max((int)p.y, 0), c.rows - 1));
            const double dm = (double)mm / 1000.0;
            d.o_depth[ob] = (unsigned short)mm;
            d.o_dkeep[ob] = (0.0 < dm && dm < c.depth_min_dist) ? 0 : 1;
        }
    }
    if (threadIdx.x == 0) {
        d.n_pts[seq] = ntot;
        d.n_id[seq] = n_id0 + totnew;
        int *hdr = d.out_hdr + (size_t)blockIdx.x * 8;
        hdr[0] = ntot;
        hdr[1] = n_id0 + totnew;
        hdr[2] = d.n_lk[seq];
        hdr[3] = d.n_unstable[seq];
        hdr[4] = (n + nnew >= VRF_CAP || ntot > d.out_pitch) ? VRF_ERR_CAPACITY : 0;
    }
}

// ---------------------------------------------------------------------------
// launchers (called from api.cu)
// ---------------------------------------------------------------------------
size_t post_b_smem_bytes()
{
    return (size_t)VRF_CAP * (3 * sizeof(float2) + 2 * sizeof(int) + sizeof(SortItem) + sizeof(int2) + sizeof(int));
}

size_t fast_smem_bytes(const FrontCfg &c)
{
    int grw = c.cols - (c.gcols - 1) * c.gw, grh = c.rows - (c.grows - 1) * c.gh;
    int rw = max(c.gw + 6, grw + 3), rh = max(c.gh + 6, grh + 3);
    size_t npx = (size_t)rw * rh;
    size_t a = (npx + 15) & ~(size_t)15;
    return 2 * a + (npx / 4 + 64) * sizeof(unsigned) + (size_t)2 * VRF_CAP * sizeof(int2);
}

// k_pyr strip height (destination rows per CTA) and dynamic shared memory: 2 R1 + 3 gray rows + as many packed row-sum rows
static int pyr_strip_rows(const FrontCfg &c, int level)
{
    int r1 = 16;
    if (const char *e = getenv("VRF_PYR_R1")) { const int v = atoi(e); if (v >= 1 && v <= 64) r1 = v; }
    while (r1 > 1 && (size_t)(2 * r1 + 3) * (c.lp[level] + 32 + ((c.lw[level + 1] + 7) >> 3) * 16) > 96 * 1024) r1 >>= 1;
    return r1;
}
static size_t pyr_smem_bytes(const FrontCfg &c, int level, int r1)
{
    const size_t b = (size_t)(2 * r1 + 3) * (c.lp[level] + 32 + ((c.lw[level + 1] + 7) >> 3) * 16);
    return b < 1024 ? 1024 : b;        // the LK work-list prefix borrows 1 KB
}

int front_configure_kernels(const FrontCfg &c)
{
    cudaError_t e = cudaFuncSetAttribute(k_post_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_b_smem_bytes());
    if (e != cudaSuccess) return (int)e;
    if (c.levels > 1) {
        const int smem0 = (int)pyr_smem_bytes(c, 0, pyr_strip_rows(c, 0));
        if ((e = cudaFuncSetAttribute(k_pyr<SRC_RGB8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(k_pyr<SRC_GRAY8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(k_pyr<SRC_PYR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0)) != cudaSuccess) return (int)e;
    }
    e = cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(c));
    return (int)e;
}

int front_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, const LkMaps &maps,
                 const uint8_t *d_frames, size_t frame_bytes, int fmt, int any_pub, int sm_count, LaunchCtx &lc)
{
    cudaStream_t st = lc.st;
    // Level 0 (+ level 1): the frame is read once by the fused ingest + pyrDown kernel.  With EQUALIZE the whole level-0
    // image has to exist before CLAHE, so the frame is ingested alone and level 1 is built from the equalised image.
    int l_first = 0;
    if (c.equalize || c.levels < 2) {
        dim3 gi((c.rows * (c.cols >> 4) + 255) / 256, ncalls);
        if (gi.x > 64) gi.x = 64;
        lc.begin(K_INGEST);
        k_ingest<<<gi, 256, 0, st>>>(c, d_calls, ncalls, d, d_frames, frame_bytes, fmt);
        lc.end();
    } else {
        const int r1 = pyr_strip_rows(c, 0);
        dim3 g((c.lh[1] + r1 - 1) / r1, ncalls);
        lc.begin(K_PYR);
        if (fmt == VRF_FMT_RGB8) k_pyr<SRC_RGB8><<<g, 256, pyr_smem_bytes(c, 0, r1), st>>>(c, d_calls, ncalls, d, d_frames, frame_bytes, 0, r1);
        else k_pyr<SRC_GRAY8><<<g, 256, pyr_smem_bytes(c, 0, r1), st>>>(c, d_calls, ncalls, d, d_frames, frame_bytes, 0, r1);
        lc.end();
        l_first = 1;
    }
    if (c.equalize) {
        lc.begin(K_CLAHE_LUT);
        k_clahe_lut<<<dim3(64, ncalls), 256, 0, st>>>(c, d_calls, d);
        lc.end();
        dim3 ga((c.rows * (c.cols >> 2) + 255) / 256, ncalls);
        if (ga.x > 64) ga.x = 64;
        lc.begin(K_CLAHE_APPLY);
        k_clahe_apply<<<ga, 256, 0, st>>>(c, d_calls, d);
        lc.end();
    }
    for (int l = l_first; l + 1 < c.levels; ++l) {
        const int r1 = pyr_strip_rows(c, l);
        dim3 g((c.lh[l + 1] + r1 - 1) / r1, ncalls);
        lc.begin(K_PYR);
        k_pyr<SRC_PYR><<<g, 256, pyr_smem_bytes(c, l, r1), st>>>(c, d_calls, ncalls, d, d_frames, frame_bytes, l, r1);
        lc.end();
    }
    lk_launch(c, d_calls, ncalls, d, maps, sm_count, lc);
    lc.begin(K_POST_A);
    k_post_a<<<ncalls, 256, 0, st>>>(c, d_calls, d);
    lc.end();
    return 0;
}

int front_launch_tail(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, int any_pub,
                      const uint8_t *d_depth, size_t depth_frame_bytes, int depth_fmt, LaunchCtx &lc)
{
    cudaStream_t st = lc.st;
    lc.begin(K_POST_B);
    k_post_b<<<ncalls, 256, post_b_smem_bytes(), st>>>(c, d_calls, d);
    lc.end();
    if (any_pub) {
        dim3 g(c.ncells, ncalls);
        lc.begin(K_FAST);
        k_fast<<<g, 256, fast_smem_bytes(c), st>>>(c, d_calls, d);
        lc.end();
    }
    lc.begin(K_FINISH);
    k_finish<<<ncalls, 256, 0, st>>>(c, d_calls, d, d_depth, depth_frame_bytes, depth_fmt);
    lc.end();
    return 0;
}

}  // namespace vrf
