// lk_kernels.cu -- k_lk: predictPtsInNextFrame (feature_tracker.cpp:595-608) + the per-point body of
// cv::calcOpticalFlowPyrLK (call sites feature_tracker.cpp:302-310; LKTrackerInvoker semantics pinned against cv2 4.13
// in oracle/frontend_spec.py::lk_track).  Compiled with -fmad=false (x86 rounding of the float/double expressions).
//
// One warp per feature, all pyramid levels.  Round-2 design:
//   * the 24x24 patch of the previous level (I) and a 32x32 tile of the next level (J) around the start position are
//     staged in shared memory by TMA (cp.async.bulk.tensor.3d on a [sequence][row][col] u8 tensor map per pyramid
//     level; one elected lane issues, the warp waits on its own mbarrier).  The TMA unit wants the box origin 16-byte
//     aligned in the innermost dimension (measured on B200: an unaligned x raises "illegal instruction",
//     tools/scratch/tma_guide.cu), so the box is 48 x 32 with x0 rounded down to a multiple of 16 and the kernel indexes
//     the tile with the residual offset.  Tiles that touch the image border are gathered with BORDER_REFLECT_101
//     addressing instead (= the padded pyramid of buildOpticalFlowPyramid).  The J tile has a 5-pixel margin, so all
//     <= 30 iterations of a level normally run out of the one tile (restaged if the window leaves it): no global-memory
//     traffic inside the iteration loop.
//   * a lane owns a fixed set of window pixels for the whole level -- 6 (5) consecutive rows of the column pair
//     (x, x+4), x < 16 (OpenCV's v_dotprod pair) plus <= 4 rows of one column of the scalar tail x = 16..20 -- and keeps
//     their I, Ix, Iy in registers.  Consecutive rows share the bilinear taps: the interpolation is
//     jv = (dp2a(Wtop, row y) + dp2a(Wbot, row y+1) + 256) >> 9 with IDP.2A (2 x (s16 weight x u8 pixel)), two
//     instructions per pixel, exact integer arithmetic => identical to OpenCV's.
//   * Scharr derivatives of the patch with IDP.4A row kernels ((-1,0,1) and (3,10,3)), 0 outside the image
//     (BORDER_CONSTANT of the derivative pyramid), interpolated with the same integer weights.
//   * the 2x2 normal equations / mismatch vector are accumulated in float32 in *exactly OpenCV's SIMD lane order*:
//     every product is written as a float addend into a per-chain array, and one lane per chain adds them in order
//     (16-byte loads, 4 adds per load) => tracks are bit-identical to cv2.calcOpticalFlowPyrLK.
#include "common.cuh"
#include "handle.h"
#include "tma.cuh"

namespace vrf {

#define LK_TW 48                            // staged tile: 32 rows x 48 bytes (TMA box 48 x 32 x 1, x origin a multiple of 16)
#define LK_TP (LK_TW / 4)                   // tile row pitch in 32-bit words
#define LK_T_BYTES (32 * LK_TW)
#define LK_U_BYTES 5632                     // union region: derivative planes | A addends | b addends
#define LK_WARP_BYTES (2 * LK_T_BYTES + LK_U_BYTES)     // 8704 = 68 * 128

// b-sum addends (floats in U): SIMD chain (q, r) at (4 q + r) * LK_FS + s (s = 2 y + hq, 42 steps + 2 zeros), tail chain q at
// LK_FT_BASE + q * LK_FT_STRIDE + t (t = 5 y + x - 16, 105 steps + 3 zeros).  LK_FS = 8 mod 32: the 32 lanes of a store hit 32 banks.
#define LK_FS 72
#define LK_FT_BASE (8 * LK_FS)
#define LK_FT_STRIDE 108
// A-sum addends (floats in U): SIMD chain (q, r) at (4 q + r) * LK_AS + 4 y + (x >> 2) (84 steps), tail chain q at
// LK_AT_BASE + q * LK_AT_STRIDE + 5 y + x - 16 (105 steps + 3 zeros).  All chain bases are 16-byte aligned.
#define LK_AS 84
#define LK_AT_BASE (12 * LK_AS + 28)
#define LK_AT_STRIDE 108
// derivative planes (ints in U): Dx[22][22] at 0, Dy[22][22] at LK_DY_OFF
#define LK_DP 22
#define LK_DY_OFF 512

static_assert((LK_AT_BASE + 3 * LK_AT_STRIDE) * 4 <= LK_U_BYTES, "A addends exceed the union region");
static_assert((LK_FT_BASE + 2 * LK_FT_STRIDE) * 4 <= LK_U_BYTES, "b addends exceed the union region");
static_assert((LK_DY_OFF + LK_DP * LK_DP) * 4 <= LK_U_BYTES, "derivative planes exceed the union region");

// OpenCV's float accumulation order (lkpyramid.cpp SSE path; pinned bit-exactly against cv2 4.13 in
// oracle/frontend_spec.py::_sum_a_opencv/_sum_b_opencv): per window row, pixels 0..15 feed 4 SIMD-lane accumulators, pixels
// 16..20 a scalar one; total = scalar + ((l0 + l2) + (l1 + l3)).
// A-sums: chains live on lanes 5 q + r (r = 0..3 SIMD accumulators, r = 4 scalar tail)
__device__ __forceinline__ float lk_combine(float acc, int q)
{
    float l0 = __shfl_sync(0xffffffffu, acc, 5 * q + 0);
    float l1 = __shfl_sync(0xffffffffu, acc, 5 * q + 1);
    float l2 = __shfl_sync(0xffffffffu, acc, 5 * q + 2);
    float l3 = __shfl_sync(0xffffffffu, acc, 5 * q + 3);
    float tl = __shfl_sync(0xffffffffu, acc, 5 * q + 4);
    return tl + ((l0 + l2) + (l1 + l3));
}

// b-sums: chains live on lanes 4 q + r (SIMD accumulators) and 8 + q (scalar tail)
__device__ __forceinline__ float lk_combine_b(float acc, int q)
{
    float l0 = __shfl_sync(0xffffffffu, acc, 4 * q + 0);
    float l1 = __shfl_sync(0xffffffffu, acc, 4 * q + 1);
    float l2 = __shfl_sync(0xffffffffu, acc, 4 * q + 2);
    float l3 = __shfl_sync(0xffffffffu, acc, 4 * q + 3);
    float tl = __shfl_sync(0xffffffffu, acc, 8 + q);
    return tl + ((l0 + l2) + (l1 + l3));
}

// Stage columns [c0, c0 + 32) x rows [0, nrows) of the tile with origin (x0, y0) of pyramid level `img` (REFLECT_101 outside
// the image) without TMA.
__device__ __forceinline__ void lk_stage_reflect(uint8_t *tile, const uint8_t *img, int pitch, int cols, int rows, int x0, int y0,
                                                 int c0, int nrows, int lane)
{
    const int gx = reflect101(x0 + c0 + lane, cols);
#pragma unroll 8
    for (int rr = 0; rr < nrows; ++rr) {
        const int gy = reflect101(y0 + rr, rows);
        tile[rr * LK_TW + c0 + lane] = __ldg(img + (size_t)gy * pitch + gx);
    }
}

__global__ void __launch_bounds__(LK_WPB * 32, 2)
k_lk(FrontCfg c, const SeqCall *calls, int ncalls, FrontDev d, const __grid_constant__ LkMaps maps)
{
    extern __shared__ unsigned char lk_smem_raw[];
    // 128-byte aligned carve-up: [0, 128) the warps' mbarriers, then LK_WARP_BYTES per warp
    unsigned char *lk_smem = lk_smem_raw + ((128u - (smem_u32(lk_smem_raw) & 127u)) & 127u);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(lk_smem) + wib;
    unsigned char *wbase = lk_smem + 128 + wib * LK_WARP_BYTES;
    uint8_t *tileI = wbase, *tileJ = wbase + LK_T_BYTES;
    const unsigned *TI = reinterpret_cast<const unsigned *>(tileI);
    const unsigned *TJ = reinterpret_cast<const unsigned *>(tileJ);
    float *sF = reinterpret_cast<float *>(wbase + 2 * LK_T_BYTES);       // addend arrays
    int *sD = reinterpret_cast<int *>(wbase + 2 * LK_T_BYTES);           // derivative planes
    unsigned ph = 0;                                                       // mbarrier phase parity of this warp
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncwarp();

    const int total = d.work_prefix[ncalls];
    const int maxLevel = c.levels - 1;
    // window pixels owned by this lane.  SIMD role: rows y0 .. y0+4 (+5 for the last group) of columns xa and xa + 4.
    const int g = lane >> 3, hq = (lane >> 2) & 1, r = lane & 3;
    const int y0 = 5 * g, xa = 8 * hq + r;
    const bool g3 = (g == 3);
    // tail role (lanes 0..29): column 16 + tx, rows ty0 .. ty0+2 (+3 for the odd groups); the group starts 0,3,7,10,14,17
    // put the six groups' tile rows on disjoint shared-memory banks (row pitch 12 words)
    const bool tl_on = lane < 30;
    const int tg = tl_on ? lane / 5 : 0, tx = tl_on ? lane - 5 * tg : 0;
    const int ty0 = 3 * tg + (tg >> 1);
    const bool t4 = tl_on && (tg & 1);
    // accumulation chain owned by this lane in the A phase
    const int cq = lane / 5, cr = lane - 5 * cq;

    for (int gi = blockIdx.x * LK_WPB + wib; gi < total; gi += gridDim.x * LK_WPB) {
        // locate the batch item: largest ci with prefix[ci] <= gi
        int lo = 0, hi = ncalls;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (d.work_prefix[mid] <= gi) lo = mid; else hi = mid;
        }
        const int ci = lo;
        const int idx = gi - d.work_prefix[ci];
        const SeqCall call = calls[ci];
        const size_t base = (size_t)call.seq * VRF_CAP + idx;
        const uint8_t *pyrI = d.pyr[call.buf_prev] + (size_t)call.seq * c.pyr_bytes;
        const uint8_t *pyrJ = d.pyr[call.buf_cur] + (size_t)call.seq * c.pyr_bytes;

        const float2 prev = d.cur_pts[base];
        float2 init = prev;
        if (c.use_imu) {
            double mx, my;
            cam_lift(c, (double)prev.x, (double)prev.y, mx, my);
            const double *R = call.R;
            double X = R[0] * mx + R[1] * my + R[2];
            double Y = R[3] * mx + R[4] * my + R[5];
            double Z = R[6] * mx + R[7] * my + R[8];
            double u, v;
            cam_project(c, X, Y, Z, u, v);
            init.x = (float)u; init.y = (float)v;
        }
        if (lane == 0) d.pred_pts[base] = init;

        int st = 1;
        float2 nextStored = init;
        for (int level = maxLevel; level >= 0; --level) {
            const int cols = c.lw[level], rows = c.lh[level], pitch = c.lp[level];
            const uint8_t *I = pyrI + c.loff[level];
            const uint8_t *J = pyrJ + c.loff[level];
            const float scale = 1.0f / (float)(1 << level);
            float2 prevPt = make_float2(prev.x * scale, prev.y * scale);
            float2 nextPt;
            if (level == maxLevel) {
                if (c.use_imu) nextPt = make_float2(init.x * scale, init.y * scale);
                else nextPt = prevPt;
            } else
                nextPt = make_float2(nextStored.x * 2.f, nextStored.y * 2.f);
            nextStored = nextPt;
            prevPt.x -= VRF_LK_HALF; prevPt.y -= VRF_LK_HALF;
            const int ix = (int)floorf(prevPt.x), iy = (int)floorf(prevPt.y);
            if (ix < -VRF_LK_WIN || ix >= cols || iy < -VRF_LK_WIN || iy >= rows) {
                if (level == 0) st = 0;
                continue;
            }
            float a = prevPt.x - (float)ix, b = prevPt.y - (float)iy;
            const int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
            const int iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
            const int iw10 = __float2int_rn((1.f - a) * b * 16384.f);
            const int iw11 = 16384 - iw00 - iw01 - iw10;

            // ---- stage the I patch (origin ix-1, iy-1; 24 x 24 used) and the J tile (origin jx-5, jy-5; 32 x 32 used) ----
            nextPt.x -= VRF_LK_HALF; nextPt.y -= VRF_LK_HALF;
            int tx0, tyo, jal;          // origin of the staged J tile (logical), its 16-aligned x origin
            const int ial = (ix - 1) & ~15, ioff = (ix - 1) - ial;
            {
                const int jx = (int)floorf(nextPt.x), jy = (int)floorf(nextPt.y);
                const bool jstaged = !(jx < -VRF_LK_WIN || jx >= cols || jy < -VRF_LK_WIN || jy >= rows);   // else iteration 0 leaves at once
                tx0 = jx - 5; tyo = jy - 5; jal = tx0 & ~15;
                const bool tmaI = ix >= 1 && iy >= 1 && ix + 23 <= cols && iy + 23 <= rows;
                const bool tmaJ = jstaged && tx0 >= 0 && tyo >= 0 && tx0 + 32 <= cols && tyo + 32 <= rows;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && (tmaI || tmaJ)) {
                    mbar_arrive_expect_tx(bar, (tmaI ? LK_T_BYTES : 0) + (tmaJ ? LK_T_BYTES : 0));
                    if (tmaI) tma_load_3d(tileI, &maps.m[call.buf_prev][level], bar, ial, iy - 1, call.seq);
                    if (tmaJ) tma_load_3d(tileJ, &maps.m[call.buf_cur][level], bar, jal, tyo, call.seq);
                }
                // (only the 32 columns that can be read are gathered: the 24-wide patch resp. the 32-wide window range, both
                // starting at the residual offset inside the 16-aligned tile)
                if (!tmaI) lk_stage_reflect(tileI, I, pitch, cols, rows, ial, iy - 1, ioff, 24, lane);
                if (jstaged && !tmaJ) lk_stage_reflect(tileJ, J, pitch, cols, rows, jal, tyo, tx0 - jal, 32, lane);
                if (tmaI || tmaJ) { mbar_wait(bar, ph); ph ^= 1u; }
                __syncwarp();
            }

            // ---- I window (5 fractional bits) of the owned pixels ----
            const int Wt = (iw00 & 0xFFFF) | (iw01 << 16), Wb = (iw10 & 0xFFFF) | (iw11 << 16);
            int IvA[6], IvB[6], IvT[4];
            {
                const int cA = ioff + xa + 1, sh = (cA & 3) * 8;
                const unsigned *rp = TI + (y0 + 1) * LK_TP + (cA >> 2);
                int tA = 0, tB = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const unsigned w0 = rp[LK_TP * k], w1 = rp[LK_TP * k + 1], w2 = rp[LK_TP * k + 2];
                    const unsigned pa = __funnelshift_r(w0, w1, sh), pb = __funnelshift_r(w1, w2, sh);
                    if (k > 0) { IvA[k - 1] = dp2a_lo_su(Wb, pa, tA) >> 9; IvB[k - 1] = dp2a_lo_su(Wb, pb, tB) >> 9; }
                    if (k < 6) { tA = dp2a_lo_su(Wt, pa, 256); tB = dp2a_lo_su(Wt, pb, 256); }
                }
                const int cT = ioff + 17 + tx, shT = (cT & 3) * 8;
                const unsigned *rt = TI + (ty0 + 1) * LK_TP + (cT >> 2);
                int tT = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const unsigned w0 = rt[LK_TP * k], w1 = rt[LK_TP * k + 1];
                    const unsigned p = __funnelshift_r(w0, w1, shT);
                    if (k > 0) IvT[k - 1] = dp2a_lo_su(Wb, p, tT) >> 9;
                    if (k < 4) tT = dp2a_lo_su(Wt, p, 256);
                }
            }
            // ---- Scharr derivative planes at the 22 x 22 positions (ix + dx, iy + dy); 0 outside the image ----
            if (lane < LK_DP) {
                const int sh = ((ioff + lane) & 3) * 8;
                const unsigned *cp = TI + ((ioff + lane) >> 2);
                const bool xin = (ix + lane >= 0) && (ix + lane < cols);
                int hx0 = 0, hx1 = 0, sm0 = 0, sm1 = 0;
#pragma unroll
                for (int R = 0; R < 24; ++R) {
                    const unsigned v = __funnelshift_r(cp[LK_TP * R], cp[LK_TP * R + 1], sh);   // columns dx, dx+1, dx+2 of patch row R
                    const int hx = dp4a_us(v, 0x000100FF, 0);          // p[dx+2] - p[dx]
                    const int sm = dp4a_us(v, 0x00030A03, 0);          // 3 p[dx] + 10 p[dx+1] + 3 p[dx+2]
                    if (R >= 2) {
                        const int dy_ = R - 2;
                        const bool in = xin && (iy + dy_ >= 0) && (iy + dy_ < rows);
                        sD[dy_ * LK_DP + lane] = in ? 3 * (hx0 + hx) + 10 * hx1 : 0;
                        sD[LK_DY_OFF + dy_ * LK_DP + lane] = in ? sm - sm0 : 0;
                    }
                    hx0 = hx1; hx1 = hx; sm0 = sm1; sm1 = sm;
                }
            }
            __syncwarp();
            // ---- Ix, Iy of the owned pixels: the same integer bilinear interpolation, (sum + 2^13) >> 14 ----
            int IxA[6], IyA[6], IxB[6], IyB[6], IxT[4], IyT[4];
            {
                const int *dp = sD + y0 * LK_DP + xa;
                int sxa = 0, sya = 0, sxb = 0, syb = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int ax0 = dp[k * LK_DP], ax1 = dp[k * LK_DP + 1], bx0 = dp[k * LK_DP + 4], bx1 = dp[k * LK_DP + 5];
                    const int ay0 = dp[LK_DY_OFF + k * LK_DP], ay1 = dp[LK_DY_OFF + k * LK_DP + 1];
                    const int by0 = dp[LK_DY_OFF + k * LK_DP + 4], by1 = dp[LK_DY_OFF + k * LK_DP + 5];
                    if (k > 0) {
                        IxA[k - 1] = (sxa + ax0 * iw10 + ax1 * iw11) >> 14; IyA[k - 1] = (sya + ay0 * iw10 + ay1 * iw11) >> 14;
                        IxB[k - 1] = (sxb + bx0 * iw10 + bx1 * iw11) >> 14; IyB[k - 1] = (syb + by0 * iw10 + by1 * iw11) >> 14;
                    }
                    if (k < 6) {
                        sxa = ax0 * iw00 + ax1 * iw01 + (1 << 13); sya = ay0 * iw00 + ay1 * iw01 + (1 << 13);
                        sxb = bx0 * iw00 + bx1 * iw01 + (1 << 13); syb = by0 * iw00 + by1 * iw01 + (1 << 13);
                    }
                }
                const int *dt = sD + ty0 * LK_DP + 16 + tx;
                int sxt = 0, syt = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int x0_ = dt[k * LK_DP], x1_ = dt[k * LK_DP + 1];
                    const int y0_ = dt[LK_DY_OFF + k * LK_DP], y1_ = dt[LK_DY_OFF + k * LK_DP + 1];
                    if (k > 0) { IxT[k - 1] = (sxt + x0_ * iw10 + x1_ * iw11) >> 14; IyT[k - 1] = (syt + y0_ * iw10 + y1_ * iw11) >> 14; }
                    if (k < 4) { sxt = x0_ * iw00 + x1_ * iw01 + (1 << 13); syt = y0_ * iw00 + y1_ * iw01 + (1 << 13); }
                }
            }
            __syncwarp();
            // ---- A11, A12, A22: products as float addends in chain order, then 15 ordered float chains ----
            {
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if (k < 5 || g3) {
                        const int o = 4 * (y0 + k) + 2 * hq;
                        *reinterpret_cast<float2 *>(sF + r * LK_AS + o) = make_float2((float)(IxA[k] * IxA[k]), (float)(IxB[k] * IxB[k]));
                        *reinterpret_cast<float2 *>(sF + (4 + r) * LK_AS + o) = make_float2((float)(IxA[k] * IyA[k]), (float)(IxB[k] * IyB[k]));
                        *reinterpret_cast<float2 *>(sF + (8 + r) * LK_AS + o) = make_float2((float)(IyA[k] * IyA[k]), (float)(IyB[k] * IyB[k]));
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (tl_on && (k < 3 || t4)) {
                        const int o = LK_AT_BASE + 5 * (ty0 + k) + tx;
                        sF[o] = (float)(IxT[k] * IxT[k]);
                        sF[o + LK_AT_STRIDE] = (float)(IxT[k] * IyT[k]);
                        sF[o + 2 * LK_AT_STRIDE] = (float)(IyT[k] * IyT[k]);
                    }
                }
                if (lane < 9) sF[LK_AT_BASE + (lane / 3) * LK_AT_STRIDE + 105 + lane % 3] = 0.f;
            }
            __syncwarp();
            float A11, A12, A22;
            {
                float acc = 0.f;
                if (lane < 15) {
                    const float4 *src = reinterpret_cast<const float4 *>(sF + (cr < 4 ? (4 * cq + cr) * LK_AS : LK_AT_BASE + cq * LK_AT_STRIDE));
                    const int nq = cr < 4 ? 21 : 27;
#pragma unroll 3
                    for (int qd = 0; qd < nq; ++qd) {
                        const float4 f4 = src[qd];
                        acc += f4.x; acc += f4.y; acc += f4.z; acc += f4.w;
                    }
                }
                A11 = lk_combine(acc, 0);
                A12 = lk_combine(acc, 1);
                A22 = lk_combine(acc, 2);
            }
            const float FLT_SCALE = 1.f / (float)(1 << 20);
            A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / 882.f;
            if (minEig < 1e-4f || D < 1.1920929e-07f) {
                if (level == 0) st = 0;
                continue;
            }
            D = 1.f / D;
            __syncwarp();
            // the union region now holds the b addends: zero padding of the chains (x + 0.0f is exact)
            if (lane < 16) sF[(lane >> 1) * LK_FS + 42 + (lane & 1)] = 0.f;                                   // SIMD chains: steps 42, 43
            else if (lane < 22) sF[LK_FT_BASE + ((lane - 16) / 3) * LK_FT_STRIDE + 105 + (lane - 16) % 3] = 0.f;   // tails: 105..107
            float2 prevDelta = make_float2(0.f, 0.f);
            for (int j = 0; j < 30; ++j) {
                const int jx = (int)floorf(nextPt.x), jy = (int)floorf(nextPt.y);
                if (jx < -VRF_LK_WIN || jx >= cols || jy < -VRF_LK_WIN || jy >= rows) {
                    if (level == 0) st = 0;
                    break;
                }
                a = nextPt.x - (float)jx; b = nextPt.y - (float)jy;
                const int w00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
                const int w01 = __float2int_rn(a * (1.f - b) * 16384.f);
                const int w10 = __float2int_rn((1.f - a) * b * 16384.f);
                const int w11 = 16384 - w00 - w01 - w10;
                const int Vt = (w00 & 0xFFFF) | (w01 << 16), Vb = (w10 & 0xFFFF) | (w11 << 16);
                int offx = jx - tx0, offy = jy - tyo;
                __syncwarp();                   // the chain lanes are done with the previous iteration's addends
                if (offx < 0 || offx > 10 || offy < 0 || offy > 10) {
                    // the window left the staged tile: restage around the current position
                    tx0 = jx - 5; tyo = jy - 5; offx = 5; offy = 5; jal = tx0 & ~15;
                    const bool tmaJ = tx0 >= 0 && tyo >= 0 && tx0 + 32 <= cols && tyo + 32 <= rows;
                    fence_proxy_async();
                    __syncwarp();
                    if (tmaJ) {
                        if (lane == 0) {
                            mbar_arrive_expect_tx(bar, LK_T_BYTES);
                            tma_load_3d(tileJ, &maps.m[call.buf_cur][level], bar, jal, tyo, call.seq);
                        }
                        mbar_wait(bar, ph); ph ^= 1u;
                    } else
                        lk_stage_reflect(tileJ, J, pitch, cols, rows, jal, tyo, tx0 - jal, 32, lane);
                    __syncwarp();
                }
                offx += tx0 - jal;              // column of the window origin inside the 16-aligned tile
                {
                    const int cA = offx + xa, sh = (cA & 3) * 8;
                    const unsigned *rp = TJ + (offy + y0) * LK_TP + (cA >> 2);
                    float *f1 = sF + r * LK_FS + 2 * y0 + hq;
                    int tA = 0, tB = 0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const unsigned w0 = rp[LK_TP * k], w1 = rp[LK_TP * k + 1], w2 = rp[LK_TP * k + 2];
                        const unsigned pa = __funnelshift_r(w0, w1, sh), pb = __funnelshift_r(w1, w2, sh);
                        if (k > 0) {
                            const int dA = (dp2a_lo_su(Vb, pa, tA) >> 9) - IvA[k - 1];
                            const int dB = (dp2a_lo_su(Vb, pb, tB) >> 9) - IvB[k - 1];
                            const int d1 = dA * IxA[k - 1] + dB * IxB[k - 1];      // v_dotprod pair sum, exact in int32
                            const int d2 = dA * IyA[k - 1] + dB * IyB[k - 1];
                            if (k - 1 < 5 || g3) {
                                f1[2 * (k - 1)] = (float)d1;
                                f1[2 * (k - 1) + 4 * LK_FS] = (float)d2;
                            }
                        }
                        if (k < 6) { tA = dp2a_lo_su(Vt, pa, 256); tB = dp2a_lo_su(Vt, pb, 256); }
                    }
                    const int cT = offx + 16 + tx, shT = (cT & 3) * 8;
                    const unsigned *rt = TJ + (offy + ty0) * LK_TP + (cT >> 2);
                    float *ft = sF + LK_FT_BASE + 5 * ty0 + tx;
                    int tT = 0;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const unsigned w0 = rt[LK_TP * k], w1 = rt[LK_TP * k + 1];
                        const unsigned p = __funnelshift_r(w0, w1, shT);
                        if (k > 0) {
                            const int dT = (dp2a_lo_su(Vb, p, tT) >> 9) - IvT[k - 1];
                            if (tl_on && (k - 1 < 3 || t4)) {
                                ft[5 * (k - 1)] = (float)(dT * IxT[k - 1]);
                                ft[5 * (k - 1) + LK_FT_STRIDE] = (float)(dT * IyT[k - 1]);
                            }
                        }
                        if (k < 4) tT = dp2a_lo_su(Vt, p, 256);
                    }
                }
                __syncwarp();
                // b1, b2: 10 ordered float chains (lanes 0..3 / 4..7: SIMD accumulators of b1 / b2, lanes 8, 9: scalar tails);
                // every chain is a contiguous, zero-padded array, read 4 addends at a time
                float acc = 0.f;
                if (lane < 10) {
                    const float4 *src = reinterpret_cast<const float4 *>(sF + (lane < 8 ? lane * LK_FS : LK_FT_BASE + (lane - 8) * LK_FT_STRIDE));
                    const int nq = lane < 8 ? 11 : 27;
#pragma unroll 3
                    for (int qd = 0; qd < nq; ++qd) {
                        const float4 f4 = src[qd];
                        acc += f4.x; acc += f4.y; acc += f4.z; acc += f4.w;
                    }
                }
                float fb1 = lk_combine_b(acc, 0) * FLT_SCALE;
                float fb2 = lk_combine_b(acc, 1) * FLT_SCALE;
                float2 delta = make_float2((A12 * fb2 - A22 * fb1) * D, (A12 * fb1 - A11 * fb2) * D);
                nextPt.x += delta.x; nextPt.y += delta.y;
                nextStored = make_float2(nextPt.x + VRF_LK_HALF, nextPt.y + VRF_LK_HALF);
                if ((double)delta.x * (double)delta.x + (double)delta.y * (double)delta.y <= 0.01 * 0.01) break;
                if (j > 0 && (double)fabsf(delta.x + prevDelta.x) < 0.01 && (double)fabsf(delta.y + prevDelta.y) < 0.01) {
                    nextStored.x -= delta.x * 0.5f; nextStored.y -= delta.y * 0.5f;
                    break;
                }
                prevDelta = delta;
            }
            if (st && level == 0) {
                // `err` is requested by the reference => final bounds check (lkpyramid.cpp)
                float qx = nextStored.x - VRF_LK_HALF, qy = nextStored.y - VRF_LK_HALF;
                int kx = (int)floorf(qx), ky = (int)floorf(qy);
                if (kx < -VRF_LK_WIN || kx >= cols || ky < -VRF_LK_WIN || ky >= rows) st = 0;
            }
        }
        if (lane == 0) {
            d.lk_pts[base] = nextStored;
            d.lk_status[base] = (uint8_t)st;
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t lk_smem_bytes() { return (size_t)LK_WPB * LK_WARP_BYTES + 128 + 128; }

int lk_configure(const FrontCfg &c, const FrontDev &d, int n_seq, LkMaps *maps)
{
    cudaError_t e = cudaFuncSetAttribute(k_lk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_smem_bytes());
    if (e != cudaSuccess) return (int)e;
    memset(maps, 0, sizeof(*maps));
    for (int bsel = 0; bsel < 2; ++bsel)
        for (int l = 0; l < c.levels; ++l) {
            int rc = tma_encode_u8_3d(&maps->m[bsel][l], d.pyr[bsel] + c.loff[l], (uint64_t)c.lw[l], (uint64_t)c.lh[l], (uint64_t)n_seq,
                                      (uint64_t)c.lp[l], (uint64_t)c.pyr_bytes, LK_TW, 32);
            if (rc != 0) return rc;
        }
    return 0;
}

int lk_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, const LkMaps &maps, int sm_count, LaunchCtx &lc)
{
    long long maxwork = (long long)ncalls * VRF_CAP;
    long long want = ((long long)ncalls * (c.max_cnt + 2 * c.ncells) + LK_WPB - 1) / LK_WPB;
    long long cap = (long long)sm_count * 8;
    int grid = (int)max(1LL, min(min(want, cap), maxwork));
    lc.begin(K_LK);
    k_lk<<<grid, LK_WPB * 32, lk_smem_bytes(), lc.st>>>(c, d_calls, ncalls, d, maps);
    lc.end();
    return 0;
}

}  // namespace vrf
