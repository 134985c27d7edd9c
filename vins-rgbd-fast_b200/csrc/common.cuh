// common.cuh -- shared definitions of the vrf CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vrf.h"

#define VRF_CAP        VRF_TRACK_CAP   // per-sequence capacity of the track arrays
#define VRF_MAX_CELLS  256
#define VRF_MAX_LEVELS 4
#define VRF_MAX_BATCH  4096            // sequences per batched call
#define VRF_LK_WIN     21
#define VRF_LK_HALF    10.0f

namespace vrf {

// Immutable per-handle front-end configuration, passed by value to kernels.
struct FrontCfg {
    int rows, cols;
    int levels;                 // maxLevel + 1
    int max_cnt, min_dist;
    int grows, gcols, gh, gw, ncells, thr, kmax;
    int use_imu, use_ransac;
    int lw[VRF_MAX_LEVELS], lh[VRF_MAX_LEVELS], lp[VRF_MAX_LEVELS];
    unsigned loff[VRF_MAX_LEVELS];
    unsigned pyr_bytes;         // bytes of one pyramid (all levels) per sequence
    double fx, fy, cx, cy, k1, k2, p1, p2;
    double ik11, ik13, ik22, ik23;
    double focal, f_thr;
    double depth_min_dist;
    int nodist;
    int equalize;               // EQUALIZE: cv::createCLAHE(3.0, Size(8, 8)) on the incoming frame (feature_tracker.cpp:269-275)
};

// One batch item of a tracker call.
struct SeqCall {
    int seq;
    int pub;
    int buf_prev, buf_cur;      // which pyramid buffer holds cur_img / receives forw_img
    int first;                  // 1: no previous image (forw_img.empty())
    int dslot;                  // index of this item's depth frame in the depth batch, -1 = no depth frame
    double dt;                  // cur_time - prev_time
    double R[9];                // relative_R row-major
};

// Device-resident per-sequence front-end state (SoA over sequences, VRF_CAP per sequence).
struct FrontDev {
    uint8_t *pyr[2];            // [S][pyr_bytes]
    float2 *cur_pts;            // positions in cur_img (previous frame)
    float2 *prev_un;            // undistorted normalised point stored by the previous undistortedPoints
    int *ids, *cnt;
    int *n_pts;                 // [S]
    int *n_id;                  // [S]
    // per-call scratch
    float2 *pred_pts, *lk_pts;  // LK init / raw output
    uint8_t *lk_status;
    int *n_lk;                  // [S] LK inputs of this call
    float2 *t_prev, *t_forw, *t_prevun;   // compacted working arrays between post kernels
    int *t_ids, *t_cnt, *t_n;
    uint8_t *t_keep;            // RANSAC inlier mask over the compacted arrays
    float2 *unstable;           // [S][CAP]
    int *n_unstable;
    int2 *maskpts;              // [S][2*CAP] rounded centres of every circle drawn by setMask
    int *n_maskpts;
    int *grid_cnt;              // [S][cells]   grids_track_num (persistent)
    uint8_t *tex_status;        // [S][cells]   grids_texture_status (persistent)
    int *cell_k;                // [S][cells]   K for selected cells, 0 otherwise
    float *cand;                // [S][cells][kmax][3] (x, y, response)
    int *ncand;                 // [S][cells]
    // outputs, indexed by batch position (not by sequence id): [batch][out_pitch] each, so that the result copy of
    // a batch is one contiguous transfer per array
    int out_pitch;
    float2 *o_pts, *o_un, *o_vel;
    int *o_ids, *o_cnt;
    uint16_t *o_depth;          // depth_img.at<ushort>((int)v, (int)u) in mm
    uint8_t *o_dkeep;           // 0: erased by the DEPTH_MIN_DIST test of addFeatureCheckParallax
    int *out_hdr;               // [batch][8]: n, n_id, n_predict, n_unstable, status
    int *work_prefix;           // [MAX_BATCH+1] prefix of LK work items for the current call
    uint8_t *clahe_lut;         // [S][64][256] per-tile CLAHE look-up tables (EQUALIZE only)
    const uint8_t *fisheye;     // [rows][cols] fisheye_mask (FISHEYE only, shared by all sequences), else NULL
};

__device__ __forceinline__ int reflect101(int i, int n)
{
    // cv::borderInterpolate(BORDER_REFLECT_101), single reflection (|overshoot| < n)
    if (i < 0) i = -i;
    else if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// 64-bit exact warp sum of 32-bit partials using two redux.sync instructions.
__device__ __forceinline__ long long warp_sum_i64(int part)
{
    int hi = part >> 16;
    unsigned lo = (unsigned)part & 0xFFFFu;
    int shi = __reduce_add_sync(0xffffffffu, hi);
    unsigned slo = __reduce_add_sync(0xffffffffu, lo);
    return (long long)shi * 65536 + (long long)slo;
}

// PinholeCamera::distortion (camera_model/src/camera_models/PinholeCamera.cc:646-663)
__device__ __forceinline__ void cam_distortion(const FrontCfg &c, double x, double y, double &dx, double &dy)
{
    double mx2 = x * x, my2 = y * y, mxy = x * y;
    double rho2 = mx2 + my2;
    double rad = c.k1 * rho2 + c.k2 * rho2 * rho2;
    dx = x * rad + 2.0 * c.p1 * mxy + c.p2 * (rho2 + 2.0 * mx2);
    dy = y * rad + 2.0 * c.p2 * mxy + c.p1 * (rho2 + 2.0 * my2);
}

// PinholeCamera::liftProjective (PinholeCamera.cc:450-510): 8 fixed-point iterations.
__device__ __forceinline__ void cam_lift(const FrontCfg &c, double u, double v, double &mx, double &my)
{
    double mx_d = c.ik11 * u + c.ik13;
    double my_d = c.ik22 * v + c.ik23;
    if (c.nodist) { mx = mx_d; my = my_d; return; }
    double dx, dy;
    cam_distortion(c, mx_d, my_d, dx, dy);
    double mx_u = mx_d - dx, my_u = my_d - dy;
#pragma unroll 1
    for (int i = 1; i < 8; ++i) {
        cam_distortion(c, mx_u, my_u, dx, dy);
        mx_u = mx_d - dx;
        my_u = my_d - dy;
    }
    mx = mx_u; my = my_u;
}

// PinholeCamera::spaceToPlane (PinholeCamera.cc:520-543)
__device__ __forceinline__ void cam_project(const FrontCfg &c, double X, double Y, double Z, double &u, double &v)
{
    double x = X / Z, y = Y / Z;
    if (!c.nodist) {
        double dx, dy;
        cam_distortion(c, x, y, dx, dy);
        x += dx; y += dy;
    }
    u = c.fx * x + c.cx;
    v = c.fy * y + c.cy;
}

}  // namespace vrf
