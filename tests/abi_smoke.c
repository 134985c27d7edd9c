/* C-ABI smoke test (compiled by tests/test_abi_c.py with gcc -std=c99): the public headers are plain C, every entry
 * point links, and without a CUDA device the library refuses to create a handle instead of falling back to a CPU path. */
#include <stdio.h>
#include <string.h>

#include "vrf.h"

int main(void)
{
    VrfConfig cfg;
    vrf_handle *h = NULL;
    VrfBaProblem pb;
    VrfFmProblem fm;
    VrfImuSegment seg;
    memset(&pb, 0, sizeof pb); memset(&fm, 0, sizeof fm); memset(&seg, 0, sizeof seg);
    vrf_config_default(&cfg);
    printf("cfg %dx%d max_cnt %d grid %dx%d\n", cfg.col, cfg.row, cfg.max_cnt, cfg.num_grid_rows, cfg.num_grid_cols);
    printf("sizeof VrfConfig %zu VrfBaProblem %zu VrfPrior %zu VrfImuPreint %zu VrfFmProblem %zu\n", sizeof(VrfConfig),
           sizeof(VrfBaProblem), sizeof(VrfPrior), sizeof(VrfImuPreint), sizeof(VrfFmProblem));
    {
        int rc = vrf_create(&cfg, 1, 0, &h);
        printf("vrf_create rc %d (%s) handle %s\n", rc, vrf_strerror(rc), h ? "set" : "null");
        if (rc == VRF_OK) {
            /* a GPU is present: the stateless calls accept empty batches */
            int r1 = vrf_fm_triangulate_with_depth_batch(h, 0, &fm), r2 = vrf_imu_preintegrate_batch(h, 0, &seg, NULL);
            printf("empty batches rc %d %d\n", r1, r2);
            vrf_destroy(h);
        }
    }
    /* NULL handles are rejected, never dereferenced */
    printf("null handle rc %d %d %d\n", vrf_ba_solve(NULL, 0, &pb, NULL), vrf_fm_moving_consistency_check_batch(NULL, 1, &fm),
           vrf_set_fisheye_mask(NULL, NULL, 0));
    return 0;
}
