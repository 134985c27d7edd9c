// ransac_kernels.cu -- rejectWithF (feature_tracker.cpp:441-473): placeholder, filled in next.
#include "common.cuh"
#include "handle.h"
namespace vrf {
int ransac_launch(const FrontCfg &, const SeqCall *, int, const FrontDev &, LaunchCtx &) { return 0; }
}
