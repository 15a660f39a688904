// sink.cuh -- internal interface of the on-device sample sink (sink.cu) used by the sampler handles.
#pragma once
#include "plan.cuh"

struct gi_stats;

namespace gi {

struct StatsMap {
    signed char fin[64];  // chains whose proposal finished in this step
};

// For every chain c < C with map.fin[c] set and (st == nullptr or st[c].res.accept): offer the
// chain's model scale .* mw[c] to slot slot0 + c (device-side gating, two launches, no host sync).
int stats_add_chains(gi_stats *s, const StatsMap &map, const DevState *st, const double *mw,
                     int64_t mw_ld, int C, int slot0, const double *scale, cudaStream_t stream);

}  // namespace gi
