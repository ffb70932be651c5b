// crt_fused.cuh — fused tile kernel (placeholder plan: staged path only for now).
#pragma once
#include "crt_stages.cuh"

namespace crt {

struct FusedPlan {
    bool ok = false;
    const char* why = "fused kernel not built yet";
};

inline FusedPlan plan_fused(const Dev&, bool, int) { return FusedPlan{}; }

inline int run_fused(const FusedPlan&, const Dev&, const FrameDev&, const uint8_t*, uint8_t*, float*, int, cudaStream_t, int*) { return 4; }

}  // namespace crt
