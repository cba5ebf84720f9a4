// Tensor-core conv engine (tcgen05 + TMA) - placeholder until the engine lands.
#pragma once
#include <vector>
#include "ptd_internal.h"
struct TcConvDesc {
    const float* src0; const float* src1; int c0p, c1p; int upsample; int H, W; int coutp; float* out; bool lrelu_first;
    const float* scale; const float* shift; const float* bias; float* pool_out;
};
struct TcConvPlan { int valid = 0; };
inline ptd_status tc_plan_create(const TcConvDesc&, const std::vector<float>&, int, TcConvPlan&, std::vector<void*>&) {
    PTD_FAIL(PTD_ERR_UNSUPPORTED, "tensor-core conv engine not built yet");
}
inline void tc_plan_destroy(TcConvPlan&) {}
inline ptd_status tc_conv_launch(TcConvPlan&, cudaStream_t, int*, bool*) { PTD_FAIL(PTD_ERR_UNSUPPORTED, "tensor-core conv engine not built yet"); }
