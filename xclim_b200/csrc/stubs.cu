// Entry points declared in include/xclim_b200.h whose kernels are not written yet.
// They fail loudly (XC_ERR_UNSUPPORTED -> NotImplementedError); no CPU fallback exists.
#include "common.cuh"
using namespace xc;
#define XC_STUB(name) set_error(name " is not implemented yet"); return XC_ERR_UNSUPPORTED

extern "C" int32_t xc_rolling_period_reduce_f32(const float*, int64_t, int64_t, int64_t, const int32_t*, int32_t,
                                                int32_t, int32_t, int32_t, int32_t, float*, void*) {
  XC_STUB("xc_rolling_period_reduce_f32");
}
extern "C" int64_t xc_eqm_train_workspace_bytes(int64_t, int64_t, int32_t) { return 0; }
extern "C" int32_t xc_eqm_train_f32(const float*, const float*, int64_t, int64_t, int64_t, int32_t, int32_t, float*,
                                    float*, void*, int64_t, void*) {
  XC_STUB("xc_eqm_train_f32");
}
extern "C" int32_t xc_eqm_adjust_f32(const float*, int64_t, int64_t, int64_t, const float*, const float*, int32_t,
                                     int32_t, int32_t, float*, void*) {
  XC_STUB("xc_eqm_adjust_f32");
}
