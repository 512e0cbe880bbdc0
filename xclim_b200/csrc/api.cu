// Library-wide pieces of the C ABI: error reporting, version, threshold folding.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace xc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  (void)cudaGetLastError();  // clear the sticky-less error state
  return XC_ERR_CUDA;
}

float fold_threshold(int32_t op, double thr, int32_t cmp_f64) {
  float f = (float)thr;  // round to nearest even, as numpy does for a weak Python scalar
  if (!cmp_f64 || thr != thr || op > XC_OP_NE) return f;
  const double fd = (double)f;
  switch (op) {
    case XC_OP_GT:  // x > thr  <=>  x > rd(thr)   (largest float32 <= thr)
    case XC_OP_LE:  // x <= thr <=>  x <= rd(thr)
      if (fd > thr) f = nextafterf(f, -INFINITY);
      return f;
    case XC_OP_GE:  // x >= thr <=>  x >= ru(thr)  (smallest float32 >= thr)
    case XC_OP_LT:  // x < thr  <=>  x < ru(thr)
      if (fd < thr) f = nextafterf(f, INFINITY);
      return f;
    case XC_OP_EQ:  // never equal to an unrepresentable threshold: NaN compares False
    case XC_OP_NE:  // always different:                              NaN compares True
      return fd == thr ? f : NAN;
  }
  return f;
}

}  // namespace xc

extern "C" int32_t xc_version(void) { return XC_VERSION; }
extern "C" const char* xc_last_error(void) { return xc::g_err; }

extern "C" int32_t xc_device_sm_count(int32_t* out_sm_count) {
  XC_REQUIRE(out_sm_count != nullptr, "null pointer argument");
  int dev = 0, n = 0;
  XC_CHECK_CUDA(cudaGetDevice(&dev));
  XC_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out_sm_count = n;
  return XC_OK;
}
