// Empirical quantile mapping (sdba / xsdba `EmpiricalQuantileMapping`, group="time").
//
// The arithmetic lives in the third-party package xsdba (re-exported by sdba.py:11; floor pin
// xsdba>=0.4.0, pyproject.toml:111) whose sources are NOT under /root/reference: this file restates
// the published algorithm (PARITY UNPINNED):
//   train : quantile nodes q_j = (j + 1/2)/nq; hist_q[j], ref_q[j] = NaN-aware linear (type 7)
//           quantiles of the whole series; af = ref_q - hist_q ("+") or ref_q / hist_q ("*").
//   adjust: scen = sim (+|*) interp(sim; hist_q -> af), nearest or linear, constant extrapolation.
// Reference call sites: tests/test_xsdba.py:21-34, 143-150.
//
// Design (B200).  train needs ~2*nq order statistics spread over the WHOLE distribution of a
// 10950-sample series, i.e. a real per-cell sort: one CTA owns one cell, gathers its series into
// shared memory (64 KB for 16384 keys) and runs a bitonic sort there; neighbouring cells are
// handled by neighbouring CTAs at the same time so that the 32-byte sectors fetched by the strided
// gather are shared through L2.  adjust is a streaming pass: a lane owns one cell, keeps the 2*nq
// table entries of its cell in a conflict-free shared-memory column and does a binary search per
// element.
#include "common.cuh"

namespace xc {
namespace {

constexpr int kSortThreads = 256;

template <int NPAD>
__global__ void __launch_bounds__(kSortThreads)
eqm_train_kernel(const float* __restrict__ ref, const float* __restrict__ hist, int32_t T, int64_t C, int64_t ldx,
                 int32_t nq, int32_t kind, float* __restrict__ af, float* __restrict__ hist_q) {
  extern __shared__ float keys[];  // NPAD sort keys + nq quantiles of ref
  __shared__ int s_nan;
  float* refq = keys + NPAD;
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  for (int pass = 0; pass < 2; ++pass) {
    const float* src = (pass == 0 ? ref : hist) + c;
    if (tid == 0) s_nan = 0;
    __syncthreads();
    int my_nan = 0;
    for (int t = tid; t < NPAD; t += kSortThreads) {
      float v = INFINITY;  // padding sorts last
      if (t < T) {
        v = ld_stream(src + (int64_t)t * ldx);
        if (v != v) { v = INFINITY; ++my_nan; }
      }
      keys[t] = v;
    }
    if (my_nan) atomicAdd(&s_nan, my_nan);
    __syncthreads();
    // bitonic sort, ascending
    for (int k = 2; k <= NPAD; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < NPAD / 2; i += kSortThreads) {
          // element pair (a, a ^ j) with a's j-bit clear
          const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
          const int b = a | j;
          const bool up = ((a & k) == 0);
          const float x0 = keys[a], x1 = keys[b];
          const bool swap = up ? (x0 > x1) : (x0 < x1);
          if (swap) { keys[a] = x1; keys[b] = x0; }
        }
        __syncthreads();
      }
    }
    const int n = T - s_nan;  // NaNs (and padding) are +inf at the end; genuine +inf data are not supported
    for (int j = tid; j < nq; j += kSortThreads) {
      float qv = NAN;
      if (n > 0) {
        // nodes are cast to the data dtype before use (xsdba: equally_spaced_nodes(n).astype(ref.dtype))
        const double q = (double)(float)(((double)j + 0.5) / (double)nq);
        const double pos = q * (double)(n - 1);
        const double lo = floor(pos);
        const int ilo = (int)lo;
        const int ihi = min(ilo + 1, n - 1);
        const double g = pos - lo;
        const double a0 = (double)keys[ilo], a1 = (double)keys[ihi];
        const double d = a1 - a0;
        qv = (float)((g >= 0.5) ? (a1 - d * (1.0 - g)) : (a0 + d * g));  // numpy's _lerp
      }
      if (pass == 0) {
        refq[j] = qv;
      } else {
        hist_q[(int64_t)j * C + c] = qv;
        af[(int64_t)j * C + c] = (kind == 0) ? (refq[j] - qv) : (refq[j] / qv);
      }
    }
    __syncthreads();
  }
}

constexpr int kAdjThreads = 128;

template <int INTERP, int KIND>
__global__ void __launch_bounds__(kAdjThreads)
eqm_adjust_kernel(const float* __restrict__ sim, int64_t T, int64_t C, int64_t ldx, const float* __restrict__ af,
                  const float* __restrict__ hist_q, int32_t nq, int32_t rows_per_block, float* __restrict__ scen) {
  extern __shared__ float tab[];  // [2][nq][kAdjThreads]: hist_q then af, column = lane
  float* hq = tab;
  float* fa = tab + (size_t)nq * kAdjThreads;
  const int lane = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * kAdjThreads + lane;
  if (c >= C) return;
  bool bad = false;
  for (int j = 0; j < nq; ++j) {
    const float h = hist_q[(int64_t)j * C + c], a = af[(int64_t)j * C + c];
    hq[j * kAdjThreads + lane] = h;
    fa[j * kAdjThreads + lane] = a;
    bad = bad || (h != h) || (a != a);
  }
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t t1 = min(T, t0 + rows_per_block);
  const float h_first = hq[lane], h_last = hq[(nq - 1) * kAdjThreads + lane];
  const float a_first = fa[lane], a_last = fa[(nq - 1) * kAdjThreads + lane];
  int top_step = 1;  // largest power of two < nq
  while (top_step * 2 < nq) top_step *= 2;
  auto factor = [&](float x) -> float {
    if (bad || x != x) return NAN;
    if (x <= h_first) return a_first;
    if (x >= h_last) return a_last;
    // idx = #{hq < x} (searchsorted side="left") by a fixed-depth, branch-free bisection
    int idx = 0;
    for (int step = top_step; step >= 1; step >>= 1) {
      const int probe = idx + step;
      const float hv = hq[min(probe, nq) * kAdjThreads - kAdjThreads + lane];  // hq[probe - 1]
      idx = (probe <= nq && hv < x) ? probe : idx;
    }
    idx = max(1, min(idx, nq - 1));
    const float x0 = hq[(idx - 1) * kAdjThreads + lane], x1 = hq[idx * kAdjThreads + lane];
    const float y0 = fa[(idx - 1) * kAdjThreads + lane], y1 = fa[idx * kAdjThreads + lane];
    if (INTERP == 1) {
      const float slope = __fdiv_rn(__fsub_rn(y1, y0), __fsub_rn(x1, x0));
      return __fadd_rn(__fmul_rn(slope, __fsub_rn(x, x0)), y0);
    }
    // nearest: boundaries at the mid-points, ties go to the lower node
    const float mid = __fmul_rn(__fadd_rn(x0, x1), 0.5f);
    return (x <= mid) ? y0 : y1;
  };
  constexpr int U = 4;  // rows in flight per lane
  int64_t t = t0;
  for (; t + U <= t1; t += U) {
    float xv[U], fv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = ld_stream(sim + (t + u) * ldx + c);
#pragma unroll
    for (int u = 0; u < U; ++u) fv[u] = factor(xv[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) scen[(t + u) * C + c] = (KIND == 0) ? __fadd_rn(xv[u], fv[u]) : __fmul_rn(xv[u], fv[u]);
  }
  for (; t < t1; ++t) {
    const float x = ld_stream(sim + t * ldx + c);
    const float f = factor(x);
    scen[t * C + c] = (KIND == 0) ? __fadd_rn(x, f) : __fmul_rn(x, f);
  }
}

}  // namespace
}  // namespace xc

using namespace xc;

extern "C" int64_t xc_eqm_train_workspace_bytes(int64_t T, int64_t C, int32_t nq) {
  (void)T; (void)C; (void)nq;
  return 0;  // the sort runs in shared memory
}

extern "C" int32_t xc_eqm_train_f32(const float* ref, const float* hist, int64_t T, int64_t C, int64_t ldx,
                                    int32_t nq, int32_t kind, float* af, float* hist_q, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  XC_REQUIRE(ref && hist && af && hist_q, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C, "bad shape");
  XC_REQUIRE(nq >= 1 && nq <= 1024, "nquantiles must be in [1, 1024]");
  XC_REQUIRE(kind == 0 || kind == 1, "kind must be 0 ('+') or 1 ('*')");
  XC_REQUIRE(C <= 2147483647LL, "too many cells for one launch");
  if (T > 32768) {
    set_error("eqm_train: series longer than 32768 steps do not fit the shared-memory sort");
    return XC_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int npad = 1024;
  while (npad < T) npad <<= 1;
  const size_t smem = ((size_t)npad + nq) * 4;
#define XC_TRAIN(NP)                                                                                               \
  do {                                                                                                             \
    if (smem > 48 * 1024) {                                                                                        \
      cudaError_t e_ = cudaFuncSetAttribute(eqm_train_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                            (int)smem);                                                            \
      if (e_ != cudaSuccess) return cuda_fail(e_, "cudaFuncSetAttribute(eqm_train_kernel)");                       \
    }                                                                                                              \
    eqm_train_kernel<NP><<<(unsigned)C, kSortThreads, smem, st>>>(ref, hist, (int32_t)T, C, ldx, nq, kind, af,     \
                                                                  hist_q);                                         \
  } while (0)
  switch (npad) {
    case 1024: XC_TRAIN(1024); break;
    case 2048: XC_TRAIN(2048); break;
    case 4096: XC_TRAIN(4096); break;
    case 8192: XC_TRAIN(8192); break;
    case 16384: XC_TRAIN(16384); break;
    default: XC_TRAIN(32768); break;
  }
#undef XC_TRAIN
  return launch_status("eqm_train_kernel");
}

extern "C" int32_t xc_eqm_adjust_f32(const float* sim, int64_t T, int64_t C, int64_t ldx, const float* af,
                                     const float* hist_q, int32_t nq, int32_t kind, int32_t interp, float* scen,
                                     void* stream) {
  XC_REQUIRE(sim && af && hist_q && scen, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C, "bad shape");
  XC_REQUIRE(nq >= 2 && nq <= 200, "nquantiles must be in [2, 200]");
  XC_REQUIRE(kind == 0 || kind == 1, "kind must be 0 ('+') or 1 ('*')");
  XC_REQUIRE(interp == 0 || interp == 1, "interp must be 0 (nearest) or 1 (linear)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cblocks = (C + kAdjThreads - 1) / kAdjThreads;
  int tchunks = (int)((148 * 16 + cblocks - 1) / cblocks);
  tchunks = tchunks < 1 ? 1 : tchunks;
  int rows = (int)((T + tchunks - 1) / tchunks);
  if (rows < 64) rows = 64;
  tchunks = (int)((T + rows - 1) / rows);
  const size_t smem = (size_t)2 * nq * kAdjThreads * 4;
  dim3 grid((unsigned)cblocks, (unsigned)tchunks, 1);
#define XC_ADJ(I, K)                                                                                               \
  do {                                                                                                             \
    if (smem > 48 * 1024) {                                                                                        \
      cudaError_t e_ = cudaFuncSetAttribute(eqm_adjust_kernel<I, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                            (int)smem);                                                            \
      if (e_ != cudaSuccess) return cuda_fail(e_, "cudaFuncSetAttribute(eqm_adjust_kernel)");                      \
    }                                                                                                              \
    eqm_adjust_kernel<I, K><<<grid, kAdjThreads, smem, st>>>(sim, T, C, ldx, af, hist_q, nq, rows, scen);          \
  } while (0)
  if (interp == 0 && kind == 0) XC_ADJ(0, 0);
  else if (interp == 0) XC_ADJ(0, 1);
  else if (kind == 0) XC_ADJ(1, 0);
  else XC_ADJ(1, 1);
#undef XC_ADJ
  return launch_status("eqm_adjust_kernel");
}
