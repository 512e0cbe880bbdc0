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

#ifndef XC_EQM_THREADS
#define XC_EQM_THREADS 256
#endif
#ifndef XC_EQM_CAP
#define XC_EQM_CAP 128
#endif
constexpr int kSortThreads = XC_EQM_THREADS;
constexpr int kBins = 1024;      // histogram bins over [min, max] of the series
constexpr int kCap = XC_EQM_CAP; // candidates kept per needed bin
constexpr int kMaxTargets = 128; // order statistics per series (2 per quantile)

// Block-wide bitonic sort of keys[0..NPAD) ascending (fallback path).
template <int NPAD>
__device__ void block_bitonic_sort(float* keys, int tid) {
  for (int k = 2; k <= NPAD; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NPAD / 2; i += kSortThreads) {
        const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));  // pair (a, a | j)
        const int b = a | j;
        const bool up = ((a & k) == 0);
        const float x0 = keys[a], x1 = keys[b];
        const bool swap = up ? (x0 > x1) : (x0 < x1);
        if (swap) { keys[a] = x1; keys[b] = x0; }
      }
      __syncthreads();
    }
  }
}

// One CTA = one cell.  The nq quantiles need only ~2*nq order statistics, so instead of sorting the
// series the CTA (1) histograms it over kBins equal-width bins between its min and max (binning by a
// monotone float expression keeps bins ordered), (2) locates the bin and the in-bin rank of every
// wanted order statistic from the prefix sums, (3) gathers only the elements of those bins and (4)
// ranks them by counting.  Degenerate distributions (a wanted bin with more than kCap elements that
// are not all equal, e.g. heavy ties next to other values) fall back to the full bitonic sort.
template <int NPAD>
__global__ void __launch_bounds__(kSortThreads)
eqm_train_kernel(const float* __restrict__ ref, const float* __restrict__ hist, int32_t T, int64_t C, int64_t ldx,
                 int32_t nq, int32_t kind, float* __restrict__ af, float* __restrict__ hist_q) {
  extern __shared__ float keys[];             // NPAD keys | nq ref quantiles | candidate lists
  float* refq = keys + NPAD;
  float* cand = refq + ((nq + 3) & ~3);       // [2 * nq][kCap]
  __shared__ int hist_s[kBins + 1];           // counts, then exclusive prefix sums
  __shared__ int slot_of_bin[kBins];          // -1 or candidate-list slot
  __shared__ int cand_n[kMaxTargets];         // elements stored per slot
  __shared__ int tgt_rank[kMaxTargets], tgt_bin[kMaxTargets];
  __shared__ float tgt_val[kMaxTargets];
  __shared__ float red_min[kSortThreads / 32], red_max[kSortThreads / 32];
  __shared__ int red_cnt[kSortThreads / 32];
  __shared__ int s_flag, s_nslots;
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ntg = 2 * nq;
  for (int pass = 0; pass < 2; ++pass) {
    const float* src = (pass == 0 ? ref : hist) + c;
    // ---- load, count valid, min / max
    float mn = INFINITY, mx = -INFINITY;
    int nvalid = 0;
    for (int t = tid; t < T; t += kSortThreads) {
      const float v = ld_stream(src + (int64_t)t * ldx);
      keys[t] = v;
      if (v == v) { mn = fminf(mn, v); mx = fmaxf(mx, v); ++nvalid; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    }
    if (lane == 0) { red_min[wid] = mn; red_max[wid] = mx; red_cnt[wid] = nvalid; }
    for (int b = tid; b <= kBins; b += kSortThreads) hist_s[b] = 0;
    for (int b = tid; b < kBins; b += kSortThreads) slot_of_bin[b] = -1;
    if (tid == 0) { s_flag = 0; s_nslots = 0; }
    __syncthreads();
    mn = red_min[0]; mx = red_max[0]; int n = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) { mn = fminf(mn, red_min[w]); mx = fmaxf(mx, red_max[w]); n += red_cnt[w]; }
    const bool degenerate = !(mx > mn) || !(mx - mn < INFINITY);  // constant series, or infinite range
    const float scale = degenerate ? 0.f : (float)kBins / (mx - mn);
    auto bin_of = [&](float v) -> int { return min(kBins - 1, (int)((v - mn) * scale)); };
    bool use_sort = false;
    if (n > 0 && !degenerate) {
      // ---- histogram + prefix sums
      for (int t = tid; t < T; t += kSortThreads) {
        const float v = keys[t];
        if (v == v) atomicAdd(&hist_s[bin_of(v)], 1);
      }
      __syncthreads();
      if (wid == 0) {  // exclusive scan of kBins counters by one warp (kBins / 32 per lane)
        constexpr int per = kBins / 32;
        int loc[per], sum = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) { loc[i] = hist_s[lane * per + i]; sum += loc[i]; }
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for (int i = 0; i < per; ++i) { hist_s[lane * per + i] = run; run += loc[i]; }
        if (lane == 31) hist_s[kBins] = run;
      }
      __syncthreads();
      // ---- wanted order statistics: ranks ilo, ihi of every quantile
      for (int g = tid; g < ntg; g += kSortThreads) {
        const int j = g >> 1;
        // nodes are cast to the data dtype before use (xsdba: equally_spaced_nodes(n).astype(ref.dtype))
        const double q = (double)(float)(((double)j + 0.5) / (double)nq);
        const int ilo = (int)floor(q * (double)(n - 1));
        const int r = (g & 1) ? min(ilo + 1, n - 1) : ilo;
        int lo = 0, hi = kBins;  // bin b with prefix[b] <= r < prefix[b+1]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (hist_s[mid] <= r) lo = mid; else hi = mid;
        }
        tgt_bin[g] = lo;
        tgt_rank[g] = r - hist_s[lo];
        if (atomicCAS(&slot_of_bin[lo], -1, -2) == -1) {  // first to claim the bin allocates its slot
          const int sl = atomicAdd(&s_nslots, 1);
          cand_n[sl] = 0;
          slot_of_bin[lo] = sl;
        }
      }
      __syncthreads();
      // ---- gather the elements of the wanted bins
      for (int t = tid; t < T; t += kSortThreads) {
        const float v = keys[t];
        if (v == v) {
          const int sl = slot_of_bin[bin_of(v)];
          if (sl >= 0) {
            const int p = atomicAdd(&cand_n[sl], 1);
            if (p < kCap) cand[sl * kCap + p] = v;
          }
        }
      }
      __syncthreads();
      // ---- heavy bins (more than kCap elements): fine when all their elements are equal (ties, e.g. dry
      // days), which one extra block-wide min/max over that bin establishes; otherwise use the sort
      for (int sl = 0; sl < s_nslots; ++sl) {
        if (cand_n[sl] <= kCap) continue;  // uniform: shared memory value
        float lo_v = INFINITY, hi_v = -INFINITY;
        for (int t = tid; t < T; t += kSortThreads) {
          const float v = keys[t];
          if (v == v && slot_of_bin[bin_of(v)] == sl) { lo_v = fminf(lo_v, v); hi_v = fmaxf(hi_v, v); }
        }
        for (int o = 16; o > 0; o >>= 1) {
          lo_v = fminf(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, o));
          hi_v = fmaxf(hi_v, __shfl_xor_sync(0xffffffffu, hi_v, o));
        }
        __syncthreads();  // red_* are free again
        if (lane == 0) { red_min[wid] = lo_v; red_max[wid] = hi_v; }
        __syncthreads();
        if (tid == 0) {
          float a = red_min[0], b = red_max[0];
          for (int w = 1; w < kSortThreads / 32; ++w) { a = fminf(a, red_min[w]); b = fmaxf(b, red_max[w]); }
          if (a == b) { cand[sl * kCap] = a; cand_n[sl] = -1; }  // constant bin
          else s_flag = 1;
        }
        __syncthreads();
      }
      // ---- rank inside the bin by counting (one warp per target)
      for (int g = wid; g < ntg; g += kSortThreads / 32) {
        const int sl = slot_of_bin[tgt_bin[g]];
        const int m = cand_n[sl];
        const float* cl = cand + sl * kCap;
        if (m < 0) {  // constant bin
          if (lane == 0) tgt_val[g] = cl[0];
          continue;
        }
        if (m > kCap) continue;  // resolved by the sort below
        const int want = tgt_rank[g];
        float found = NAN;
        for (int i = lane; i < m; i += 32) {
          const float vi = cl[i];
          int less = 0;
          for (int k = 0; k < m; ++k) {
            const float vk = cl[k];
            less += (vk < vi || (vk == vi && k < i)) ? 1 : 0;
          }
          if (less == want) found = vi;
        }
        for (int o = 16; o > 0; o >>= 1) {
          const float other = __shfl_xor_sync(0xffffffffu, found, o);
          found = (found == found) ? found : other;
        }
        if (lane == 0) tgt_val[g] = found;
      }
      __syncthreads();
      use_sort = (s_flag != 0);
    }
    if (use_sort) {
      // fallback: NaN -> +inf, pad with +inf, full sort; targets read from the sorted keys
      for (int t = tid; t < NPAD; t += kSortThreads) {
        float v = (t < T) ? keys[t] : INFINITY;
        if (v != v) v = INFINITY;
        keys[t] = v;
      }
      __syncthreads();
      block_bitonic_sort<NPAD>(keys, tid);
      for (int g = tid; g < ntg; g += kSortThreads) tgt_val[g] = keys[tgt_rank[g] + hist_s[tgt_bin[g]]];
      __syncthreads();
    }
    // ---- quantiles (numpy's _lerp on the two neighbours)
    for (int j = tid; j < nq; j += kSortThreads) {
      float qv = NAN;
      if (n > 0) {
        if (degenerate) {
          qv = mn;  // every valid value equals mn (or the range is not finite: not supported, gives mn)
        } else {
          const double q = (double)(float)(((double)j + 0.5) / (double)nq);
          const double pos = q * (double)(n - 1);
          const double lo = floor(pos);
          const double g = pos - lo;
          const double a0 = (double)tgt_val[2 * j], a1 = (double)tgt_val[2 * j + 1];
          const double d = a1 - a0;
          qv = (float)((g >= 0.5) ? (a1 - d * (1.0 - g)) : (a0 + d * g));
        }
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
  extern __shared__ float tab[];  // [3][nq][kAdjThreads]: hist_q, af, slope of the segment ending at j; column = lane
  float* hq = tab;
  float* fa = tab + (size_t)nq * kAdjThreads;
  float* sl = fa + (size_t)nq * kAdjThreads;
  const int lane = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * kAdjThreads + lane;
  if (c >= C) return;
  bool bad = false;
  float hp = 0.f, ap = 0.f;
  for (int j = 0; j < nq; ++j) {
    const float h = hist_q[(int64_t)j * C + c], a = af[(int64_t)j * C + c];
    hq[j * kAdjThreads + lane] = h;
    fa[j * kAdjThreads + lane] = a;
    // the division of the linear interpolation is done once per segment instead of once per element
    // (same operands, same rounding: the results are bit-identical)
    sl[j * kAdjThreads + lane] = (j > 0) ? __fdiv_rn(__fsub_rn(a, ap), __fsub_rn(h, hp)) : 0.f;
    hp = h;
    ap = a;
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
    const float x0 = hq[(idx - 1) * kAdjThreads + lane];
    const float y0 = fa[(idx - 1) * kAdjThreads + lane];
    if (INTERP == 1) {
      const float slope = sl[idx * kAdjThreads + lane];
      return __fadd_rn(__fmul_rn(slope, __fsub_rn(x, x0)), y0);
    }
    // nearest: boundaries at the mid-points, ties go to the lower node
    const float x1 = hq[idx * kAdjThreads + lane], y1 = fa[idx * kAdjThreads + lane];
    const float mid = __fmul_rn(__fadd_rn(x0, x1), 0.5f);
    return (x <= mid) ? y0 : y1;
  };
  constexpr int U = 8;  // rows in flight per lane
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
  XC_REQUIRE(nq >= 1 && nq <= 64, "nquantiles must be in [1, 64]");
  XC_REQUIRE(kind == 0 || kind == 1, "kind must be 0 ('+') or 1 ('*')");
  XC_REQUIRE(C <= 2147483647LL, "too many cells for one launch");
  if (T > 32768) {
    set_error("eqm_train: series longer than 32768 steps do not fit the shared-memory sort");
    return XC_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int npad = 1024;
  while (npad < T) npad <<= 1;
  const size_t smem = ((size_t)npad + ((nq + 3) & ~3) + (size_t)2 * nq * kCap) * 4;
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
  const size_t smem = (size_t)3 * nq * kAdjThreads * 4;
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
