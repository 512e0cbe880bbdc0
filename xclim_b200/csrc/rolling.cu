// Rolling-window primitives: (1) rolling statistic then per-period reduction, (2) spell statistics
// for spells defined by a rolling window (window > 1).
//
// Replaces:
//   indices/generic.py:128-174   select_rolling_resample_op   (da.rolling(time=w).<op>() then resample)
//   indices/_simple.py:485-525   max_n_day_precipitation_amount
//   indices/generic.py:434-540   spell_mask, general path (:519-535) and min/max fast path (:503-518)
//   indices/generic.py:543-585   _spell_length_statistics (window > 1)
//
// Design (B200): a lane owns one cell and walks the time steps of one period; the w values of a
// window are re-read from L1/L2 (the rows were just fetched by the same warp), so DRAM traffic stays
// one read of the input.  These are the low-volume members of the family (the window == 1 fast
// kernels in period_stats.cu carry the headline configurations); operators are runtime switches.
#include "common.cuh"

namespace xc {
namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ bool cmp_rt(int op, float x, float t) {
  switch (op) {
    case XC_OP_GT: return x > t;
    case XC_OP_LT: return x < t;
    case XC_OP_GE: return x >= t;
    case XC_OP_LE: return x <= t;
    case XC_OP_EQ: return x == t;
    default: return x != t;
  }
}

// statistic of x[i .. i+w-1] (NaN if any element is NaN) -- xarray rolling with min_periods == w
__device__ __forceinline__ float window_stat(const float* __restrict__ col, int64_t ldx, int i, int w, int stat) {
  double s = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  bool nan = false;
  for (int k = 0; k < w; ++k) {
    const float v = __ldg(col + (int64_t)(i + k) * ldx);
    nan = nan || (v != v);
    s += (double)v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  if (nan) return NAN;
  switch (stat) {
    case XC_STAT_SUM: return (float)s;
    case XC_STAT_MEAN: return (float)(s / (double)w);
    case XC_STAT_MIN: return mn;
    default: return mx;
  }
}

// Sliding window state for sum / mean windows: running float64 sum of the non-NaN values plus the
// number of NaNs in the window.  Sums of float32 values whose exponents lie within a few decades of
// each other are exact in float64, so the running sum equals the direct sum of the w values.
struct SlideSum {
  double s;
  int nan;
  __device__ __forceinline__ void add(float v) {
    const bool bad = (v != v);
    nan += bad ? 1 : 0;
    s += bad ? 0.0 : (double)v;
  }
  __device__ __forceinline__ void drop(float v) {
    const bool bad = (v != v);
    nan -= bad ? 1 : 0;
    s -= bad ? 0.0 : (double)v;
  }
};

constexpr int kChunk = 8;

__global__ void __launch_bounds__(kThreads)
rolling_period_reduce_kernel(const float* __restrict__ x, int64_t T, int64_t C, int64_t ldx,
                             const int32_t* __restrict__ poff, int32_t w, int32_t wstat, int32_t shift,
                             int32_t stat, float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  const float* col = x + c;
  double s = 0.0, q = 0.0;
  float m = (stat == XC_STAT_MIN) ? INFINITY : -INFINITY;
  int n = 0;
  auto take = [&](float r) {
    if (r == r) {
      ++n;
      s += (double)r;
      q += (double)r * (double)r;
      m = (stat == XC_STAT_MIN) ? fminf(m, r) : fmaxf(m, r);
    }
  };
  // the value labelled t covers x[i .. i+w-1], i = t + shift - w + 1 (shift 0: right-aligned, w/2: centred)
  if (wstat == XC_STAT_SUM || wstat == XC_STAT_MEAN) {
    const double inv = (wstat == XC_STAT_MEAN) ? 1.0 / (double)w : 1.0;
    int t = t0;
    // labels whose window is incomplete at the series start give NaN
    while (t < t1 && t + shift - w + 1 < 0) ++t;
    SlideSum win{0.0, 0};
    int i = t + shift - w + 1;          // first complete window of this period
    if (t < t1 && i + w <= (int)T)
      for (int k = 0; k < w; ++k) win.add(__ldg(col + (int64_t)(i + k) * ldx));
    while (t < t1 && i + w <= (int)T) {
      take(win.nan ? NAN : (float)(wstat == XC_STAT_MEAN ? win.s / (double)w : win.s));
      // slide by up to kChunk steps with all loads issued first
      const int steps = min(kChunk, min(t1 - 1 - t, (int)T - (i + w)));
      if (steps <= 0) break;
      float vin[kChunk], vout[kChunk];
#pragma unroll
      for (int k = 0; k < kChunk; ++k) {
        if (k < steps) {
          vin[k] = ld_stream(col + (int64_t)(i + w + k) * ldx);
          vout[k] = __ldg(col + (int64_t)(i + k) * ldx);
        }
      }
#pragma unroll
      for (int k = 0; k < kChunk; ++k) {
        if (k < steps) {
          win.add(vin[k]);
          win.drop(vout[k]);
          if (k + 1 < steps) take(win.nan ? NAN : (float)(wstat == XC_STAT_MEAN ? win.s / (double)w : win.s));
        }
      }
      t += steps;
      i += steps;
    }
    (void)inv;
  } else {
    for (int t = t0; t < t1; ++t) {
      const int i = t + shift - w + 1;
      float r = NAN;
      if (i >= 0 && i + w <= (int)T) r = window_stat(col, ldx, i, w, wstat);
      take(r);
    }
  }
  float res;
  const double nn = (double)n;
  switch (stat) {
    case XC_STAT_SUM: res = (float)s; break;
    case XC_STAT_COUNT: res = (float)n; break;
    case XC_STAT_MEAN: res = n ? (float)(s / nn) : NAN; break;
    case XC_STAT_MIN:
    case XC_STAT_MAX: res = n ? m : NAN; break;
    default: {
      if (!n) { res = NAN; break; }
      const double mean = s / nn;
      double var = q / nn - mean * mean;
      var = var > 0.0 ? var : 0.0;
      res = (stat == XC_STAT_STD) ? (float)sqrt(var) : (float)var;
    }
  }
  out[(int64_t)p * C + c] = res;
}

// mask[t] = any length-w block [i, i+w-1] containing t, fully inside the series, whose window
// statistic satisfies (stat op thr)   (indices/generic.py:519-535); then run-length statistics of
// the mask per period (window == 1 on the mask, generic.py:562-570).
//
// The block statistics are produced in time order by BlockStream: for sum / mean windows a sliding
// float64 sum (+ NaN counter), two loads per step; for min / max windows the w values are re-read
// (L1 hits).
struct BlockStream {
  const float* col;
  int64_t ldx;
  int T, w, wstat, op;
  float thr;
  int i;          // start of the next block to evaluate
  SlideSum win;   // window [i, i+w-1] when sliding
  bool primed;

  __device__ __forceinline__ void init(int start) {
    i = start;
    primed = false;
    win.s = 0.0;
    win.nan = 0;
  }
  // short windows are cheaper to re-read (L1 hits) than to slide with float64 adds
  __device__ __forceinline__ bool sliding() const { return (wstat == XC_STAT_SUM || wstat == XC_STAT_MEAN) && w > 6; }
  // qualifies(block starting at i) ; advances to i + 1.  Caller guarantees i + w <= T.
  __device__ __forceinline__ bool next() {
    float r;
    if (sliding()) {
      if (!primed) {
        for (int k = 0; k < w; ++k) win.add(__ldg(col + (int64_t)(i + k) * ldx));
        primed = true;
      }
      r = win.nan ? NAN : (float)(wstat == XC_STAT_MEAN ? win.s / (double)w : win.s);
      if (i + 1 + w <= T) {  // slide to i + 1: one new row, one row re-read from L1/L2
        win.add(__ldg(col + (int64_t)(i + w) * ldx));
        win.drop(__ldg(col + (int64_t)i * ldx));
      }
    } else {
      r = window_stat(col, ldx, i, w, wstat);
    }
    ++i;
    return cmp_rt(op, r, thr);
  }
};

__global__ void __launch_bounds__(kThreads)
spell_runstat_kernel(const float* __restrict__ x, int64_t T, int64_t C, int64_t ldx,
                     const int32_t* __restrict__ poff, int32_t w, int32_t wstat, int32_t op, float thr,
                     int32_t reducer, int32_t after, float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  BlockStream bs;
  bs.col = x + c; bs.ldx = ldx; bs.T = (int)T; bs.w = w; bs.wstat = wstat; bs.op = op; bs.thr = thr;
  // the mask at t needs the blocks starting in [t - w + 1, t]; with `after` the mask at t0 - 1 too
  const int first_t = after ? max(0, t0 - 1) : t0;
  bs.init(max(0, first_t - w + 1));
  int last_true = -1;
  auto mask_at = [&](int t) -> bool {
    const int imax = min(t, (int)T - w);
    while (bs.i <= imax) {
      const int ii = bs.i;
      if (bs.next()) last_true = ii;
    }
    return last_true >= 0 && last_true >= t - w + 1;
  };
  bool skip = false;
  if (after && t0 > 0) skip = mask_at(t0 - 1);  // a run already open belongs to an earlier period
  int cur = 0, mx = 0, mn = 0x7fffffff, sum = 0, cnt = 0;
  unsigned long long sq = 0ull;
  auto close_run = [&](int L) {
    if (L >= 1) {
      mx = max(mx, L);
      mn = min(mn, L);
      sum += L;
      cnt += 1;
      sq += (unsigned long long)L * (unsigned long long)L;
    }
  };
  for (int t = t0; t < t1; ++t) {
    bool m = mask_at(t);
    if (after) {
      skip = skip && m;
      m = m && !skip;
    }
    if (m) {
      ++cur;
    } else {
      close_run(cur);
      cur = 0;
    }
  }
  if (after) {
    int t = t1;
    while (cur > 0 && t < (int)T) {
      if (mask_at(t)) {
        ++cur;
      } else {
        close_run(cur);
        cur = 0;
      }
      ++t;
    }
  }
  close_run(cur);
  float res;
  switch (reducer) {
    case XC_RL_MAX: res = (float)mx; break;
    case XC_RL_MIN: res = cnt ? (float)mn : 0.f; break;
    case XC_RL_SUM: res = (float)sum; break;
    case XC_RL_COUNT: res = (float)cnt; break;
    case XC_RL_MEAN: res = cnt ? (float)((double)sum / (double)cnt) : 0.f; break;
    default: {
      if (!cnt) { res = 0.f; break; }
      const double n = (double)cnt, mean = (double)sum / n;
      double var = (double)sq / n - mean * mean;
      res = (float)sqrt(var > 0.0 ? var : 0.0);
    }
  }
  out[(int64_t)p * C + c] = res;
}

}  // namespace
}  // namespace xc

using namespace xc;

extern "C" int32_t xc_rolling_period_reduce_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                                const int32_t* period_offsets, int32_t P, int32_t window,
                                                int32_t window_stat, int32_t center, int32_t stat, float* out,
                                                void* stream) {
  XC_REQUIRE(x && period_offsets && out, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C && P > 0 && P <= 65535 && T < 2147483647LL, "bad shape");
  XC_REQUIRE(window >= 1 && window <= T, "window must be in [1, T]");
  XC_REQUIRE(window_stat == XC_STAT_SUM || window_stat == XC_STAT_MEAN || window_stat == XC_STAT_MIN ||
                 window_stat == XC_STAT_MAX,
             "window statistic must be sum, mean, min or max");
  XC_REQUIRE(stat >= XC_STAT_SUM && stat <= XC_STAT_COUNT, "unknown reduction %d", stat);
  if (center && window % 2 == 0) {
    set_error("centred rolling windows of even length are not supported");
    return XC_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)((C + kThreads - 1) / kThreads), (unsigned)P, 1);
  rolling_period_reduce_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      x, T, C, ldx, period_offsets, window, window_stat, center ? window / 2 : 0, stat, out);
  return launch_status("rolling_period_reduce_kernel");
}

extern "C" int32_t xc_spell_runstat_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                        const int32_t* period_offsets, int32_t P, int32_t window,
                                        int32_t window_stat, int32_t op, double thr, int32_t reducer,
                                        int32_t resample_before_rl, float* out, void* stream) {
  XC_REQUIRE(x && period_offsets && out, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C && P > 0 && P <= 65535 && T < 2147483647LL, "bad shape");
  XC_REQUIRE(window >= 1 && window <= T, "window must be in [1, T]");
  XC_REQUIRE(window_stat == XC_STAT_SUM || window_stat == XC_STAT_MEAN || window_stat == XC_STAT_MIN ||
                 window_stat == XC_STAT_MAX,
             "window reducer must be sum, mean, min or max");
  XC_REQUIRE(op >= XC_OP_GT && op <= XC_OP_NE, "Operation `%d` not recognized.", op);
  XC_REQUIRE(reducer >= XC_RL_MAX && reducer <= XC_RL_STD, "unknown run-length reducer %d", reducer);
  dim3 grid((unsigned)((C + kThreads - 1) / kThreads), (unsigned)P, 1);
  spell_runstat_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, T, C, ldx, period_offsets, window,
                                                                    window_stat, op, (float)thr, reducer,
                                                                    resample_before_rl ? 0 : 1, out);
  return launch_status("spell_runstat_kernel");
}
