// Streaming per-period statistics: threshold counts, run-length statistics, reductions.
//
// Replaces (reference paths relative to /root/reference/src/xclim):
//   indices/generic.py:301-361      compare, threshold_count
//   indices/generic.py:543-585      _spell_length_statistics (window == 1)
//   indices/run_length.py:87-132    resample_and_rl
//   indices/run_length.py:143-335   _cumsum_reset_np, rle, rle_statistics
//   indices/run_length.py:381-488   windowed_run_events / windowed_run_count
//   indices/generic.py:83-125       select_resample_op ; 1514-1552 cumulative_difference
//   core/missing.py:296-322         MissingAny's valid-step count (fused as `valid_count`)
//
// Design (B200): the (time, lat, lon) buffer is coalesced along cells, so a THREAD owns VEC=4
// adjacent cells (one 128-bit load per time step; a warp reads 512 contiguous bytes per step) and
// marches through the time steps of ONE period keeping the run-length state machine in
// registers -- no shuffles, no shared memory, no intermediate arrays.  The grid is
// (cell-vectors, periods): every (period, cell) unit is independent when runs are cut at period
// edges; with resample_before_rl == 0 a unit additionally looks one step back (to skip a run
// that started earlier) and reads past the period end until its open runs close.  UNROLL
// independent 16-byte loads are issued before any use so that each thread keeps >= 128 B in
// flight; data is read exactly once (ld.global.nc.L1::no_allocate).
#include <type_traits>

#include "common.cuh"

namespace xc {
namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 8;

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = ld_stream(p); }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 q = ld_stream4(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
};

template <int VEC, typename T>
__device__ __forceinline__ void store_vec(T* dst, const T (&v)[VEC]) {
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    uint4 q;
    q.x = *reinterpret_cast<const uint32_t*>(&v[0]);
    q.y = *reinterpret_cast<const uint32_t*>(&v[1]);
    q.z = *reinterpret_cast<const uint32_t*>(&v[2]);
    q.w = *reinterpret_cast<const uint32_t*>(&v[3]);
    *reinterpret_cast<uint4*>(dst) = q;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = v[i];
  }
}

// Iterate f(t, Vec) over t in [t0, t1) with kUnroll loads in flight.
template <int VEC, typename F>
__device__ __forceinline__ void stream_rows(const float* __restrict__ col, int64_t ldx, int t0, int t1,
                                            F&& f) {
  int t = t0;
  const float* p = col + (int64_t)t0 * ldx;
  for (; t + kUnroll <= t1; t += kUnroll) {
    Vec<VEC> r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) r[u].load(p + (int64_t)u * ldx);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) f(r[u]);
    p += (int64_t)kUnroll * ldx;
  }
  for (; t < t1; ++t) {
    Vec<VEC> r;
    r.load(p);
    f(r);
    p += ldx;
  }
}

// ------------------------------------------------------------------------------------------------
// threshold count (+ valid count)
// ------------------------------------------------------------------------------------------------
template <int OP, int VEC, bool VALID>
__global__ void __launch_bounds__(kThreads)
period_count_kernel(const float* __restrict__ x, int64_t C, int64_t ldx,
                    const int32_t* __restrict__ poff, float thr,
                    int32_t* __restrict__ out, int32_t* __restrict__ valid) {
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  int32_t cnt[VEC], nv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { cnt[i] = 0; nv[i] = 0; }
  stream_rows<VEC>(x + c0, ldx, t0, t1, [&](const Vec<VEC>& r) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      cnt[i] += cmp<OP>(r.v[i], thr) ? 1 : 0;
      if constexpr (VALID) nv[i] += (r.v[i] == r.v[i]) ? 1 : 0;
    }
  });
  store_vec<VEC>(out + (int64_t)p * C + c0, cnt);
  if constexpr (VALID) store_vec<VEC>(valid + (int64_t)p * C + c0, nv);
}

// ------------------------------------------------------------------------------------------------
// run-length statistics
// ------------------------------------------------------------------------------------------------
// Per-cell accumulator over the run lengths L >= window attributed to the period.
template <int RED>
struct RunAcc {
  int32_t a;        // MAX: max L ; MIN: min L ; SUM: sum L ; COUNT/MEAN/STD: n runs
  int32_t s;        // MEAN/STD: sum L
  unsigned long long q;  // STD: sum L^2
  __device__ __forceinline__ void init() {
    a = (RED == XC_RL_MIN) ? 0x7fffffff : 0;
    s = 0;
    q = 0ull;
  }
  __device__ __forceinline__ void add(int32_t L, bool take) {
    if constexpr (RED == XC_RL_MAX) a = take ? max(a, L) : a;
    if constexpr (RED == XC_RL_MIN) a = take ? min(a, L) : a;
    if constexpr (RED == XC_RL_SUM) a += take ? L : 0;
    if constexpr (RED == XC_RL_COUNT) a += take ? 1 : 0;
    if constexpr (RED == XC_RL_MEAN || RED == XC_RL_STD) {
      a += take ? 1 : 0;
      s += take ? L : 0;
    }
    if constexpr (RED == XC_RL_STD) q += take ? (unsigned long long)L * (unsigned long long)L : 0ull;
  }
  __device__ __forceinline__ float result() const {
    if constexpr (RED == XC_RL_MAX || RED == XC_RL_SUM || RED == XC_RL_COUNT) return (float)a;
    if constexpr (RED == XC_RL_MIN) return a == 0x7fffffff ? 0.f : (float)a;
    if constexpr (RED == XC_RL_MEAN) return a == 0 ? 0.f : (float)((double)s / (double)a);
    if constexpr (RED == XC_RL_STD) {
      if (a == 0) return 0.f;
      double n = (double)a, m = (double)s / n;
      double var = (double)q / n - m * m;
      return (float)sqrt(var > 0.0 ? var : 0.0);
    }
    return 0.f;
  }
};

// FASTMAX: reducer == max and window == 1 needs no run-end detection: max over t of the running
// length (this is the maximum_consecutive_dry_days configuration).
template <int OP, int RED, int VEC, bool VALID, bool AFTER, bool FASTMAX>
__global__ void __launch_bounds__(kThreads)
period_runstat_kernel(const float* __restrict__ x, int64_t T, int64_t C, int64_t ldx,
                      const int32_t* __restrict__ poff, float thr, int32_t window,
                      float* __restrict__ out, int32_t* __restrict__ valid) {
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  const float* col = x + c0;

  int32_t cur[VEC], nv[VEC];
  bool skip[VEC];
  RunAcc<RED> acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { cur[i] = 0; nv[i] = 0; skip[i] = false; acc[i].init(); }

  if constexpr (AFTER) {
    // A run already under way at the period start belongs to an earlier period
    // (indices/run_length.py:329-334: run lengths sit on the run's FIRST element).
    if (t0 > 0) {
      Vec<VEC> r;
      r.load(col + (int64_t)(t0 - 1) * ldx);
#pragma unroll
      for (int i = 0; i < VEC; ++i) skip[i] = cmp<OP>(r.v[i], thr);
    }
  }

  auto step = [&](const Vec<VEC>& r) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bool c = cmp<OP>(r.v[i], thr);
      if constexpr (VALID) nv[i] += (r.v[i] == r.v[i]) ? 1 : 0;
      if constexpr (AFTER) {
        skip[i] = skip[i] && c;
        c = c && !skip[i];
      }
      if constexpr (FASTMAX) {
        cur[i] = c ? cur[i] + 1 : 0;
        acc[i].a = max(acc[i].a, cur[i]);
      } else {
        const int32_t L = cur[i];
        acc[i].add(L, !c && L >= window);
        cur[i] = c ? L + 1 : 0;
      }
    }
  };
  stream_rows<VEC>(col, ldx, t0, t1, step);

  if constexpr (AFTER) {
    // Runs still open at the period end keep their FULL length: read on until they all close.
    // (VALID counts only the period's own steps.)
    int t = t1;
    bool open = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) open = open || (cur[i] > 0);
    const float* pp = col + (int64_t)t1 * ldx;
    while (open && t < (int)T) {
      Vec<VEC> r;
      r.load(pp);
      open = false;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        // a cell whose run has closed must not start a new one here (it belongs to a later period)
        bool c = (cur[i] > 0) && cmp<OP>(r.v[i], thr);
        if constexpr (FASTMAX) {
          cur[i] = c ? cur[i] + 1 : 0;
          acc[i].a = max(acc[i].a, cur[i]);
        } else {
          const int32_t L = cur[i];
          acc[i].add(L, !c && L >= window);
          cur[i] = c ? L + 1 : 0;
        }
        open = open || c;
      }
      pp += ldx;
      ++t;
    }
  }
  float res[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if constexpr (!FASTMAX) acc[i].add(cur[i], cur[i] >= window);  // run closed by the period/series end
    res[i] = acc[i].result();
  }
  store_vec<VEC>(out + (int64_t)p * C + c0, res);
  if constexpr (VALID) store_vec<VEC>(valid + (int64_t)p * C + c0, nv);
}

// ------------------------------------------------------------------------------------------------
// windowed_max_run_sum of the excess over a threshold (indices/run_length.py:491-540 applied to
// `(x - thr).clip(0)`, indices/_threshold.py:2064-2073 `hot_spell_max_magnitude`): per period the
// largest sum of (x - thr) over a run of x > thr (x < thr: thr - x) at least `window` long, 0 if none.
// The excess is formed in float32 like the reference's `tasmax - thresh`; run sums are float64.
// ------------------------------------------------------------------------------------------------
template <int OP, int VEC, bool AFTER>
__global__ void __launch_bounds__(kThreads)
period_run_maxsum_kernel(const float* __restrict__ x, int64_t T, int64_t C, int64_t ldx,
                         const int32_t* __restrict__ poff, float thr, int32_t window, float* __restrict__ out) {
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  const float* col = x + c0;
  int32_t cur[VEC];
  double rs[VEC], best[VEC];
  bool skip[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { cur[i] = 0; rs[i] = 0.0; best[i] = 0.0; skip[i] = false; }
  auto excess = [&](float v) -> float {
    const float d = (OP == XC_OP_GT || OP == XC_OP_GE) ? (v - thr) : (thr - v);
    return (d > 0.f) ? d : 0.f;  // NaN -> 0: not part of a run (rle: da > 0)
  };
  if constexpr (AFTER) {
    if (t0 > 0) {
      Vec<VEC> r;
      r.load(col + (int64_t)(t0 - 1) * ldx);
#pragma unroll
      for (int i = 0; i < VEC; ++i) skip[i] = excess(r.v[i]) > 0.f;
    }
  }
  auto step = [&](const Vec<VEC>& r, bool only_open) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float e = excess(r.v[i]);
      bool in = e > 0.f;
      if constexpr (AFTER) {
        if (!only_open) {
          skip[i] = skip[i] && in;
          in = in && !skip[i];
        } else {
          in = in && (cur[i] > 0);
        }
      }
      if (in) {
        cur[i] += 1;
        rs[i] += (double)e;
      } else {
        if (cur[i] >= window) best[i] = fmax(best[i], rs[i]);
        cur[i] = 0;
        rs[i] = 0.0;
      }
    }
  };
  stream_rows<VEC>(col, ldx, t0, t1, [&](const Vec<VEC>& r) { step(r, false); });
  if constexpr (AFTER) {
    int t = t1;
    bool open = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) open = open || (cur[i] > 0);
    while (open && t < (int)T) {
      Vec<VEC> r;
      r.load(col + (int64_t)t * ldx);
      step(r, true);
      open = false;
#pragma unroll
      for (int i = 0; i < VEC; ++i) open = open || (cur[i] > 0);
      ++t;
    }
  }
  float res[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (cur[i] >= window) best[i] = fmax(best[i], rs[i]);
    res[i] = (float)best[i];
  }
  store_vec<VEC>(out + (int64_t)p * C + c0, res);
}

// ------------------------------------------------------------------------------------------------
// per-period reductions with optional fused transform
// ------------------------------------------------------------------------------------------------
template <int STAT, int TF, int OP, int VEC, bool VALID>
__global__ void __launch_bounds__(kThreads)
period_reduce_kernel(const float* __restrict__ x, int64_t C, int64_t ldx,
                     const int32_t* __restrict__ poff, float thr,
                     float* __restrict__ out, int32_t* __restrict__ valid) {
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  double s[VEC], q[VEC];
  float m[VEC];
  int32_t n[VEC], nv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    s[i] = 0.0; q[i] = 0.0; n[i] = 0; nv[i] = 0;
    m[i] = (STAT == XC_STAT_MIN) ? INFINITY : -INFINITY;
  }
  stream_rows<VEC>(x + c0, ldx, t0, t1, [&](const Vec<VEC>& r) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float v = r.v[i];
      if constexpr (VALID) nv[i] += (v == v) ? 1 : 0;
      if constexpr (TF == XC_TF_EXCESS) {
        // (x - t).clip(0) / (t - x).clip(0) in float32 (indices/generic.py:1545-1549)
        float d = (OP == XC_OP_GT || OP == XC_OP_GE) ? (v - thr) : (thr - v);
        v = (d != d) ? d : fmaxf(d, 0.f);
      }
      if constexpr (TF == XC_TF_WHERE) v = cmp<OP>(v, thr) ? v : NAN;
      const bool ok = (v == v);
      n[i] += ok ? 1 : 0;
      if constexpr (STAT == XC_STAT_SUM || STAT == XC_STAT_MEAN || STAT == XC_STAT_STD ||
                    STAT == XC_STAT_VAR)
        s[i] += ok ? (double)v : 0.0;
      if constexpr (STAT == XC_STAT_STD || STAT == XC_STAT_VAR) q[i] += ok ? (double)v * (double)v : 0.0;
      if constexpr (STAT == XC_STAT_MIN) m[i] = fminf(m[i], v);  // fminf ignores NaN
      if constexpr (STAT == XC_STAT_MAX) m[i] = fmaxf(m[i], v);
    }
  });
  float res[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const double nn = (double)n[i];
    if constexpr (STAT == XC_STAT_SUM) res[i] = (float)s[i];
    if constexpr (STAT == XC_STAT_COUNT) res[i] = (float)n[i];
    if constexpr (STAT == XC_STAT_MEAN) res[i] = n[i] ? (float)(s[i] / nn) : NAN;
    if constexpr (STAT == XC_STAT_MIN || STAT == XC_STAT_MAX) res[i] = n[i] ? m[i] : NAN;
    if constexpr (STAT == XC_STAT_STD || STAT == XC_STAT_VAR) {
      if (n[i] == 0) {
        res[i] = NAN;
      } else {
        double mean = s[i] / nn;
        double var = q[i] / nn - mean * mean;
        var = var > 0.0 ? var : 0.0;
        res[i] = (STAT == XC_STAT_STD) ? (float)sqrt(var) : (float)var;
      }
    }
  }
  store_vec<VEC>(out + (int64_t)p * C + c0, res);
  if constexpr (VALID) store_vec<VEC>(valid + (int64_t)p * C + c0, nv);
}

inline bool can_vec4(const void* x, int64_t C, int64_t ldx, const void* o1, const void* o2) {
  return (C % 4 == 0) && (ldx % 4 == 0) && aligned16(x) && aligned16(o1) && (o2 == nullptr || aligned16(o2));
}

inline dim3 grid_for(int64_t C, int vec, int32_t P) {
  int64_t nvec = (C + vec - 1) / vec;
  return dim3((unsigned)((nvec + kThreads - 1) / kThreads), (unsigned)P, 1);
}

int32_t check_common(const void* x, int64_t T, int64_t C, int64_t ldx, const void* poff, int32_t P,
                     const void* out) {
  XC_REQUIRE(x != nullptr && out != nullptr && poff != nullptr, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C, "bad shape: T=%lld C=%lld ldx=%lld", (long long)T, (long long)C,
             (long long)ldx);
  XC_REQUIRE(T < 2147483647LL, "time axis too long");
  XC_REQUIRE(P > 0 && P <= 65535, "number of periods must be in [1, 65535], got %d", P);
  return XC_OK;
}


// ------------------------------------------------------------------------------------------------
// threshold count against an ARRAY threshold (indices/generic.py:301-361 with a DataArray threshold:
// numpy compares float32 data with a float64 array in float64); thr_tstride = 0 for a
// per-cell threshold (lat, lon), C for one that also varies in time.
// ------------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(kThreads)
period_count_arr_kernel(const float* __restrict__ x, int64_t C, int64_t ldx, const int32_t* __restrict__ poff,
                        const double* __restrict__ thr, int64_t thr_tstride, int32_t* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  int32_t n = 0;
  const float* px = x + (int64_t)t0 * ldx + c;
  const double* pt = thr + (int64_t)t0 * thr_tstride + c;
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
    n += cmpd<OP>((double)ld_stream(px), *pt) ? 1 : 0;
    px += ldx;
    pt += thr_tstride;
  }
  out[(int64_t)p * C + c] = n;
}


// ------------------------------------------------------------------------------------------------
// fused multi-output pass: every count / run statistic / reduction that shares (x, periods)
// ------------------------------------------------------------------------------------------------
// One thread = 4 adjacent cells x one period, like the single-output kernels; the state of NL "lite"
// conditions (count + longest run), NF "full" conditions (+ two windowed run sums / counts + largest run
// sum of an excess) and NS conditional sums lives in registers next to the plain statistics.  The
// template arguments are the (rounded-up) numbers of each kind: unused entries are computed and dropped.
template <int NL, int NF, int NS>
__global__ void __launch_bounds__(kThreads)
period_multi_kernel(const float* __restrict__ x, int64_t C, int64_t ldx, const int32_t* __restrict__ poff,
                    const XcMultiPlan plan, float* __restrict__ out, int64_t slot_stride) {
  constexpr int VEC = 4;
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], t1 = poff[p + 1];
  double s[VEC];
  float mn[VEC], mx[VEC];
  int32_t nok[VEC];
  int32_t l_len[NL > 0 ? NL : 1][VEC], l_n[NL > 0 ? NL : 1][VEC], l_mx[NL > 0 ? NL : 1][VEC];
  int32_t f_len[NF > 0 ? NF : 1][VEC], f_n[NF > 0 ? NF : 1][VEC], f_mx[NF > 0 ? NF : 1][VEC];
  int32_t f_sa[NF > 0 ? NF : 1][VEC], f_ca[NF > 0 ? NF : 1][VEC], f_sb[NF > 0 ? NF : 1][VEC], f_cb[NF > 0 ? NF : 1][VEC];
  double f_rs[NF > 0 ? NF : 1][VEC], f_best[NF > 0 ? NF : 1][VEC];
  double q[NS > 0 ? NS : 1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    s[i] = 0.0; mn[i] = INFINITY; mx[i] = -INFINITY; nok[i] = 0;
#pragma unroll
    for (int j = 0; j < NL; ++j) { l_len[j][i] = 0; l_n[j][i] = 0; l_mx[j][i] = 0; }
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      f_len[j][i] = 0; f_n[j][i] = 0; f_mx[j][i] = 0; f_sa[j][i] = 0; f_ca[j][i] = 0; f_sb[j][i] = 0; f_cb[j][i] = 0;
      f_rs[j][i] = 0.0; f_best[j][i] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) q[j][i] = 0.0;
  }
  auto close_run = [&](int j, int i, int32_t L, double rs) {   // a run of L steps of full condition j has ended
    f_sa[j][i] += (L >= plan.full[j].wa) ? L : 0;
    f_ca[j][i] += (L >= plan.full[j].wa) ? 1 : 0;
    f_sb[j][i] += (L >= plan.full[j].wb) ? L : 0;
    f_cb[j][i] += (L >= plan.full[j].wb) ? 1 : 0;
    f_best[j][i] = (L >= plan.full[j].wms) ? fmax(f_best[j][i], rs) : f_best[j][i];
  };
  stream_rows<VEC>(x + c0, ldx, t0, t1, [&](const Vec<VEC>& r) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float v = r.v[i];
      const bool ok = (v == v);
      nok[i] += ok ? 1 : 0;
      s[i] += ok ? (double)v : 0.0;
      mn[i] = fminf(mn[i], v);
      mx[i] = fmaxf(mx[i], v);
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        const bool c = plan.lite[j].sgn * v > plan.lite[j].thr;
        l_len[j][i] = c ? l_len[j][i] + 1 : 0;
        l_n[j][i] += c ? 1 : 0;
        l_mx[j][i] = max(l_mx[j][i], l_len[j][i]);
      }
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const bool c = plan.full[j].sgn * v > plan.full[j].thr;
        const int32_t L = f_len[j][i];
        // windows are >= 1, so a closing "run" of length 0 never qualifies
        close_run(j, i, c ? 0 : L, f_rs[j][i]);
        const float e = plan.full[j].ms_sgn * (v - plan.full[j].ms_thr0);   // float32 excess, like `tasmax - thresh`
        f_rs[j][i] = c ? f_rs[j][i] + (double)e : 0.0;
        f_len[j][i] = c ? L + 1 : 0;
        f_n[j][i] += c ? 1 : 0;
        f_mx[j][i] = max(f_mx[j][i], f_len[j][i]);
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (plan.sums[j].mode == 0) {
          const float d = plan.sums[j].off_sgn * (v - plan.sums[j].off);     // (x - t) / (t - x) in float32
          q[j][i] += ok ? (double)fmaxf(d, 0.f) : 0.0;
        } else {
          q[j][i] += (plan.sums[j].sgn * v > plan.sums[j].thr) ? (double)v : 0.0;
        }
      }
    }
  });
  auto put_f = [&](int slot, const float (&v)[VEC]) {
    if (slot >= 0) store_vec<VEC>(out + (int64_t)slot * slot_stride + (int64_t)p * C + c0, v);
  };
  auto put_i = [&](int slot, const int32_t (&v)[VEC]) {
    if (slot >= 0) store_vec<VEC>(reinterpret_cast<int32_t*>(out) + (int64_t)slot * slot_stride + (int64_t)p * C + c0, v);
  };
  float rf[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) rf[i] = (float)s[i];
  put_f(plan.slot_sum, rf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) rf[i] = nok[i] ? (float)(s[i] / (double)nok[i]) : NAN;
  put_f(plan.slot_mean, rf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) rf[i] = nok[i] ? mn[i] : NAN;
  put_f(plan.slot_min, rf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) rf[i] = nok[i] ? mx[i] : NAN;
  put_f(plan.slot_max, rf);
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    put_i(plan.lite[j].slot_n, l_n[j]);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (l_mx[j][i] >= plan.lite[j].wmax) ? (float)l_mx[j][i] : 0.f;
    put_f(plan.lite[j].slot_max, rf);
  }
#pragma unroll
  for (int j = 0; j < NF; ++j) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) close_run(j, i, f_len[j][i], f_rs[j][i]);     // runs closed by the period end
    put_i(plan.full[j].slot_n, f_n[j]);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (f_mx[j][i] >= plan.full[j].wmax) ? (float)f_mx[j][i] : 0.f;
    put_f(plan.full[j].slot_max, rf);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)f_sa[j][i];
    put_f(plan.full[j].slot_sum_a, rf);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)f_ca[j][i];
    put_f(plan.full[j].slot_cnt_a, rf);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)f_sb[j][i];
    put_f(plan.full[j].slot_sum_b, rf);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)f_cb[j][i];
    put_f(plan.full[j].slot_cnt_b, rf);
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)f_best[j][i];
    put_f(plan.full[j].slot_ms, rf);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) rf[i] = (float)q[j][i];
    put_f(plan.sums[j].slot, rf);
  }
}

template <int NL, int NF, int NS>
int32_t launch_multi(const float* x, int64_t C, int64_t ldx, const int32_t* poff, int32_t P, const XcMultiPlan& plan,
                     float* out, cudaStream_t st) {
  period_multi_kernel<NL, NF, NS><<<grid_for(C, 4, P), kThreads, 0, st>>>(x, C, ldx, poff, plan, out, (int64_t)P * C);
  return launch_status("period_multi_kernel");
}

}  // namespace
}  // namespace xc

using namespace xc;

extern "C" int32_t xc_period_count_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                       const int32_t* period_offsets, int32_t P, int32_t op, double thr,
                                       int32_t cmp_f64, int32_t* out_count, int32_t* valid_count,
                                       void* stream) {
  if (int32_t e = check_common(x, T, C, ldx, period_offsets, P, out_count)) return e;
  const float t32 = fold_threshold(op, thr, cmp_f64);
  const bool v4 = can_vec4(x, C, ldx, out_count, valid_count);
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_op_nan(op, [&](auto OPC) -> int32_t {
    constexpr int OP = decltype(OPC)::value;
    auto go = [&](auto VECC, auto VALC) -> int32_t {
      constexpr int VEC = decltype(VECC)::value;
      constexpr bool VAL = decltype(VALC)::value;
      period_count_kernel<OP, VEC, VAL><<<grid_for(C, VEC, P), kThreads, 0, st>>>(
          x, C, ldx, period_offsets, t32, out_count, valid_count);
      return launch_status("period_count_kernel");
    };
    if (v4) return valid_count ? go(std::integral_constant<int, 4>{}, std::true_type{})
                               : go(std::integral_constant<int, 4>{}, std::false_type{});
    return valid_count ? go(std::integral_constant<int, 1>{}, std::true_type{})
                       : go(std::integral_constant<int, 1>{}, std::false_type{});
  });
}

namespace {
template <int OP, int RED, bool FASTMAX>
int32_t launch_runstat(const float* x, int64_t T, int64_t C, int64_t ldx, const int32_t* poff, int32_t P,
                       float t32, int32_t window, bool after, float* out, int32_t* valid, cudaStream_t st) {
  const bool v4 = can_vec4(x, C, ldx, out, valid);
  auto go = [&](auto VECC, auto VALC, auto AFTC) -> int32_t {
    constexpr int VEC = decltype(VECC)::value;
    constexpr bool VAL = decltype(VALC)::value;
    constexpr bool AFT = decltype(AFTC)::value;
    period_runstat_kernel<OP, RED, VEC, VAL, AFT, FASTMAX><<<grid_for(C, VEC, P), kThreads, 0, st>>>(
        x, T, C, ldx, poff, t32, window, out, valid);
    return launch_status("period_runstat_kernel");
  };
  auto go2 = [&](auto VECC) -> int32_t {
    if (valid) return after ? go(VECC, std::true_type{}, std::true_type{}) : go(VECC, std::true_type{}, std::false_type{});
    return after ? go(VECC, std::false_type{}, std::true_type{}) : go(VECC, std::false_type{}, std::false_type{});
  };
  return v4 ? go2(std::integral_constant<int, 4>{}) : go2(std::integral_constant<int, 1>{});
}
}  // namespace

extern "C" int32_t xc_period_runstat_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                         const int32_t* period_offsets, int32_t P, int32_t op, double thr,
                                         int32_t cmp_f64, int32_t reducer, int32_t window,
                                         int32_t resample_before_rl, float* out, int32_t* valid_count,
                                         void* stream) {
  if (int32_t e = check_common(x, T, C, ldx, period_offsets, P, out)) return e;
  XC_REQUIRE(window >= 1, "window must be >= 1, got %d", window);
  XC_REQUIRE(reducer >= XC_RL_MAX && reducer <= XC_RL_STD, "unknown run-length reducer %d", reducer);
  const float t32 = fold_threshold(op, thr, cmp_f64);
  const bool after = resample_before_rl == 0;
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_op_nan(op, [&](auto OPC) -> int32_t {
    constexpr int OP = decltype(OPC)::value;
#define XC_RS(RED, FAST) \
  return launch_runstat<OP, RED, FAST>(x, T, C, ldx, period_offsets, P, t32, window, after, out, valid_count, st)
    switch (reducer) {
      case XC_RL_MAX:
        if (window == 1) { XC_RS(XC_RL_MAX, true); }
        XC_RS(XC_RL_MAX, false);
      case XC_RL_MIN: XC_RS(XC_RL_MIN, false);
      case XC_RL_SUM: XC_RS(XC_RL_SUM, false);
      case XC_RL_COUNT: XC_RS(XC_RL_COUNT, false);
      case XC_RL_MEAN: XC_RS(XC_RL_MEAN, false);
      default: XC_RS(XC_RL_STD, false);
    }
#undef XC_RS
  });
}

namespace {
template <int STAT, int TF, int OP>
int32_t launch_reduce(const float* x, int64_t C, int64_t ldx, const int32_t* poff, int32_t P, float t32,
                      float* out, int32_t* valid, cudaStream_t st) {
  const bool v4 = can_vec4(x, C, ldx, out, valid);
  auto go = [&](auto VECC, auto VALC) -> int32_t {
    constexpr int VEC = decltype(VECC)::value;
    constexpr bool VAL = decltype(VALC)::value;
    period_reduce_kernel<STAT, TF, OP, VEC, VAL><<<grid_for(C, VEC, P), kThreads, 0, st>>>(x, C, ldx, poff, t32,
                                                                                         out, valid);
    return launch_status("period_reduce_kernel");
  };
  if (v4) return valid ? go(std::integral_constant<int, 4>{}, std::true_type{})
                       : go(std::integral_constant<int, 4>{}, std::false_type{});
  return valid ? go(std::integral_constant<int, 1>{}, std::true_type{})
               : go(std::integral_constant<int, 1>{}, std::false_type{});
}

template <int TF, int OP>
int32_t reduce_stat(int32_t stat, const float* x, int64_t C, int64_t ldx, const int32_t* poff, int32_t P, float t32,
                    float* out, int32_t* valid, cudaStream_t st) {
  switch (stat) {
    case XC_STAT_SUM: return launch_reduce<XC_STAT_SUM, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_MEAN: return launch_reduce<XC_STAT_MEAN, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_MIN: return launch_reduce<XC_STAT_MIN, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_MAX: return launch_reduce<XC_STAT_MAX, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_STD: return launch_reduce<XC_STAT_STD, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_VAR: return launch_reduce<XC_STAT_VAR, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
    case XC_STAT_COUNT: return launch_reduce<XC_STAT_COUNT, TF, OP>(x, C, ldx, poff, P, t32, out, valid, st);
  }
  set_error("unknown reduction %d", stat);
  return XC_ERR_INVALID;
}
}  // namespace

extern "C" int32_t xc_period_reduce_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                        const int32_t* period_offsets, int32_t P, int32_t stat,
                                        int32_t transform, int32_t op, double thr, float* out,
                                        int32_t* valid_count, void* stream) {
  if (int32_t e = check_common(x, T, C, ldx, period_offsets, P, out)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (transform == XC_TF_NONE)
    return reduce_stat<XC_TF_NONE, XC_OP_GT>(stat, x, C, ldx, period_offsets, P, 0.f, out, valid_count, st);
  // thresholds of the fused transforms are applied in float32 (float32 data op python float)
  const float t32 = (float)thr;
  if (transform == XC_TF_EXCESS) {
    XC_REQUIRE(op >= XC_OP_GT && op <= XC_OP_LE, "Operation `%d` not permitted for indice.", op);
    if (op == XC_OP_GT || op == XC_OP_GE)
      return reduce_stat<XC_TF_EXCESS, XC_OP_GT>(stat, x, C, ldx, period_offsets, P, t32, out, valid_count, st);
    return reduce_stat<XC_TF_EXCESS, XC_OP_LT>(stat, x, C, ldx, period_offsets, P, t32, out, valid_count, st);
  }
  if (transform == XC_TF_WHERE) {
    return dispatch_op(op, [&](auto OPC) -> int32_t {
      constexpr int OP = decltype(OPC)::value;
      return reduce_stat<XC_TF_WHERE, OP>(stat, x, C, ldx, period_offsets, P, t32, out, valid_count, st);
    });
  }
  set_error("unknown transform %d", transform);
  return XC_ERR_INVALID;
}

extern "C" int32_t xc_period_run_maxsum_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                            const int32_t* period_offsets, int32_t P, int32_t op, double thr,
                                            int32_t window, int32_t resample_before_rl, float* out, void* stream) {
  if (int32_t e = check_common(x, T, C, ldx, period_offsets, P, out)) return e;
  XC_REQUIRE(window >= 1, "window must be >= 1, got %d", window);
  XC_REQUIRE(op >= XC_OP_GT && op <= XC_OP_LE, "Operation `%d` not permitted for indice.", op);
  const float t32 = (float)thr;
  const bool after = resample_before_rl == 0;
  const bool v4 = can_vec4(x, C, ldx, out, nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto OPC, auto VECC, auto AFTC) -> int32_t {
    constexpr int OP = decltype(OPC)::value;
    constexpr int VEC = decltype(VECC)::value;
    constexpr bool AFT = decltype(AFTC)::value;
    period_run_maxsum_kernel<OP, VEC, AFT><<<grid_for(C, VEC, P), kThreads, 0, st>>>(x, T, C, ldx, period_offsets,
                                                                                   t32, window, out);
    return launch_status("period_run_maxsum_kernel");
  };
  auto go2 = [&](auto OPC) -> int32_t {
    if (v4) return after ? go(OPC, std::integral_constant<int, 4>{}, std::true_type{})
                         : go(OPC, std::integral_constant<int, 4>{}, std::false_type{});
    return after ? go(OPC, std::integral_constant<int, 1>{}, std::true_type{})
                 : go(OPC, std::integral_constant<int, 1>{}, std::false_type{});
  };
  return (op == XC_OP_GT || op == XC_OP_GE) ? go2(std::integral_constant<int, XC_OP_GT>{})
                                            : go2(std::integral_constant<int, XC_OP_LT>{});
}

extern "C" int32_t xc_period_multi_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                       const int32_t* period_offsets, int32_t P, const XcMultiPlan* plan_host,
                                       void* out, int32_t n_slots, void* stream) {
  int32_t e = check_common(x, T, C, ldx, period_offsets, P, out);
  if (e) return e;
  XC_REQUIRE(plan_host != nullptr, "null pointer argument");
  const XcMultiPlan& pl = *plan_host;
  XC_REQUIRE(pl.n_lite >= 0 && pl.n_lite <= 4 && pl.n_full >= 0 && pl.n_full <= 2 && pl.n_sums >= 0 && pl.n_sums <= 3,
             "plan holds at most 4 lite, 2 full and 3 sum conditions");
  XC_REQUIRE(can_vec4(x, C, ldx, out, nullptr), "xc_period_multi_f32 needs C, ldx multiples of 4 and 16-byte aligned buffers");
  auto slot_ok = [&](int32_t sl) { return sl >= -1 && sl < n_slots; };
  bool ok = slot_ok(pl.slot_sum) && slot_ok(pl.slot_mean) && slot_ok(pl.slot_min) && slot_ok(pl.slot_max);
  for (int j = 0; j < 4; ++j) ok = ok && slot_ok(pl.lite[j].slot_n) && slot_ok(pl.lite[j].slot_max);
  for (int j = 0; j < 2; ++j)
    ok = ok && slot_ok(pl.full[j].slot_n) && slot_ok(pl.full[j].slot_max) && slot_ok(pl.full[j].slot_sum_a) &&
         slot_ok(pl.full[j].slot_cnt_a) && slot_ok(pl.full[j].slot_sum_b) && slot_ok(pl.full[j].slot_cnt_b) &&
         slot_ok(pl.full[j].slot_ms) && pl.full[j].wa >= 1 && pl.full[j].wb >= 1 && pl.full[j].wms >= 1;
  for (int j = 0; j < 3; ++j) ok = ok && slot_ok(pl.sums[j].slot) && (pl.sums[j].mode == 0 || pl.sums[j].mode == 1);
  XC_REQUIRE(ok, "plan: output slot outside [-1, n_slots) or window < 1");
  cudaStream_t st = (cudaStream_t)stream;
  float* o = (float*)out;
  // rounded-up instantiations: unused entries of the plan must carry slot -1 (their work is dropped)
  const int nl = pl.n_lite == 0 ? 0 : (pl.n_lite <= 2 ? 2 : 4);
  const int nf = pl.n_full;
  const int ns = pl.n_sums == 0 ? 0 : 3;
#define XC_MULTI(L, F, S) \
  if (nl == L && nf == F && ns == S) return launch_multi<L, F, S>(x, C, ldx, period_offsets, P, pl, o, st)
  XC_MULTI(0, 0, 0); XC_MULTI(0, 0, 3); XC_MULTI(0, 1, 0); XC_MULTI(0, 1, 3); XC_MULTI(0, 2, 0); XC_MULTI(0, 2, 3);
  XC_MULTI(2, 0, 0); XC_MULTI(2, 0, 3); XC_MULTI(2, 1, 0); XC_MULTI(2, 1, 3); XC_MULTI(2, 2, 0); XC_MULTI(2, 2, 3);
  XC_MULTI(4, 0, 0); XC_MULTI(4, 0, 3); XC_MULTI(4, 1, 0); XC_MULTI(4, 1, 3); XC_MULTI(4, 2, 0); XC_MULTI(4, 2, 3);
#undef XC_MULTI
  set_error("plan shape not instantiated");
  return XC_ERR_UNSUPPORTED;
}

extern "C" int32_t xc_period_count_arr_f32(const float* x, int64_t T, int64_t C, int64_t ldx,
                                           const int32_t* period_offsets, int32_t P, int32_t op,
                                           const double* thr, int64_t thr_tstride, int32_t* out_count, void* stream) {
  int32_t e = check_common(x, T, C, ldx, period_offsets, P, out_count);
  if (e) return e;
  XC_REQUIRE(thr != nullptr, "null pointer argument");
  XC_REQUIRE(thr_tstride == 0 || thr_tstride >= C, "threshold time stride must be 0 (per cell) or >= C");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((C + kThreads - 1) / kThreads), (unsigned)P, 1);
  return dispatch_op(op, [&](auto OPC) -> int32_t {
    constexpr int OP = decltype(OPC)::value;
    period_count_arr_kernel<OP><<<grid, kThreads, 0, st>>>(x, C, ldx, period_offsets, thr, thr_tstride, out_count);
    return launch_status("period_count_arr_kernel");
  });
}
