// Run-length / count statistics of a condition on TWO variables.
//
// Replaces indices/_multivariate.py:646-880 (`heat_wave_frequency` / `_max_length` / `_total_length`:
// cond = (tasmin op t1) & (tasmax op t2) -> rl.resample_and_rl(<run statistic>)), :1653-1716
// (`tx_tn_days_above`) and indices/generic.py:1002-1073 (`bivariate_count_occurrences`, var_reducer
// all / any).  Same lane-owns-4-cells streaming as period_stats.cu with two 128-bit loads per step;
// operators and the reducer are runtime switches (the kernel moves 8 B per element and is HBM-bound).
#include "common.cuh"

namespace xc {
namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 4;

// Runtime operator without a branch: (x op t) == (((x<t)&L) | ((x==t)&E) | ((x>t)&G)) != N with the
// masks below (NaN makes the three tests false, so only != is true on NaN, as in numpy).
struct OpMask {
  bool L, E, G, N;
};
inline OpMask op_mask(int op) {
  switch (op) {
    case XC_OP_GT: return {false, false, true, false};
    case XC_OP_LT: return {true, false, false, false};
    case XC_OP_GE: return {false, true, true, false};
    case XC_OP_LE: return {true, true, false, false};
    case XC_OP_EQ: return {false, true, false, false};
    default: return {false, true, false, true};
  }
}
__device__ __forceinline__ bool cmp_rt(const OpMask& m, float x, float t) {
  const bool r = ((x < t) & m.L) | ((x == t) & m.E) | ((x > t) & m.G);
  return r != m.N;
}

struct Acc {
  int cur, mx, mn, sum, cnt;
  unsigned long long sq;
  bool skip;
};

// RED: XC_RL_MAX / SUM / COUNT get branch-free single-accumulator updates (the heat-wave indices);
// any other value keeps every accumulator and switches on `reducer` at the end.
// GTGT: both operators are ">" (the heat-wave indices and tx_tn_days_above): two compares and one
// predicate op per element pair instead of the six + six of the operator masks.
template <int VEC, int RED, bool GTGT>
__global__ void __launch_bounds__(kThreads)
period_runstat2_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int64_t T, int64_t C, int64_t ldx,
                       const int32_t* __restrict__ poff, OpMask op1, float t1, OpMask op2, float t2, int any,
                       int reducer, int window, int after, float* __restrict__ out) {
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * VEC;
  if (c0 >= C) return;
  const int p = blockIdx.y;
  const int t0 = poff[p], tend = poff[p + 1];
  Acc a[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] = Acc{0, 0, 0x7fffffff, 0, 0, 0ull, false};
  auto cond = [&](float u, float v) -> bool {
    bool c1, c2;
    if constexpr (GTGT) {
      c1 = u > t1;
      c2 = v > t2;
    } else {
      c1 = cmp_rt(op1, u, t1);
      c2 = cmp_rt(op2, v, t2);
    }
    return any ? (c1 | c2) : (c1 & c2);
  };
  auto close_run = [&](Acc& s) {
    const int L = s.cur;
    const bool take = L >= window;
    if constexpr (RED == XC_RL_MAX) {
      s.mx = take ? max(s.mx, L) : s.mx;
    } else if constexpr (RED == XC_RL_SUM) {
      s.sum += take ? L : 0;
    } else if constexpr (RED == XC_RL_COUNT) {
      s.cnt += take ? 1 : 0;
    } else {
      if (take) {
        s.mx = max(s.mx, L);
        s.mn = min(s.mn, L);
        s.sum += L;
        s.cnt += 1;
        s.sq += (unsigned long long)L * (unsigned long long)L;
      }
    }
    s.cur = 0;
  };
  auto load = [&](const float* base, int t, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
      const float4 q = ld_stream4(base + (int64_t)t * ldx + c0);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      v[0] = ld_stream(base + (int64_t)t * ldx + c0);
    }
  };
  if (after && t0 > 0) {
    float u[VEC], v[VEC];
    load(x1, t0 - 1, u);
    load(x2, t0 - 1, v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) a[i].skip = cond(u[i], v[i]);
  }
  auto step = [&](const float (&u)[VEC], const float (&v)[VEC], bool only_open) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bool m = cond(u[i], v[i]);
      if (after) {
        if (only_open) {
          m = m && (a[i].cur > 0);
        } else {
          a[i].skip = a[i].skip && m;
          m = m && !a[i].skip;
        }
      }
      if constexpr (RED == XC_RL_MAX || RED == XC_RL_SUM || RED == XC_RL_COUNT) {
        // branch-free: closing an empty run (cur == 0 < window) is a no-op
        Acc closing = a[i];
        close_run(closing);
        const int next = a[i].cur + 1;
        a[i].mx = m ? a[i].mx : closing.mx;
        a[i].sum = m ? a[i].sum : closing.sum;
        a[i].cnt = m ? a[i].cnt : closing.cnt;
        a[i].cur = m ? next : 0;
      } else {
        if (m) a[i].cur += 1; else close_run(a[i]);
      }
    }
  };
  int t = t0;
  for (; t + kUnroll <= tend; t += kUnroll) {
    float u[kUnroll][VEC], v[kUnroll][VEC];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) { load(x1, t + k, u[k]); load(x2, t + k, v[k]); }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) step(u[k], v[k], false);
  }
  for (; t < tend; ++t) {
    float u[VEC], v[VEC];
    load(x1, t, u);
    load(x2, t, v);
    step(u, v, false);
  }
  if (after) {
    bool open = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) open = open || (a[i].cur > 0);
    while (open && t < (int)T) {
      float u[VEC], v[VEC];
      load(x1, t, u);
      load(x2, t, v);
      step(u, v, true);
      open = false;
#pragma unroll
      for (int i = 0; i < VEC; ++i) open = open || (a[i].cur > 0);
      ++t;
    }
  }
  float res[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    close_run(a[i]);
    const Acc& s = a[i];
    switch (reducer) {
      case XC_RL_MAX: res[i] = (float)s.mx; break;
      case XC_RL_MIN: res[i] = s.cnt ? (float)s.mn : 0.f; break;
      case XC_RL_SUM: res[i] = (float)s.sum; break;
      case XC_RL_COUNT: res[i] = (float)s.cnt; break;
      case XC_RL_MEAN: res[i] = s.cnt ? (float)((double)s.sum / (double)s.cnt) : 0.f; break;
      default: {
        if (!s.cnt) { res[i] = 0.f; break; }
        const double n = (double)s.cnt, mean = (double)s.sum / n;
        const double var = (double)s.sq / n - mean * mean;
        res[i] = (float)sqrt(var > 0.0 ? var : 0.0);
      }
    }
  }
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(out + (int64_t)p * C + c0) = make_float4(res[0], res[1], res[2], res[3]);
  } else {
    out[(int64_t)p * C + c0] = res[0];
  }
}

}  // namespace
}  // namespace xc

using namespace xc;

extern "C" int32_t xc_period_runstat2_f32(const float* x1, const float* x2, int64_t T, int64_t C, int64_t ldx,
                                          const int32_t* period_offsets, int32_t P, int32_t op1, double thr1,
                                          int32_t op2, double thr2, int32_t var_any, int32_t reducer,
                                          int32_t window, int32_t resample_before_rl, float* out, void* stream) {
  XC_REQUIRE(x1 && x2 && period_offsets && out, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C && P > 0 && P <= 65535 && T < 2147483647LL, "bad shape");
  XC_REQUIRE(window >= 1, "window must be >= 1");
  XC_REQUIRE(op1 >= XC_OP_GT && op1 <= XC_OP_NE && op2 >= XC_OP_GT && op2 <= XC_OP_NE, "Operation not recognized.");
  XC_REQUIRE(reducer >= XC_RL_MAX && reducer <= XC_RL_STD, "unknown run-length reducer %d", reducer);
  const bool v4 = (C % 4 == 0) && (ldx % 4 == 0) && aligned16(x1) && aligned16(x2) && aligned16(out);
  const int vec = v4 ? 4 : 1;
  dim3 grid((unsigned)(((C + vec - 1) / vec + kThreads - 1) / kThreads), (unsigned)P, 1);
  cudaStream_t st = (cudaStream_t)stream;
  // thresholds are Python floats in the reference: compared in float32
  const OpMask m1 = op_mask(op1), m2 = op_mask(op2);
  const int anyf = var_any ? 1 : 0, afterf = resample_before_rl ? 0 : 1;
  const bool gtgt = (op1 == XC_OP_GT) && (op2 == XC_OP_GT);
#define XC_LAUNCH2(VEC, RED)                                                                                       \
  do {                                                                                                             \
    if (gtgt)                                                                                                      \
      period_runstat2_kernel<VEC, RED, true><<<grid, kThreads, 0, st>>>(                                           \
          x1, x2, T, C, ldx, period_offsets, m1, (float)thr1, m2, (float)thr2, anyf, reducer, window, afterf, out); \
    else                                                                                                           \
      period_runstat2_kernel<VEC, RED, false><<<grid, kThreads, 0, st>>>(                                          \
          x1, x2, T, C, ldx, period_offsets, m1, (float)thr1, m2, (float)thr2, anyf, reducer, window, afterf, out); \
  } while (0)
  if (v4) {
    switch (reducer) {
      case XC_RL_MAX: XC_LAUNCH2(4, XC_RL_MAX); break;
      case XC_RL_SUM: XC_LAUNCH2(4, XC_RL_SUM); break;
      case XC_RL_COUNT: XC_LAUNCH2(4, XC_RL_COUNT); break;
      default: XC_LAUNCH2(4, XC_RL_STD);
    }
  } else {
    XC_LAUNCH2(1, XC_RL_STD);
  }
#undef XC_LAUNCH2
  return launch_status("period_runstat2_kernel");
}
