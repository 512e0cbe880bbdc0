// Stateless synthetic inputs for the BASELINE.json configurations.
// value(t, global_cell) is a pure function of (seed, t, global_cell): any lat tile generated on
// any rank equals the same slab of the global grid.  No reference counterpart (the reference's
// test data come from files / `xclim.testing.helpers.test_timeseries`).
#include "common.cuh"

namespace xc {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
  return mix64(mix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ (b * 0xA24BAED4963EE407ull));
}
__device__ __forceinline__ float u01(uint32_t bits) {  // (0, 1)
  return ((float)(bits >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
// N(0,1) from one 64-bit hash (Box-Muller)
__device__ __forceinline__ float normal(uint64_t h) {
  float u1 = u01((uint32_t)h), u2 = u01((uint32_t)(h >> 32));
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

template <int KIND>
__global__ void __launch_bounds__(256)
synth_kernel(float* __restrict__ out, int64_t T, int64_t C, int64_t ldx, int64_t cell_offset,
             int64_t cells_per_lat, int64_t n_lat_global, int32_t year_len, uint64_t seed) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t gc = c + cell_offset;
  const int rows_per_block = (int)((T + gridDim.y - 1) / gridDim.y);
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t t1 = min(T, t0 + rows_per_block);
  // per-cell NaN block: 1e-4 of the cells get a 10-step NaN block at a hashed position
  const uint64_t hc = hash3(seed + 7777, (uint64_t)gc, 0);
  const bool has_nan = (hc % 10000ull) == 0ull;
  const int64_t nan0 = has_nan ? (int64_t)((hc >> 20) % (uint64_t)(T > 10 ? T - 10 : 1)) : -1;
  float lat_term = 0.f;
  if (KIND == 1) {
    const int64_t j = gc / cells_per_lat;
    const float lat = -90.f + 180.f * (float)j / (float)(n_lat_global > 1 ? n_lat_global - 1 : 1);
    lat_term = 30.f * fabsf(__sinf(lat * 0.01745329252f));
  }
  for (int64_t t = t0; t < t1; ++t) {
    float v;
    if (KIND == 0) {
      // 10-day regimes: dry regime -> P(dry day) = 0.8, wet regime -> 0.3; wet amounts Exp(mean 6 mm/d)
      const uint64_t hr = hash3(seed + 100, (uint64_t)gc, (uint64_t)(t / 10));
      const float pdry = (hr & 1ull) ? 0.8f : 0.3f;
      const uint64_t h = hash3(seed, (uint64_t)gc, (uint64_t)t);
      const float u = u01((uint32_t)h);
      v = (u < pdry) ? 0.f : -6.f * __logf(u01((uint32_t)(h >> 32)));
    } else {
      const int doy = (int)(t % year_len);
      const float season = 12.f * __sinf(6.28318530718f * (float)(doy - 109) / (float)year_len);
      const float daily = 3.f * normal(hash3(seed, (uint64_t)gc, (uint64_t)t));
      const float weekly = 2.f * normal(hash3(seed + 100, (uint64_t)gc, (uint64_t)(t / 7)));
      v = 288.f - lat_term + season + daily + weekly;
    }
    if (has_nan && t >= nan0 && t < nan0 + 10) v = NAN;
    out[t * ldx + c] = v;
  }
}

}  // namespace
}  // namespace xc

using namespace xc;

extern "C" int32_t xc_synth_f32(float* out, int64_t T, int64_t C, int64_t ldx, int64_t cell_offset,
                                int64_t cells_per_lat, int64_t n_lat_global, int32_t year_len, int32_t kind,
                                uint64_t seed, void* stream) {
  XC_REQUIRE(out != nullptr, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C && cells_per_lat > 0 && year_len > 0, "bad shape");
  XC_REQUIRE(kind == 0 || kind == 1, "unknown synthetic kind %d", kind);
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)(T < 64 ? T : 64), 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == 0)
    synth_kernel<0><<<grid, 256, 0, st>>>(out, T, C, ldx, cell_offset, cells_per_lat, n_lat_global, year_len, seed);
  else
    synth_kernel<1><<<grid, 256, 0, st>>>(out, T, C, ldx, cell_offset, cells_per_lat, n_lat_global, year_len, seed);
  return launch_status("synth_kernel");
}

// ---------------------------------------------------------------------------------------------
// select_time as a step mask (core/calendar.py:1259-1376 with drop=False: `da.where(mask)`):
// out[t, c] = keep[t] ? x[t, c] : NaN.  One streaming pass; the kernels downstream already treat NaN
// as "not selected" (NaN compares False, reductions skip NaN).
// ---------------------------------------------------------------------------------------------
namespace xc {
namespace {
__global__ void __launch_bounds__(256)
mask_steps_kernel(const float* __restrict__ x, int64_t T, int64_t C, int64_t ldx, const uint8_t* __restrict__ keep,
                  float* __restrict__ out) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  const int rows = (int)((T + gridDim.y - 1) / gridDim.y);
  const int64_t t0 = (int64_t)blockIdx.y * rows, t1 = min(T, t0 + rows);
  const bool v4 = (c + 3 < C) && ((ldx & 3) == 0) && ((C & 3) == 0) &&
                  ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  for (int64_t t = t0; t < t1; ++t) {
    const bool k = keep[t] != 0;
    if (v4) {
      float4 v = make_float4(NAN, NAN, NAN, NAN);
      if (k) v = ld_stream4(x + t * ldx + c);
      *reinterpret_cast<float4*>(out + t * C + c) = v;
    } else {
      for (int i = 0; i < 4 && c + i < C; ++i) out[t * C + c + i] = k ? x[t * ldx + c + i] : NAN;
    }
  }
}
}  // namespace
}  // namespace xc

extern "C" int32_t xc_mask_steps_f32(const float* x, int64_t T, int64_t C, int64_t ldx, const uint8_t* keep,
                                     float* out, void* stream) {
  XC_REQUIRE(x && keep && out, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && ldx >= C, "bad shape");
  dim3 grid((unsigned)(((C + 3) / 4 + 255) / 256), (unsigned)(T < 128 ? T : 128), 1);
  mask_steps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, T, C, ldx, keep, out);
  return launch_status("mask_steps_kernel");
}
