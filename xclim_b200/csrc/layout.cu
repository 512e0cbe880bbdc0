// Layout changes of small per-cell tables (no arithmetic): the percentile kernels keep their tables
// doy-major -- (n_per, n_doy, C), coalesced along cells -- while the reference hands out
// (*space, dayofyear, percentiles) (core/calendar.py:479-483).  A shared-memory tiled transpose moves
// one to the other at copy speed so that the host never transposes gigabytes.
#include "common.cuh"

using namespace xc;

namespace {

constexpr int TILE = 32;

// in : rows r = p * n_doy + d (R = n_per * n_doy rows) of C doubles
// out: out[c * R + d * n_per + p]
__global__ void __launch_bounds__(TILE * 8) table_cell_major_kernel(const double* __restrict__ in, int n_per,
                                                                      int n_doy, int64_t C,
                                                                      double* __restrict__ out) {
  __shared__ double tile[TILE][TILE + 1];
  const int R = n_per * n_doy;
  const int64_t c0 = (int64_t)blockIdx.x * TILE;
  const int r0 = blockIdx.y * TILE;
  for (int j = threadIdx.y; j < TILE; j += 8) {
    const int r = r0 + j;
    const int64_t c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = in[(int64_t)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < TILE; j += 8) {
    const int64_t c = c0 + j;
    const int r = r0 + threadIdx.x;
    if (r < R && c < C) {
      const int p = r / n_doy, d = r - p * n_doy;
      out[c * R + (int64_t)d * n_per + p] = tile[threadIdx.x][j];
    }
  }
}

}  // namespace

extern "C" int32_t xc_table_cell_major_f64(const double* table, int32_t n_per, int32_t n_doy, int64_t C,
                                           double* out, void* stream) {
  XC_REQUIRE(table && out, "null pointer argument");
  XC_REQUIRE(n_per > 0 && n_doy > 0 && C > 0, "bad shape");
  const int R = n_per * n_doy;
  const int64_t gx = (C + TILE - 1) / TILE;
  XC_REQUIRE(gx <= 2147483647LL, "too many cells");
  dim3 grid((unsigned)gx, (unsigned)((R + TILE - 1) / TILE)), block(TILE, 8);
  table_cell_major_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(table, n_per, n_doy, C, out);
  return launch_status("table_cell_major_kernel");
}
