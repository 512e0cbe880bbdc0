// Host-buffer entry point: the end-to-end path a caller with data in host memory takes.
//
// The reference hands numpy/dask arrays in host memory to its index functions
// (core/indicator.py:884-886 -> indices/_threshold.py:2927-2937).  Here the (time, cell) float32
// host buffer is streamed through HBM one period slab at a time: slab p+1 is copied (H2D, copy
// engine) while the run-length kernel works on slab p (double buffering on two streams), and the
// small (P, C) outputs are copied back at the end.  With resample_before_rl semantics every period
// is independent, so each slab is a complete unit of work.
#include <vector>

#include "common.cuh"

using namespace xc;

extern "C" int64_t xc_host_stream_workspace_bytes(int64_t T, int64_t C, const int32_t* period_offsets_host,
                                                  int32_t P) {
  if (!period_offsets_host || P <= 0 || C <= 0) return -1;
  (void)T;
  int64_t max_rows = 0;
  for (int p = 0; p < P; ++p) {
    const int64_t r = (int64_t)period_offsets_host[p + 1] - period_offsets_host[p];
    if (r > max_rows) max_rows = r;
  }
  const int64_t slab = ((max_rows * C * 4 + 255) / 256) * 256;
  const int64_t outs = (((int64_t)P * C * 4 + 255) / 256) * 256;
  const int64_t tbl = (((int64_t)P * 8 + 255) / 256) * 256;  // one {0, rows} table per period
  return 2 * slab + 2 * outs + tbl;
}

extern "C" int32_t xc_period_runstat_f32_host(const float* x_host, int64_t T, int64_t C,
                                              const int32_t* period_offsets_host, int32_t P, int32_t op,
                                              double thr, int32_t cmp_f64, int32_t reducer, int32_t window,
                                              float* out_host, int32_t* valid_count_host, void* workspace,
                                              int64_t workspace_bytes) {
  XC_REQUIRE(x_host && period_offsets_host && out_host && workspace, "null pointer argument");
  XC_REQUIRE(T > 0 && C > 0 && P > 0, "bad shape");
  const int64_t need = xc_host_stream_workspace_bytes(T, C, period_offsets_host, P);
  XC_REQUIRE(workspace_bytes >= need, "workspace too small: need %lld bytes", (long long)need);
  int64_t max_rows = 0;
  for (int p = 0; p < P; ++p) {
    const int64_t r = (int64_t)period_offsets_host[p + 1] - period_offsets_host[p];
    XC_REQUIRE(r > 0, "empty period %d", p);
    if (r > max_rows) max_rows = r;
  }
  XC_REQUIRE(period_offsets_host[0] >= 0 && period_offsets_host[P] <= T, "period offsets outside the series");
  const int64_t slab = ((max_rows * C * 4 + 255) / 256) * 256;
  const int64_t outs = (((int64_t)P * C * 4 + 255) / 256) * 256;
  char* ws = (char*)workspace;
  float* buf[2] = {(float*)ws, (float*)(ws + slab)};
  float* out_d = (float*)(ws + 2 * slab);
  int32_t* valid_d = (int32_t*)(ws + 2 * slab + outs);
  int32_t* poff_d = (int32_t*)(ws + 2 * slab + 2 * outs);

  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
  int32_t rc = XC_OK;
  auto cleanup = [&]() {
    for (int i = 0; i < 2; ++i) {
      if (copied[i]) cudaEventDestroy(copied[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
    if (s_copy) cudaStreamDestroy(s_copy);
    if (s_comp) cudaStreamDestroy(s_comp);
  };
#define XC_TRY(expr)                                \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      rc = cuda_fail(_e, #expr);                    \
      cudaDeviceSynchronize();                      \
      cleanup();                                    \
      return rc;                                    \
    }                                               \
  } while (0)
  XC_TRY(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
  XC_TRY(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    XC_TRY(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    XC_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  // one {0, rows} period table per slab (a slab is a one-period series for the kernel)
  std::vector<int32_t> tables(2 * (size_t)P);
  for (int p = 0; p < P; ++p) {
    tables[2 * p] = 0;
    tables[2 * p + 1] = period_offsets_host[p + 1] - period_offsets_host[p];
  }
  int32_t* tbl_d = poff_d;
  XC_TRY(cudaMemcpyAsync(tbl_d, tables.data(), tables.size() * 4, cudaMemcpyHostToDevice, s_comp));

  for (int p = 0; p < P; ++p) {
    const int b = p & 1;
    const int64_t rows = tables[2 * p + 1];
    if (p >= 2) XC_TRY(cudaStreamWaitEvent(s_copy, done[b], 0));  // kernel p-2 has released the buffer
    XC_TRY(cudaMemcpyAsync(buf[b], x_host + (int64_t)period_offsets_host[p] * C, (size_t)(rows * C * 4),
                           cudaMemcpyHostToDevice, s_copy));
    XC_TRY(cudaEventRecord(copied[b], s_copy));
    XC_TRY(cudaStreamWaitEvent(s_comp, copied[b], 0));
    const int32_t* tbl = tbl_d + 2 * p;
    int32_t e = xc_period_runstat_f32(buf[b], rows, C, C, tbl, 1, op, thr, cmp_f64, reducer, window, 1,
                                      out_d + (int64_t)p * C, valid_count_host ? valid_d + (int64_t)p * C : nullptr,
                                      (void*)s_comp);
    if (e != XC_OK) {
      cudaDeviceSynchronize();
      cleanup();
      return e;
    }
    XC_TRY(cudaEventRecord(done[b], s_comp));
  }
  XC_TRY(cudaMemcpyAsync(out_host, out_d, (size_t)P * C * 4, cudaMemcpyDeviceToHost, s_comp));
  if (valid_count_host)
    XC_TRY(cudaMemcpyAsync(valid_count_host, valid_d, (size_t)P * C * 4, cudaMemcpyDeviceToHost, s_comp));
  XC_TRY(cudaStreamSynchronize(s_comp));
  XC_TRY(cudaStreamSynchronize(s_copy));
#undef XC_TRY
  cleanup();
  return XC_OK;
}

extern "C" int32_t xc_copy_box_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                                     int64_t width_bytes, int64_t height, int32_t to_device, void* stream) {
  XC_REQUIRE(dst && src, "null pointer argument");
  XC_REQUIRE(width_bytes > 0 && height > 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes,
             "bad box: width %lld, height %lld, pitches %lld / %lld", (long long)width_bytes, (long long)height,
             (long long)dst_pitch, (long long)src_pitch);
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (dst_pitch == width_bytes && src_pitch == width_bytes) {
    XC_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)(width_bytes * height), kind, (cudaStream_t)stream));
  } else {
    XC_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes,
                                    (size_t)height, kind, (cudaStream_t)stream));
  }
  return XC_OK;
}

extern "C" int32_t xc_host_pinned(const void* host_ptr) {
  if (host_ptr == nullptr) return 0;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}
