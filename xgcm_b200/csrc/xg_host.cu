// xg_stencil2_host — the fused stencil on HOST buffers, streamed through the GPU.
//
// This is the call the reference-facing API makes for numpy-backed fields: the
// whole of xgcm/padding.py:575-616 + gridops.py + the metric passes for one axis,
// with host<->device copies inside.  The field is cut into slabs along dim 0
// (contiguous in host memory).  Three streams form a pipeline
//     H2D(slab s+1)  ||  kernel(slab s)  ||  D2H(slab s-1)
// so PCIe runs full duplex and the kernel time hides entirely behind the copies.
// When dim 0 is the operated axis the slabs overlap by the one-cell halo and the
// exterior halo plane (periodic wrap) is uploaded once.
//
// Workspace (device slabs + events + streams) is cached per device and reused;
// xg_host_workspace_release() frees it.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "xg_common.cuh"

namespace {

constexpr int kSlots = 3;

struct Workspace {
  int device = -1;
  std::mutex mu;  // one host call at a time per device; different devices run concurrently
  size_t slab_in_bytes = 0, slab_out_bytes = 0, metric_bytes[2] = {0, 0}, halo_bytes = 0;
  void* d_in[kSlots] = {nullptr, nullptr, nullptr};
  void* d_out[kSlots] = {nullptr, nullptr, nullptr};
  void* d_metric[2] = {nullptr, nullptr};
  void* d_halo[2] = {nullptr, nullptr};
  cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
  cudaEvent_t e_up[kSlots], e_done[kSlots], e_down[kSlots];
  bool events = false;
};

std::mutex g_ws_mutex;
std::vector<Workspace*> g_ws;

#define XG_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return xg_fail(XG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
  } while (0)

int ensure(void** p, size_t* have, size_t want) {
  if (*have >= want && *p) return XG_OK;
  if (*p) XG_CUDA(cudaFree(*p));
  *p = nullptr;
  *have = 0;
  if (want == 0) return XG_OK;
  XG_CUDA(cudaMalloc(p, want));
  *have = want;
  return XG_OK;
}

int get_workspace(int device, Workspace** out) {
  for (Workspace* w : g_ws)
    if (w->device == device) {
      *out = w;
      return XG_OK;
    }
  Workspace* w = new Workspace();
  w->device = device;
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_h2d, cudaStreamNonBlocking));
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_k, cudaStreamNonBlocking));
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    XG_CUDA(cudaEventCreateWithFlags(&w->e_up[i], cudaEventDisableTiming));
    XG_CUDA(cudaEventCreateWithFlags(&w->e_done[i], cudaEventDisableTiming));
    XG_CUDA(cudaEventCreateWithFlags(&w->e_down[i], cudaEventDisableTiming));
  }
  w->events = true;
  g_ws.push_back(w);
  *out = w;
  return XG_OK;
}

// bytes spanned by a broadcast operand laid out with `strides` over `shape`
size_t operand_span(const int64_t* strides, const int64_t* shape, int ndim, size_t es) {
  int64_t last = 0;
  for (int d = 0; d < ndim; ++d)
    if (shape[d] > 1) last += (shape[d] - 1) * strides[d];
  return (size_t)(last + 1) * es;
}

}  // namespace

void xg_host_pipe_release();  // xg_host_pipe.cu

extern "C" int xg_host_workspace_release(void) {
  xg_host_pipe_release();
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (Workspace* w : g_ws) {
    cudaSetDevice(w->device);
    for (int i = 0; i < kSlots; ++i) {
      if (w->d_in[i]) cudaFree(w->d_in[i]);
      if (w->d_out[i]) cudaFree(w->d_out[i]);
      if (w->events) {
        cudaEventDestroy(w->e_up[i]);
        cudaEventDestroy(w->e_done[i]);
        cudaEventDestroy(w->e_down[i]);
      }
    }
    for (int i = 0; i < 2; ++i) {
      if (w->d_metric[i]) cudaFree(w->d_metric[i]);
      if (w->d_halo[i]) cudaFree(w->d_halo[i]);
    }
    if (w->s_h2d) cudaStreamDestroy(w->s_h2d);
    if (w->s_k) cudaStreamDestroy(w->s_k);
    if (w->s_d2h) cudaStreamDestroy(w->s_d2h);
    delete w;
  }
  g_ws.clear();
  return XG_OK;
}

extern "C" int xg_stencil2_host(int op, int dtype, const void* in, void* out, int ndim,
                                const int64_t* shape, int axis, int lo, int hi, int bc,
                                double fill_value, const void* pre_metric,
                                const int64_t* pre_strides, const void* post_metric,
                                const int64_t* post_strides, int device) {
  if (!in || !out || !shape) return xg_fail(XG_EINVAL, "xg_stencil2_host: null pointer");
  if (dtype != XG_F32 && dtype != XG_F64)
    return xg_fail(XG_EINVAL, "xg_stencil2_host: dtype must be XG_F32 or XG_F64");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_stencil2_host: bad ndim");
  if (axis < 0 || axis >= ndim) return xg_fail(XG_EINVAL, "xg_stencil2_host: axis out of range");
  if (lo < 0 || lo > 1 || hi < 0 || hi > 1)
    return xg_fail(XG_EINVAL, "xg_stencil2_host: halo widths must be 0 or 1");
  if ((lo || hi) && (bc <= XG_BC_NONE || bc > XG_BC_EXTRAPOLATE))
    return xg_fail(XG_EINVAL,
                   "xg_stencil2_host: no boundary condition was specified but the operation "
                   "needs to pad the axis");
  const size_t es = dtype == XG_F32 ? 4 : 8;
  XG_CUDA(cudaSetDevice(device));
  Workspace* w = nullptr;
  int rc;
  {
    std::lock_guard<std::mutex> reg(g_ws_mutex);  // registry only
    rc = get_workspace(device, &w);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(w->mu);

  int64_t out_shape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) out_shape[d] = shape[d];
  out_shape[axis] = shape[axis] + lo + hi - 1;
  if (shape[axis] == 0) return xg_fail(XG_EINVAL, "xg_stencil2_host: empty operated axis");
  int64_t row_in = 1, row_out = 1;  // elements per index of dim 0
  for (int d = 1; d < ndim; ++d) {
    row_in *= shape[d];
    row_out *= out_shape[d];
  }
  const int64_t n0_out = out_shape[0];
  if (n0_out <= 0 || row_out == 0 || row_in == 0) return XG_OK;
  const bool ax0 = axis == 0;

  // slab height along dim 0: ~128 MiB of input per slab, at least 4 slabs if possible
  int64_t target_bytes = 128ll << 20;
  if (const char* env = getenv("XG_HOST_SLAB_MB")) {  // tuning knob (benchmarks only)
    const long mb = atol(env);
    if (mb >= 1 && mb <= 4096) target_bytes = (int64_t)mb << 20;
  }
  int64_t rows = target_bytes / (int64_t)(row_in * es);
  if (rows < 1) rows = 1;
  if (rows > (n0_out + 3) / 4) rows = (n0_out + 3) / 4;
  if (rows < 1) rows = 1;
  if (ax0 && bc == XG_BC_EXTRAPOLATE && (lo || hi)) {
    // the extrapolated halo is 2 A[edge] - A[next]: the slab that touches an edge of the axis must hold two
    // source planes, i.e. no one-row slab at either end (a one-row tail is merged by growing the slab height)
    if (rows < 2) rows = 2;
    while (rows < n0_out && n0_out % rows == 1) ++rows;
    if (rows > n0_out) rows = n0_out;
  }
  const int64_t nslab = xg_ceil_div(n0_out, rows);
  const int64_t in_rows_max = ax0 ? rows + 1 : rows;

  for (int i = 0; i < kSlots; ++i) {
    size_t have_in = w->slab_in_bytes, have_out = w->slab_out_bytes;
    rc = ensure(&w->d_in[i], &have_in, (size_t)(in_rows_max * row_in) * es);
    if (rc) return rc;
    rc = ensure(&w->d_out[i], &have_out, (size_t)(rows * row_out) * es);
    if (rc) return rc;
    if (i == kSlots - 1) {
      w->slab_in_bytes = have_in;
      w->slab_out_bytes = have_out;
    }
  }
  // (all three slots share one recorded capacity: grow them together)
  // metrics: uploaded whole, once
  const void* hm[2] = {pre_metric, post_metric};
  const int64_t* ms[2] = {pre_strides, post_strides};
  const int64_t* mshape[2] = {shape, out_shape};
  for (int k = 0; k < 2; ++k) {
    if (!hm[k]) continue;
    if (!ms[k]) return xg_fail(XG_EINVAL, "xg_stencil2_host: metric strides missing");
    const size_t span = operand_span(ms[k], mshape[k], ndim, es);
    rc = ensure(&w->d_metric[k], &w->metric_bytes[k], span);
    if (rc) return rc;
    XG_CUDA(cudaMemcpyAsync(w->d_metric[k], hm[k], span, cudaMemcpyHostToDevice, w->s_h2d));
  }
  // exterior halo planes when dim 0 is the operated axis and the halo is data (periodic wrap)
  const char* hin = static_cast<const char*>(in);
  char* hout = static_cast<char*>(out);
  const int64_t n0 = shape[0];
  bool wrap_planes = ax0 && bc == XG_BC_PERIODIC && (pre_metric == nullptr);
  if (ax0 && bc == XG_BC_PERIODIC && pre_metric != nullptr)
    return xg_fail(XG_ENOTIMPL,
                   "xg_stencil2_host: periodic halo with a pre-metric along the outermost axis; "
                   "use the device entry point");
  if (wrap_planes) {
    size_t hb = w->halo_bytes;
    for (int k = 0; k < 2; ++k) {
      size_t have = hb;
      rc = ensure(&w->d_halo[k], &have, (size_t)row_in * es);
      if (rc) return rc;
      if (k == 1) w->halo_bytes = have;
    }
    if (lo)  // below the first plane sits the last plane
      XG_CUDA(cudaMemcpyAsync(w->d_halo[0], hin + (size_t)(n0 - 1) * row_in * es, row_in * es,
                              cudaMemcpyHostToDevice, w->s_h2d));
    if (hi)
      XG_CUDA(cudaMemcpyAsync(w->d_halo[1], hin, row_in * es, cudaMemcpyHostToDevice, w->s_h2d));
  }
  XG_CUDA(cudaEventRecord(w->e_up[0], w->s_h2d));
  XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_up[0], 0));  // metrics + halo planes before any kernel

  int64_t slab_shape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) slab_shape[d] = shape[d];

  int64_t prev_last_row = -1;          // global index of the last plane of the previous slab
  const char* prev_last_ptr = nullptr;  // ... and where it sits on the device
  for (int64_t s = 0; s < nslab; ++s) {
    const int slot = (int)(s % kSlots);
    const int64_t j0 = s * rows;                                  // first output row of the slab
    const int64_t j1 = (j0 + rows < n0_out) ? j0 + rows : n0_out;  // one past the last
    // input rows needed: non-operated dim 0 -> [j0, j1); operated -> P[j0 .. j1] i.e.
    // source rows [j0 - lo, j1 - lo] clipped to [0, n0)
    int64_t i0 = j0, i1 = j1;
    int slab_lo = lo, slab_hi = hi;
    if (ax0) {
      i0 = j0 - lo;
      i1 = j1 - lo + 1;
      slab_lo = 0;
      slab_hi = 0;
      if (i0 < 0) { i0 = 0; slab_lo = 1; }
      if (i1 > n0) { i1 = n0; slab_hi = 1; }
    }
    // slot reuse: the previous D2H out of this slot must have drained, and the kernel that read
    // the slot's input must have finished before we overwrite it
    if (s >= kSlots) {
      XG_CUDA(cudaStreamWaitEvent(w->s_h2d, w->e_done[slot], 0));
      XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_down[slot], 0));
    }
    if (ax0 && s > 0 && prev_last_row == i0 && i1 - i0 > 1) {
      // consecutive slabs of the operated axis overlap by exactly one plane: carry it over on the
      // device (same in-order stream as the uploads) instead of sending it over PCIe again
      XG_CUDA(cudaMemcpyAsync(w->d_in[slot], prev_last_ptr, (size_t)row_in * es,
                              cudaMemcpyDeviceToDevice, w->s_h2d));
      XG_CUDA(cudaMemcpyAsync((char*)w->d_in[slot] + (size_t)row_in * es,
                              hin + (size_t)(i0 + 1) * row_in * es,
                              (size_t)(i1 - i0 - 1) * row_in * es, cudaMemcpyHostToDevice, w->s_h2d));
    } else {
      XG_CUDA(cudaMemcpyAsync(w->d_in[slot], hin + (size_t)i0 * row_in * es,
                              (size_t)(i1 - i0) * row_in * es, cudaMemcpyHostToDevice, w->s_h2d));
    }
    prev_last_row = i1 - 1;
    prev_last_ptr = (const char*)w->d_in[slot] + (size_t)(i1 - 1 - i0) * row_in * es;
    XG_CUDA(cudaEventRecord(w->e_up[slot], w->s_h2d));
    XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_up[slot], 0));

    slab_shape[0] = i1 - i0;
    const char* pm = (const char*)w->d_metric[0];
    const char* qm = (const char*)w->d_metric[1];
    if (pre_metric) pm += (size_t)(i0 * pre_strides[0]) * es;
    if (post_metric) qm += (size_t)(j0 * post_strides[0]) * es;
    const void* hl = nullptr;
    const void* hh = nullptr;
    int slab_bc = bc;
    if (ax0 && wrap_planes) {
      if (slab_lo) hl = w->d_halo[0];
      if (slab_hi) hh = w->d_halo[1];
    }
    rc = xg_stencil2(op, dtype, w->d_in[slot], w->d_out[slot], ndim, slab_shape, axis, slab_lo,
                     slab_hi, (slab_lo || slab_hi) ? slab_bc : XG_BC_NONE, fill_value,
                     pre_metric ? pm : nullptr, pre_strides, post_metric ? qm : nullptr,
                     post_strides, hl, hh, w->s_k);
    if (rc) {
      cudaDeviceSynchronize();
      return rc;
    }
    XG_CUDA(cudaEventRecord(w->e_done[slot], w->s_k));
    XG_CUDA(cudaStreamWaitEvent(w->s_d2h, w->e_done[slot], 0));
    XG_CUDA(cudaMemcpyAsync(hout + (size_t)j0 * row_out * es, w->d_out[slot],
                            (size_t)(j1 - j0) * row_out * es, cudaMemcpyDeviceToHost, w->s_d2h));
    XG_CUDA(cudaEventRecord(w->e_down[slot], w->s_d2h));
  }
  XG_CUDA(cudaStreamSynchronize(w->s_d2h));
  XG_CUDA(cudaStreamSynchronize(w->s_k));
  XG_CUDA(cudaStreamSynchronize(w->s_h2d));
  return XG_OK;
}
