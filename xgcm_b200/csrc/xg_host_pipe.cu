// Host-buffer entry points beyond xg_stencil2_host: the SAME three-stream slab pipeline
//     H2D(slab s+1)  ||  kernel(s)(slab s)  ||  D2H(slab s-1)
// generalised to (a) several results per uploaded slab and (b) slabs cut along a dimension that is
// not the outermost one (strided 2-D copies), so that every device entry point has a host twin:
//
//   xg_stencil2_host_multi   one field up, K (op, axis, shift, boundary) results down: a `Grid.diff` +
//                            `Grid.interp` sweep over X, Y, Z moves the field over PCIe once, not six
//                            times (xgcm/grid.py:796-832 would re-read it per call).
//   xg_cumscan_host          xgcm/grid.py:1306-1414 on numpy-backed fields
//   xg_wreduce_host          xgcm/grid.py:1598-1605, :1680-1685
//   xg_vinterp_linear_host   xgcm/transform.py:233-249
//
// The three single-result twins cut slabs along the first NON-operated dimension: a slab of
// (Z, y0:y1, X) is Z pieces of (y1-y0)*X contiguous elements -> one cudaMemcpy2DAsync each way.  Lines
// along the operated axis stay whole, so no halo exchange between slabs is needed and summation order
// is untouched.  Metric / weight / theta / target operands are small and uploaded whole, once.
//
// Workspaces (device slabs, streams, events) are cached per device and guarded by a per-device mutex:
// calls on different GPUs run concurrently, calls on one GPU serialise (they would fight for PCIe anyway).
#include <stdlib.h>

#include <functional>
#include <mutex>
#include <vector>

#include "xg_common.cuh"

namespace {

constexpr int kSlots = 3;
constexpr int kMaxOut = 8;
constexpr int kMaxAux = 4;

struct PipeWorkspace {
  int device = -1;
  std::mutex mu;
  size_t in_cap[kSlots] = {0, 0, 0};
  void* d_in[kSlots] = {nullptr, nullptr, nullptr};
  size_t out_cap[kSlots][kMaxOut] = {};
  void* d_out[kSlots][kMaxOut] = {};
  size_t aux_cap[kMaxAux] = {0, 0, 0, 0};
  void* d_aux[kMaxAux] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
  cudaEvent_t e_up[kSlots], e_done[kSlots], e_down[kSlots], e_aux;
  bool ready = false;
};

std::mutex g_reg_mutex;
std::vector<PipeWorkspace*> g_pipes;

#define XG_CUDA(call)                                                               \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return xg_fail(XG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
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

int get_pipe(int device, PipeWorkspace** out) {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  for (PipeWorkspace* w : g_pipes)
    if (w->device == device) {
      *out = w;
      return XG_OK;
    }
  PipeWorkspace* w = new PipeWorkspace();
  w->device = device;
  g_pipes.push_back(w);
  *out = w;
  return XG_OK;
}

int init_pipe(PipeWorkspace* w) {  // caller holds w->mu and has set the device
  if (w->ready) return XG_OK;
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_h2d, cudaStreamNonBlocking));
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_k, cudaStreamNonBlocking));
  XG_CUDA(cudaStreamCreateWithFlags(&w->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    XG_CUDA(cudaEventCreateWithFlags(&w->e_up[i], cudaEventDisableTiming));
    XG_CUDA(cudaEventCreateWithFlags(&w->e_done[i], cudaEventDisableTiming));
    XG_CUDA(cudaEventCreateWithFlags(&w->e_down[i], cudaEventDisableTiming));
  }
  XG_CUDA(cudaEventCreateWithFlags(&w->e_aux, cudaEventDisableTiming));
  w->ready = true;
  return XG_OK;
}

size_t operand_span(const int64_t* strides, const int64_t* shape, int ndim, size_t es) {
  int64_t last = 0;
  for (int d = 0; d < ndim; ++d)
    if (shape[d] > 1) last += (shape[d] - 1) * strides[d];
  return (size_t)(last + 1) * es;
}

// [C][L][R] view of a C-contiguous array around the slab dimension (extent L)
struct View3 {
  int64_t C, L, R;
};

View3 view3(int ndim, const int64_t* shape, int sd) {
  View3 v{1, ndim ? shape[sd] : 1, 1};
  for (int d = 0; d < sd; ++d) v.C *= shape[d];
  for (int d = sd + 1; d < ndim; ++d) v.R *= shape[d];
  return v;
}

// rows [j0, j1) of the slab dim <-> a dense [C][j1-j0][R] device block
int copy_slab(void* dev, const void* host_base, const View3& v, int64_t j0, int64_t j1, size_t es, bool to_device,
              cudaStream_t st) {
  const size_t width = (size_t)(j1 - j0) * v.R * es;
  if (width == 0 || v.C == 0) return XG_OK;
  const char* h = static_cast<const char*>(host_base) + (size_t)j0 * v.R * es;
  const size_t hpitch = (size_t)v.L * v.R * es;
  if (v.C == 1) {
    if (to_device) XG_CUDA(cudaMemcpyAsync(dev, h, width, cudaMemcpyHostToDevice, st));
    else XG_CUDA(cudaMemcpyAsync(const_cast<char*>(h), dev, width, cudaMemcpyDeviceToHost, st));
    return XG_OK;
  }
  if (to_device) XG_CUDA(cudaMemcpy2DAsync(dev, width, h, hpitch, width, (size_t)v.C, cudaMemcpyHostToDevice, st));
  else XG_CUDA(cudaMemcpy2DAsync(const_cast<char*>(h), hpitch, dev, width, width, (size_t)v.C, cudaMemcpyDeviceToHost, st));
  return XG_OK;
}

int64_t slab_rows(const View3& in, size_t es, int64_t extra_rows) {
  int64_t target_bytes = 128ll << 20;
  if (const char* env = getenv("XG_HOST_SLAB_MB")) {  // tuning knob (benchmarks only)
    const long mb = atol(env);
    if (mb >= 1 && mb <= 4096) target_bytes = (int64_t)mb << 20;
  }
  const int64_t row_bytes = (int64_t)(in.C * in.R * (int64_t)es);
  int64_t rows = row_bytes > 0 ? target_bytes / row_bytes : in.L;
  if (rows < 1) rows = 1;
  if (rows > (in.L + 3) / 4) rows = (in.L + 3) / 4;  // at least 4 slabs when the dim allows: overlap
  if (rows < 1 + extra_rows) rows = 1 + extra_rows;
  return rows;
}

// launch(j0, j1, i0, i1, d_in, d_out[], stream): kernels for output rows [j0, j1) given input rows [i0, i1)
typedef std::function<int(int64_t, int64_t, int64_t, int64_t, void*, void* const*, cudaStream_t)> LaunchFn;

// The pipeline.  `halo` = extra input rows wanted on each side of a slab (0, or 1 when some result is
// operated along the slab dim).  Result k has view out[k] with the same L as the input.
int run_pipe(PipeWorkspace* w, size_t es, const void* hin, const View3& in, int nout, void* const* hout,
             const View3* out, int halo, const LaunchFn& launch) {
  if (in.L == 0 || in.C == 0 || in.R == 0) return XG_OK;
  const int64_t rows = slab_rows(in, es, 0);
  const int64_t nslab = xg_ceil_div(in.L, rows);
  for (int i = 0; i < kSlots; ++i) {
    int rc = ensure(&w->d_in[i], &w->in_cap[i], (size_t)(in.C * (rows + 2 * halo) * in.R) * es);
    if (rc) return rc;
    for (int k = 0; k < nout; ++k) {
      rc = ensure(&w->d_out[i][k], &w->out_cap[i][k], (size_t)(out[k].C * rows * out[k].R) * es);
      if (rc) return rc;
    }
  }
  for (int64_t s = 0; s < nslab; ++s) {
    const int slot = (int)(s % kSlots);
    const int64_t j0 = s * rows, j1 = (j0 + rows < in.L) ? j0 + rows : in.L;
    const int64_t i0 = (j0 - halo < 0) ? 0 : j0 - halo, i1 = (j1 + halo > in.L) ? in.L : j1 + halo;
    if (s >= kSlots) {
      XG_CUDA(cudaStreamWaitEvent(w->s_h2d, w->e_done[slot], 0));  // kernels that read this slot's input
      XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_down[slot], 0));    // downloads out of this slot's results
    }
    int rc = copy_slab(w->d_in[slot], hin, in, i0, i1, es, true, w->s_h2d);
    if (rc) return rc;
    XG_CUDA(cudaEventRecord(w->e_up[slot], w->s_h2d));
    XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_up[slot], 0));
    rc = launch(j0, j1, i0, i1, w->d_in[slot], w->d_out[slot], w->s_k);
    if (rc) {
      cudaDeviceSynchronize();
      return rc;
    }
    XG_CUDA(cudaEventRecord(w->e_done[slot], w->s_k));
    XG_CUDA(cudaStreamWaitEvent(w->s_d2h, w->e_done[slot], 0));
    for (int k = 0; k < nout; ++k) {
      rc = copy_slab(w->d_out[slot][k], hout[k], out[k], j0, j1, es, false, w->s_d2h);
      if (rc) return rc;
    }
    XG_CUDA(cudaEventRecord(w->e_down[slot], w->s_d2h));
  }
  XG_CUDA(cudaStreamSynchronize(w->s_d2h));
  XG_CUDA(cudaStreamSynchronize(w->s_k));
  XG_CUDA(cudaStreamSynchronize(w->s_h2d));
  return XG_OK;
}

// upload a small broadcast operand whole (aux slot `slot`); *dev = its device address (nullptr if absent)
int upload_aux(PipeWorkspace* w, int slot, const void* host, size_t bytes, const void** dev) {
  *dev = nullptr;
  if (!host) return XG_OK;
  int rc = ensure(&w->d_aux[slot], &w->aux_cap[slot], bytes);
  if (rc) return rc;
  XG_CUDA(cudaMemcpyAsync(w->d_aux[slot], host, bytes, cudaMemcpyHostToDevice, w->s_h2d));
  *dev = w->d_aux[slot];
  return XG_OK;
}

int aux_fence(PipeWorkspace* w) {  // kernels must see the aux uploads
  XG_CUDA(cudaEventRecord(w->e_aux, w->s_h2d));
  XG_CUDA(cudaStreamWaitEvent(w->s_k, w->e_aux, 0));
  return XG_OK;
}

struct Session {  // device selected, workspace locked and initialised
  PipeWorkspace* w = nullptr;
  std::unique_lock<std::mutex> lock;
  int open(int device) {
    int rc = get_pipe(device, &w);
    if (rc) return rc;
    lock = std::unique_lock<std::mutex>(w->mu);
    XG_CUDA(cudaSetDevice(device));
    return init_pipe(w);
  }
};

int first_free_dim(int ndim, int axis) { return (ndim == 1) ? -1 : (axis == 0 ? 1 : 0); }

}  // namespace

void xg_host_pipe_release() {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  for (PipeWorkspace* w : g_pipes) {
    std::lock_guard<std::mutex> l2(w->mu);
    cudaSetDevice(w->device);
    for (int i = 0; i < kSlots; ++i) {
      if (w->d_in[i]) cudaFree(w->d_in[i]);
      w->d_in[i] = nullptr;
      w->in_cap[i] = 0;
      for (int k = 0; k < kMaxOut; ++k) {
        if (w->d_out[i][k]) cudaFree(w->d_out[i][k]);
        w->d_out[i][k] = nullptr;
        w->out_cap[i][k] = 0;
      }
    }
    for (int i = 0; i < kMaxAux; ++i) {
      if (w->d_aux[i]) cudaFree(w->d_aux[i]);
      w->d_aux[i] = nullptr;
      w->aux_cap[i] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
extern "C" int xg_stencil2_host_multi(int nout, const int* op, int dtype, const void* in, void* const* out, int ndim,
                                      const int64_t* shape, const int* axis, const int* lo, const int* hi,
                                      const int* bc, const double* fill_value, int device) {
  if (!in || !out || !shape || !op || !axis || !lo || !hi || !bc || !fill_value)
    return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: null pointer");
  if (nout < 1 || nout > kMaxOut)
    return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: between 1 and " + std::to_string(kMaxOut) + " results per call");
  if (dtype != XG_F32 && dtype != XG_F64)
    return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: dtype must be XG_F32 or XG_F64");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: bad ndim");
  const size_t es = dtype == XG_F32 ? 4 : 8;
  bool any0 = false;
  for (int k = 0; k < nout; ++k) {
    if (!out[k]) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: null result pointer");
    if (axis[k] < 0 || axis[k] >= ndim) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: axis out of range");
    if (lo[k] < 0 || lo[k] > 1 || hi[k] < 0 || hi[k] > 1)
      return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: halo widths must be 0 or 1");
    if ((lo[k] || hi[k]) && (bc[k] <= XG_BC_NONE || bc[k] > XG_BC_EXTRAPOLATE))
      return xg_fail(XG_EINVAL,
                     "xg_stencil2_host_multi: no boundary condition was specified but the operation needs to "
                     "pad the axis");
    if (shape[axis[k]] == 0) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: empty operated axis");
    if (axis[k] == 0) {
      // slabs are cut along dim 0: a result operated along it must keep that extent (center <-> left / right
      // shifts), and its halo must be expressible per slab
      if (lo[k] + hi[k] != 1)
        return xg_fail(XG_ENOTIMPL,
                       "xg_stencil2_host_multi: outer / inner shifts along the outermost dimension; use "
                       "xg_stencil2_host for that result");
      if (bc[k] == XG_BC_EXTRAPOLATE)
        return xg_fail(XG_ENOTIMPL, "xg_stencil2_host_multi: extrapolate along the outermost dimension");
      any0 = true;
    }
  }
  Session ss;
  int rc = ss.open(device);
  if (rc) return rc;
  PipeWorkspace* w = ss.w;

  const View3 vin = view3(ndim, shape, 0);
  View3 vout[kMaxOut];
  int64_t out_shape[kMaxOut][XG_MAX_NDIM];
  for (int k = 0; k < nout; ++k) {
    for (int d = 0; d < ndim; ++d) out_shape[k][d] = shape[d];
    out_shape[k][axis[k]] = shape[axis[k]] + lo[k] + hi[k] - 1;
    vout[k] = view3(ndim, out_shape[k], 0);
    if (vout[k].R == 0 || vout[k].L == 0) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: empty result");
  }
  const int64_t n0 = shape[0];
  // periodic wrap planes for results operated along dim 0: plane n0-1 below the first slab, plane 0 above the last
  const void* d_wrap[2] = {nullptr, nullptr};
  bool need_wrap = false;
  for (int k = 0; k < nout; ++k) need_wrap = need_wrap || (axis[k] == 0 && bc[k] == XG_BC_PERIODIC);
  if (need_wrap) {
    const size_t pb = (size_t)vin.R * es;
    rc = upload_aux(w, 0, static_cast<const char*>(in) + (size_t)(n0 - 1) * pb, pb, &d_wrap[0]);
    if (rc) return rc;
    rc = upload_aux(w, 1, in, pb, &d_wrap[1]);
    if (rc) return rc;
    rc = aux_fence(w);
    if (rc) return rc;
  }
  auto launch = [&](int64_t j0, int64_t j1, int64_t i0, int64_t i1, void* d_in, void* const* d_out,
                    cudaStream_t st) -> int {
    int64_t sshape[XG_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) sshape[d] = shape[d];
    for (int k = 0; k < nout; ++k) {
      const char* src = static_cast<const char*>(d_in);
      int slo = lo[k], shi = hi[k], sbc = bc[k];
      const void* hl = nullptr;
      const void* hh = nullptr;
      if (axis[k] == 0) {
        // output rows [j0, j1) need P[j0 .. j1], i.e. source planes [j0 - lo, j1 - lo] clipped to the field
        int64_t s0 = j0 - lo[k], s1 = j1 - lo[k] + 1;
        slo = shi = 0;
        if (s0 < 0) { s0 = 0; slo = 1; }
        if (s1 > n0) { s1 = n0; shi = 1; }
        if (s0 < i0 || s1 > i1) return xg_fail(XG_EINVAL, "xg_stencil2_host_multi: internal slab window error");
        src += (size_t)(s0 - i0) * vin.R * es;
        sshape[0] = s1 - s0;
        if (bc[k] == XG_BC_PERIODIC) {
          if (slo) hl = d_wrap[0];
          if (shi) hh = d_wrap[1];
        }
        if (!slo && !shi) sbc = XG_BC_NONE;
      } else {
        src += (size_t)(j0 - i0) * vin.R * es;
        sshape[0] = j1 - j0;
      }
      const int rc2 = xg_stencil2(op[k], dtype, src, d_out[k], ndim, sshape, axis[k], slo, shi, sbc, fill_value[k],
                                  nullptr, nullptr, nullptr, nullptr, hl, hh, st);
      if (rc2) return rc2;
    }
    return XG_OK;
  };
  return run_pipe(w, es, in, vin, nout, out, vout, any0 ? 1 : 0, launch);
}

// ---------------------------------------------------------------------------------------------------
extern "C" int xg_cumscan_host(int dtype, const void* in, void* out, int ndim, const int64_t* shape, int axis,
                               int reverse, int trim, int pad_lo, int pad_hi, int bc, double fill_value,
                               const void* pre_metric, const int64_t* pre_strides, const void* post_metric,
                               const int64_t* post_strides, int skipna, int device) {
  if (!in || !out || !shape) return xg_fail(XG_EINVAL, "xg_cumscan_host: null pointer");
  if (dtype != XG_F32 && dtype != XG_F64) return xg_fail(XG_EINVAL, "xg_cumscan_host: dtype must be XG_F32 or XG_F64");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_cumscan_host: bad ndim");
  if (axis < 0 || axis >= ndim) return xg_fail(XG_EINVAL, "xg_cumscan_host: axis out of range");
  if ((pre_metric && !pre_strides) || (post_metric && !post_strides))
    return xg_fail(XG_EINVAL, "xg_cumscan_host: metric strides missing");
  const size_t es = dtype == XG_F32 ? 4 : 8;
  const int64_t kept = shape[axis] - (trim != XG_TRIM_NONE ? 1 : 0);
  if (kept < 0) return xg_fail(XG_EINVAL, "xg_cumscan_host: operated axis too short to trim");
  int64_t out_shape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) out_shape[d] = shape[d];
  out_shape[axis] = kept + pad_lo + pad_hi;
  Session ss;
  int rc = ss.open(device);
  if (rc) return rc;
  PipeWorkspace* w = ss.w;
  const int sd = first_free_dim(ndim, axis);
  const void* d_pre = nullptr;
  const void* d_post = nullptr;
  if (pre_metric) rc = upload_aux(w, 0, pre_metric, operand_span(pre_strides, shape, ndim, es), &d_pre);
  if (rc) return rc;
  if (post_metric) rc = upload_aux(w, 1, post_metric, operand_span(post_strides, out_shape, ndim, es), &d_post);
  if (rc) return rc;
  rc = aux_fence(w);
  if (rc) return rc;
  if (sd < 0) {  // 1-D: one slab = the whole line
    const View3 v{1, 1, shape[0]}, vo{1, 1, out_shape[0]};
    void* outs[1] = {out};
    auto launch = [&](int64_t, int64_t, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
      return xg_cumscan(dtype, d_in, d_out[0], ndim, shape, axis, reverse, trim, pad_lo, pad_hi, bc, fill_value, d_pre,
                        pre_strides, d_post, post_strides, skipna, st);
    };
    return run_pipe(w, es, in, v, 1, outs, &vo, 0, launch);
  }
  const View3 vin = view3(ndim, shape, sd), vout = view3(ndim, out_shape, sd);
  void* outs[1] = {out};
  auto launch = [&](int64_t j0, int64_t j1, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
    int64_t sshape[XG_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) sshape[d] = shape[d];
    sshape[sd] = j1 - j0;
    const char* pm = static_cast<const char*>(d_pre);
    const char* qm = static_cast<const char*>(d_post);
    if (pm) pm += (size_t)(j0 * pre_strides[sd]) * es;
    if (qm) qm += (size_t)(j0 * post_strides[sd]) * es;
    return xg_cumscan(dtype, d_in, d_out[0], ndim, sshape, axis, reverse, trim, pad_lo, pad_hi, bc, fill_value, pm,
                      pre_strides, qm, post_strides, skipna, st);
  };
  return run_pipe(w, es, in, vin, 1, outs, &vout, 0, launch);
}

// ---------------------------------------------------------------------------------------------------
extern "C" int xg_wreduce_host(int dtype, const void* in, const void* weight, const int64_t* w_strides, void* out,
                               int ndim, const int64_t* shape, int axis, int mode, int skipna, int device) {
  if (!in || !out || !shape) return xg_fail(XG_EINVAL, "xg_wreduce_host: null pointer");
  if (dtype != XG_F32 && dtype != XG_F64) return xg_fail(XG_EINVAL, "xg_wreduce_host: dtype must be XG_F32 or XG_F64");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_wreduce_host: bad ndim");
  if (axis < 0 || axis >= ndim) return xg_fail(XG_EINVAL, "xg_wreduce_host: axis out of range");
  if (weight && !w_strides) return xg_fail(XG_EINVAL, "xg_wreduce_host: weight strides missing");
  const size_t es = dtype == XG_F32 ? 4 : 8;
  Session ss;
  int rc = ss.open(device);
  if (rc) return rc;
  PipeWorkspace* w = ss.w;
  const void* d_w = nullptr;
  if (weight) rc = upload_aux(w, 0, weight, operand_span(w_strides, shape, ndim, es), &d_w);
  if (rc) return rc;
  rc = aux_fence(w);
  if (rc) return rc;
  const int sd = first_free_dim(ndim, axis);
  void* outs[1] = {out};
  if (sd < 0) {
    const View3 v{1, 1, shape[0]}, vo{1, 1, 1};
    auto launch = [&](int64_t, int64_t, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
      return xg_wreduce(dtype, d_in, d_w, w_strides, d_out[0], ndim, shape, axis, mode, skipna, st);
    };
    return run_pipe(w, es, in, v, 1, outs, &vo, 0, launch);
  }
  // result shape = shape without `axis`; the slab dim keeps its extent
  int64_t out_shape[XG_MAX_NDIM];
  int nd_o = 0, sd_o = 0;
  for (int d = 0; d < ndim; ++d) {
    if (d == axis) continue;
    if (d == sd) sd_o = nd_o;
    out_shape[nd_o++] = shape[d];
  }
  const View3 vin = view3(ndim, shape, sd), vout = view3(nd_o, out_shape, sd_o);
  auto launch = [&](int64_t j0, int64_t j1, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
    int64_t sshape[XG_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) sshape[d] = shape[d];
    sshape[sd] = j1 - j0;
    const char* wm = static_cast<const char*>(d_w);
    if (wm) wm += (size_t)(j0 * w_strides[sd]) * es;
    return xg_wreduce(dtype, d_in, wm, w_strides, d_out[0], ndim, sshape, axis, mode, skipna, st);
  };
  return run_pipe(w, es, in, vin, 1, outs, &vout, 0, launch);
}

// ---------------------------------------------------------------------------------------------------
extern "C" int xg_vinterp_linear_host(int dtype, const void* phi, const void* theta, const int64_t* theta_strides,
                                      const void* target, const int64_t* target_strides, int64_t m, void* out,
                                      int ndim, const int64_t* shape, int axis, int mask_edges, int bypass_checks,
                                      int logarithmic, int device) {
  if (!phi || !theta || !theta_strides || !target || !out || !shape)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear_host: null pointer");
  if (dtype != XG_F32 && dtype != XG_F64)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear_host: dtype must be XG_F32 or XG_F64");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_vinterp_linear_host: bad ndim");
  if (axis < 0 || axis >= ndim) return xg_fail(XG_EINVAL, "xg_vinterp_linear_host: axis out of range");
  if (m < 0) return xg_fail(XG_EINVAL, "xg_vinterp_linear_host: negative number of target levels");
  const size_t es = dtype == XG_F32 ? 4 : 8;
  Session ss;
  int rc = ss.open(device);
  if (rc) return rc;
  PipeWorkspace* w = ss.w;
  // theta (shared coordinate or a whole field) and target are uploaded whole: the field case costs one extra
  // resident copy of theta, it is not the streamed operand
  const void* d_theta = nullptr;
  const void* d_target = nullptr;
  rc = upload_aux(w, 0, theta, operand_span(theta_strides, shape, ndim, es), &d_theta);
  if (rc) return rc;
  int64_t tshape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) tshape[d] = shape[d];
  tshape[axis] = m;
  const size_t tbytes = target_strides ? operand_span(target_strides, tshape, ndim, es) : (size_t)m * es;
  rc = upload_aux(w, 1, target, tbytes ? tbytes : es, &d_target);
  if (rc) return rc;
  rc = aux_fence(w);
  if (rc) return rc;
  const int sd = first_free_dim(ndim, axis);
  void* outs[1] = {out};
  if (sd < 0) {
    const View3 v{1, 1, shape[0]}, vo{1, 1, m};
    auto launch = [&](int64_t, int64_t, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
      return xg_vinterp_linear(dtype, d_in, d_theta, theta_strides, d_target, target_strides, m, d_out[0], ndim, shape,
                               axis, mask_edges, bypass_checks, logarithmic, st);
    };
    return run_pipe(w, es, phi, v, 1, outs, &vo, 0, launch);
  }
  // result shape = shape without `axis`, then m appended last
  int64_t out_shape[XG_MAX_NDIM + 1];
  int nd_o = 0, sd_o = 0;
  for (int d = 0; d < ndim; ++d) {
    if (d == axis) continue;
    if (d == sd) sd_o = nd_o;
    out_shape[nd_o++] = shape[d];
  }
  out_shape[nd_o++] = m;
  const View3 vin = view3(ndim, shape, sd), vout = view3(nd_o, out_shape, sd_o);
  auto launch = [&](int64_t j0, int64_t j1, int64_t, int64_t, void* d_in, void* const* d_out, cudaStream_t st) -> int {
    int64_t sshape[XG_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) sshape[d] = shape[d];
    sshape[sd] = j1 - j0;
    const char* th = static_cast<const char*>(d_theta) + (size_t)(j0 * theta_strides[sd]) * es;
    const char* tg = static_cast<const char*>(d_target);
    if (target_strides) tg += (size_t)(j0 * target_strides[sd]) * es;
    return xg_vinterp_linear(dtype, d_in, th, theta_strides, tg, target_strides, m, d_out[0], ndim, sshape, axis,
                             mask_edges, bypass_checks, logarithmic, st);
  };
  return run_pipe(w, es, phi, vin, 1, outs, &vout, 0, launch);
}
