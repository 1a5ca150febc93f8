// xg_stencil2 — fused halo-pad + 2-point stencil + metric multiply/divide.
//
// Replaces, in ONE pass over HBM (read n, write n):
//   xgcm/padding.py:575-616   np.pad copy of the whole field
//   xgcm/gridops.py:23-24,76-77,123-126,172-175   the pairwise operator
//   xgcm/grid.py:806-808,830-832,1576-1578        metric multiply / divide
//
// Any C-contiguous field collapses to (outer, n, inner) around the operated
// axis.  Three kernels:
//   k_stencil_strided  inner > 1 (Y, Z, ...): a warp owns 32 x VEC contiguous
//       columns and marches J cells along the axis keeping the previous row in
//       registers, so each input element is read once; U independent 16-byte
//       loads are in flight per thread.
//   k_stencil_row_vec  inner == 1 (X), aligned rows, n_out == n: a warp owns a
//       512 B x U chunk of one row, the missing neighbour of each 16-byte
//       vector comes from a warp shuffle, and only the chunk edge does one
//       extra scalar load (or takes the boundary value).
//   k_stencil_row_scalar  inner == 1, any length / alignment / n_out != n.
//
// Roofline: HBM.  Algorithmic bytes = 2 * sizeof(T) per output cell
// (+ metric bytes), see DESIGN.md.
#include <stdlib.h>

#include "xg_common.cuh"

namespace {

template <typename T>
struct StencilArgs {
  const T* in;
  T* out;
  int64_t outer, n, inner, n_out;
  int lo, hi, bc;
  T fill;
  int J;               // cells marched per warp-unit (strided kernel)
  int64_t nseg, nwc;   // segments along the axis, warp-columns (or row chunks)
  int64_t nunits;      // total warp-units
  bool small_units;    // nunits < 2^31: 32-bit unit decomposition
  bool seg_fast;       // unit order: segment index fastest (else warp-column fastest)
  XgOperand pre, post;
  int pre_axis_vec_ok, post_axis_vec_ok;  // row kernels: metric vector loads along x
  const T* halo_lo;
  const T* halo_hi;
};

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

// ---------------------------------------------------------------------------
// strided-axis kernel
// ---------------------------------------------------------------------------
template <typename T, int VEC, int OP, bool MET, int U>
__global__ void __launch_bounds__(kThreads, 4)  // <= 64 registers: 4 CTAs (1024 threads) per SM
k_stencil_strided(const StencilArgs<T> a) {
  typedef XgPack<T, VEC> Pack;
  const int64_t unit =
      (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (unit >= a.nunits) return;
  const int lane = threadIdx.x & 31;
  int64_t wc, t, seg, o;
  if (a.seg_fast) {  // segments of one column group adjacent in launch order
    xg_divmod(unit, a.nseg, a.small_units, t, seg);
    xg_divmod(t, a.nwc, a.small_units, o, wc);
  } else {
    xg_divmod(unit, a.nwc, a.small_units, t, wc);
    xg_divmod(t, a.nseg, a.small_units, o, seg);
  }
  const int64_t i = (wc * 32 + lane) * VEC;
  if (i >= a.inner) return;

  const int64_t j0 = seg * a.J;
  const int64_t j1 = (j0 + a.J < a.n_out) ? (j0 + a.J) : a.n_out;

  const T* ibase = a.in + o * a.n * a.inner + i;
  T* obase = a.out + o * a.n_out * a.inner + i;
  const bool has_pre = MET && a.pre.ptr != nullptr;
  const bool has_post = MET && a.post.ptr != nullptr;
  XgOperandView<T, VEC> pre_v, post_v;
  if (MET) {
    // everything that depends only on (o, i) is hoisted out of the march
    if (has_pre) pre_v = xg_operand_view<T, VEC>(a.pre, xg_groups_offset(a.pre.outer, o), i);
    if (has_post) post_v = xg_operand_view<T, VEC>(a.post, xg_groups_offset(a.post.outer, o), i);
  }

  // A[s] = in[s] * pre[s], s in range
  auto loadA = [&](int64_t s) -> Pack {
    Pack v = xg_ld_stream<T, VEC>(ibase + s * a.inner);
    if (has_pre) {
      Pack m = xg_ld_view<T, VEC>(pre_v, s * a.pre.axis_stride);
#pragma unroll
      for (int k = 0; k < VEC; ++k) v.v[k] = v.v[k] * m.v[k];
    }
    return v;
  };
  auto emit = [&](int64_t j, const Pack& lo_v, const Pack& hi_v) {
    Pack r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = xg_apply_op<T, OP>(lo_v.v[k], hi_v.v[k]);
    if (has_post) {
      Pack m = xg_ld_view<T, VEC>(post_v, j * a.post.axis_stride);
#pragma unroll
      for (int k = 0; k < VEC; ++k) r.v[k] = r.v[k] / m.v[k];
    }
    xg_st_stream<T, VEC>(obase + j * a.inner, r);
  };

  // P[k] with the boundary rule; only ever needed for the first row of the first segment
  // (s = -1) and the last row of the last segment (s = n)
  auto loadP = [&](int64_t k) -> Pack {
    int64_t s = k - a.lo;
    if (s >= 0 && s < a.n) return loadA(s);
    const bool low = s < 0;
    const T* halo = low ? a.halo_lo : a.halo_hi;
    Pack r;
    if (halo) return xg_ld_cached<T, VEC>(halo + o * a.inner + i);
    if (a.bc == XG_BC_FILL) {
#pragma unroll
      for (int k2 = 0; k2 < VEC; ++k2) r.v[k2] = a.fill;
      return r;
    }
    if (a.bc == XG_BC_PERIODIC) return loadA(low ? s + a.n : s - a.n);
    if (a.bc == XG_BC_EXTEND) return loadA(low ? 0 : a.n - 1);
    // extrapolate: 2*A[edge] - A[next]
    const int64_t e = low ? 0 : a.n - 1;
    const int64_t e2 = a.n > 1 ? (low ? 1 : a.n - 2) : e;
    Pack a0 = loadA(e), a1 = loadA(e2);
#pragma unroll
    for (int q = 0; q < VEC; ++q) r.v[q] = T(2) * a0.v[q] - a1.v[q];
    return r;
  };

  // Row j needs P[j] (carried in registers) and P[j+1] = A[j+1-lo].  Since lo <= 1 the source
  // index j+1-lo is never negative; it is in range while j < n+lo-1.  So the whole march is the
  // branch-free unrolled loop, plus one boundary-aware load at each end.
  Pack prev = loadP(j0);
  const int64_t jm = (j1 < a.n + a.lo - 1) ? j1 : (a.n + a.lo - 1);
  int64_t j = j0;
  for (; j + U <= jm; j += U) {
    Pack cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = loadA(j + u + 1 - a.lo);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      emit(j + u, prev, cur[u]);  // (hoisting the divisor loads here was measured slower: it costs
      prev = cur[u];              //  16 registers, i.e. one resident CTA per SM)
    }
  }
  for (; j < jm; ++j) {
    Pack cur = loadA(j + 1 - a.lo);
    emit(j, prev, cur);
    prev = cur;
  }
#pragma unroll 1
  for (; j < j1; ++j) {  // at most one row: the upper halo
    Pack cur = loadP(j + 1);
    emit(j, prev, cur);
    prev = cur;
  }
}

// ---------------------------------------------------------------------------
// row kernels (operated axis is the innermost one)
// ---------------------------------------------------------------------------

// scalar A[r, s] = in * pre and the boundary values of row r
template <typename T, bool MET>
struct RowAccess {
  const StencilArgs<T>& a;
  const T* row;
  int64_t r;
  int64_t pre_base;
  __device__ __forceinline__ RowAccess(const StencilArgs<T>& a_, int64_t r_)
      : a(a_), row(a_.in + r_ * a_.n), r(r_), pre_base(0) {
    if (MET && a.pre.ptr) pre_base = xg_groups_offset(a.pre.outer, r);
  }
  __device__ __forceinline__ T A(int64_t s) const {
    T v = __ldg(row + s);
    if (MET && a.pre.ptr)
      v = v * __ldg(reinterpret_cast<const T*>(a.pre.ptr) + pre_base +
                    s * a.pre.axis_stride);
    return v;
  }
  __device__ __forceinline__ T below() const {  // P at s = -1
    if (a.halo_lo) return __ldg(a.halo_lo + r);
    if (a.bc == XG_BC_FILL) return a.fill;
    if (a.bc == XG_BC_PERIODIC) return A(a.n - 1);
    if (a.bc == XG_BC_EXTEND) return A(0);
    return T(2) * A(0) - A(a.n > 1 ? 1 : 0);
  }
  __device__ __forceinline__ T above() const {  // P at s = n
    if (a.halo_hi) return __ldg(a.halo_hi + r);
    if (a.bc == XG_BC_FILL) return a.fill;
    if (a.bc == XG_BC_PERIODIC) return A(0);
    if (a.bc == XG_BC_EXTEND) return A(a.n - 1);
    return T(2) * A(a.n - 1) - A(a.n > 1 ? a.n - 2 : 0);
  }
  __device__ __forceinline__ T P(int64_t k) const {
    int64_t s = k - a.lo;
    if (s < 0) return below();
    if (s >= a.n) return above();
    return A(s);
  }
};

template <typename T, int OP, bool MET>
__global__ void __launch_bounds__(kThreads)
k_stencil_row_scalar(const StencilArgs<T> a) {
  const int64_t total = a.outer * a.n_out;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * kThreads) {
    const int64_t r = g / a.n_out;
    const int64_t xo = g - r * a.n_out;
    RowAccess<T, MET> ra(a, r);
    T lo_v = ra.P(xo);
    T hi_v = ra.P(xo + 1);
    T res = xg_apply_op<T, OP>(lo_v, hi_v);
    if (MET && a.post.ptr) {
      int64_t pb = xg_groups_offset(a.post.outer, r);
      res = res / __ldg(reinterpret_cast<const T*>(a.post.ptr) + pb +
                        xo * a.post.axis_stride);
    }
    a.out[g] = res;
  }
}

// metric values for the VEC elements starting at x0 of a row
template <typename T, int VEC>
__device__ __forceinline__ XgPack<T, VEC> row_metric(const XgOperand& m,
                                                    int64_t base, int64_t x0,
                                                    int vec_ok) {
  const T* p = reinterpret_cast<const T*>(m.ptr) + base;
  XgPack<T, VEC> r;
  if (m.axis_stride == 0) {
    T s = __ldg(p);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = s;
  } else if (vec_ok) {
    r = xg_ld_cached<T, VEC>(p + x0);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = __ldg(p + (x0 + k) * m.axis_stride);
  }
  return r;
}

template <typename T, int VEC, int OP, bool MET, int U>
__global__ void __launch_bounds__(kThreads)
k_stencil_row_vec(const StencilArgs<T> a) {
  typedef XgPack<T, VEC> Pack;
  const unsigned FULL = 0xffffffffu;
  const int64_t unit =
      (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (unit >= a.nunits) return;  // warp-uniform
  const int lane = threadIdx.x & 31;
  int64_t c, r;
  xg_divmod(unit, a.nwc, a.small_units, r, c);
  const int64_t nv = a.n / VEC;
  RowAccess<T, MET> ra(a, r);
  T* orow = a.out + r * a.n;  // n_out == n
  int64_t post_base = 0;
  if (MET && a.post.ptr) post_base = xg_groups_offset(a.post.outer, r);

  Pack v[U], pm[U];
  bool act[U];
  int64_t x0[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t q = (c * U + u) * 32 + lane;
    act[u] = q < nv;
    x0[u] = q * VEC;
    if (act[u]) {
      v[u] = xg_ld_stream<T, VEC>(ra.row + x0[u]);
      // the divisor is fetched together with the field so its latency overlaps the field's
      if (MET && a.post.ptr) pm[u] = row_metric<T, VEC>(a.post, post_base, x0[u], a.post_axis_vec_ok);
      if (MET && a.pre.ptr) {
        Pack m = row_metric<T, VEC>(a.pre, ra.pre_base, x0[u], a.pre_axis_vec_ok);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[u].v[k] = v[u].v[k] * m.v[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[u].v[k] = T(0);
    }
  }

#pragma unroll
  for (int u = 0; u < U; ++u) {
    Pack res;
    if (a.lo == 1) {
      // out[x] = OP(A[x-1], A[x])
      T left = __shfl_up_sync(FULL, v[u].v[VEC - 1], 1);
      T wrap = (u > 0) ? __shfl_sync(FULL, v[u > 0 ? u - 1 : 0].v[VEC - 1], 31) : T(0);
      if (lane == 0 && act[u]) {
        if (x0[u] == 0) left = ra.below();
        else if (u > 0) left = wrap;
        else left = ra.A(x0[u] - 1);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        res.v[k] = xg_apply_op<T, OP>(k == 0 ? left : v[u].v[k > 0 ? k - 1 : 0], v[u].v[k]);
    } else {
      // out[x] = OP(A[x], A[x+1])
      T right = __shfl_down_sync(FULL, v[u].v[0], 1);
      T wrap = (u < U - 1) ? __shfl_sync(FULL, v[u < U - 1 ? u + 1 : u].v[0], 0) : T(0);
      if (act[u]) {
        if (x0[u] + VEC >= a.n) right = ra.above();
        else if (lane == 31) right = (u < U - 1) ? wrap : ra.A(x0[u] + VEC);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        res.v[k] = xg_apply_op<T, OP>(v[u].v[k], k == VEC - 1 ? right : v[u].v[k < VEC - 1 ? k + 1 : k]);
    }
    if (act[u]) {
      if (MET && a.post.ptr) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) res.v[k] = res.v[k] / pm[u].v[k];
      }
      xg_st_stream<T, VEC>(orow + x0[u], res);
    }
  }
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T, int VEC, int OP, bool MET>
int launch_strided(StencilArgs<T>& a, cudaStream_t st) {
  const int64_t nvec = xg_ceil_div(a.inner, VEC);
  a.nwc = xg_ceil_div(nvec, 32);
  // march length: long enough that the re-read halo row is a few % of traffic,
  // short enough that there are plenty of warps for 148 SMs.
  static const int tune_j = env_int("XG_STRIDED_J", 0);  // tuning knobs (benchmarks only)
  static const int tune_u = env_int("XG_STRIDED_U", 0);
  // Tuning notes (profiles/r1b_tune_strided.txt): in a loop of identical launches short marches
  // (J = 4) look 8 % faster for Y, but per-launch ncu timings and the mixed sequence of bench.py
  // show no gain, and the fused-metric variants lose 20 % (per-segment operand setup is amortised
  // over fewer rows) — so 32 stays; a plane-strided axis marches as far as possible.
  int J;
  if (tune_j > 0) J = tune_j;
  else J = (a.n_out <= 96) ? (int)a.n_out : 32;
  if (J > a.n_out) J = (int)a.n_out;
  a.J = J;
  a.nseg = xg_ceil_div(a.n_out, J);
  a.nunits = a.outer * a.nseg * a.nwc;
  a.small_units = a.nunits < (1ll << 31);
  static const int tune_sf = env_int("XG_STRIDED_SEGFAST", 0);
  a.seg_fast = tune_sf != 0;
  const int64_t blocks = xg_ceil_div(a.nunits, kWarpsPerBlock);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_stencil2: grid too large");
  if (!MET && tune_u == 8)
    k_stencil_strided<T, VEC, OP, MET, 8><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  else if (!MET && tune_u == 2)
    k_stencil_strided<T, VEC, OP, MET, 2><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  else
    k_stencil_strided<T, VEC, OP, MET, 4><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  return xg_check_launch("xg_stencil2(strided)");
}

template <typename T, int VEC, int OP, bool MET>
int launch_row_vec(StencilArgs<T>& a, cudaStream_t st) {
  constexpr int U = 4;
  const int64_t nv = a.n / VEC;
  a.nwc = xg_ceil_div(nv, 32 * U);
  a.nunits = a.outer * a.nwc;
  a.small_units = a.nunits < (1ll << 31);
  const int64_t blocks = xg_ceil_div(a.nunits, kWarpsPerBlock);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_stencil2: grid too large");
  k_stencil_row_vec<T, VEC, OP, MET, U><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  return xg_check_launch("xg_stencil2(row_vec)");
}

template <typename T, int OP, bool MET>
int launch_row_scalar(StencilArgs<T>& a, cudaStream_t st) {
  const int64_t total = a.outer * a.n_out;
  int64_t blocks = xg_ceil_div(total, kThreads);
  if (blocks > 148 * 64) blocks = 148 * 64;  // grid-stride beyond that
  k_stencil_row_scalar<T, OP, MET><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  return xg_check_launch("xg_stencil2(row_scalar)");
}

template <typename T, int OP, bool MET>
int dispatch_layout(StencilArgs<T>& a, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  const bool ptr_ok = ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0) &&
                      (!a.halo_lo || (uintptr_t)a.halo_lo % 16 == 0) &&
                      (!a.halo_hi || (uintptr_t)a.halo_hi % 16 == 0);
  if (a.inner > 1) {
    if (ptr_ok && a.inner % VEC == 0) return launch_strided<T, VEC, OP, MET>(a, st);
    a.pre.vec_ok = 0;
    a.post.vec_ok = 0;
    return launch_strided<T, 1, OP, MET>(a, st);
  }
  if (ptr_ok && a.n_out == a.n && a.n % VEC == 0 && a.n / VEC >= 32)
    return launch_row_vec<T, VEC, OP, MET>(a, st);
  return launch_row_scalar<T, OP, MET>(a, st);
}

template <typename T, int OP>
int dispatch_met(StencilArgs<T>& a, cudaStream_t st) {
  if (a.pre.ptr || a.post.ptr) return dispatch_layout<T, OP, true>(a, st);
  return dispatch_layout<T, OP, false>(a, st);
}

template <typename T>
int dispatch_op(int op, StencilArgs<T>& a, cudaStream_t st) {
  switch (op) {
    case XG_OP_DIFF: return dispatch_met<T, XG_OP_DIFF>(a, st);
    case XG_OP_INTERP: return dispatch_met<T, XG_OP_INTERP>(a, st);
    case XG_OP_MIN: return dispatch_met<T, XG_OP_MIN>(a, st);
    case XG_OP_MAX: return dispatch_met<T, XG_OP_MAX>(a, st);
  }
  return xg_fail(XG_EINVAL, "xg_stencil2: unknown op");
}

// vector loads of a metric along x in the row kernels
static int axis_vec_ok(const XgOperand& m, int vec, size_t es) {
  if (!m.ptr || m.axis_stride != 1) return 0;
  if ((uintptr_t)m.ptr % (vec * es) != 0) return 0;
  for (int k = 0; k < m.outer.n; ++k)
    if (m.outer.stride[k] % vec != 0) return 0;
  return 1;
}

template <typename T>
int stencil2_typed(int op, const void* in, void* out, int ndim, const int64_t* shape,
                   int axis, int lo, int hi, int bc, double fill_value,
                   const void* pre_metric, const int64_t* pre_strides,
                   const void* post_metric, const int64_t* post_strides,
                   const void* halo_lo, const void* halo_hi, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  StencilArgs<T> a;
  a.in = static_cast<const T*>(in);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.n_out = v.n + lo + hi - 1;
  a.lo = lo;
  a.hi = hi;
  a.bc = bc;
  a.fill = static_cast<T>(fill_value);
  a.halo_lo = static_cast<const T*>(halo_lo);
  a.halo_hi = static_cast<const T*>(halo_hi);
  a.J = 0;
  a.nseg = a.nwc = a.nunits = 0;
  a.small_units = false;
  a.seg_fast = false;
  if (v.n == 0) return xg_fail(XG_EINVAL, "xg_stencil2: empty operated axis");
  if (v.outer == 0 || v.inner == 0 || a.n_out <= 0) return XG_OK;  // nothing to write

  int64_t out_shape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) out_shape[d] = shape[d];
  out_shape[axis] = a.n_out;
  rc = xg_make_operand(pre_metric, pre_strides, ndim, shape, axis, VEC, sizeof(T), &a.pre,
                       "xg_stencil2(pre_metric)");
  if (rc) return rc;
  rc = xg_make_operand(post_metric, post_strides, ndim, out_shape, axis, VEC, sizeof(T),
                       &a.post, "xg_stencil2(post_metric)");
  if (rc) return rc;
  a.pre_axis_vec_ok = axis_vec_ok(a.pre, VEC, sizeof(T));
  a.post_axis_vec_ok = axis_vec_ok(a.post, VEC, sizeof(T));
  return dispatch_op<T>(op, a, st);
}

}  // namespace

extern "C" int xg_stencil2(int op, int dtype, const void* in, void* out, int ndim,
                           const int64_t* shape, int axis, int lo, int hi, int bc,
                           double fill_value, const void* pre_metric,
                           const int64_t* pre_strides, const void* post_metric,
                           const int64_t* post_strides, const void* halo_lo,
                           const void* halo_hi, void* stream) {
  if (!in || !out) return xg_fail(XG_EINVAL, "xg_stencil2: null field pointer");
  if (!shape) return xg_fail(XG_EINVAL, "xg_stencil2: null shape");
  if (lo < 0 || lo > 1 || hi < 0 || hi > 1)
    return xg_fail(XG_EINVAL, "xg_stencil2: halo widths must be 0 or 1");
  if (bc < XG_BC_NONE || bc > XG_BC_EXTRAPOLATE)
    return xg_fail(XG_EINVAL, "xg_stencil2: unknown boundary condition");
  if ((lo && !halo_lo && bc == XG_BC_NONE) || (hi && !halo_hi && bc == XG_BC_NONE))
    // padding.py:601-608
    return xg_fail(XG_EINVAL,
                   "xg_stencil2: no boundary condition was specified but the "
                   "operation needs to pad the axis");
  if (in == out) return xg_fail(XG_EINVAL, "xg_stencil2: in-place operation is not supported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return stencil2_typed<float>(op, in, out, ndim, shape, axis, lo, hi, bc, fill_value,
                                 pre_metric, pre_strides, post_metric, post_strides,
                                 halo_lo, halo_hi, st);
  if (dtype == XG_F64)
    return stencil2_typed<double>(op, in, out, ndim, shape, axis, lo, hi, bc, fill_value,
                                  pre_metric, pre_strides, post_metric, post_strides,
                                  halo_lo, halo_hi, st);
  return xg_fail(XG_EINVAL, "xg_stencil2: dtype must be XG_F32 or XG_F64");
}
