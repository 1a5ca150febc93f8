// xg_cumscan — cumulative sum with xgcm's position-shift bookkeeping, one HBM pass.
//
// Replaces xgcm/grid.py:1306-1414:  (x metric) -> [flip] -> DataArray.cumsum -> [flip]
// -> trim (table :1326-1383) -> pad() of the cumsum'd data (:1385-1391) -> (/ metric).
//
// Summation order is STRICTLY SEQUENTIAL along the axis, because np.cumsum is
// (on every layout) and a tree scan already differs by 2e-6 rel in fp32 at
// n = 3600 — outside the 1e-6 parity budget.  Parallelism comes from
// the independent lines instead:
//   k_scan_strided  inner > 1: one thread per 16-byte column vector marches the
//       axis; U independent loads in flight, only the adds are serial.
//   k_scan_rows     inner == 1: a warp owns 32 rows; 32x32 tiles go through
//       shared memory (coalesced 128 B row segments in, lane-per-row serial scan,
//       coalesced segments out) with the running sums carried in registers.
//
// Roofline: HBM, 2 * sizeof(T) bytes per cell.
#include "xg_common.cuh"

namespace {

constexpr int kThreads = 128;

template <typename T>
struct ScanArgs {
  const T* in;
  T* out;
  int64_t outer, n, inner, n_out;
  int64_t k_first, k_last;  // kept range of the cumsum (after trim), inclusive
  int reverse, pad_lo, pad_hi, bc, skipna;
  T fill;
  XgOperand pre, post;
  int64_t nvec_inner;
};

template <typename T>
__device__ __forceinline__ T nan_to_zero(T v, int skipna) {
  return (skipna && xg_isnan(v)) ? T(0) : v;
}

// value of a halo cell from the recorded ends of the trimmed cumsum
template <typename T>
__device__ __forceinline__ T halo_value(bool low, int bc, T fill, T cf, T cf1, T cl1, T cl) {
  if (bc == XG_BC_FILL) return fill;
  if (bc == XG_BC_PERIODIC) return low ? cl : cf;
  if (bc == XG_BC_EXTEND) return low ? cf : cl;
  return low ? (T(2) * cf - cf1) : (T(2) * cl - cl1);  // extrapolate
}

// ------------------------------------------------------------------ strided axis
template <typename T, int VEC, bool MET, int U>
__global__ void __launch_bounds__(kThreads) k_scan_strided(const ScanArgs<T> a) {
  typedef XgPack<T, VEC> Pack;
  const int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (g >= a.outer * a.nvec_inner) return;
  const int64_t o = g / a.nvec_inner;
  const int64_t i = (g - o * a.nvec_inner) * VEC;
  const T* ibase = a.in + o * a.n * a.inner + i;
  T* obase = a.out + o * a.n_out * a.inner + i;
  int64_t pre_base = 0, post_base = 0;
  if (MET) {
    if (a.pre.ptr) pre_base = xg_groups_offset(a.pre.outer, o);
    if (a.post.ptr) post_base = xg_groups_offset(a.post.outer, o);
  }
  auto loadA = [&](int64_t k) -> Pack {
    Pack v = xg_ld_stream<T, VEC>(ibase + k * a.inner);
    if (MET && a.pre.ptr) {
      Pack m = xg_ld_operand<T, VEC>(a.pre, pre_base + k * a.pre.axis_stride, i);
#pragma unroll
      for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] * m.v[q];
    }
    return v;
  };
  auto store = [&](int64_t j_out, Pack v) {
    if (MET && a.post.ptr) {
      Pack m = xg_ld_operand<T, VEC>(a.post, post_base + j_out * a.post.axis_stride, i);
#pragma unroll
      for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] / m.v[q];
    }
    xg_st_stream<T, VEC>(obase + j_out * a.inner, v);
  };

  Pack acc, cf, cf1, cl1, cl;
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc.v[q] = cf.v[q] = cf1.v[q] = cl1.v[q] = cl.v[q] = T(0);

  auto step = [&](int64_t k, const Pack& v) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc.v[q] = acc.v[q] + nan_to_zero(v.v[q], a.skipna);
    if (k == a.k_first) cf = acc;
    if (k == a.k_first + 1) cf1 = acc;
    if (k == a.k_last - 1) cl1 = acc;
    if (k == a.k_last) cl = acc;
    if (k >= a.k_first && k <= a.k_last) store(a.pad_lo + (k - a.k_first), acc);
  };

  int64_t kk = 0;
  for (; kk + U <= a.n; kk += U) {
    Pack v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = loadA(a.reverse ? (a.n - 1 - (kk + u)) : (kk + u));
#pragma unroll
    for (int u = 0; u < U; ++u) step(a.reverse ? (a.n - 1 - (kk + u)) : (kk + u), v[u]);
  }
  for (; kk < a.n; ++kk) {
    const int64_t k = a.reverse ? (a.n - 1 - kk) : kk;
    step(k, loadA(k));
  }
  if (a.k_last - a.k_first < 1) {  // a single kept cell: "next" is the edge itself
    cf1 = cf;
    cl1 = cl;
  }
  if (a.pad_lo) {
    Pack h;
#pragma unroll
    for (int q = 0; q < VEC; ++q)
      h.v[q] = halo_value<T>(true, a.bc, a.fill, cf.v[q], cf1.v[q], cl1.v[q], cl.v[q]);
    store(0, h);
  }
  if (a.pad_hi) {
    Pack h;
#pragma unroll
    for (int q = 0; q < VEC; ++q)
      h.v[q] = halo_value<T>(false, a.bc, a.fill, cf.v[q], cf1.v[q], cl1.v[q], cl.v[q]);
    store(a.n_out - 1, h);
  }
}

// ------------------------------------------------------------------ innermost axis
constexpr int kRowWarps = 4;
constexpr int kTile = 32;

template <typename T, bool MET>
__global__ void __launch_bounds__(kRowWarps * 32) k_scan_rows(const ScanArgs<T> a) {
  __shared__ T tile_s[kRowWarps][kTile][kTile + 1];
  __shared__ int64_t pre_off_s[kRowWarps][kTile];
  __shared__ int64_t post_off_s[kRowWarps][kTile];
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * kRowWarps + w;
  const int64_t r0 = unit * kTile;
  if (r0 >= a.outer) return;  // warp-uniform
  T(*tile)[kTile + 1] = tile_s[w];
  const int64_t my_row = r0 + lane;
  const bool row_ok = my_row < a.outer;
  const int nrows = (int)((a.outer - r0 < kTile) ? (a.outer - r0) : kTile);
  if (MET) {
    pre_off_s[w][lane] = (a.pre.ptr && row_ok) ? xg_groups_offset(a.pre.outer, my_row) : 0;
    post_off_s[w][lane] = (a.post.ptr && row_ok) ? xg_groups_offset(a.post.outer, my_row) : 0;
  }
  __syncwarp();
  const T* prep = reinterpret_cast<const T*>(a.pre.ptr);
  const T* postp = reinterpret_cast<const T*>(a.post.ptr);

  T acc = T(0), cf = T(0), cf1 = T(0), cl1 = T(0), cl = T(0);
  const int64_t ntile = xg_ceil_div(a.n, kTile);
  for (int64_t tt = 0; tt < ntile; ++tt) {
    const int64_t c0 = (a.reverse ? (ntile - 1 - tt) : tt) * kTile;
    const int64_t kcol = c0 + lane;
    const bool col_ok = kcol < a.n;
    // coalesced load: one 32-element row segment per instruction, 32 independent loads
#pragma unroll 8
    for (int rr = 0; rr < kTile; ++rr) {
      T v = T(0);
      if (rr < nrows && col_ok) {
        v = __ldcs(a.in + (r0 + rr) * a.n + kcol);
        if (MET && prep) v = v * __ldg(prep + pre_off_s[w][rr] + kcol * a.pre.axis_stride);
      }
      tile[rr][lane] = v;
    }
    __syncwarp();
    // serial scan: lane = row
    if (row_ok) {
#pragma unroll 8
      for (int cc = 0; cc < kTile; ++cc) {
        const int c = a.reverse ? (kTile - 1 - cc) : cc;
        const int64_t k = c0 + c;
        if (k < a.n) {
          acc = acc + nan_to_zero(tile[lane][c], a.skipna);
          tile[lane][c] = acc;
          if (k == a.k_first) cf = acc;
          if (k == a.k_first + 1) cf1 = acc;
          if (k == a.k_last - 1) cl1 = acc;
          if (k == a.k_last) cl = acc;
        }
      }
    }
    __syncwarp();
    // coalesced store of the kept cells (shifted by pad_lo - k_first)
    if (col_ok && kcol >= a.k_first && kcol <= a.k_last) {
      const int64_t j_out = a.pad_lo + (kcol - a.k_first);
#pragma unroll 8
      for (int rr = 0; rr < kTile; ++rr) {
        if (rr < nrows) {
          T v = tile[rr][lane];
          if (MET && postp) v = v / __ldg(postp + post_off_s[w][rr] + j_out * a.post.axis_stride);
          __stcs(a.out + (r0 + rr) * a.n_out + j_out, v);
        }
      }
    }
    __syncwarp();
  }
  if (!row_ok) return;
  if (a.k_last - a.k_first < 1) {
    cf1 = cf;
    cl1 = cl;
  }
  if (a.pad_lo) {
    T h = halo_value<T>(true, a.bc, a.fill, cf, cf1, cl1, cl);
    if (MET && postp) h = h / __ldg(postp + post_off_s[w][lane]);
    a.out[my_row * a.n_out] = h;
  }
  if (a.pad_hi) {
    T h = halo_value<T>(false, a.bc, a.fill, cf, cf1, cl1, cl);
    if (MET && postp)
      h = h / __ldg(postp + post_off_s[w][lane] + (a.n_out - 1) * a.post.axis_stride);
    a.out[my_row * a.n_out + a.n_out - 1] = h;
  }
}

template <typename T, bool MET>
int scan_launch(ScanArgs<T>& a, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  constexpr int U = 8;
  if (a.inner > 1) {
    bool vec_ok = (a.inner % VEC == 0) && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
    // few columns: prefer 4x more (scalar) threads over 16-byte accesses
    if (vec_ok && a.outer * (a.inner / VEC) < 148 * 512) vec_ok = false;
    if (vec_ok) {
      a.nvec_inner = a.inner / VEC;
      const int64_t blocks = xg_ceil_div(a.outer * a.nvec_inner, kThreads);
      if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_cumscan: grid too large");
      k_scan_strided<T, VEC, MET, U><<<(unsigned)blocks, kThreads, 0, st>>>(a);
    } else {
      a.pre.vec_ok = 0;
      a.post.vec_ok = 0;
      a.nvec_inner = a.inner;
      const int64_t blocks = xg_ceil_div(a.outer * a.nvec_inner, kThreads);
      if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_cumscan: grid too large");
      k_scan_strided<T, 1, MET, U><<<(unsigned)blocks, kThreads, 0, st>>>(a);
    }
    return xg_check_launch("xg_cumscan(strided)");
  }
  const int64_t units = xg_ceil_div(a.outer, kTile);
  const int64_t blocks = xg_ceil_div(units, kRowWarps);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_cumscan: grid too large");
  k_scan_rows<T, MET><<<(unsigned)blocks, kRowWarps * 32, 0, st>>>(a);
  return xg_check_launch("xg_cumscan(rows)");
}

template <typename T>
int cumscan_typed(const void* in, void* out, int ndim, const int64_t* shape, int axis, int reverse,
                  int trim, int pad_lo, int pad_hi, int bc, double fill, const void* pre_metric,
                  const int64_t* pre_strides, const void* post_metric,
                  const int64_t* post_strides, int skipna, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  ScanArgs<T> a;
  a.in = static_cast<const T*>(in);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.k_first = (trim == XG_TRIM_DROP_FIRST) ? 1 : 0;
  a.k_last = v.n - 1 - ((trim == XG_TRIM_DROP_LAST) ? 1 : 0);
  const int64_t kept = a.k_last - a.k_first + 1;
  if (kept < 0) return xg_fail(XG_EINVAL, "xg_cumscan: operated axis too short to trim");
  a.n_out = kept + pad_lo + pad_hi;
  a.reverse = reverse ? 1 : 0;
  a.pad_lo = pad_lo;
  a.pad_hi = pad_hi;
  a.bc = bc;
  a.skipna = skipna ? 1 : 0;
  a.fill = static_cast<T>(fill);
  a.nvec_inner = 0;
  if (kept == 0 && (pad_lo || pad_hi) && bc != XG_BC_FILL)
    return xg_fail(XG_EINVAL, "xg_cumscan: cannot wrap/extend an empty axis");
  if (v.outer == 0 || v.inner == 0 || a.n_out == 0) return XG_OK;
  int64_t out_shape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) out_shape[d] = shape[d];
  out_shape[axis] = a.n_out;
  rc = xg_make_operand(pre_metric, pre_strides, ndim, shape, axis, VEC, sizeof(T), &a.pre,
                       "xg_cumscan(pre_metric)");
  if (rc) return rc;
  rc = xg_make_operand(post_metric, post_strides, ndim, out_shape, axis, VEC, sizeof(T), &a.post,
                       "xg_cumscan(post_metric)");
  if (rc) return rc;
  if (a.pre.ptr || a.post.ptr) return scan_launch<T, true>(a, st);
  return scan_launch<T, false>(a, st);
}

}  // namespace

extern "C" int xg_cumscan(int dtype, const void* in, void* out, int ndim, const int64_t* shape,
                          int axis, int reverse, int trim, int pad_lo, int pad_hi, int bc,
                          double fill_value, const void* pre_metric, const int64_t* pre_strides,
                          const void* post_metric, const int64_t* post_strides, int skipna,
                          void* stream) {
  if (!in || !out || !shape) return xg_fail(XG_EINVAL, "xg_cumscan: null pointer");
  if (pad_lo < 0 || pad_lo > 1 || pad_hi < 0 || pad_hi > 1)
    return xg_fail(XG_EINVAL, "xg_cumscan: halo widths must be 0 or 1");
  if (trim < XG_TRIM_NONE || trim > XG_TRIM_DROP_FIRST)
    return xg_fail(XG_EINVAL, "xg_cumscan: unknown trim mode");
  if ((pad_lo || pad_hi) && (bc <= XG_BC_NONE || bc > XG_BC_EXTRAPOLATE))
    // padding.py:601-608
    return xg_fail(XG_EINVAL,
                   "xg_cumscan: no boundary condition was specified but the operation needs to "
                   "pad the axis");
  if (in == out) return xg_fail(XG_EINVAL, "xg_cumscan: in-place operation is not supported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return cumscan_typed<float>(in, out, ndim, shape, axis, reverse, trim, pad_lo, pad_hi, bc,
                                fill_value, pre_metric, pre_strides, post_metric, post_strides,
                                skipna, st);
  if (dtype == XG_F64)
    return cumscan_typed<double>(in, out, ndim, shape, axis, reverse, trim, pad_lo, pad_hi, bc,
                                 fill_value, pre_metric, pre_strides, post_metric, post_strides,
                                 skipna, st);
  return xg_fail(XG_EINVAL, "xg_cumscan: dtype must be XG_F32 or XG_F64");
}
