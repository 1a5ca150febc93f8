// xg_pad, xg_binary, xg_fill_uniform(_host): the copy-shaped helpers around the hot path.
//
//   xg_pad     xgcm/padding.py:575-616 (_pad_basic -> DataArray.pad -> np.pad), one axis
//   xg_binary  the broadcast multiply / divide xarray performs around the kernels
//              (xgcm/grid.py:808,832,1578,1599,1657) for device-resident label algebra
//   xg_fill_uniform  synthetic U(0,1) fields, same bits on host and device (bench / tests)
//
// All HBM-bound: 2 * sizeof(T) bytes per output element, 16-byte accesses where
// alignment allows.
#include "xg_common.cuh"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ pad
template <typename T>
struct PadArgs {
  const T* in;
  T* out;
  int64_t outer, n, inner, n_out;
  int lo, hi, bc;
  T fill;
  int64_t nvec_inner;  // vectors per row of `inner`
  bool small;          // total output vectors < 2^31: 32-bit index math
  XgFastDiv fd_nvi, fd_nout;  // multiply-high forms of nvec_inner / n_out (valid with small)
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) k_pad(const PadArgs<T> a) {
  typedef XgPack<T, VEC> Pack;
  const int64_t total = a.outer * a.n_out * a.nvec_inner;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * kThreads) {
    int64_t iv, t, k, o;
    xg_divmod(g, a.nvec_inner, a.fd_nvi, a.small, t, iv);
    xg_divmod(t, a.n_out, a.fd_nout, a.small, o, k);
    const int64_t i = iv * VEC;
    const T* base = a.in + o * a.n * a.inner + i;
    int64_t s = k - a.lo;
    Pack v;
    bool done = false;
    if (s < 0 || s >= a.n) {
      const bool low = s < 0;
      if (a.bc == XG_BC_FILL) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) v.v[q] = a.fill;
        done = true;
      } else if (a.bc == XG_BC_PERIODIC) {
        s = low ? s + a.n : s - a.n;
        // np.pad(mode="wrap") with a halo wider than the array wraps repeatedly
        s %= a.n;
        if (s < 0) s += a.n;
      } else if (a.bc == XG_BC_EXTEND) {
        s = low ? 0 : a.n - 1;
      } else {  // extrapolate, halo width 1 only
        const int64_t e = low ? 0 : a.n - 1;
        const int64_t e2 = a.n > 1 ? (low ? 1 : a.n - 2) : e;
        Pack p0 = xg_ld_cached<T, VEC>(base + e * a.inner);
        Pack p1 = xg_ld_cached<T, VEC>(base + e2 * a.inner);
#pragma unroll
        for (int q = 0; q < VEC; ++q) v.v[q] = T(2) * p0.v[q] - p1.v[q];
        done = true;
      }
    }
    if (!done) v = xg_ld_stream<T, VEC>(base + s * a.inner);
    xg_st_stream<T, VEC>(a.out + (o * a.n_out + k) * a.inner + i, v);
  }
}

// Padding the INNERMOST dim (inner == 1): output rows are n + lo + hi long, so their starts lose the 16-byte
// alignment of the input rows whatever the widths.  The output is therefore written as ONE flat stream of aligned
// 16-byte vectors (a vector may straddle two rows); each of its elements is gathered with a scalar load (neighbouring
// lanes read neighbouring addresses: the L1 absorbs the 4 passes over each line).  0.36 -> the copy rate of the
// strided-axis pad for odd widths.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) k_pad_rows(const PadArgs<T> a, int64_t total_vec, int64_t total) {
  typedef XgPack<T, VEC> Pack;
  const int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (g >= total_vec) return;
  const int64_t e0 = g * VEC;
  int64_t o, k;
  xg_divmod(e0, a.n_out, a.fd_nout, a.small, o, k);
  Pack v;
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    if (e0 + q < total) {
      int64_t s = k - a.lo;
      const T* row = a.in + o * a.n;
      T val;
      if (s >= 0 && s < a.n) {
        val = __ldg(row + s);
      } else if (a.bc == XG_BC_FILL) {
        val = a.fill;
      } else if (a.bc == XG_BC_PERIODIC) {
        s %= a.n;  // np.pad(mode="wrap") with a halo wider than the array wraps repeatedly
        if (s < 0) s += a.n;
        val = __ldg(row + s);
      } else if (a.bc == XG_BC_EXTEND) {
        val = __ldg(row + (s < 0 ? 0 : a.n - 1));
      } else {  // extrapolate, halo width 1 only
        const int64_t e = s < 0 ? 0 : a.n - 1;
        const int64_t e2 = a.n > 1 ? (s < 0 ? 1 : a.n - 2) : e;
        val = T(2) * __ldg(row + e) - __ldg(row + e2);
      }
      v.v[q] = val;
    } else {
      v.v[q] = T(0);
    }
    if (++k == a.n_out) {  // next element starts the next row
      k = 0;
      ++o;
    }
  }
  if (e0 + VEC <= total) {
    xg_st_stream<T, VEC>(a.out + e0, v);
  } else {
#pragma unroll
    for (int q = 0; q < VEC; ++q)
      if (e0 + q < total) a.out[e0 + q] = v.v[q];
  }
}

template <typename T>
int pad_typed(const void* in, void* out, int ndim, const int64_t* shape, int axis, int lo,
              int hi, int bc, double fill, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  if (v.n == 0 && bc != XG_BC_FILL)
    return xg_fail(XG_EINVAL, "xg_pad: cannot wrap/extend an empty axis");
  PadArgs<T> a;
  a.in = static_cast<const T*>(in);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.n_out = v.n + lo + hi;
  a.lo = lo;
  a.hi = hi;
  a.bc = bc;
  a.fill = static_cast<T>(fill);
  if (v.outer == 0 || v.inner == 0 || a.n_out == 0) return XG_OK;
  if (v.inner == 1 && v.n > 0 && ((uintptr_t)out % 16 == 0) && a.n_out >= 8) {
    const int64_t total = a.outer * a.n_out;
    const int64_t total_vec = xg_ceil_div(total, VEC);
    a.nvec_inner = 1;
    a.small = total < (1ll << 31);
    a.fd_nvi = xg_fastdiv_make(1);
    a.fd_nout = xg_fastdiv_make(a.small ? a.n_out : 1);
    const int64_t blocks = xg_ceil_div(total_vec, kThreads);
    if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_pad: grid too large");
    k_pad_rows<T, VEC><<<(unsigned)blocks, kThreads, 0, st>>>(a, total_vec, total);
    return xg_check_launch("xg_pad(rows)");
  }
  const bool vec_ok = (v.inner % VEC == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
  a.nvec_inner = vec_ok ? v.inner / VEC : v.inner;
  const int64_t total = a.outer * a.n_out * a.nvec_inner;
  a.small = total < (1ll << 31);
  a.fd_nvi = xg_fastdiv_make(a.small ? a.nvec_inner : 1);
  a.fd_nout = xg_fastdiv_make(a.small ? a.n_out : 1);
  int64_t blocks = xg_ceil_div(total, kThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec_ok)
    k_pad<T, VEC><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  else
    k_pad<T, 1><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  return xg_check_launch("xg_pad");
}

// ------------------------------------------------------------------ binary
template <typename T>
struct BinArgs {
  const T* a;
  T* out;
  int64_t rows, n, nvec;
  XgOperand b;  // outer groups over rows, axis_stride along the last dim
  int b_vec_ok;
  bool small;
  XgFastDiv fd_nvec;  // multiply-high form of nvec (valid with small)
};

template <typename T, int OP>
__device__ __forceinline__ T bin_apply(T x, T y) {
  if constexpr (OP == XG_BIN_MUL) return x * y;
  else if constexpr (OP == XG_BIN_DIV) return x / y;
  else if constexpr (OP == XG_BIN_DIVNZ) return (y != T(0)) ? x / y : T(NAN);
  else if constexpr (OP == XG_BIN_ADD) return x + y;
  else return x - y;
}

template <typename T, int VEC, int OP>
__global__ void __launch_bounds__(kThreads) k_binary(const BinArgs<T> a) {
  typedef XgPack<T, VEC> Pack;
  const int64_t total = a.rows * a.nvec;
  const T* bp = reinterpret_cast<const T*>(a.b.ptr);
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * kThreads) {
    int64_t r, xq;
    xg_divmod(g, a.nvec, a.fd_nvec, a.small, r, xq);
    const int64_t x0 = xq * VEC;
    const int64_t boff = xg_groups_offset(a.b.outer, r);
    Pack va = xg_ld_stream<T, VEC>(a.a + r * a.n + x0);
    Pack vb;
    if (a.b.axis_stride == 0) {
      T s = __ldg(bp + boff);
#pragma unroll
      for (int q = 0; q < VEC; ++q) vb.v[q] = s;
    } else if (VEC > 1 && a.b_vec_ok) {
      vb = xg_ld_cached<T, VEC>(bp + boff + x0);
    } else {
#pragma unroll
      for (int q = 0; q < VEC; ++q) vb.v[q] = __ldg(bp + boff + (x0 + q) * a.b.axis_stride);
    }
    Pack r_;
#pragma unroll
    for (int q = 0; q < VEC; ++q) r_.v[q] = bin_apply<T, OP>(va.v[q], vb.v[q]);
    xg_st_stream<T, VEC>(a.out + r * a.n + x0, r_);
  }
}

template <typename T, int VEC>
int binary_launch(int binop, BinArgs<T>& a, cudaStream_t st) {
  const int64_t total = a.rows * a.nvec;
  a.small = total < (1ll << 31);
  a.fd_nvec = xg_fastdiv_make(a.small ? a.nvec : 1);
  int64_t blocks = xg_ceil_div(total, kThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  switch (binop) {
    case XG_BIN_MUL: k_binary<T, VEC, XG_BIN_MUL><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
    case XG_BIN_DIV: k_binary<T, VEC, XG_BIN_DIV><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
    case XG_BIN_ADD: k_binary<T, VEC, XG_BIN_ADD><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
    case XG_BIN_SUB: k_binary<T, VEC, XG_BIN_SUB><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
    case XG_BIN_DIVNZ: k_binary<T, VEC, XG_BIN_DIVNZ><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
    default: return xg_fail(XG_EINVAL, "xg_binary: unknown operator");
  }
  return xg_check_launch("xg_binary");
}

template <typename T>
int binary_typed(int binop, const void* pa, const void* pb, const int64_t* b_strides, void* out,
                 int ndim, const int64_t* shape, cudaStream_t st) {
  constexpr int VEC = XgVecWidth<T>::value;
  XgView v;
  int rc = xg_collapse_view(ndim, shape, ndim - 1, &v);
  if (rc) return rc;
  BinArgs<T> a;
  a.a = static_cast<const T*>(pa);
  a.out = static_cast<T*>(out);
  a.rows = v.outer;
  a.n = v.n;
  rc = xg_make_operand(pb, b_strides, ndim, shape, ndim - 1, VEC, sizeof(T), &a.b, "xg_binary(b)");
  if (rc) return rc;
  if (a.rows == 0 || a.n == 0) return XG_OK;
  bool vec_ok = (a.n % VEC == 0) && ((uintptr_t)pa % 16 == 0) && ((uintptr_t)out % 16 == 0);
  bool bvec = a.b.axis_stride == 1 && ((uintptr_t)pb % 16 == 0);
  for (int k = 0; k < a.b.outer.n; ++k) bvec = bvec && (a.b.outer.stride[k] % VEC == 0);
  a.b_vec_ok = bvec ? 1 : 0;
  if (vec_ok) {
    a.nvec = a.n / VEC;
    return binary_launch<T, VEC>(binop, a, st);
  }
  a.nvec = a.n;
  a.b_vec_ok = 0;
  return binary_launch<T, 1>(binop, a, st);
}

// ------------------------------------------------------------------ uniform fill
__host__ __device__ __forceinline__ uint64_t xg_mix64(uint64_t z) {
  // splitmix64 finaliser: counter-based, so any sub-block can be generated anywhere
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T>
__host__ __device__ __forceinline__ T xg_uniform(uint64_t seed, uint64_t idx) {
  const uint64_t h = xg_mix64(xg_mix64(seed) ^ (idx * 0x9E3779B97F4A7C15ull));
  if (sizeof(T) == 4) return (T)((float)(h >> 40) * (1.0f / 16777216.0f));
  return (T)((double)(h >> 11) * (1.0 / 9007199254740992.0));
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_fill_uniform(T* out, int64_t count, uint64_t seed, uint64_t offset) {
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < count;
       g += (int64_t)gridDim.x * kThreads)
    out[g] = xg_uniform<T>(seed, offset + (uint64_t)g);
}

}  // namespace

extern "C" int xg_pad(int dtype, const void* in, void* out, int ndim, const int64_t* shape,
                      int axis, int lo, int hi, int bc, double fill_value, void* stream) {
  if (!in || !out || !shape) return xg_fail(XG_EINVAL, "xg_pad: null pointer");
  if (lo < 0 || hi < 0) return xg_fail(XG_EINVAL, "xg_pad: negative halo width");
  if ((lo || hi) && (bc <= XG_BC_NONE || bc > XG_BC_EXTRAPOLATE))
    return xg_fail(XG_EINVAL,
                   "xg_pad: no boundary condition was specified but the operation needs to pad "
                   "the axis");
  if (bc == XG_BC_EXTRAPOLATE && (lo > 1 || hi > 1))
    return xg_fail(XG_ENOTIMPL, "xg_pad: extrapolate supports halo width 1 only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32) return pad_typed<float>(in, out, ndim, shape, axis, lo, hi, bc, fill_value, st);
  if (dtype == XG_F64) return pad_typed<double>(in, out, ndim, shape, axis, lo, hi, bc, fill_value, st);
  return xg_fail(XG_EINVAL, "xg_pad: dtype must be XG_F32 or XG_F64");
}

extern "C" int xg_binary(int binop, int dtype, const void* a, const void* b,
                         const int64_t* b_strides, void* out, int ndim, const int64_t* shape,
                         void* stream) {
  if (!a || !b || !out || !shape || !b_strides) return xg_fail(XG_EINVAL, "xg_binary: null pointer");
  if (ndim < 1) return xg_fail(XG_EINVAL, "xg_binary: ndim must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32) return binary_typed<float>(binop, a, b, b_strides, out, ndim, shape, st);
  if (dtype == XG_F64) return binary_typed<double>(binop, a, b, b_strides, out, ndim, shape, st);
  return xg_fail(XG_EINVAL, "xg_binary: dtype must be XG_F32 or XG_F64");
}

extern "C" int xg_fill_uniform(int dtype, void* out, int64_t count, uint64_t seed, uint64_t offset,
                               void* stream) {
  if (!out && count) return xg_fail(XG_EINVAL, "xg_fill_uniform: null pointer");
  if (count <= 0) return XG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = xg_ceil_div(count, kThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == XG_F32)
    k_fill_uniform<float><<<(unsigned)blocks, kThreads, 0, st>>>((float*)out, count, seed, offset);
  else if (dtype == XG_F64)
    k_fill_uniform<double><<<(unsigned)blocks, kThreads, 0, st>>>((double*)out, count, seed, offset);
  else
    return xg_fail(XG_EINVAL, "xg_fill_uniform: dtype must be XG_F32 or XG_F64");
  return xg_check_launch("xg_fill_uniform");
}

extern "C" int xg_fill_uniform_host(int dtype, void* out, int64_t count, uint64_t seed,
                                    uint64_t offset) {
  if (!out && count) return xg_fail(XG_EINVAL, "xg_fill_uniform_host: null pointer");
  if (dtype == XG_F32) {
    float* p = (float*)out;
    for (int64_t g = 0; g < count; ++g) p[g] = xg_uniform<float>(seed, offset + (uint64_t)g);
  } else if (dtype == XG_F64) {
    double* p = (double*)out;
    for (int64_t g = 0; g < count; ++g) p[g] = xg_uniform<double>(seed, offset + (uint64_t)g);
  } else {
    return xg_fail(XG_EINVAL, "xg_fill_uniform_host: dtype must be XG_F32 or XG_F64");
  }
  return XG_OK;
}
