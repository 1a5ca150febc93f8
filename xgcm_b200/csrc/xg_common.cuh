// Shared device/host helpers for the xgcm_b200 kernels (sm_100a).
//
// Everything here is plain LD/ST work: the path is HBM-bound (<= 0.25 flop/B),
// so there are no tensor-core instructions by design.  Arithmetic must round
// exactly like numpy does one ufunc at a time, therefore the library is built
// with --fmad=false and IEEE division (no fast-math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "xgcm_b200.h"

#define XG_MAXG 4  // max collapsed dim groups when decomposing a flat index

// ---------------------------------------------------------------------------
// error plumbing (thread-local message, never abort)
// ---------------------------------------------------------------------------
void xg_set_error(const std::string& msg);
int xg_fail(int code, const std::string& msg);
int xg_check_launch(const char* what);

// ---------------------------------------------------------------------------
// broadcast operand descriptor
// ---------------------------------------------------------------------------
struct XgGroups {
  int n;
  int small;  // every size and the total extent fit in 31 bits: 32-bit index math
  int64_t size[XG_MAXG];
  int64_t stride[XG_MAXG];
};

enum { XG_IM_BCAST = 0, XG_IM_CONTIG = 1, XG_IM_GENERIC = 2 };

// A metric (or weight / theta) operand broadcast against a field collapsed to
// (outer, n, inner).
struct XgOperand {
  const void* ptr;      // nullptr = absent
  XgGroups outer;       // flat outer index -> element offset
  int64_t axis_stride;  // element stride along the operated axis
  XgGroups inner;       // flat inner index -> element offset
  int inner_mode;       // XG_IM_*
  int vec_ok;           // CONTIG and every offset is a multiple of the vector width
};

struct XgView {
  int64_t outer, n, inner;
};

// host: collapse (shape, axis) and an operand's per-dim strides.
int xg_collapse_view(int ndim, const int64_t* shape, int axis, XgView* v);
int xg_make_operand(const void* ptr, const int64_t* strides, int ndim,
                    const int64_t* shape, int axis, int vec, size_t elem_size,
                    XgOperand* op, const char* what);

__host__ __device__ __forceinline__ int64_t xg_groups_offset(const XgGroups& g,
                                                             int64_t flat) {
  int64_t off = 0;
  if (g.n == 0) return 0;
  if (g.small) {  // 32-bit divisions are ~5x cheaper than 64-bit ones on the SM
    uint32_t f = (uint32_t)flat;
#pragma unroll
    for (int k = XG_MAXG - 1; k >= 0; --k) {
      if (k < g.n) {
        const uint32_t sz = (uint32_t)g.size[k];
        const uint32_t q = f / sz;
        off += (int64_t)(f - q * sz) * g.stride[k];
        f = q;
      }
    }
    return off;
  }
#pragma unroll
  for (int k = XG_MAXG - 1; k >= 0; --k) {
    if (k < g.n) {
      int64_t q = flat / g.size[k];
      off += (flat - q * g.size[k]) * g.stride[k];
      flat = q;
    }
  }
  return off;
}

// flat -> (q, r) with a 32-bit fast path (warp-unit decomposition)
__device__ __forceinline__ void xg_divmod(int64_t x, int64_t d, bool small, int64_t& q, int64_t& r) {
  if (small) {
    const uint32_t qq = (uint32_t)x / (uint32_t)d;
    q = qq;
    r = (uint32_t)x - qq * (uint32_t)d;
  } else {
    q = x / d;
    r = x - q * d;
  }
}

// ---------------------------------------------------------------------------
// vector types: 16-byte accesses for both dtypes
// ---------------------------------------------------------------------------
template <typename T, int VEC>
struct XgVec;
template <>
struct XgVec<float, 4> {
  typedef float4 type;
};
template <>
struct XgVec<double, 2> {
  typedef double2 type;
};
template <>
struct XgVec<float, 1> {
  typedef float type;
};
template <>
struct XgVec<double, 1> {
  typedef double type;
};

template <typename T>
struct XgVecWidth;
template <>
struct XgVecWidth<float> {
  static const int value = 4;
};
template <>
struct XgVecWidth<double> {
  static const int value = 2;
};

template <typename T, int VEC>
struct XgPack {
  T v[VEC];
};

// streaming (evict-first) global accesses: every field element is touched once.
template <typename T, int VEC>
__device__ __forceinline__ XgPack<T, VEC> xg_ld_stream(const T* p) {
  XgPack<T, VEC> r;
  if constexpr (VEC == 1) {
    r.v[0] = __ldcs(p);
  } else {
    typedef typename XgVec<T, VEC>::type V;
    V t = __ldcs(reinterpret_cast<const V*>(p));
    const T* tp = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = tp[k];
  }
  return r;
}

// cached (read-only path) accesses: metrics and halo rows are re-used.
template <typename T, int VEC>
__device__ __forceinline__ XgPack<T, VEC> xg_ld_cached(const T* p) {
  XgPack<T, VEC> r;
  if constexpr (VEC == 1) {
    r.v[0] = __ldg(p);
  } else {
    typedef typename XgVec<T, VEC>::type V;
    V t = __ldg(reinterpret_cast<const V*>(p));
    const T* tp = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = tp[k];
  }
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void xg_st_stream(T* p, const XgPack<T, VEC>& r) {
  if constexpr (VEC == 1) {
    __stcs(p, r.v[0]);
  } else {
    typedef typename XgVec<T, VEC>::type V;
    V t;
    T* tp = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int k = 0; k < VEC; ++k) tp[k] = r.v[k];
    __stcs(reinterpret_cast<V*>(p), t);
  }
}

// Per-thread view of a broadcast operand for the VEC elements starting at flat inner index i:
// computed ONCE per thread (i is fixed while a thread marches along the axis), so the per-row
// cost of a fused metric is one (vector) load and VEC multiplies / divides.  Kept small on
// purpose (a pointer + VEC-1 32-bit deltas): register pressure decides the occupancy of the
// fused kernels.
template <typename T, int VEC>
struct XgOperandView {
  const T* p0;       // operand pointer + outer offset + inner offset of element 0
  int d[VEC];        // element k sits at p0[d[k]] (d[0] = 0); generic mode only
  int mode;          // XG_IM_* ; CONTIG with `vec` -> one 16-byte load
  bool vec;
};

template <typename T, int VEC>
__device__ __forceinline__ XgOperandView<T, VEC> xg_operand_view(const XgOperand& m, int64_t outer_off,
                                                                 int64_t i) {
  XgOperandView<T, VEC> r;
  const T* p = reinterpret_cast<const T*>(m.ptr) + outer_off;
  r.mode = m.inner_mode;
  r.vec = false;
#pragma unroll
  for (int k = 0; k < VEC; ++k) r.d[k] = 0;
  if (m.inner_mode == XG_IM_BCAST) {
    r.p0 = p;
  } else if (m.inner_mode == XG_IM_CONTIG) {
    r.p0 = p + i;
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.d[k] = k;
    r.vec = VEC > 1 && m.vec_ok;
  } else {
    const int64_t o0 = xg_groups_offset(m.inner, i);
    r.p0 = p + o0;
#pragma unroll
    for (int k = 1; k < VEC; ++k) r.d[k] = (int)(xg_groups_offset(m.inner, i + k) - o0);
  }
  return r;
}

// the operand's VEC values `row_off` elements further along the operated axis
template <typename T, int VEC>
__device__ __forceinline__ XgPack<T, VEC> xg_ld_view(const XgOperandView<T, VEC>& v, int64_t row_off) {
  const T* p = v.p0 + row_off;
  XgPack<T, VEC> r;
  if (v.mode == XG_IM_BCAST) {
    const T s = __ldg(p);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = s;
  } else if (v.vec) {
    r = xg_ld_cached<T, VEC>(p);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = __ldg(p + v.d[k]);
  }
  return r;
}

// ---------------------------------------------------------------------------
// the four pairwise operators, rounding exactly like numpy
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool xg_isnan(T x) {
  return x != x;
}

// a = P[j] (lower neighbour), b = P[j+1] (upper neighbour)
template <typename T, int OP>
__device__ __forceinline__ T xg_apply_op(T a, T b) {
  if constexpr (OP == XG_OP_DIFF) {
    return b - a;  // gridops.py:24  a[...,1:] - a[...,:-1]
  } else if constexpr (OP == XG_OP_INTERP) {
    return (a + b) * T(0.5);  // gridops.py:77 (a[:-1]+a[1:])/2.0 ; x*0.5 == x/2 in IEEE
  } else if constexpr (OP == XG_OP_MIN) {
    // gridops.py:123-126 np.min over the stacked pair: NaN propagates
    return (a < b || xg_isnan(a)) ? a : b;
  } else {
    return (a > b || xg_isnan(a)) ? a : b;
  }
}

__host__ __device__ static inline int64_t xg_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
