// Shared pieces of the vertical-interpolation kernels (xg_vinterp.cu, xg_vinterp_tma.cu).
#pragma once
#include "xg_common.cuh"

namespace xgvi {


constexpr int kTile = 32;
constexpr int LIKELY_IN_CACHE_SIZE = 8;
constexpr int kPrefetchRows = 6;  // theta-field kernel: phi / theta rows pulled into L1 ahead of the walk

template <typename T>
struct InterpArgs {
  const T* phi;
  T* out;
  int64_t outer, n, inner, m;
  XgOperand theta;
  XgOperand target;  // levels: shared 1-D vector or one vector per column (axis_stride = level stride)
  int mask_edges, bypass_checks, logarithmic;
  int64_t ntiles;    // column tiles of 32
  bool small_cols;   // outer * inner < 2^31
  bool theta_full;   // theta is a C-contiguous field of phi's shape (same strides)
};

template <typename T>
__device__ __forceinline__ T xg_log(T x);
// float32 natural log with numpy's bits.  `method="log"` takes np.log of theta and of the target levels in the
// FIELD dtype before interpolating (xgcm/transform.py:82-84); a 1-ulp difference in that log is amplified by
// 1 / (log spacing of the levels) in the interpolation weight, so CUDA's logf (<= 1 ulp, but not numpy's
// rounding) only reached ~5e-5.  numpy's float32 log on x86 (AVX2+FMA3 and AVX512F dispatch targets, the only
// ones a GPU host has; numpy/_core/src/umath/loops_exponent_log.dispatch.c.src) is a rational minimax
// approximation evaluated with FMAs; this is the same computation, operation for operation — checked bit for
// bit against np.log on 88 M float32 values covering every exponent, denormals, the reduction threshold and the
// neighbourhood of 1 (tests/test_oracle_golden.py::test_log32_port_is_numpys holds the numpy restatement).
template <>
__device__ __forceinline__ float xg_log<float>(float x) {
  if (x != x) return x;
  if (x < 0.0f) return NAN;
  if (x == 0.0f) return -INFINITY;
  if (x == INFINITY) return x;
  int e;
  float m = frexpf(x, &e);  // m in [0.5, 1), denormals normalised
  if (m <= 0.70710678118654752440f) {  // mantissa in (1/sqrt2, sqrt2]
    m = m * 2.0f;
    e -= 1;
  }
  const float t = m - 1.0f;
  float num = 2.589979117907922693523e-02f;
  num = fmaf(num, t, 3.808837741388407920751e-01f);
  num = fmaf(num, t, 1.480000633576506585156e+00f);
  num = fmaf(num, t, 2.112677543073053063722e+00f);
  num = fmaf(num, t, 9.999999999999998702752e-01f);
  num = fmaf(num, t, 0.0f);
  float den = 5.875095403124574342950e-03f;
  den = fmaf(den, t, 1.546476374983906719538e-01f);
  den = fmaf(den, t, 9.864942958519418960339e-01f);
  den = fmaf(den, t, 2.453006071784736363091e+00f);
  den = fmaf(den, t, 2.612677543073109236779e+00f);
  den = fmaf(den, t, 1.0f);
  return fmaf((float)e, 0.693147180559945309417232121458176568f, num / den);
}
template <>
__device__ __forceinline__ double xg_log<double>(double x) { return log(x); }

// numba/np binary_search_with_guess (compiled_base.c), literal port over an accessor X(k)
template <typename F>
__device__ __forceinline__ int search_with_guess(double key, F X, int len, int guess) {
  int imin = 0, imax = len;
  if (key > X(len - 1)) return len;
  if (key < X(0)) return -1;
  if (len <= 4) {
    int i = 1;
    while (i < len && key >= X(i)) ++i;
    return i - 1;
  }
  if (guess > len - 3) guess = len - 3;
  if (guess < 1) guess = 1;
  if (key < X(guess)) {
    if (key < X(guess - 1)) {
      imax = guess - 1;
      if (guess > LIKELY_IN_CACHE_SIZE && key >= X(guess - LIKELY_IN_CACHE_SIZE))
        imin = guess - LIKELY_IN_CACHE_SIZE;
    } else {
      return guess - 1;
    }
  } else {
    if (key < X(guess + 1)) return guess;
    if (key < X(guess + 2)) return guess + 1;
    imin = guess + 2;
    if (guess < len - LIKELY_IN_CACHE_SIZE - 1 && key < X(guess + LIKELY_IN_CACHE_SIZE))
      imax = guess + LIKELY_IN_CACHE_SIZE;
  }
  while (imin < imax) {
    const int imid = imin + ((imax - imin) >> 1);
    if (key >= X(imid)) imin = imid + 1;
    else imax = imid;
  }
  return imin - 1;
}

// the arithmetic of np.interp for one target given its interval (fp64, no contraction)
__device__ __forceinline__ double interp_value(double x, double xj, double xj1, double yj, double yj1,
                                               double slope) {
  double res = slope * (x - xj) + yj;
  if (res != res) {
    res = slope * (x - xj1) + yj1;
    if (res != res && yj == yj1) res = yj;
  }
  return res;
}

// correctly rounded division from a shared reciprocal: see xg_common.cuh
__device__ __forceinline__ double div_with_recip(double a, double b, double r) {
  return xg_div_with_recip(a, b, r);
}
__device__ __forceinline__ bool exponent_safe(double v) { return xg_exponent_safe(v); }

// TMA-staged shared-theta kernel (xg_vinterp_tma.cu).  Returns 1 when the launch was made, 0 when the
// layout does not qualify (caller falls back to k_vinterp_shared), < 0 on error.
template <typename T>
int vinterp_shared_tma(const InterpArgs<T>& a, cudaStream_t st);
// TMA-staged theta-field kernel; same return convention.
template <typename T>
int vinterp_columns_tma(const InterpArgs<T>& a, cudaStream_t st);

}  // namespace xgvi
