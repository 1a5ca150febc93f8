// xg_vinterp_linear — per-column linear interpolation onto target levels.
//
// Replaces the numba gufunc xgcm/transform.py:15-41 (_interp_1d_linear) and its
// wrapper :44-85 (optional log), i.e. per column: NaN-filtered monotonicity test
// and flip (:27-31), np.interp (:33; numba's port of numpy's arr_interp incl.
// binary_search_with_guess and the NaN retry), edge masking (:35-41).  All
// arithmetic is fp64 and rounded once to the field dtype, as in the reference.
// The guess-carrying search is reproduced literally so that columns whose theta
// contains NaNs (where np.interp's answer depends on the probe sequence) give
// the same values.
//
// Layout: a warp owns 32 adjacent columns (lane = column, coalesced along the
// contiguous dim for every level); the outputs of 32 targets x 32 columns are
// transposed through shared memory so each column's new vertical dim — appended
// LAST like xr.apply_ufunc does (transform.py:233-249) — is written as
// contiguous 128-byte segments.
//
// Roofline: HBM, (n + m) * sizeof(T) bytes per column (shared 1-D theta) or
// (2n + m) * sizeof(T) (theta field).
#include "xg_common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kTile = 32;
constexpr int LIKELY_IN_CACHE_SIZE = 8;

template <typename T>
struct InterpArgs {
  const T* phi;
  const T* target;
  T* out;
  int64_t outer, n, inner, m;
  XgOperand theta;
  int mask_edges, bypass_checks, logarithmic;
};

template <typename T>
__device__ __forceinline__ T xg_log(T x);
template <>
__device__ __forceinline__ float xg_log<float>(float x) { return logf(x); }
template <>
__device__ __forceinline__ double xg_log<double>(double x) { return log(x); }

template <typename T>
struct Column {
  const T* phi;     // + k * phi_stride
  const T* theta;   // + k * theta_stride
  int64_t phi_stride, theta_stride, n;
  bool flip, logarithmic;
  __device__ __forceinline__ T theta_raw(int64_t k) const {
    T v = __ldg(theta + k * theta_stride);
    return logarithmic ? xg_log<T>(v) : v;  // transform.py:82-84, in the field dtype
  }
  __device__ __forceinline__ double X(int64_t k) const {
    return (double)theta_raw(flip ? n - 1 - k : k);
  }
  __device__ __forceinline__ double Y(int64_t k) const {
    return (double)__ldg(phi + (flip ? n - 1 - k : k) * phi_stride);
  }
};

// numba/np: binary_search_with_guess (compiled_base.c), literal port
template <typename T>
__device__ int64_t search_with_guess(double key, const Column<T>& c, int64_t len, int64_t guess) {
  int64_t imin = 0, imax = len;
  if (key > c.X(len - 1)) return len;
  if (key < c.X(0)) return -1;
  if (len <= 4) {
    int64_t i = 1;
    while (i < len && key >= c.X(i)) ++i;
    return i - 1;
  }
  if (guess > len - 3) guess = len - 3;
  if (guess < 1) guess = 1;
  if (key < c.X(guess)) {
    if (key < c.X(guess - 1)) {
      imax = guess - 1;
      if (guess > LIKELY_IN_CACHE_SIZE && key >= c.X(guess - LIKELY_IN_CACHE_SIZE))
        imin = guess - LIKELY_IN_CACHE_SIZE;
    } else {
      return guess - 1;
    }
  } else {
    if (key < c.X(guess + 1)) return guess;
    if (key < c.X(guess + 2)) return guess + 1;
    imin = guess + 2;
    if (guess < len - LIKELY_IN_CACHE_SIZE - 1 && key < c.X(guess + LIKELY_IN_CACHE_SIZE))
      imax = guess + LIKELY_IN_CACHE_SIZE;
  }
  while (imin < imax) {
    const int64_t imid = imin + ((imax - imin) >> 1);
    if (key >= c.X(imid)) imin = imid + 1;
    else imax = imid;
  }
  return imin - 1;
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) k_vinterp(const InterpArgs<T> a) {
  __shared__ T tile_s[kWarps][kTile][kTile + 1];
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t ncols = a.outer * a.inner;
  const int64_t col0 = ((int64_t)blockIdx.x * kWarps + w) * kTile;
  if (col0 >= ncols) return;  // warp-uniform
  T(*tile)[kTile + 1] = tile_s[w];
  const int64_t col = col0 + lane;
  const bool col_ok = col < ncols;
  const int ncol_here = (int)((ncols - col0 < kTile) ? (ncols - col0) : kTile);

  Column<T> c;
  c.n = a.n;
  c.flip = false;
  c.logarithmic = a.logarithmic != 0;
  c.phi_stride = a.inner;
  c.theta_stride = a.theta.axis_stride;
  c.phi = a.phi;
  c.theta = reinterpret_cast<const T*>(a.theta.ptr);
  double tmin = 0.0, tmax = 0.0;
  if (col_ok) {
    const int64_t o = col / a.inner;
    const int64_t i = col - o * a.inner;
    c.phi = a.phi + o * a.n * a.inner + i;
    int64_t toff = xg_groups_offset(a.theta.outer, o);
    if (a.theta.inner_mode == XG_IM_CONTIG) toff += i;
    else if (a.theta.inner_mode == XG_IM_GENERIC) toff += xg_groups_offset(a.theta.inner, i);
    c.theta = reinterpret_cast<const T*>(a.theta.ptr) + toff;
    if (!a.bypass_checks) {
      // transform.py:27-31: sign test on the NaN-filtered theta
      int64_t kf = 0, kl = a.n - 1;
      while (kf < a.n && xg_isnan(c.theta_raw(kf))) ++kf;
      while (kl >= 0 && xg_isnan(c.theta_raw(kl))) --kl;
      if (kf < a.n && c.theta_raw(kl) < c.theta_raw(kf)) c.flip = true;
    }
    if (a.mask_edges) {
      // transform.py:36-37 nanmax / nanmin (NaN if the column is all-NaN)
      bool any = false;
      for (int64_t k = 0; k < a.n; ++k) {
        const double v = (double)c.theta_raw(k);
        if (v != v) continue;
        if (!any) { tmin = tmax = v; any = true; }
        else { tmin = v < tmin ? v : tmin; tmax = v > tmax ? v : tmax; }
      }
      if (!any) tmin = tmax = NAN;
    }
  }

  int64_t guess = 0;
  for (int64_t t0 = 0; t0 < a.m; t0 += kTile) {
    const int nt = (int)((a.m - t0 < kTile) ? (a.m - t0) : kTile);
    if (col_ok) {
      for (int tt = 0; tt < nt; ++tt) {
        T xt = __ldg(a.target + t0 + tt);
        if (c.logarithmic) xt = xg_log<T>(xt);
        const double x = (double)xt;
        double res;
        if (a.n == 1) {
          res = c.Y(0);  // np.interp: dx.size == 1 -> full(dy[0])
        } else if (x != x) {
          res = x;
        } else {
          const int64_t j = search_with_guess<T>(x, c, a.n, guess);
          guess = j;
          if (j == -1) res = c.Y(0);
          else if (j == a.n) res = c.Y(a.n - 1);
          else if (j == a.n - 1) res = c.Y(j);
          else {
            const double xj = c.X(j);
            const double yj = c.Y(j);
            if (xj == x) {
              res = yj;
            } else {
              const double xj1 = c.X(j + 1);
              const double yj1 = c.Y(j + 1);
              const double slope = (yj1 - yj) / (xj1 - xj);
              res = slope * (x - xj) + yj;
              if (res != res) {
                res = slope * (x - xj1) + yj1;
                if (res != res && yj == yj1) res = yj;
              }
            }
          }
        }
        if (a.mask_edges && (x < tmin || x > tmax)) res = NAN;  // transform.py:38-41
        tile[lane][tt] = (T)res;
      }
    }
    __syncwarp();
    if (lane < nt) {
      for (int cc = 0; cc < ncol_here; ++cc)
        __stcs(a.out + (col0 + cc) * a.m + t0 + lane, tile[cc][lane]);
    }
    __syncwarp();
  }
}

template <typename T>
int vinterp_typed(const void* phi, const void* theta, const int64_t* theta_strides,
                  const void* target, int64_t m, void* out, int ndim, const int64_t* shape,
                  int axis, int mask_edges, int bypass_checks, int logarithmic, cudaStream_t st) {
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  if (v.n == 0) return xg_fail(XG_EINVAL, "xg_vinterp_linear: array of sample points is empty");
  InterpArgs<T> a;
  a.phi = static_cast<const T*>(phi);
  a.target = static_cast<const T*>(target);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.m = m;
  a.mask_edges = mask_edges;
  a.bypass_checks = bypass_checks;
  a.logarithmic = logarithmic;
  rc = xg_make_operand(theta, theta_strides, ndim, shape, axis, 1, sizeof(T), &a.theta,
                       "xg_vinterp_linear(theta)");
  if (rc) return rc;
  if (shape[axis] > 1 && a.theta.axis_stride == 0)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear: theta must vary along the operated axis");
  const int64_t ncols = v.outer * v.inner;
  if (ncols == 0 || m == 0) return XG_OK;
  const int64_t blocks = xg_ceil_div(xg_ceil_div(ncols, kTile), kWarps);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_vinterp_linear: grid too large");
  k_vinterp<T><<<(unsigned)blocks, kWarps * 32, 0, st>>>(a);
  return xg_check_launch("xg_vinterp_linear");
}

}  // namespace

extern "C" int xg_vinterp_linear(int dtype, const void* phi, const void* theta,
                                 const int64_t* theta_strides, const void* target, int64_t m,
                                 void* out, int ndim, const int64_t* shape, int axis,
                                 int mask_edges, int bypass_checks, int logarithmic,
                                 void* stream) {
  if (!phi || !theta || !theta_strides || !target || !out || !shape)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear: null pointer");
  if (m < 0) return xg_fail(XG_EINVAL, "xg_vinterp_linear: negative number of target levels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return vinterp_typed<float>(phi, theta, theta_strides, target, m, out, ndim, shape, axis,
                                mask_edges, bypass_checks, logarithmic, st);
  if (dtype == XG_F64)
    return vinterp_typed<double>(phi, theta, theta_strides, target, m, out, ndim, shape, axis,
                                 mask_edges, bypass_checks, logarithmic, st);
  return xg_fail(XG_EINVAL, "xg_vinterp_linear: dtype must be XG_F32 or XG_F64");
}
