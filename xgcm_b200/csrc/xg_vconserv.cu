// xg_vinterp_conservative — conservative (integral preserving) vertical remapping.
//
// Replaces the numba gufunc xgcm/transform.py:88-136 (_interp_1d_conservative): every source cell
// i, bounded by theta[i] and theta[i+1], spreads its extensive content phi[i] over the target bins
// it overlaps, proportionally to the overlap; NaN rules as in the reference (a cell with both
// bounds NaN is skipped, one NaN bound makes the cell a point, NaN phi contributes nothing, bins
// that receive nothing stay NaN).  Arithmetic is in the field dtype and, per bin, contributions
// are added in source-cell order — bit-identical to the reference's double loop.  The reference
// visits all n x (m-1) (cell, bin) pairs; bins are sorted (transform.py:167-176 enforces monotonic
// targets), so the bins a cell overlaps form one contiguous range found by bisection, and cells
// that do not overlap a bin do not touch it: same result in O(n log m + overlaps).
//
// Layout: lane = column (coalesced along the contiguous dim for every level); the bins of 32
// columns accumulate in shared memory and are written out transposed, new dim LAST.
#include "xg_common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kTile = 32;

template <typename T>
struct ConsArgs {
  const T* phi;
  const T* bins;   // m sorted (ascending) bin edges, device
  T* out;
  int64_t outer, n, inner, m;
  XgOperand theta;  // n + 1 bounds along the axis
  int flip_out;     // the caller passed decreasing bins: reverse the output bins
  bool small_cols;
  int bins_per_pass;  // accumulators held in shared memory at a time (>= 1; == m - 1: a single pass)
  int edges_in_smem;  // the m bin edges are staged in shared memory (else read through the read-only cache)
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) k_vconserv(const ConsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = (int)a.n, m = (int)a.m, nb = m - 1;
  const int W = blockDim.x >> 5, nbp = a.bins_per_pass;
  T* smem_t = reinterpret_cast<T*>(smem_raw);
  const T* edges = a.bins;                                  // [m]
  T* acc_all = smem_t;                                      // [W][nbp][33]
  if (a.edges_in_smem) {
    for (int k = threadIdx.x; k < m; k += blockDim.x) smem_t[k] = __ldg(a.bins + k);
    edges = smem_t;
    acc_all = smem_t + m;
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* acc = acc_all + (size_t)w * nbp * (kTile + 1);
  const int64_t ncols = a.outer * a.inner;
  const int64_t col0 = ((int64_t)blockIdx.x * W + w) * kTile;
  if (col0 >= ncols) return;  // warp-uniform
  const int64_t col = col0 + lane;
  const bool col_ok = col < ncols;
  const int ncol_here = (int)((ncols - col0 < kTile) ? (ncols - col0) : kTile);
  const T* phi = a.phi;
  const T* theta = reinterpret_cast<const T*>(a.theta.ptr);
  const int64_t ts = a.theta.axis_stride;
  if (col_ok) {
    int64_t o, i;
    xg_divmod(col, a.inner, a.small_cols, o, i);
    phi = a.phi + o * a.n * a.inner + i;
    int64_t toff = xg_groups_offset(a.theta.outer, o);
    if (a.theta.inner_mode == XG_IM_CONTIG) toff += i;
    else if (a.theta.inner_mode == XG_IM_GENERIC) toff += xg_groups_offset(a.theta.inner, i);
    theta += toff;
  }
  // Target grids with more bins than one shared-memory tile holds are done in passes over bin ranges
  // [jb, je): every pass walks the source cells again and only touches its own bins, so each bin still
  // receives its contributions in source-cell order (bit-identical to a single pass).
  for (int jb = 0; jb < nb; jb += nbp) {
    const int je = (jb + nbp < nb) ? jb + nbp : nb;
    for (int j = 0; j < je - jb; ++j) acc[j * (kTile + 1) + lane] = T(NAN);  // transform.py:98
    if (col_ok) {
      T t_lo = __ldg(theta);
      for (int c = 0; c < n; ++c) {
        const T t_hi = __ldg(theta + (int64_t)(c + 1) * ts);
        const T t1 = t_lo, t2 = t_hi;
        t_lo = t_hi;
        const T p = __ldg(phi + (int64_t)c * a.inner);
        T tmin, tmax;
        if (xg_isnan(t1) && xg_isnan(t2)) continue;        // transform.py:105-106
        else if (xg_isnan(t1)) tmin = tmax = t2;           // :109-110 homogeneous cell
        else if (xg_isnan(t2)) tmin = tmax = t1;           // :111-112
        else if (t1 < t2) { tmin = t1; tmax = t2; }        // :114-116
        else { tmin = t2; tmax = t1; }                     // :117-119 non-monotonic stratification
        if (xg_isnan(p)) continue;                         // :122-125 missing data adds nothing
        // bins j overlapping [tmin, tmax]: edges[j] <= tmax and edges[j+1] >= tmin (:126-128)
        int lo = 0, hi = nb;  // first j with edges[j+1] >= tmin
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (edges[mid + 1] >= tmin) hi = mid; else lo = mid + 1;
        }
        int j_lo = lo;
        lo = 0; hi = nb;      // first j with edges[j] > tmax
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (edges[mid] > tmax) hi = mid; else lo = mid + 1;
        }
        int j_hi = lo;  // exclusive
        if (j_lo < jb) j_lo = jb;
        if (j_hi > je) j_hi = je;
        for (int j = j_lo; j < j_hi; ++j) {
          T* slot = acc + (j - jb) * (kTile + 1) + lane;
          const T cur = *slot;
          T add;
          if (tmax == tmin) {
            add = p;                                        // :129-133
          } else {
            const T e1 = edges[j], e2 = edges[j + 1];
            const T hmin = (tmin >= e1) ? tmin : e1;        // max(theta_min, theta_hat_1[j])
            const T hmax = (tmax <= e2) ? tmax : e2;        // min(theta_max, theta_hat_2[j])
            const T alpha = (hmax - hmin) / (tmax - tmin);  // :136-138
            add = alpha * p;
          }
          *slot = xg_isnan(cur) ? add : cur + add;          // :140-143
        }
      }
    }
    __syncwarp();
    // transposed write-out: lanes run along the bins of one column
    for (int j0 = jb; j0 < je; j0 += kTile) {
      const int j = j0 + lane;
      if (j < je) {
        const int jo = a.flip_out ? (nb - 1 - j) : j;
        for (int cc = 0; cc < ncol_here; ++cc)
          __stcs(a.out + (col0 + cc) * (int64_t)nb + jo, acc[(j - jb) * (kTile + 1) + cc]);
      }
    }
    __syncwarp();
  }
}

template <typename T>
int vconserv_typed(const void* phi, const void* theta, const int64_t* theta_strides, const void* bins,
                   int64_t m, int flip_out, void* out, int ndim, const int64_t* shape, int axis,
                   cudaStream_t st) {
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  if (m < 2) return xg_fail(XG_EINVAL, "xg_vinterp_conservative: need at least two bin edges");
  if (v.n >= (1ll << 31) || m >= (1ll << 31))
    return xg_fail(XG_EINVAL, "xg_vinterp_conservative: more than 2^31 levels");
  ConsArgs<T> a;
  a.phi = static_cast<const T*>(phi);
  a.bins = static_cast<const T*>(bins);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.m = m;
  a.flip_out = flip_out;
  int64_t tshape[XG_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) tshape[d] = shape[d];
  tshape[axis] = v.n + 1;
  rc = xg_make_operand(theta, theta_strides, ndim, tshape, axis, 1, sizeof(T), &a.theta,
                       "xg_vinterp_conservative(theta)");
  if (rc) return rc;
  const int64_t ncols = v.outer * v.inner;
  a.small_cols = ncols < (1ll << 31);
  if (ncols == 0) return XG_OK;
  // shared memory: [m edges (if they fit in 32 KiB)] + W warps x bins_per_pass x 33 accumulators.  Four warps
  // and a single pass when that fits; otherwise one warp per block and as many bins per pass as fit.
  const size_t budget = 200 * 1024;
  const size_t edge_bytes = (size_t)m * sizeof(T);
  a.edges_in_smem = edge_bytes <= 32 * 1024 ? 1 : 0;
  const size_t head = a.edges_in_smem ? edge_bytes : 0;
  const size_t per_bin = (size_t)(kTile + 1) * sizeof(T);
  int W = kWarps;
  int64_t nbp = m - 1;
  if (head + (size_t)W * nbp * per_bin > budget) {
    W = 1;
    nbp = (int64_t)((budget - head) / per_bin);
    if (nbp > m - 1) nbp = m - 1;
  }
  a.bins_per_pass = (int)nbp;
  const size_t smem = head + (size_t)W * nbp * per_bin;
  cudaError_t e = cudaFuncSetAttribute(k_vconserv<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return xg_fail(XG_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  const int64_t blocks = xg_ceil_div(xg_ceil_div(ncols, kTile), W);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_vinterp_conservative: grid too large");
  k_vconserv<T><<<(unsigned)blocks, W * 32, smem, st>>>(a);
  return xg_check_launch("xg_vinterp_conservative");
}

}  // namespace

extern "C" int xg_vinterp_conservative(int dtype, const void* phi, const void* theta,
                                       const int64_t* theta_strides, const void* target_bins,
                                       int64_t m, int flip_out, void* out, int ndim,
                                       const int64_t* shape, int axis, void* stream) {
  if (!phi || !theta || !theta_strides || !target_bins || !out || !shape)
    return xg_fail(XG_EINVAL, "xg_vinterp_conservative: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return vconserv_typed<float>(phi, theta, theta_strides, target_bins, m, flip_out, out, ndim, shape, axis, st);
  if (dtype == XG_F64)
    return vconserv_typed<double>(phi, theta, theta_strides, target_bins, m, flip_out, out, ndim, shape, axis, st);
  return xg_fail(XG_EINVAL, "xg_vinterp_conservative: dtype must be XG_F32 or XG_F64");
}
