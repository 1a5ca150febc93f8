// xg_vinterp_linear — per-column linear interpolation onto target levels.
//
// Replaces the numba gufunc xgcm/transform.py:15-41 (_interp_1d_linear) and its
// wrapper :44-85 (optional log), i.e. per column: NaN-filtered monotonicity test
// and flip (:27-31), np.interp (:33; numba's port of numpy's arr_interp incl.
// binary_search_with_guess and the NaN retry), edge masking (:35-41).  All
// arithmetic is fp64 and rounded once to the field dtype, as in the reference.
//
// Two kernels:
//   k_vinterp_shared   theta AND target are shared 1-D vectors (the default
//       `Grid.transform(da, 'Z', levels)` case: theta = the depth coordinate).
//       The search is column independent, so each block computes the plan
//       (interval index j, fp64 x, mask) of every target ONCE into shared memory —
//       by parallel bisection when theta is NaN-free and sorted (then the guess
//       path of np.interp cannot change the answer), otherwise by one thread
//       replaying binary_search_with_guess literally — and the columns only do
//       slope * (x - xj) + yj with the interval memoised (one fp64 divide per
//       interval actually used, not per target).
//   k_vinterp_columns  theta (or target) varies per column: lane = column,
//       literal guess-carrying search per column so that NaN-containing or
//       non-monotonic theta give the reference's (path dependent) values.
//
// Layout: a warp owns 32 adjacent columns (lane = column, coalesced along the
// contiguous dim for every level); the outputs of 32 targets x 32 columns are
// transposed through shared memory so each column's new vertical dim — appended
// LAST like xr.apply_ufunc does (transform.py:233-249) — is written as
// contiguous 128-byte segments.
//
// Roofline: HBM, (n + m) * sizeof(T) bytes per column (shared 1-D theta) or
// (2n + m) * sizeof(T) (theta field).
#include "xg_vinterp.cuh"

namespace {
using namespace xgvi;

constexpr int kWarps = 8;
template <typename T>
constexpr int kCPLv = 1;  // columns per lane in the shared-theta kernel (2 was measured slower: 3.0 vs 2.7 ms at C5)

// transposed write-out of a 32-column x nt-target tile
template <typename T>
__device__ __forceinline__ void store_tile(T (*tile)[kTile + 1], T* out, int64_t col0, int ncol_here,
                                           int64_t m, int64_t t0, int nt, int lane) {
  __syncwarp();
  if (lane < nt) {
    T* o = out + col0 * m + t0 + lane;
#pragma unroll 4
    for (int cc = 0; cc < ncol_here; ++cc) __stcs(o + (int64_t)cc * m, tile[cc][lane]);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// shared theta / shared target
// ---------------------------------------------------------------------------
// what a column has to do for a run of consecutive targets, decided once per block
enum { PK_INTERP = 0, PK_EXACT = 1, PK_FIRST = 2, PK_LAST = 3, PK_NAN = 4 };

struct __align__(16) Run {
  int t_begin;   // targets [t_begin, t_begin + len) — never crosses a multiple of 32
  int len_kind;  // len | kind << 8
  int j;         // interval index (PK_INTERP / PK_EXACT)
  int next_j;    // set on runs that enter a new interval: the interval the NEXT such run enters
                 // (-1: none), so its two nodes are requested one interval early
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) k_vinterp_shared(const InterpArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = (int)a.n, m = (int)a.m;
  const int nchunk = (m + kTile - 1) / kTile;
  // layout: runs[m + nchunk] | xt[m] | X[n] | rdx[n] | tile[kWarps][kCPLv<T>][32][33] | tj[m] tk[m] chunk_run[nchunk+1] flags[4]
  Run* runs = reinterpret_cast<Run*>(smem_raw);
  double* xt = reinterpret_cast<double*>(runs + (m + nchunk));
  double* Xs = xt + m;
  double* rdx = Xs + n;  // RN(1 / (X[j+1] - X[j])) or 0 when the fast division must not be used
  T(*tiles)[kCPLv<T>][kTile][kTile + 1] = reinterpret_cast<T(*)[kCPLv<T>][kTile][kTile + 1]>(rdx + n);
  int* tj = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(tiles) +
                                   sizeof(T) * kWarps * kCPLv<T> * kTile * (kTile + 1));
  int* tk = tj + m;
  int* chunk_run = tk + m;
  int* flags = chunk_run + nchunk + 1;
  const int tid = threadIdx.x;
  const T* theta = reinterpret_cast<const T*>(a.theta.ptr);
  const T* target = reinterpret_cast<const T*>(a.target.ptr);
  const int64_t tstride = a.theta.axis_stride;

  // ---- plan, once per block -------------------------------------------------
  for (int k = tid; k < n; k += blockDim.x) {
    T v = __ldg(theta + k * tstride);
    if (a.logarithmic) v = xg_log<T>(v);  // transform.py:82-84, in the field dtype
    Xs[k] = (double)v;
  }
  __syncthreads();
  __shared__ double s_tmin, s_tmax;
  if (tid == 0) {
    int flip = 0;
    if (!a.bypass_checks) {  // transform.py:27-31
      int kf = 0, kl = n - 1;
      while (kf < n && Xs[kf] != Xs[kf]) ++kf;
      while (kl >= 0 && Xs[kl] != Xs[kl]) --kl;
      if (kf < n && Xs[kl] < Xs[kf]) flip = 1;
    }
    bool any = false, nan = false, sorted = true;
    double tmin = 0.0, tmax = 0.0, prev = 0.0;
    for (int k = 0; k < n; ++k) {
      const double v = Xs[flip ? n - 1 - k : k];
      if (v != v) { nan = true; continue; }
      if (!any) { tmin = tmax = v; any = true; }
      else { tmin = v < tmin ? v : tmin; tmax = v > tmax ? v : tmax; if (v < prev) sorted = false; }
      prev = v;
    }
    if (!any) tmin = tmax = NAN;
    s_tmin = tmin;
    s_tmax = tmax;
    flags[0] = flip;
    flags[1] = (!nan && sorted) ? 1 : 0;
  }
  __syncthreads();
  const int flip = flags[0];
  if (flip) {  // reverse in place
    for (int k = tid; k < n / 2; k += blockDim.x) {
      const double t0 = Xs[k];
      Xs[k] = Xs[n - 1 - k];
      Xs[n - 1 - k] = t0;
    }
    __syncthreads();
  }
  for (int j = tid; j < n; j += blockDim.x) {
    double r = 0.0;
    if (j + 1 < n) {
      const double dxj = Xs[j + 1] - Xs[j];
      if (exponent_safe(dxj)) r = 1.0 / dxj;
    }
    rdx[j] = r;
  }
  auto X = [&](int k) -> double { return Xs[k]; };
  auto classify = [&](int t, double x, int j) {
    int kind, jj = 0;
    if (a.mask_edges && (x < s_tmin || x > s_tmax)) kind = PK_NAN;  // transform.py:38-41
    else if (n == 1) kind = PK_FIRST;  // np.interp: dx.size == 1 -> full(dy[0])
    else if (x != x) kind = PK_NAN;    // np.interp: a NaN target stays NaN
    else if (j == -1) kind = PK_FIRST;
    else if (j >= n - 1) kind = PK_LAST;  // right of the range, or exactly the last node
    else {
      jj = j;
      kind = (Xs[j] == x) ? PK_EXACT : PK_INTERP;
    }
    xt[t] = x;
    tj[t] = jj;
    tk[t] = kind;
  };
  auto load_target = [&](int t) -> double {
    T v = __ldg(target + t * a.target.axis_stride);
    if (a.logarithmic) v = xg_log<T>(v);
    return (double)v;
  };
  if (flags[1] || n == 1) {
    // NaN-free sorted theta: binary_search_with_guess returns the largest j with X[j] <= x
    // whatever the guess, so every target can be searched independently
    for (int t = tid; t < m; t += blockDim.x) {
      const double x = load_target(t);
      int j = 0;
      if (n > 1 && x == x) {
        if (x > Xs[n - 1]) j = n;
        else if (x < Xs[0]) j = -1;
        else {
          int lo = 0, hi = n;
          while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if (x >= Xs[mid]) lo = mid + 1;
            else hi = mid;
          }
          j = lo - 1;
        }
      }
      classify(t, x, j);
    }
  } else if (tid == 0) {  // literal replay (guess carried from target to target)
    int guess = 0;
    for (int t = 0; t < m; ++t) {
      const double x = load_target(t);
      int j = 0;
      if (x == x) {
        j = search_with_guess(x, X, n, guess);
        guess = j;
      }
      classify(t, x, j);
    }
  }
  __syncthreads();
  if (tid == 0) {  // group consecutive targets with the same work into runs, split at tile edges
    int nr = 0;
    for (int t = 0; t < m; ++t) {
      const bool fresh = (t % kTile == 0) || tk[t] != (runs[nr - 1].len_kind >> 8) ||
                         (tk[t] <= PK_EXACT && tj[t] != runs[nr - 1].j);
      if (t % kTile == 0) chunk_run[t / kTile] = nr;
      if (fresh) {
        runs[nr].t_begin = t;
        runs[nr].len_kind = 1 | (tk[t] << 8);
        runs[nr].j = tj[t];
        runs[nr].next_j = -1;
        ++nr;
      } else {
        runs[nr - 1].len_kind += 1;
      }
    }
    chunk_run[nchunk] = nr;
    // link the runs that enter a new interval (the column loop memoises the current one)
    int cjs = -2, last = -1;
    flags[3] = -1;
    for (int r = 0; r < nr; ++r) {
      if ((runs[r].len_kind >> 8) > PK_EXACT || runs[r].j == cjs) continue;
      cjs = runs[r].j;
      if (last >= 0) runs[last].next_j = cjs;
      else flags[3] = cjs;
      last = r;
    }
  }
  __syncthreads();

  // ---- columns ----------------------------------------------------------------
  // Every column runs the SAME sequence of runs (theta is shared), so a lane carries kCPLv<T> columns in
  // lock-step: twice the independent load -> divide -> interpolate chains per warp, and the plan
  // reads / loop control are paid once for both.
  const int w = tid >> 5, lane = tid & 31;
  const int64_t ncols = a.outer * a.inner;
  const int64_t group_stride = (int64_t)gridDim.x * kWarps;
  const int64_t ngroups = (a.ntiles + kCPLv<T> - 1) / kCPLv<T>;
  const int64_t step = flip ? -a.inner : a.inner;  // phi pointer step for j -> j + 1
  const int j_first = flags[3];                    // first interval any column enters (-1: none)
  for (int64_t cg = (int64_t)blockIdx.x * kWarps + w; cg < ngroups; cg += group_stride) {
    const T* phi0[kCPLv<T>];
    T raw_a[kCPLv<T>], raw_b[kCPLv<T>];  // nodes j, j + 1 of the NEXT interval to enter, in flight
    T raw_first[kCPLv<T>], raw_last[kCPLv<T>];
    double yj[kCPLv<T>], yj1[kCPLv<T>], slope[kCPLv<T>];
#pragma unroll
    for (int c = 0; c < kCPLv<T>; ++c) {
      const int64_t tile_id = cg * kCPLv<T> + c;
      int64_t col = tile_id * kTile + lane;
      if (col >= ncols) col = ncols - 1;  // spare lanes (and a spare tile) shadow the last column
      int64_t o, i;
      xg_divmod(col, a.inner, a.small_cols, o, i);
      phi0[c] = a.phi + o * a.n * a.inner + i + (flip ? (int64_t)(n - 1) * a.inner : 0);
      raw_first[c] = __ldg(phi0[c]);
      raw_last[c] = __ldg(phi0[c] + (int64_t)(n - 1) * step);
      raw_a[c] = raw_b[c] = T(0);
      if (j_first >= 0) {
        raw_a[c] = __ldg(phi0[c] + (int64_t)j_first * step);
        raw_b[c] = __ldg(phi0[c] + (int64_t)(j_first + 1) * step);
      }
      yj[c] = yj1[c] = slope[c] = 0.0;
    }
    int cj = -2;  // memoised interval (common to the kCPLv<T> columns)
    double xj = 0.0, xj1 = 0.0;
    for (int cc = 0; cc < nchunk; ++cc) {
      const int t0 = cc * kTile;
      const int nt = (m - t0 < kTile) ? (m - t0) : kTile;
      const int r_end = chunk_run[cc + 1];
      for (int r = chunk_run[cc]; r < r_end; ++r) {
        const Run run = runs[r];  // one 16-byte broadcast read
        const int len = run.len_kind & 0xff;
        const int kind = run.len_kind >> 8;
        const int d0 = run.t_begin - t0;
        if (kind <= PK_EXACT) {
          if (run.j != cj) {
            const bool seq = (run.j == cj + 1);  // node j is the old node j + 1
            cj = run.j;
            xj = seq ? xj1 : Xs[cj];
            xj1 = Xs[cj + 1];
            const double dxj = xj1 - xj, rr = rdx[cj];
            const int nj = run.next_j;
#pragma unroll
            for (int c = 0; c < kCPLv<T>; ++c) {
              yj[c] = seq ? yj1[c] : (double)raw_a[c];
              yj1[c] = (double)raw_b[c];
              // The plan knows which interval is entered next: request its nodes now, a whole
              // interval of work before they are converted above (the per-interval wait on this
              // load was 37 % of all stall samples, profiles/r1b_vinterp_stalls.txt).
              if (nj >= 0) {
                const T* pn = phi0[c] + (int64_t)nj * step;
                raw_b[c] = __ldg(pn + step);
                if (nj != cj + 1) raw_a[c] = __ldg(pn);
              }
            }
#pragma unroll
            for (int c = 0; c < kCPLv<T>; ++c) {
              const double dyj = yj1[c] - yj[c];
              slope[c] = (rr != 0.0 && exponent_safe(dyj)) ? div_with_recip(dyj, dxj, rr) : dyj / dxj;
            }
          }
          if (kind == PK_INTERP) {
            const double* xp = xt + run.t_begin;
#pragma unroll 1  // runs hold one or two targets
            for (int q = 0; q < len; ++q) {
              const double x = xp[q];
#pragma unroll
              for (int c = 0; c < kCPLv<T>; ++c)
                tiles[w][c][lane][d0 + q] = (T)interp_value(x, xj, xj1, yj[c], yj1[c], slope[c]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < kCPLv<T>; ++c) {
              const T v = (T)yj[c];
#pragma unroll 1
              for (int q = 0; q < len; ++q) tiles[w][c][lane][d0 + q] = v;
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < kCPLv<T>; ++c) {
            const T v = (kind == PK_FIRST) ? raw_first[c] : (kind == PK_LAST) ? raw_last[c] : T(NAN);
#pragma unroll 1
            for (int q = 0; q < len; ++q) tiles[w][c][lane][d0 + q] = v;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kCPLv<T>; ++c) {
        const int64_t col0 = (cg * kCPLv<T> + c) * kTile;
        if (col0 < ncols) {  // warp-uniform
          const int ncol_here = (int)((ncols - col0 < kTile) ? (ncols - col0) : kTile);
          store_tile<T>(tiles[w][c], a.out, col0, ncol_here, a.m, t0, nt, lane);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// per-column theta and / or target
// ---------------------------------------------------------------------------
constexpr int kWarpsCol = 4;  // static smem: 4 x 32 x 33 x 8 B = 33 KiB for fp64

template <typename T>
__global__ void __launch_bounds__(kWarpsCol * 32) k_vinterp_columns(const InterpArgs<T> a) {
  __shared__ T tile_s[kWarpsCol][kTile][kTile + 1];
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = (int)a.n, m = (int)a.m;
  const int64_t ncols = a.outer * a.inner;
  const int64_t col0 = ((int64_t)blockIdx.x * kWarpsCol + w) * kTile;
  if (col0 >= ncols) return;  // warp-uniform
  T(*tile)[kTile + 1] = tile_s[w];
  const int64_t col = col0 + lane;
  const bool col_ok = col < ncols;
  const int ncol_here = (int)((ncols - col0 < kTile) ? (ncols - col0) : kTile);

  const T* phi = a.phi;
  const T* theta = reinterpret_cast<const T*>(a.theta.ptr);
  const T* tgt = reinterpret_cast<const T*>(a.target.ptr);
  const int64_t ts = a.theta.axis_stride, ps = a.inner;
  const bool logarithmic = a.logarithmic != 0;
  bool flip = false;
  T tmin = T(0), tmax = T(0);
  if (col_ok) {
    int64_t o, i;
    xg_divmod(col, a.inner, a.small_cols, o, i);
    phi = a.phi + o * a.n * a.inner + i;
    int64_t toff = xg_groups_offset(a.theta.outer, o);
    if (a.theta.inner_mode == XG_IM_CONTIG) toff += i;
    else if (a.theta.inner_mode == XG_IM_GENERIC) toff += xg_groups_offset(a.theta.inner, i);
    theta += toff;
    int64_t goff = xg_groups_offset(a.target.outer, o);
    if (a.target.inner_mode == XG_IM_CONTIG) goff += i;
    else if (a.target.inner_mode == XG_IM_GENERIC) goff += xg_groups_offset(a.target.inner, i);
    tgt += goff;
  }
  // raw (unflipped) theta in the field dtype; comparisons in T are exact for T -> double
  auto TH = [&](int k) -> T {
    T v = __ldg(theta + (int64_t)k * ts);
    return logarithmic ? xg_log<T>(v) : v;
  };
  bool walk = false;  // this column's theta is NaN-free and sorted: the search is path independent
  if (col_ok) {
    if (!a.bypass_checks) {  // transform.py:27-31: sign test on the NaN-filtered theta
      int kf = 0, kl = n - 1;
      while (kf < n && xg_isnan(TH(kf))) ++kf;
      while (kl >= 0 && xg_isnan(TH(kl))) --kl;
      if (kf < n && TH(kl) < TH(kf)) flip = true;
    }
    // one pass over the column: nanmax / nanmin (transform.py:36-37; NaN if all-NaN) and whether
    // the (possibly flipped) theta is NaN-free and non-decreasing
    bool any = false, nan = false, sorted = true;
    T prev = T(0);
    for (int k = 0; k < n; ++k) {
      const T v = TH(flip ? n - 1 - k : k);
      if (xg_isnan(v)) { nan = true; continue; }
      if (!any) { tmin = tmax = v; any = true; }
      else { tmin = v < tmin ? v : tmin; tmax = v > tmax ? v : tmax; if (v < prev) sorted = false; }
      prev = v;
    }
    if (!any) tmin = tmax = T(NAN);
    walk = !nan && sorted && n > 1;
  }
  auto X = [&](int k) -> double { return (double)TH(flip ? n - 1 - k : k); };
  auto Y = [&](int k) -> double { return (double)__ldg(phi + (int64_t)(flip ? n - 1 - k : k) * ps); };
  const double x_first = (col_ok && walk) ? X(0) : 0.0;
  const double x_last = (col_ok && walk) ? X(n - 1) : 0.0;

  int guess = 0, cj = -2;
  double yj = 0.0, yj1 = 0.0, slope = 0.0, xj = 0.0, xj1 = 0.0;
  for (int t0 = 0; t0 < m; t0 += kTile) {
    const int nt = (m - t0 < kTile) ? (m - t0) : kTile;
    if (col_ok) {
      for (int tt = 0; tt < nt; ++tt) {
        T xv = __ldg(tgt + (int64_t)(t0 + tt) * a.target.axis_stride);
        if (logarithmic) xv = xg_log<T>(xv);
        const double x = (double)xv;
        double res;
        if (n == 1) {
          res = Y(0);  // np.interp: dx.size == 1 -> full(dy[0])
        } else if (x != x) {
          res = x;
        } else {
          int j;
          if (walk) {
            // sorted NaN-free theta: binary_search_with_guess returns the largest j with
            // X[j] <= x whatever its guess; targets usually ascend, so walk from the current
            // interval (sequential, prefetch-friendly) and bisect only on a step backwards
            if (x > x_last) j = n;
            else if (x < x_first) j = -1;
            else if (cj >= 0 && x >= xj) {
              j = cj;
              double xn = xj1;
              while (j + 1 < n && xn <= x) {
                ++j;
                if (j + 1 < n) xn = X(j + 1);
              }
            } else {
              int lo = 0, hi = n;
              while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (x >= X(mid)) lo = mid + 1;
                else hi = mid;
              }
              j = lo - 1;
            }
          } else {
            j = search_with_guess(x, X, n, guess);  // literal replay, guess carried along
            guess = j;
          }
          if (j == -1) res = Y(0);
          else if (j >= n - 1) res = Y(n - 1);
          else {
            if (j != cj) {
              cj = j;
              if (walk && j + 1 + kPrefetchRows < n) {
                const int64_t ahead = flip ? n - 1 - (j + 1 + kPrefetchRows) : j + 1 + kPrefetchRows;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(phi + ahead * ps));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(theta + ahead * ts));
              }
              xj = X(j);
              xj1 = X(j + 1);
              yj = Y(j);
              yj1 = Y(j + 1);
              slope = (yj1 - yj) / (xj1 - xj);
            }
            res = (xj == x) ? yj : interp_value(x, xj, xj1, yj, yj1, slope);
          }
        }
        if (a.mask_edges && (xv < tmin || xv > tmax)) res = NAN;  // transform.py:38-41
        tile[lane][tt] = (T)res;
      }
    }
    store_tile<T>(tile, a.out, col0, ncol_here, a.m, t0, nt, lane);
  }
}

template <typename T>
int vinterp_typed(const void* phi, const void* theta, const int64_t* theta_strides,
                  const void* target, const int64_t* target_strides, int64_t m, void* out,
                  int ndim, const int64_t* shape, int axis, int mask_edges, int bypass_checks,
                  int logarithmic, cudaStream_t st) {
  XgView v;
  int rc = xg_collapse_view(ndim, shape, axis, &v);
  if (rc) return rc;
  if (v.n == 0) return xg_fail(XG_EINVAL, "xg_vinterp_linear: array of sample points is empty");
  if (v.n >= (1ll << 31) || m >= (1ll << 31))
    return xg_fail(XG_EINVAL, "xg_vinterp_linear: more than 2^31 levels");
  InterpArgs<T> a;
  a.phi = static_cast<const T*>(phi);
  a.out = static_cast<T*>(out);
  a.outer = v.outer;
  a.n = v.n;
  a.inner = v.inner;
  a.m = m;
  a.theta_full = false;
  a.mask_edges = mask_edges;
  a.bypass_checks = bypass_checks;
  a.logarithmic = logarithmic;
  rc = xg_make_operand(theta, theta_strides, ndim, shape, axis, 1, sizeof(T), &a.theta,
                       "xg_vinterp_linear(theta)");
  if (rc) return rc;
  {
    int64_t tshape[XG_MAX_NDIM], tstr[XG_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) {
      tshape[d] = shape[d];
      tstr[d] = target_strides ? target_strides[d] : 0;
    }
    tshape[axis] = m;
    if (!target_strides) tstr[axis] = 1;
    rc = xg_make_operand(target, tstr, ndim, tshape, axis, 1, sizeof(T), &a.target,
                         "xg_vinterp_linear(target)");
    if (rc) return rc;
    if (m == 1) a.target.axis_stride = 0;
  }
  if (shape[axis] > 1 && a.theta.axis_stride == 0)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear: theta must vary along the operated axis");
  const int64_t ncols = v.outer * v.inner;
  if (ncols == 0 || m == 0) return XG_OK;
  a.ntiles = xg_ceil_div(ncols, kTile);
  a.small_cols = ncols < (1ll << 31);

  auto all_bcast = [](const XgOperand& op) {
    bool outer0 = true;
    for (int k = 0; k < op.outer.n; ++k) outer0 = outer0 && op.outer.stride[k] == 0;
    return outer0 && op.inner_mode == XG_IM_BCAST;
  };
  const int64_t nchunk = xg_ceil_div(m, kTile);
  const size_t plan_bytes = (size_t)(m + nchunk) * sizeof(Run) + (size_t)(m + 2 * v.n) * sizeof(double) +
                            sizeof(T) * kWarps * kCPLv<T> * kTile * (kTile + 1) +
                            (size_t)(2 * m + nchunk + 1 + 4) * sizeof(int);
  if (all_bcast(a.theta) && all_bcast(a.target)) {
    const int r = vinterp_shared_tma<T>(a, st);  // bulk-async staged tiles when the layout allows
    if (r != 0) return r < 0 ? r : XG_OK;
  }
  if (all_bcast(a.theta) && all_bcast(a.target) && plan_bytes <= 200 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_vinterp_shared<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)plan_bytes);
    if (e != cudaSuccess)
      return xg_fail(XG_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    // persistent: the plan is amortised over many tiles, and exactly one resident wave so that no
    // SM runs a half-empty second wave
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vinterp_shared<T>, kWarps * 32, plan_bytes);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t blocks = xg_ceil_div(xg_ceil_div(a.ntiles, kCPLv<T>), kWarps);
    if (blocks > (int64_t)sms * per_sm) blocks = (int64_t)sms * per_sm;
    k_vinterp_shared<T><<<(unsigned)blocks, kWarps * 32, plan_bytes, st>>>(a);
    return xg_check_launch("xg_vinterp_linear(shared)");
  }
  {
    // theta a full field laid out like phi, one shared level vector: tiles of phi AND theta staged by TMA
    bool full = true;
    int64_t want = 1;
    for (int d = ndim - 1; d >= 0; --d) {
      if (shape[d] > 1 && theta_strides[d] != want) full = false;
      want *= shape[d];
    }
    a.theta_full = full;
    if (full && all_bcast(a.target)) {
      const int r = vinterp_columns_tma<T>(a, st);
      if (r != 0) return r < 0 ? r : XG_OK;
    }
  }
  const int64_t blocks = xg_ceil_div(a.ntiles, kWarpsCol);
  if (blocks > 0x7fffffffLL) return xg_fail(XG_EINVAL, "xg_vinterp_linear: grid too large");
  k_vinterp_columns<T><<<(unsigned)blocks, kWarpsCol * 32, 0, st>>>(a);
  return xg_check_launch("xg_vinterp_linear(columns)");
}

}  // namespace

extern "C" int xg_vinterp_linear(int dtype, const void* phi, const void* theta,
                                 const int64_t* theta_strides, const void* target,
                                 const int64_t* target_strides, int64_t m, void* out, int ndim,
                                 const int64_t* shape, int axis, int mask_edges, int bypass_checks,
                                 int logarithmic, void* stream) {
  if (!phi || !theta || !theta_strides || !target || !out || !shape)
    return xg_fail(XG_EINVAL, "xg_vinterp_linear: null pointer");
  if (m < 0) return xg_fail(XG_EINVAL, "xg_vinterp_linear: negative number of target levels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return vinterp_typed<float>(phi, theta, theta_strides, target, target_strides, m, out, ndim,
                                shape, axis, mask_edges, bypass_checks, logarithmic, st);
  if (dtype == XG_F64)
    return vinterp_typed<double>(phi, theta, theta_strides, target, target_strides, m, out, ndim,
                                 shape, axis, mask_edges, bypass_checks, logarithmic, st);
  return xg_fail(XG_EINVAL, "xg_vinterp_linear: dtype must be XG_F32 or XG_F64");
}
