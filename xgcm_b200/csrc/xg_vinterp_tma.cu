// xg_vinterp_linear, shared 1-D theta and target (config C5: `Grid.transform(da, 'Z', levels)` with the
// depth coordinate as theta) — the bulk-async (TMA) staged kernel.
//
// Replaces xgcm/transform.py:15-41 (_interp_1d_linear) + :233-249 (linear_interpolation: the new
// vertical dim is appended LAST) for the layout (outer, n, inner) -> (outer, inner, m).
//
// Data movement (the kernel is a tile mover: n x TC in, TC x m out, TC = 32 * CPL columns):
//   in : one `cp.async.bulk.shared.global` (1-D TMA, SASS UBLKCP) per level row of the tile,
//        TC * sizeof(T) bytes each, all rows of a tile completing on ONE mbarrier (expect_tx).
//        Nobody waits on a global load: the columns read their levels from shared memory.
//   out: the TC x m results of a tile are contiguous in the (outer, inner, m) output, so the tile is
//        assembled in shared memory with conflict-free 16-byte stores and leaves as ONE
//        `cp.async.bulk.global.shared` (bulk_group) of TC * m * sizeof(T) bytes.
//   A block is NT teams of WT warps; it owns a ring of NB >= NT input buffers and NT output buffers.
//   Local tile i is computed by team i % NT from buffer i % NB, the WT warps of the team splitting the
//   TARGETS of the tile between them (shared memory, not registers, limits how many columns an SM can
//   hold — ~300 — so the warps needed to hide the fp64 dependency chains have to come from splitting
//   each tile's work, profiles/r2_vtma_*).  The team that finishes tile i refills its buffer with tile
//   i + NB (it was the buffer's last reader, so no "empty" barrier is needed), which team (i + NB) % NT
//   picks up (NB - NT tiles of lookahead beyond one-buffer-per-team).
//
// Arithmetic: identical to k_vinterp_shared (fp64, rounded once; numpy's slope * (x - xj) + yj with
// the NaN retries; a correctly rounded slope per (column, interval)).  The per-block plan is
// target-major: entry t = {x_t - X[j_t] (fp64, column independent), j_t, kind}; the slope of an
// interval is computed when a target enters it.  ~1900 warp instructions per 32-column tile instead
// of ~7100 (profiles/): no run bookkeeping, no address arithmetic per level, no global-load waits.
#include "xg_vinterp.cuh"

#include <cstdlib>

#include "xg_tma.cuh"

namespace xgvi {
namespace {

enum { PK_INTERP = 0, PK_EXACT = 1, PK_FIRST = 2, PK_LAST = 3, PK_NAN = 4 };

// target t: {x_t - X[j_t] (fp64), j_t | kind << 24}.  A plain PK_INTERP entry is just its interval index,
// so `entry.jk == current` is the whole fast-path test (interval unchanged, every lane clean, nothing
// special); every other entry carries kind bits and never equals an interval index.
constexpr int kKindShift = 24;
constexpr int kUnclean = 1 << 23;  // state flag: the current interval needs the careful path in some lane
struct __align__(16) TargetPlan {
  double dxt;
  int jk;
  int pad;
};
struct __align__(16) IntervalPlan {
  double dxj;  // X[j+1] - X[j]
  double rr;   // RN(1 / dxj), or 0 when the reciprocal sequence must not be used
};

template <typename T>
__device__ __forceinline__ bool is_finite(T v);
template <>
__device__ __forceinline__ bool is_finite<float>(float v) { return fabsf(v) < INFINITY; }
template <>
__device__ __forceinline__ bool is_finite<double>(double v) { return fabs(v) < (double)INFINITY; }

template <typename T>
struct Vec16;  // 16 bytes of T
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  typedef float4 type;
};
template <>
struct Vec16<double> {
  static constexpr int N = 2;
  typedef double2 type;
};

// tile geometry shared by host and device
template <typename T, int CPL>
struct Geo {
  static constexpr int TC = 32 * CPL;  // columns per tile
};

template <typename T>
struct TmaArgs {
  InterpArgs<T> a;
  int64_t tiles_per_o;  // ceil(inner / TC)
  int64_t ntiles;       // outer * tiles_per_o
  int nb;               // input buffers in the ring (>= teams per block)
  int wt;               // warps per team
  bool small;           // tile indices fit in 32 bits
  int box_rows;         // levels per TMA box (<= 256)
  int nbox;             // boxes per tile: nbox * box_rows >= n
  unsigned in_bytes;    // nbox * box_rows * TC * sizeof(T), rounded up to 128
  unsigned out_bytes;   // TC * m * sizeof(T), rounded up to 128
  unsigned plan_bytes;  // offset of the first tile buffer
};

// np.interp's NaN retries (arr_interp): only reached when slope * (x - xj) + yj is NaN
__device__ __noinline__ double interp_retry(double slope, double x, double xj1, double yj, double yj1) {
  double res = slope * (x - xj1) + yj1;
  if (res != res && yj == yj1) res = yj;
  return res;
}

// slope of one interval when the reciprocal sequence does not apply
__device__ __noinline__ double slow_slope(double dy, double dxj) { return dy / dxj; }

template <typename T, int CPL>
__global__ void __launch_bounds__(1024, 1)
    k_vinterp_shared_tma(const __grid_constant__ CUtensorMap tmap, const TmaArgs<T> p) {
  constexpr int TC = Geo<T, CPL>::TC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const InterpArgs<T>& a = p.a;
  const int n = (int)a.n, m = (int)a.m;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int W = blockDim.x >> 5, NB = p.nb;

  // layout: tp[m] | iv[n] | Xs[n] | xt[m] | full[NB] (8 B each) | tj[m] tk[m] flags[4] wbeg[33] | pad | in[NB] | out[NT]
  TargetPlan* tp = reinterpret_cast<TargetPlan*>(smem_raw);
  IntervalPlan* iv = reinterpret_cast<IntervalPlan*>(tp + m);
  double* Xs = reinterpret_cast<double*>(iv + n);
  double* xt = Xs + n;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(xt + m);
  int* tj = reinterpret_cast<int*>(full + NB);
  int* tk = tj + m;
  int* flags = tk + m;
  int* wbeg = flags + 4;  // target range of warp q of a team: [wbeg[q], wbeg[q + 1])
  const int WT = p.wt, NT = W / WT;
  const int team = w / WT, wq = w - team * WT;  // this warp's team and its rank in the team
  unsigned char* in0 = smem_raw + p.plan_bytes;
  unsigned char* out0 = in0 + (size_t)NB * p.in_bytes;
  T* out_tile = reinterpret_cast<T*>(out0 + (size_t)team * p.out_bytes);
  const uint32_t full_u32 = smem_u32(full);

  // ---- tile geometry -----------------------------------------------------------------------------
  // local tile i of this block = global tile i * gridDim.x + blockIdx.x (blocks sweep the field side by
  // side, so concurrently processed tiles are neighbours in memory)
  const int64_t nloc = (p.ntiles > blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto tile_geom = [&](int64_t i, int64_t& o, int64_t& i0, int& ncol) {
    const int64_t g = i * gridDim.x + blockIdx.x;
    if (p.small) {
      const uint32_t oo = (uint32_t)g / (uint32_t)p.tiles_per_o;
      o = oo;
      i0 = (int64_t)((uint32_t)g - oo * (uint32_t)p.tiles_per_o) * TC;
    } else {
      o = g / p.tiles_per_o;
      i0 = (g - o * p.tiles_per_o) * TC;
    }
    const int64_t left = a.inner - i0;
    ncol = left < TC ? (int)left : TC;
  };
  // one elected lane arms the barrier and issues the tile's box copies (out-of-range rows / columns of a
  // box are zero-filled by the TMA unit and still count towards the transaction bytes)
  auto issue_load = [&](int64_t i) {
    const int b = (int)(i % NB);
    int64_t o, i0;
    int ncol;
    tile_geom(i, o, i0, ncol);
    const uint32_t bar = full_u32 + 8u * b;
    const unsigned box_bytes = (unsigned)p.box_rows * TC * sizeof(T);
    mbar_expect_tx(bar, box_bytes * (unsigned)p.nbox);
    const uint32_t dst = smem_u32(in0 + (size_t)b * p.in_bytes);
    for (int k = 0; k < p.nbox; ++k) tensor_load_3d(dst + k * box_bytes, &tmap, (int)i0, k * p.box_rows, (int)o, bar);
  };

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int b = 0; b < NB; ++b) mbar_init(full_u32 + 8u * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // prologue loads go out before the plan is built
  if (wq == 0 && lane == 0)
    for (int64_t i = team; i < NB && i < nloc; i += NT) issue_load(i);

  // ---- plan, once per block (same classification as k_vinterp_shared) ------------------------------
  const T* theta = reinterpret_cast<const T*>(a.theta.ptr);
  const T* target = reinterpret_cast<const T*>(a.target.ptr);
  const int64_t tstride = a.theta.axis_stride;
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
  const bool sorted = flags[1] != 0;
  if (flip) {  // reverse in place
    for (int k = tid; k < n / 2; k += blockDim.x) {
      const double t0 = Xs[k];
      Xs[k] = Xs[n - 1 - k];
      Xs[n - 1 - k] = t0;
    }
    __syncthreads();
  }
  for (int j = tid; j < n; j += blockDim.x) {
    IntervalPlan e;
    e.dxj = 0.0;
    e.rr = 0.0;
    if (j + 1 < n) {
      e.dxj = Xs[j + 1] - Xs[j];
      // The reciprocal sequence is only used on sorted NaN-free theta (dxj > 0, targets at or right of
      // X[j]): there a slope whose zero sign differs (dy == -0) cannot change any result, so dy needs
      // no zero test in the column loop (see advance()).
      if (sorted && exponent_safe(e.dxj)) e.rr = 1.0 / e.dxj;
    }
    iv[j] = e;
  }
  auto X = [&](int k) -> double { return Xs[k]; };
  auto classify = [&](int t, double x, int j) {
    int kind, jj = -1;
    if (a.mask_edges && (x < s_tmin || x > s_tmax)) kind = PK_NAN;  // transform.py:38-41
    else if (x != x) kind = PK_NAN;                                  // np.interp: a NaN target stays NaN
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
  if (sorted) {
    // NaN-free sorted theta: binary_search_with_guess returns the largest j with X[j] <= x whatever
    // the guess, so every target can be searched independently
    for (int t = tid; t < m; t += blockDim.x) {
      const double x = load_target(t);
      int j = 0;
      if (x == x) {
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
  for (int t = tid; t < m; t += blockDim.x) {
    const int kind = tk[t], j = tj[t] < 0 ? 0 : tj[t];
    TargetPlan e;
    e.dxt = xt[t] - Xs[j];
    e.jk = j | (kind << kKindShift);
    e.pad = 0;
    tp[t] = e;
  }
  __syncthreads();
  if (tid == 0) {
    // Split the targets between the WT warps of a team in contiguous runs of equal estimated cost: a plain
    // target costs ~1, entering a new interval ~4 more (two node loads + conversions + the slope's
    // reciprocal sequence per column).  Units of 2 targets (one 8 / 16-byte store each) when m is even.
    const int unit = (m % 2 == 0) ? 2 : 1;
    int total = 0;
    for (int t = 0; t < m; ++t) total += 1 + ((t == 0 || tp[t].jk != tp[t - 1].jk) ? 4 : 0);
    int acc = 0, q = 1;
    wbeg[0] = 0;
    for (int t = 0; t < m && q < WT; ++t) {
      acc += 1 + ((t == 0 || tp[t].jk != tp[t - 1].jk) ? 4 : 0);
      if ((t + 1) % unit == 0 && (int64_t)acc * WT >= (int64_t)total * q) wbeg[q++] = t + 1;
    }
    for (; q <= WT; ++q) wbeg[q] = m;
  }
  __syncthreads();

  // ---- tiles -----------------------------------------------------------------------------------------
  const bool pair_ok = (m % 2) == 0;
  const int t_begin = wbeg[wq], t_end = wbeg[wq + 1];
  const uint32_t team_bar = 1 + team;          // named barrier of the team
  const uint32_t team_threads = 32u * WT;
  auto team_sync = [&]() {
    if (WT > 1) asm volatile("bar.sync %0, %1;" ::"r"(team_bar), "r"(team_threads) : "memory");
    else __syncwarp();
  };
  // level j of the (possibly flipped) column sits in tile row row0 + j * rstep
  const int rstep = flip ? -TC : TC;
  const int row0 = flip ? (n - 1) * TC : 0;
  const int jmask = kUnclean - 1;
  int b = team % NB;       // ring buffer of the current tile (local tile i lives in buffer i % NB)
  uint32_t phase = 0;      // parity of that buffer's current use ((i / NB) & 1)
  for (int64_t i = team; i < nloc; i += NT) {
    const T* tile = reinterpret_cast<const T*>(in0 + (size_t)b * p.in_bytes) + lane + row0;
    // the team's previous bulk store must have finished READING the output buffer before anyone writes it
    if (wq == 0 && lane == 0) bulk_wait_read0();
    mbar_wait(full_u32 + 8u * b, phase);
    team_sync();
    b += NT;
    if (b >= NB) {
      b -= NB;
      phase ^= 1u;
    }

    // State: `cur` = j (every lane clean on interval j: plain targets take the fast path), j | kUnclean
    // (interval loaded, some lane must go the careful way), or -1 (nothing loaded).
    int cur = -1;
    double yj[CPL], slope[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) yj[c] = slope[c] = 0.0;
    // clean(c): both nodes finite and the interval has a usable reciprocal -> slope * dxt + yj cannot be
    // NaN, and (fp32 fields: |dy| is 0 or within [2^-149, 2^129]) the reciprocal sequence is exact
    auto advance = [&](int j) {
      const IntervalPlan e = iv[j];
      const bool rr_ok = __double2hiint(e.rr) != 0;  // rr is 0.0 or a normal number
      const T* row = tile + j * rstep;
      bool all = rr_ok;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const T ra = row[c * 32], rb = row[rstep + c * 32];
        yj[c] = (double)ra;
        const double dy = (double)rb - yj[c];
        const bool clean = rr_ok && is_finite<T>(ra) && is_finite<T>(rb);
        bool fast = clean;
        if constexpr (sizeof(T) == 8) fast = fast && exponent_safe(dy);
        slope[c] = fast ? div_with_recip(dy, e.dxj, e.rr) : slow_slope(dy, e.dxj);
        all = all && clean;
      }
      cur = __all_sync(0xffffffffu, all) ? j : (j | kUnclean);
    };
    // everything that is not "same clean interval, plain interpolation"; false = go on with the fast path
    auto careful = [&](int t, int jk, double dxt, T (&v)[CPL]) -> bool {
      const int kind = jk >> kKindShift, j = jk & jmask;
      if (kind >= PK_FIRST) {
        const int r = (kind == PK_FIRST) ? 0 : (n - 1) * rstep;
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = (kind == PK_NAN) ? T(NAN) : tile[r + c * 32];
        return true;
      }
      if (cur < 0 || j != (cur & jmask)) advance(j);
      if (kind == PK_EXACT) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = tile[j * rstep + c * 32];
        return true;
      }
      if (cur == j) return false;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        double res = slope[c] * dxt + yj[c];
        if (res != res) res = interp_retry(slope[c], xt[t], Xs[j + 1], yj[c], (double)tile[(j + 1) * rstep + c * 32]);
        v[c] = (T)res;
      }
      return true;
    };
    auto one_target = [&](int t, T (&v)[CPL]) {
      const TargetPlan e = tp[t];
      if (e.jk != cur) {
        if (careful(t, e.jk, e.dxt, v)) return;
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[c] = (T)(slope[c] * e.dxt + yj[c]);
    };
    if (pair_ok) {
#pragma unroll 1
      for (int t = t_begin; t < t_end; t += 2) {
        T v[2][CPL];
#pragma unroll
        for (int q = 0; q < 2; ++q) one_target(t + q, v[q]);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          T* o = out_tile + (size_t)(c * 32 + lane) * m + t;
          if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(o) = make_float2(v[0][c], v[1][c]);
          else *reinterpret_cast<double2*>(o) = make_double2(v[0][c], v[1][c]);
        }
      }
    } else {
#pragma unroll 1
      for (int t = t_begin; t < t_end; ++t) {
        T v[CPL];
        one_target(t, v);
#pragma unroll
        for (int c = 0; c < CPL; ++c) out_tile[(size_t)(c * 32 + lane) * m + t] = v[c];
      }
    }

    // ---- tile done: results out as one bulk store, buffer refilled with tile i + NB --------------------
    fence_async_smem();  // this lane's STS -> visible to the async proxy
    team_sync();         // ... for the whole team; also: everybody is done reading the input buffer
    if (wq == 0 && lane == 0) {
      int64_t o, i0;
      int ncol;
      tile_geom(i, o, i0, ncol);
      bulk_store(a.out + (o * a.inner + i0) * a.m, smem_u32(out_tile), (unsigned)ncol * (unsigned)m * sizeof(T));
      bulk_commit();
      if (i + NB < nloc) issue_load(i + NB);
    }
  }
  if (wq == 0 && lane == 0) bulk_wait_read0();
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <typename T, int CPL>
int launch_shared(const InterpArgs<T>& a, cudaStream_t st, int sms, int smem_max) {
  constexpr int TC = Geo<T, CPL>::TC;
  TmaArgs<T> p;
  p.a = a;
  const int n = (int)a.n, m = (int)a.m;
  p.tiles_per_o = xg_ceil_div(a.inner, TC);
  p.ntiles = a.outer * p.tiles_per_o;
  p.nbox = (int)xg_ceil_div(n, 256);
  p.box_rows = (int)xg_ceil_div(n, p.nbox);
  auto up128 = [](size_t v) { return (unsigned)((v + 127) / 128 * 128); };
  p.in_bytes = up128((size_t)p.nbox * p.box_rows * TC * sizeof(T));
  p.out_bytes = up128((size_t)TC * m * sizeof(T));
  // choose NT teams (tiles in flight) and NB = NT + extra ring buffers.  Lookahead first: two spare buffers keep the
  // TMA loads a full tile ahead of the teams (sweeps: 4 teams + 2 spare 1.11 ms, 4 + 1 1.22, 5 + 0 1.36 at C5), then
  // as many teams as still fit (named barriers: at most 15); WT warps per team up to 32 warps per block.
  const int want_nt = env_int("XG_VINTERP_W", 0), want_extra = env_int("XG_VINTERP_EXTRA", 2);
  int best_w = 0, best_nb = 0;
  unsigned best_plan = 0;
  for (int extra = want_extra; extra >= 0 && !best_w; --extra) {
    for (int NT = 15; NT >= 1; --NT) {
      if (want_nt && NT != want_nt) continue;
      const int NB = NT + extra;
      const size_t plan = (size_t)m * sizeof(TargetPlan) + (size_t)n * sizeof(IntervalPlan) +
                          (size_t)(n + m) * sizeof(double) + (size_t)NB * 8 + (size_t)(2 * m + 4 + 33) * sizeof(int);
      const unsigned plan_b = up128(plan);
      const size_t total = plan_b + (size_t)NB * p.in_bytes + (size_t)NT * p.out_bytes;
      if (total + 64 <= (size_t)smem_max) {  // + the static s_tmin / s_tmax
        best_w = NT;
        best_nb = NB;
        best_plan = plan_b;
        break;
      }
    }
  }
  if (!best_w) return 0;
  int wt = env_int("XG_VINTERP_WT", 0);
  if (wt <= 0) {
    wt = 32 / best_w;                     // up to 32 warps per block
    const int units = (int)(m / 2);       // a warp wants at least 2 pairs of targets
    while (wt > 1 && units / wt < 2) --wt;
  }
  if (wt < 1) wt = 1;
  while (wt > 1 && best_w * wt > 32) --wt;
  p.wt = wt;
  p.small = p.ntiles < (1ll << 31);
  p.nb = best_nb;
  p.plan_bytes = best_plan;

  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return 0;
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)a.inner, (cuuint64_t)a.n, (cuuint64_t)a.outer};
  const cuuint64_t gstr[2] = {(cuuint64_t)a.inner * sizeof(T), (cuuint64_t)a.n * a.inner * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)TC, (cuuint32_t)p.box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = enc(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                          const_cast<T*>(a.phi), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return 0;  // e.g. a stride beyond the descriptor's range: plain kernel

  const size_t smem = best_plan + (size_t)best_nb * p.in_bytes + (size_t)best_w * p.out_bytes;
  cudaError_t e = cudaFuncSetAttribute(k_vinterp_shared_tma<T, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return xg_fail(XG_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  int64_t blocks = xg_ceil_div(p.ntiles, best_w);
  if (blocks > sms) blocks = sms;
  k_vinterp_shared_tma<T, CPL><<<(unsigned)blocks, best_w * wt * 32, smem, st>>>(map, p);
  const int rc = xg_check_launch("xg_vinterp_linear(shared, tma)");
  return rc ? rc : 1;
}

}  // namespace

template <typename T>
int vinterp_shared_tma(const InterpArgs<T>& a, cudaStream_t st) {
  if (env_int("XG_VINTERP_TMA", 1) == 0) return 0;
  const int n = (int)a.n;
  if (n < 2 || n >= kUnclean) return 0;  // np.interp's single-node rule lives in the fallback kernel
  // TMA needs 16-byte aligned global strides and base; the bulk store 16-byte aligned tile starts
  if ((a.inner * sizeof(T)) % 16 != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(a.phi) | reinterpret_cast<uintptr_t>(a.out)) & 15) return 0;
  if (a.inner < 32 || a.inner >= (1ll << 31) || a.outer >= (1ll << 31)) return 0;
  int dev = 0;
  cudaGetDevice(&dev);
  // init-once device property cache (attribute queries cost microseconds each, the kernel ~1 ms)
  static int s_sms[64], s_smem[64];
  static bool s_have[64];
  int sms = 148, smem_max = 0;
  if (dev >= 0 && dev < 64 && s_have[dev]) {
    sms = s_sms[dev];
    smem_max = s_smem[dev];
  } else {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (dev >= 0 && dev < 64) {
      s_sms[dev] = sms;
      s_smem[dev] = smem_max;
      s_have[dev] = true;
    }
  }
  // columns per lane: 2 (64-column tiles) halves the per-column share of the plan reads and loop control, as
  // long as four or so tiles (n levels in + m targets out each) still fit in shared memory
  int cpl = env_int("XG_VINTERP_CPL", 0);
  if (cpl <= 0) cpl = ((size_t)(a.n + a.m) * 64 * sizeof(T) <= 48 * 1024) ? 2 : 1;
  if (cpl == 2) {
    const int r = launch_shared<T, 2>(a, st, sms, smem_max);
    if (r != 0) return r;
  }
  return launch_shared<T, 1>(a, st, sms, smem_max);
}

// =====================================================================================================
// theta FIELD (one theta column per phi column), shared target levels — `Grid.transform(da, 'Z', levels,
// target_data=<3-D field>)`, e.g. density coordinates.  Same machinery as above: tiles of 32 columns, phi AND
// theta levels of a tile staged by two TMA box loads on one mbarrier, the 32 x m results leave as one bulk
// store, teams of warps split a tile's targets.  What differs is the arithmetic: the interval search runs per
// column (lane = column, all reads from shared memory — the old kernel's dependent global loads, 0.13 of
// peak, are gone) and each (column, interval) needs a true fp64 division.
//
// np.interp's search is replayed exactly as k_vinterp_columns does: columns whose (possibly flipped) theta is
// NaN-free and sorted walk forward from the current interval (the guess cannot change the answer there);
// every other column replays binary_search_with_guess literally, carrying the guess from target 0 on (a warp
// that starts at target tb > 0 first replays the searches of targets 0 .. tb-1 for those columns).
namespace {

template <typename T>
struct ColArgs {
  InterpArgs<T> a;
  int64_t tiles_per_o, ntiles;
  int nb, wt;
  bool small;
  int box_rows, nbox;
  unsigned half_bytes;  // one field's box(es): nbox * box_rows * 32 * sizeof(T), rounded up to 128
  unsigned out_bytes, plan_bytes;
};

struct ColPartial {  // per (warp of the team, column): what its chunk of levels says about theta
  int first_k, last_k;  // first / last non-NaN level in the chunk (-1: none)
  int flags;            // 1: has NaN, 2: some pair ascends strictly, 4: some pair descends strictly
  int pad;
};

template <typename T>
__global__ void __launch_bounds__(1024, 1)
    k_vinterp_columns_tma(const __grid_constant__ CUtensorMap map_phi, const __grid_constant__ CUtensorMap map_theta,
                          const ColArgs<T> p) {
  constexpr int TC = 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const InterpArgs<T>& a = p.a;
  const int n = (int)a.n, m = (int)a.m;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int W = blockDim.x >> 5, NB = p.nb, WT = p.wt, NT = W / WT;
  const int team = w / WT, wq = w - team * WT;
  // layout: xt[m] (double) | xv[m] (T) | full[NB] | wbeg[33] | parts[NT][WT][32] | pad | in[NB] (phi, theta) | out[NT]
  double* xt = reinterpret_cast<double*>(smem_raw);
  T* xv = reinterpret_cast<T*>(xt + m);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + ((size_t)m * (8 + sizeof(T)) + 7) / 8 * 8);
  int* wbeg = reinterpret_cast<int*>(full + NB);
  ColPartial* parts = reinterpret_cast<ColPartial*>(wbeg + 36);
  unsigned char* in0 = smem_raw + p.plan_bytes;
  unsigned char* out0 = in0 + (size_t)NB * 2 * p.half_bytes;
  T* out_tile = reinterpret_cast<T*>(out0 + (size_t)team * p.out_bytes);
  const uint32_t full_u32 = smem_u32(full);

  const int64_t nloc = (p.ntiles > blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto tile_geom = [&](int64_t i, int64_t& o, int64_t& i0, int& ncol) {
    const int64_t g = i * gridDim.x + blockIdx.x;
    if (p.small) {
      const uint32_t oo = (uint32_t)g / (uint32_t)p.tiles_per_o;
      o = oo;
      i0 = (int64_t)((uint32_t)g - oo * (uint32_t)p.tiles_per_o) * TC;
    } else {
      o = g / p.tiles_per_o;
      i0 = (g - o * p.tiles_per_o) * TC;
    }
    const int64_t left = a.inner - i0;
    ncol = left < TC ? (int)left : TC;
  };
  auto issue_load = [&](int64_t i) {
    const int b = (int)(i % NB);
    int64_t o, i0;
    int ncol;
    tile_geom(i, o, i0, ncol);
    const uint32_t bar = full_u32 + 8u * b;
    const unsigned box_bytes = (unsigned)p.box_rows * TC * sizeof(T);
    mbar_expect_tx(bar, 2u * box_bytes * (unsigned)p.nbox);
    const uint32_t dst = smem_u32(in0 + (size_t)b * 2 * p.half_bytes);
    for (int k = 0; k < p.nbox; ++k) {
      tensor_load_3d(dst + k * box_bytes, &map_phi, (int)i0, k * p.box_rows, (int)o, bar);
      tensor_load_3d(dst + p.half_bytes + k * box_bytes, &map_theta, (int)i0, k * p.box_rows, (int)o, bar);
    }
  };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_phi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_theta) : "memory");
    for (int b = 0; b < NB; ++b) mbar_init(full_u32 + 8u * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wq == 0 && lane == 0)
    for (int64_t i = team; i < NB && i < nloc; i += NT) issue_load(i);
  // ---- per block: the target levels (log applied in the field dtype, transform.py:82-84) and the split
  const T* target = reinterpret_cast<const T*>(a.target.ptr);
  const bool logarithmic = a.logarithmic != 0;
  for (int t = tid; t < m; t += blockDim.x) {
    T v = __ldg(target + t * a.target.axis_stride);
    if (logarithmic) v = xg_log<T>(v);
    xv[t] = v;
    xt[t] = (double)v;
  }
  __syncthreads();
  if (tid == 0) {
    int sorted_t = 1;  // non-decreasing and NaN-free target levels: the level-major fast path applies
    for (int t = 0; t < m; ++t)
      if (xt[t] != xt[t] || (t > 0 && xt[t] < xt[t - 1])) sorted_t = 0;
    wbeg[0] = sorted_t;
  }
  __syncthreads();
  const bool tsorted = wbeg[0] != 0;
  const int k_begin = (int)((int64_t)n * wq / WT), k_end = (int)((int64_t)n * (wq + 1) / WT);
  const uint32_t team_bar = 1 + team, team_threads = 32u * WT;
  auto team_sync = [&]() {
    if (WT > 1) asm volatile("bar.sync %0, %1;" ::"r"(team_bar), "r"(team_threads) : "memory");
    else __syncwarp();
  };
  ColPartial* my_parts = parts + (size_t)team * WT * 32;

  int b = team % NB;
  uint32_t phase = 0;
  for (int64_t i = team; i < nloc; i += NT) {
    T* phi_t = reinterpret_cast<T*>(in0 + (size_t)b * 2 * p.half_bytes) + lane;
    T* th_t = reinterpret_cast<T*>(in0 + (size_t)b * 2 * p.half_bytes + p.half_bytes) + lane;
    if (wq == 0 && lane == 0) bulk_wait_read0();
    mbar_wait(full_u32 + 8u * b, phase);
    b += NT;
    if (b >= NB) {
      b -= NB;
      phase ^= 1u;
    }
    // ---- per-column facts about theta, the levels split between the warps of the team -----------------
    {
      if (logarithmic) {  // in place, once per tile: everything below reads log(theta)
        for (int k = k_begin; k < k_end; ++k) th_t[k * TC] = xg_log<T>(th_t[k * TC]);
        team_sync();
      }
      ColPartial cp;
      cp.first_k = cp.last_k = -1;
      cp.flags = 0;
      cp.pad = 0;
      T prev = T(0);
      bool have_prev = false;
      // the chunk plus the first level of the next one, so that every adjacent pair is looked at once
      const int k_stop = (k_end < n) ? k_end + 1 : n;
      for (int k = k_begin; k < k_stop; ++k) {
        const T v = th_t[k * TC];
        if (xg_isnan(v)) {
          if (k < k_end) cp.flags |= 1;
          have_prev = false;  // a NaN breaks the walk property anyway
          continue;
        }
        if (k < k_end) {
          if (cp.first_k < 0) cp.first_k = k;
          cp.last_k = k;
        }
        if (have_prev) {
          if (v > prev) cp.flags |= 2;
          if (v < prev) cp.flags |= 4;
        }
        prev = v;
        have_prev = true;
      }
      my_parts[wq * 32 + lane] = cp;
    }
    team_sync();  // partials + (log) transformed theta visible to the whole team
    bool flip = false, walk = false;
    T tmin = T(0), tmax = T(0);
    {
      int first_k = -1, last_k = -1, flags = 0;
      for (int q = 0; q < WT; ++q) {
        const ColPartial cp = my_parts[q * 32 + lane];
        if (cp.first_k >= 0 && first_k < 0) first_k = cp.first_k;
        if (cp.last_k >= 0) last_k = cp.last_k;
        flags |= cp.flags;
      }
      const bool any = first_k >= 0;
      if (!a.bypass_checks && any) flip = th_t[last_k * TC] < th_t[first_k * TC];  // transform.py:27-31
      const bool nan = (flags & 1) != 0;
      const bool sorted = flip ? (flags & 2) == 0 : (flags & 4) == 0;
      walk = !nan && sorted && n > 1;
      if (walk) {  // sorted: the extremes are the two ends
        const T e0 = th_t[0], e1 = th_t[(n - 1) * TC];
        tmin = e0 < e1 ? e0 : e1;
        tmax = e0 < e1 ? e1 : e0;
      } else if (a.mask_edges) {  // nanmin / nanmax (transform.py:36-37): rare columns, full scan
        bool got = false;
        for (int k = 0; k < n; ++k) {
          const T v = th_t[k * TC];
          if (xg_isnan(v)) continue;
          if (!got) { tmin = tmax = v; got = true; }
          else { tmin = v < tmin ? v : tmin; tmax = v > tmax ? v : tmax; }
        }
        if (!got) tmin = tmax = T(NAN);
      }
    }
    const int rstep = flip ? -TC : TC, row0 = flip ? (n - 1) * TC : 0;
    auto X = [&](int k) -> double { return (double)th_t[row0 + k * rstep]; };
    auto Y = [&](int k) -> double { return (double)phi_t[row0 + k * rstep]; };
    T* orow = out_tile + (size_t)lane * m;
    const bool fast = walk && tsorted;
    if (fast) {
      // ---- level-major: this warp owns the intervals [k_begin, k_end) of every column; a column's targets
      // that fall into them are found by position (targets are sorted), so each interval's slope is formed
      // once, uniformly across the lanes, and every target is written by exactly one warp of the team.
      // np.interp on sorted NaN-free theta: j = the largest index with X[j] <= x.
      int t = 0;
      {  // first target at or right of X[k_begin] (lower bound over the sorted levels)
        const double xb = X(k_begin);
        int lo_t = 0, hi_t = m;
        while (lo_t < hi_t) {
          const int mid = lo_t + ((hi_t - lo_t) >> 1);
          if (xt[mid] < xb) lo_t = mid + 1;
          else hi_t = mid;
        }
        t = lo_t;
      }
      if (wq == 0) {  // targets left of the first node: np.interp gives Y(0), the edge mask NaN
        const T y0 = a.mask_edges ? T(NAN) : phi_t[row0];
        for (int tt = 0; tt < t; ++tt) orow[tt] = y0;
      }
      const T* pth = th_t + row0 + k_begin * rstep;
      const T* pph = phi_t + row0 + k_begin * rstep;
      double xk = (double)pth[0], yk = (double)pph[0];
      const int k_last = (k_end < n - 1) ? k_end : n - 1;  // intervals k_begin .. k_last - 1
#pragma unroll 1
      for (int k = k_begin; k < k_last; ++k) {
        pth += rstep;
        pph += rstep;
        const double xk1 = (double)pth[0], yk1 = (double)pph[0];
        const double slope = (yk1 - yk) / (xk1 - xk);
        while (t < m) {
          const double x = xt[t];
          if (!(x < xk1)) break;
          const double res = (xk == x) ? yk : interp_value(x, xk, xk1, yk, yk1, slope);
          orow[t] = (T)res;  // inside [X(0), X(n-1)): never masked
          ++t;
        }
        xk = xk1;
        yk = yk1;
      }
      if (wq == WT - 1) {  // at or right of the last node: Y(n-1); strictly right of it the edge mask applies
        const double xl = (double)th_t[row0 + (n - 1) * rstep];
        const T yl = phi_t[row0 + (n - 1) * rstep];
        for (; t < m; ++t) orow[t] = (a.mask_edges && xt[t] > xl) ? T(NAN) : yl;
      }
    } else if (lane % WT == wq) {
      // ---- everything else (theta with NaNs or out of order, unsorted / NaN targets, a single level): the
      // literal replay of np.interp for the whole column, such columns dealt round-robin to the team's warps
      int guess = 0, cj = -2;
      double yj = 0.0, yj1 = 0.0, slope = 0.0, xj = 0.0, xj1 = 0.0;
      for (int t = 0; t < m; ++t) {
        const double x = xt[t];
        double res;
        if (n == 1) {
          res = Y(0);
        } else if (x != x) {
          res = x;
        } else {
          const int j = search_with_guess(x, X, n, guess);
          guess = j;
          if (j == -1) res = Y(0);
          else if (j >= n - 1) res = Y(n - 1);
          else {
            if (j != cj) {
              cj = j;
              xj = X(j);
              xj1 = X(j + 1);
              yj = Y(j);
              yj1 = Y(j + 1);
              slope = (yj1 - yj) / (xj1 - xj);
            }
            res = (xj == x) ? yj : interp_value(x, xj, xj1, yj, yj1, slope);
          }
        }
        if (a.mask_edges && (xv[t] < tmin || xv[t] > tmax)) res = NAN;  // transform.py:38-41
        orow[t] = (T)res;
      }
    }
    fence_async_smem();
    team_sync();
    if (wq == 0 && lane == 0) {
      int64_t o, i0;
      int ncol;
      tile_geom(i, o, i0, ncol);
      bulk_store(a.out + (o * a.inner + i0) * a.m, smem_u32(out_tile), (unsigned)ncol * (unsigned)m * sizeof(T));
      bulk_commit();
      if (i + NB < nloc) issue_load(i + NB);
    }
  }
  if (wq == 0 && lane == 0) bulk_wait_read0();
}

template <typename T>
int encode_field_map(EncodeTiledFn enc, CUtensorMap* map, const T* base, int64_t inner, int64_t n, int64_t outer,
                     int box_cols, int box_rows) {
  const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)n, (cuuint64_t)outer};
  const cuuint64_t gstr[2] = {(cuuint64_t)inner * sizeof(T), (cuuint64_t)n * inner * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
             const_cast<T*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

}  // namespace

template <typename T>
int vinterp_columns_tma(const InterpArgs<T>& a, cudaStream_t st) {
  if (env_int("XG_VINTERP_TMA", 1) == 0) return 0;
  const int n = (int)a.n, m = (int)a.m;
  if (n < 1 || n >= (1 << 23) || m < 1) return 0;
  if ((a.inner * sizeof(T)) % 16 != 0) return 0;
  const T* theta = reinterpret_cast<const T*>(a.theta.ptr);
  if ((reinterpret_cast<uintptr_t>(a.phi) | reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(a.out)) & 15)
    return 0;
  if (a.inner < 32 || a.inner >= (1ll << 31) || a.outer >= (1ll << 31)) return 0;
  int dev = 0, sms = 148, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  constexpr int TC = 32;
  ColArgs<T> p;
  p.a = a;
  p.tiles_per_o = xg_ceil_div(a.inner, TC);
  p.ntiles = a.outer * p.tiles_per_o;
  p.small = p.ntiles < (1ll << 31);
  p.nbox = (int)xg_ceil_div(n, 256);
  p.box_rows = (int)xg_ceil_div(n, p.nbox);
  auto up128 = [](size_t v) { return (unsigned)((v + 127) / 128 * 128); };
  p.half_bytes = up128((size_t)p.nbox * p.box_rows * TC * sizeof(T));
  p.out_bytes = up128((size_t)TC * m * sizeof(T));
  const int want_nt = env_int("XG_VINTERP_W", 0), want_extra = env_int("XG_VINTERP_EXTRA", 2);
  int nt = 0, nb = 0;
  unsigned plan_b = 0;
  for (int NT = 15; NT >= 1 && !nt; --NT) {
    if (want_nt && NT != want_nt) continue;
    for (int extra = want_extra; extra >= 0; --extra) {
      const int NB = NT + extra;
      // the partials array is sized for the largest team this NT allows
      const int wt_max = 32 / NT > 0 ? 32 / NT : 1;
      const size_t plan = ((size_t)m * (8 + sizeof(T)) + 7) / 8 * 8 + (size_t)NB * 8 + 36 * sizeof(int) +
                          (size_t)NT * wt_max * 32 * sizeof(ColPartial);
      const unsigned pb = up128(plan);
      const size_t total = pb + (size_t)NB * 2 * p.half_bytes + (size_t)NT * p.out_bytes;
      if (total + 64 <= (size_t)smem_max) {
        nt = NT;
        nb = NB;
        plan_b = pb;
        break;
      }
    }
  }
  if (!nt) return 0;
  int wt = env_int("XG_VINTERP_WT", 0);
  if (wt <= 0) {
    wt = 32 / nt;
    while (wt > 1 && (m / 2) / wt < 2) --wt;
  }
  if (wt < 1) wt = 1;
  while (wt > 1 && nt * wt > 32) --wt;
  p.wt = wt;
  p.nb = nb;
  p.plan_bytes = plan_b;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return 0;
  CUtensorMap map_phi, map_theta;
  if (encode_field_map<T>(enc, &map_phi, a.phi, a.inner, a.n, a.outer, TC, p.box_rows)) return 0;
  if (encode_field_map<T>(enc, &map_theta, theta, a.inner, a.n, a.outer, TC, p.box_rows)) return 0;
  const size_t smem = plan_b + (size_t)nb * 2 * p.half_bytes + (size_t)nt * p.out_bytes;
  cudaError_t e = cudaFuncSetAttribute(k_vinterp_columns_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return xg_fail(XG_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  int64_t blocks = xg_ceil_div(p.ntiles, nt);
  if (blocks > sms) blocks = sms;
  k_vinterp_columns_tma<T><<<(unsigned)blocks, nt * wt * 32, smem, st>>>(map_phi, map_theta, p);
  const int rc = xg_check_launch("xg_vinterp_linear(columns, tma)");
  return rc ? rc : 1;
}

template int vinterp_shared_tma<float>(const InterpArgs<float>&, cudaStream_t);
template int vinterp_shared_tma<double>(const InterpArgs<double>&, cudaStream_t);
template int vinterp_columns_tma<float>(const InterpArgs<float>&, cudaStream_t);
template int vinterp_columns_tma<double>(const InterpArgs<double>&, cudaStream_t);

}  // namespace xgvi
