// TMA-staged tile kernel for stencils whose metrics are shared between levels — interface.
//
// Everything is seen as (z, p, x): Zn levels, rows, n cells of the innermost dim.  One launch evaluates
//
//     out = ( OPa(pad_a(A x ma)) along x   (+|-)   OPb(pad_b(B x mb)) along p ) / post
//
// with either term alone (A absent: a metric-fused stencil along the second-to-last dim, e.g.
// derivative('Y'); both: the C-grid divergence / vorticity of xg_stencil_pair), rounding operator by
// operator exactly like the register-staged kernels.  Implementation and design notes:
// xg_stencil_tile.cu.
#pragma once
#include "xg_common.cuh"

// an operand laid out against the (z, p, x) view: element strides, 0 = broadcast along that dim
template <typename T>
struct XgTileOperand {
  const T* ptr;  // nullptr: absent
  int64_t sz, sp, sx;
};

template <typename T>
struct XgTileSpec {
  int64_t Zn, Pb, Po, n;  // levels, rows of B, rows of the output (= Pb + lo_b + hi_b - 1; = rows of A), cells per row
  // element strides (rows are x-contiguous): f_sp between rows of A, B and out; b_sz between levels of B;
  // o_sz between levels of out and A.  Contiguous (z, p, x) arrays: f_sp = n, b_sz = Pb * n, o_sz = Po * n.
  // A stencil along Z of a (Z, Y, X) field runs as rows = Z (f_sp = Y * X) and levels = Y (b_sz = o_sz = X).
  int64_t f_sp, b_sz, o_sz;
  const T* a;             // x term, nullptr when absent (then lo_a.. are ignored); lo_a + hi_a == 1
  int op_a, lo_a, bc_a;
  T fill_a;
  const T* b;             // row term (always present)
  int op_b, lo_b, hi_b, bc_b;
  T fill_b;
  const T* halo_lo;       // optional halo planes (Zn, n) of the row term, replacing bc_b on their side
  const T* halo_hi;
  int subtract;           // 0: a + b, 1: a - b, 2: b - a
  XgTileOperand<T> ma, mb, post;
  T* out;
};

// Returns XG_OK with *launched = false when the layout does not qualify (caller runs its own kernel).
template <typename T>
int xg_tile_stencil(const XgTileSpec<T>& spec, cudaStream_t st, bool* launched, const char* label);

// (outer, n, inner)-collapsed operand -> (z, p, x) strides when `inner` is exactly the x dim; false if it
// cannot be expressed (more than one index group on either side)
template <typename T>
inline bool xg_tile_operand_from(const XgOperand& m, int64_t outer, int64_t inner, XgTileOperand<T>* out) {
  out->ptr = static_cast<const T*>(m.ptr);
  out->sz = out->sp = out->sx = 0;
  if (!m.ptr) return true;
  if (m.outer.n > 1 || m.inner.n > 1) return false;
  if (m.outer.n == 1) {
    if (m.outer.size[0] != outer) return false;
    out->sz = m.outer.stride[0];
  }
  if (m.inner.n == 1) {
    if (m.inner.size[0] != inner) return false;
    out->sx = m.inner.stride[0];
  }
  out->sp = m.axis_stride;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// TMA-staged fused multi-axis stencil (xg_stencil_multi_tma.cu): one operator along x, rows and / or levels
// of a contiguous (L, P, n) array, applied innermost axis first, every op length preserving.
template <typename T>
struct XgMultiTileSpec {
  const T* in;
  T* out;
  int64_t L, P, n;
  int op;      // XG_OP_*
  int has[3];  // x, rows, levels operated?
  int lo[3];   // halo below (hi = 1 - lo) per operated axis
  int bc[3];   // XG_BC_PERIODIC / FILL / EXTEND
  T fill[3];
};
template <typename T>
int xg_multi_tile(const XgMultiTileSpec<T>& spec, cudaStream_t st, bool* launched);
