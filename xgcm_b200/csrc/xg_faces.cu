// xg_strided_copy: the data movement of face-connection padding.
//
// xgcm/padding.py:260-572 (_pad_face_connections) fills the halo of a face edge with a slab of
// the neighbouring face: sliced (padding.py:443-459), possibly with the two horizontal dims
// swapped (:466-472), flipped across and / or along the seam (:478-498) and sign-flipped for
// vector components.  Every one of those is an affine index map, so one edge is ONE copy with
// signed element strides on the source side; the host (xgcm_b200/padding.py) works the strides
// out and this kernel moves the bytes.  Halo slabs are thin, so the kernel is plain: flat index
// over the destination order (coalesced stores), at most 8 collapsed dims.
#include "xg_common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
struct CopyArgs {
  T* dst;
  const T* src;
  int ndim;
  int64_t total;
  int64_t shape[XG_MAX_NDIM];
  int64_t dstride[XG_MAX_NDIM];
  int64_t sstride[XG_MAX_NDIM];
  int negate;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) k_strided_copy(const CopyArgs<T> a) {
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < a.total;
       g += (int64_t)gridDim.x * kThreads) {
    int64_t rem = g, doff = 0, soff = 0;
#pragma unroll
    for (int d = XG_MAX_NDIM - 1; d >= 0; --d) {
      if (d < a.ndim) {
        const int64_t q = rem / a.shape[d];
        const int64_t c = rem - q * a.shape[d];
        rem = q;
        doff += c * a.dstride[d];
        soff += c * a.sstride[d];
      }
    }
    const T v = a.src[soff];
    a.dst[doff] = a.negate ? -v : v;
  }
}

template <typename T>
int copy_typed(void* dst, const int64_t* dst_strides, const void* src, const int64_t* src_strides,
               int ndim, const int64_t* shape, int negate, cudaStream_t st) {
  CopyArgs<T> a;
  a.dst = static_cast<T*>(dst);
  a.src = static_cast<const T*>(src);
  a.negate = negate;
  a.total = 1;
  // drop unit dims, merge neighbours that are contiguous on both sides
  int k = 0;
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] < 0) return xg_fail(XG_EINVAL, "xg_strided_copy: negative extent");
    a.total *= shape[d];
    if (shape[d] == 1) continue;
    if (k > 0 && a.dstride[k - 1] == dst_strides[d] * shape[d] &&
        a.sstride[k - 1] == src_strides[d] * shape[d]) {
      a.shape[k - 1] *= shape[d];
      a.dstride[k - 1] = dst_strides[d];
      a.sstride[k - 1] = src_strides[d];
    } else {
      a.shape[k] = shape[d];
      a.dstride[k] = dst_strides[d];
      a.sstride[k] = src_strides[d];
      ++k;
    }
  }
  a.ndim = k;
  for (int d = k; d < XG_MAX_NDIM; ++d) {
    a.shape[d] = 1;
    a.dstride[d] = a.sstride[d] = 0;
  }
  if (a.total == 0) return XG_OK;
  int64_t blocks = xg_ceil_div(a.total, kThreads);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_strided_copy<T><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  return xg_check_launch("xg_strided_copy");
}

}  // namespace

extern "C" int xg_strided_copy(int dtype, void* dst, const int64_t* dst_strides, const void* src,
                               const int64_t* src_strides, int ndim, const int64_t* shape,
                               int negate, void* stream) {
  if (!dst || !src || !dst_strides || !src_strides || !shape)
    return xg_fail(XG_EINVAL, "xg_strided_copy: null pointer");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_strided_copy: bad ndim");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XG_F32)
    return copy_typed<float>(dst, dst_strides, src, src_strides, ndim, shape, negate, st);
  if (dtype == XG_F64)
    return copy_typed<double>(dst, dst_strides, src, src_strides, ndim, shape, negate, st);
  return xg_fail(XG_EINVAL, "xg_strided_copy: dtype must be XG_F32 or XG_F64");
}
