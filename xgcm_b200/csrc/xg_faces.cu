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

// ---- all connected edges of a field in one launch ----------------------------------------------
constexpr int kBatchDims = 5;
constexpr int kBatchMax = 24;  // 24 x 152 B of descriptors stay under the 4 KB kernel-parameter limit

struct EdgeDesc {
  void* dst;
  const void* src;
  int64_t total;
  int64_t shape[kBatchDims], dstride[kBatchDims], sstride[kBatchDims];
  int ndim, negate;
};
struct BatchArgs {
  EdgeDesc e[kBatchMax];
};

template <typename T>
__global__ void __launch_bounds__(kThreads) k_strided_copy_batch(const __grid_constant__ BatchArgs a) {
  const EdgeDesc& e = a.e[blockIdx.y];
  T* dst = static_cast<T*>(e.dst);
  const T* src = static_cast<const T*>(e.src);
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < e.total;
       g += (int64_t)gridDim.x * kThreads) {
    int64_t rem = g, doff = 0, soff = 0;
#pragma unroll
    for (int d = kBatchDims - 1; d >= 0; --d) {
      if (d < e.ndim) {
        const int64_t q = rem / e.shape[d];
        const int64_t c = rem - q * e.shape[d];
        rem = q;
        doff += c * e.dstride[d];
        soff += c * e.sstride[d];
      }
    }
    const T v = src[soff];
    dst[doff] = e.negate ? -v : v;
  }
}

// collapse one edge into a descriptor; false when it needs more than kBatchDims dims
bool make_edge(EdgeDesc& e, void* dst, const int64_t* ds, const void* src, const int64_t* ss, int ndim,
               const int64_t* shape, int negate) {
  e.dst = dst;
  e.src = src;
  e.negate = negate;
  e.total = 1;
  int k = 0;
  for (int d = 0; d < ndim; ++d) {
    e.total *= shape[d];
    if (shape[d] == 1) continue;
    if (k > 0 && e.dstride[k - 1] == ds[d] * shape[d] && e.sstride[k - 1] == ss[d] * shape[d]) {
      e.shape[k - 1] *= shape[d];
      e.dstride[k - 1] = ds[d];
      e.sstride[k - 1] = ss[d];
    } else {
      if (k == kBatchDims) return false;
      e.shape[k] = shape[d];
      e.dstride[k] = ds[d];
      e.sstride[k] = ss[d];
      ++k;
    }
  }
  e.ndim = k;
  for (int d = k; d < kBatchDims; ++d) {
    e.shape[d] = 1;
    e.dstride[d] = e.sstride[d] = 0;
  }
  return true;
}

}  // namespace

extern "C" int xg_strided_copy_batch(int dtype, int count, void* const* dst, const void* const* src,
                                     int ndim, const int64_t* shapes, const int64_t* dst_strides,
                                     const int64_t* src_strides, const int* negate, void* stream) {
  if (count < 0) return xg_fail(XG_EINVAL, "xg_strided_copy_batch: negative count");
  if (count == 0) return XG_OK;
  if (!dst || !src || !shapes || !dst_strides || !src_strides || !negate)
    return xg_fail(XG_EINVAL, "xg_strided_copy_batch: null pointer");
  if (ndim < 1 || ndim > XG_MAX_NDIM) return xg_fail(XG_EINVAL, "xg_strided_copy_batch: bad ndim");
  if (dtype != XG_F32 && dtype != XG_F64)
    return xg_fail(XG_EINVAL, "xg_strided_copy_batch: dtype must be XG_F32 or XG_F64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int first = 0; first < count; first += kBatchMax) {
    const int nb = (count - first < kBatchMax) ? (count - first) : kBatchMax;
    BatchArgs a;
    int64_t longest = 0;
    for (int k = 0; k < nb; ++k) {
      const int i = first + k;
      if (!dst[i] || !src[i]) return xg_fail(XG_EINVAL, "xg_strided_copy_batch: null pointer");
      for (int d = 0; d < ndim; ++d)
        if (shapes[(size_t)i * ndim + d] < 0)
          return xg_fail(XG_EINVAL, "xg_strided_copy_batch: negative extent");
      if (!make_edge(a.e[k], dst[i], dst_strides + (size_t)i * ndim, src[i], src_strides + (size_t)i * ndim,
                     ndim, shapes + (size_t)i * ndim, negate[i]))
        return xg_fail(XG_ENOTIMPL, "xg_strided_copy_batch: an edge needs more than 5 collapsed dims");
      if (a.e[k].total > longest) longest = a.e[k].total;
    }
    for (int k = nb; k < kBatchMax; ++k) a.e[k] = a.e[0], a.e[k].total = 0;
    if (longest == 0) continue;
    int64_t bx = xg_ceil_div(longest, kThreads);
    if (bx > 148 * 4) bx = 148 * 4;
    dim3 grid((unsigned)bx, (unsigned)nb);
    if (dtype == XG_F32) k_strided_copy_batch<float><<<grid, kThreads, 0, st>>>(a);
    else k_strided_copy_batch<double><<<grid, kThreads, 0, st>>>(a);
    int rc = xg_check_launch("xg_strided_copy_batch");
    if (rc) return rc;
  }
  return XG_OK;
}

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
