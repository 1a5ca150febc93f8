// Entry points declared in the header but not implemented yet: fail loudly.
#include "xg_common.cuh"

#define XG_TODO(name) return xg_fail(XG_ENOTIMPL, name ": not implemented yet")

extern "C" int xg_cumscan(int, const void*, void*, int, const int64_t*, int, int, int, int, int, int,
                          double, const void*, const int64_t*, const void*, const int64_t*, int,
                          void*) { XG_TODO("xg_cumscan"); }
extern "C" int xg_wreduce(int, const void*, const void*, const int64_t*, void*, int, const int64_t*,
                          int, int, int, void*) { XG_TODO("xg_wreduce"); }
extern "C" int xg_vinterp_linear(int, const void*, const void*, const int64_t*, const void*, int64_t,
                                 void*, int, const int64_t*, int, int, int, int, void*) {
  XG_TODO("xg_vinterp_linear");
}
extern "C" int xg_pad(int, const void*, void*, int, const int64_t*, int, int, int, int, double,
                      void*) { XG_TODO("xg_pad"); }
extern "C" int xg_binary(int, int, const void*, const void*, const int64_t*, void*, int,
                         const int64_t*, void*) { XG_TODO("xg_binary"); }
extern "C" int xg_fill_uniform(int, void*, int64_t, uint64_t, uint64_t, void*) {
  XG_TODO("xg_fill_uniform");
}
extern "C" int xg_fill_uniform_host(int, void*, int64_t, uint64_t, uint64_t) {
  XG_TODO("xg_fill_uniform_host");
}
extern "C" int xg_stencil2_host(int, int, const void*, void*, int, const int64_t*, int, int, int,
                                int, double, const void*, const int64_t*, const void*,
                                const int64_t*, int) { XG_TODO("xg_stencil2_host"); }
