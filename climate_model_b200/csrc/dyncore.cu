// dyncore.cu -- libdyncore.so: the B200 (sm_100a) dynamical core behind include/dyncore.h.
// CUDA backend of dc_api_impl.h: every kernel body of dc_kernels.h is wrapped in a
// __global__ kernel with one thread per (lon, lat) column, longitude along threadIdx.x
// so that each level's loads and stores coalesce.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared ...
// -fmad=false keeps the reference's evaluation order (numba emits no FMA contraction).
#include <cuda_runtime.h>

#define DC_BACKEND_IS_CUDA 1

static int dcb_malloc(void **p, size_t n) { return (int)cudaMalloc(p, n); }
static int dcb_free(void *p) { return (int)cudaFree(p); }
static int dcb_h2d(void *dst, const void *src, size_t n)
{
    return (int)cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
}
static int dcb_d2d_async(void *dst, const void *src, size_t n, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
static int dcb_last_error() { return (int)cudaGetLastError(); }
static const char *dcb_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

namespace dc {
constexpr int BX = 64, BY = 4;  // 256 threads: 64 consecutive longitudes x 4 rows

template <class Body>
__global__ void __launch_bounds__(BX *BY) k_columns(const Body b, int i0, int i1, int j0, int j1)
{
    const int i = i0 + blockIdx.x * BX + threadIdx.x;
    const int j = j0 + blockIdx.y * BY + threadIdx.y;
    if (i <= i1 && j <= j1) b(i, j);
}
}  // namespace dc

template <class Body>
static void dcb_launch(const Body &b, int i0, int i1, int j0, int j1, void *stream)
{
    dim3 block(dc::BX, dc::BY);
    dim3 grid((i1 - i0 + dc::BX) / dc::BX, (j1 - j0 + dc::BY) / dc::BY);
    dc::k_columns<Body><<<grid, block, 0, (cudaStream_t)stream>>>(b, i0, i1, j0, j1);
}

#include "dc_api_impl.h"
