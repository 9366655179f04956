// pair_common.cuh -- device helpers shared by the persistent sweep kernels (pair_kernels2.cu, pair_kernels3.cu).
#pragma once
#include "internal.cuh"

namespace rb {
namespace {

__device__ __forceinline__ double fast_rcp2(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    double t = fma(e, e, e);
    return fma(r, t, r);
}

struct __align__(16) Src2 {
    double p, q, fr, fi;
};

template <typename T>
__device__ __forceinline__ void mirror_store2(const CommView& c, T* local_ptr, T v) {
    if (c.nranks <= 1) {
        *local_ptr = v;
        return;
    }
    const size_t off = (size_t)(reinterpret_cast<char*>(local_ptr) - c.my_base);
    for (int r = 0; r < c.nranks; ++r) *reinterpret_cast<T*>(c.peer_base[r] + off) = v;
}

__device__ __forceinline__ void comm_signal2(const CommView& c) {
    const unsigned long long e = *c.signal_epoch + 1ull;
    *c.signal_epoch = e;
    for (int r = 0; r < c.nranks; ++r) {
        volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(c.peer_base[r] + c.off_flags) + c.rank;
        *f = e;
    }
    __threadfence_system();
}

// deterministic block reductions for an arbitrary (fixed per launch) thread count
__device__ __forceinline__ double block_sum_any(double v, double* sred, int T, int P2) {
    sred[threadIdx.x] = v;
    __syncthreads();
    for (int w = P2 >> 1; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && (int)threadIdx.x + w < T) sred[threadIdx.x] += sred[threadIdx.x + w];
        __syncthreads();
    }
    double r = sred[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ double block_max_any(double v, double* sred, int T, int P2) {
    sred[threadIdx.x] = v;
    __syncthreads();
    for (int w = P2 >> 1; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && (int)threadIdx.x + w < T) sred[threadIdx.x] = fmax(sred[threadIdx.x], sred[threadIdx.x + w]);
        __syncthreads();
    }
    double r = sred[0];
    __syncthreads();
    return r;
}

}  // namespace
}  // namespace rb
