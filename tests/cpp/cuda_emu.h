// cuda_emu.h -- TEST INFRASTRUCTURE: a minimal host emulation of the CUDA execution model, enough to run this library's small kernels
// (index arithmetic, shared-memory staging, __syncthreads, warp shuffles, the m8n8k4 FP64 mma fragment layout) under g++ in the
// CPU test tier, where there is no GPU.  One CUDA thread = one std::thread, one block at a time; __syncthreads is a std::barrier
// over the block's live threads, a warp shuffle is an exchange through a per-warp buffer between two warp barriers.  A thread that
// returns from the kernel drops out of its barriers (a CUDA thread that has exited does not take part in later barriers either).
//
// What it proves: that a kernel source, compiled unchanged (apart from this header standing in for the CUDA headers), computes what
// the oracle computes when executed by the CUDA programming model's rules.  What it does not prove: anything about timing, memory
// ordering between blocks, or that the hardware's mma fragment layout is the documented one (dmma_m8n8k4 below encodes the PTX ISA
// layout: A[g][t], B[t][g], C[g][2t], C[g][2t+1] with g = lane / 4, t = lane % 4).
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define RB_EMULATE 1
#ifdef RB_EMU_CUDA_HEADERS
// For sources that include the toolkit's own headers (include/cusuperhelium_compat.cuh: cuda_runtime.h, cuComplex.h, libcu++): the
// types and the __global__ / __device__ markers are the toolkit's (they compile under g++ as they stand); only the execution model
// -- the index variables and the launch loops below -- is supplied here.  No runtime call is stubbed in this mode.
#include <cuda_runtime.h>
#else
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static   // blocks run one at a time: function-local statics are shared by exactly the threads of the running block

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
#endif

namespace emu {
struct Warp {
    double dbuf[2][32];
    long long ibuf[32];
    std::unique_ptr<std::barrier<>> bar;
};
inline thread_local std::barrier<>* block_bar = nullptr;
inline thread_local Warp* warp = nullptr;
inline thread_local int lane = 0;
}  // namespace emu

inline thread_local uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() {
    if (!emu::block_bar) throw std::logic_error("__syncthreads in a kernel launched without barriers");
    emu::block_bar->arrive_and_wait();
}

inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp->bar->arrive_and_wait(); }

template <typename T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
    static_assert(sizeof(T) <= sizeof(long long), "shuffle of at most 8 bytes");
    emu::Warp& w = *emu::warp;
    std::memcpy(&w.ibuf[emu::lane], &v, sizeof(T));
    w.bar->arrive_and_wait();
    T r = v;
    if (emu::lane + delta < 32) std::memcpy(&r, &w.ibuf[emu::lane + delta], sizeof(T));
    w.bar->arrive_and_wait();
    return r;
}

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= sizeof(long long), "shuffle of at most 8 bytes");
    emu::Warp& w = *emu::warp;
    std::memcpy(&w.ibuf[emu::lane], &v, sizeof(T));
    w.bar->arrive_and_wait();
    T r;
    std::memcpy(&r, &w.ibuf[src & 31], sizeof(T));
    w.bar->arrive_and_wait();
    return r;
}

// mma.sync.aligned.m8n8k4.row.col.f64: D (8 x 8) += A (8 x 4) B (4 x 8); lane (g, t) holds A[g][t], B[t][g], D[g][2t], D[g][2t + 1]
inline void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    emu::Warp& w = *emu::warp;
    w.dbuf[0][emu::lane] = a;
    w.dbuf[1][emu::lane] = b;
    w.bar->arrive_and_wait();
    const int g = emu::lane >> 2, t = emu::lane & 3;
    double s0 = d0, s1 = d1;
    for (int k = 0; k < 4; ++k) {
        const double av = w.dbuf[0][g * 4 + k];
        s0 = std::fma(av, w.dbuf[1][(2 * t) * 4 + k], s0);
        s1 = std::fma(av, w.dbuf[1][(2 * t + 1) * 4 + k], s1);
    }
    w.bar->arrive_and_wait();
    d0 = s0;
    d1 = s1;
}

#ifndef RB_EMU_CUDA_HEADERS
inline void sincos(double x, double* s, double* c) {
    *s = std::sin(x);
    *c = std::cos(x);
}
#endif

namespace emu {

// every thread of every block, one after the other, on the calling OS thread: for kernels without barriers or shuffles
template <typename K, typename... A>
void launch_seq(K kernel, dim3 grid, dim3 block, A... args) {
    block_bar = nullptr;
    warp = nullptr;
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx) {
                            blockIdx = uint3{bx, by, bz};
                            threadIdx = uint3{tx, ty, tz};
                            kernel(args...);
                        }
}

// one std::thread per CUDA thread, one block at a time
template <typename K, typename... A>
void launch_coop(K kernel, dim3 grid, dim3 block, A... args) {
    const int nthreads = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::barrier<> bar(nthreads);
                const int nwarps = (nthreads + 31) / 32;
                std::vector<Warp> warps(nwarps);
                for (int w = 0; w < nwarps; ++w) warps[w].bar = std::make_unique<std::barrier<>>(std::min(32, nthreads - 32 * w));
                std::vector<std::thread> pool;
                pool.reserve(nthreads);
                std::string error;
                for (int tid = 0; tid < nthreads; ++tid)
                    pool.emplace_back([&, tid] {
                        gridDim = grid;
                        blockDim = block;
                        blockIdx = uint3{bx, by, bz};
                        threadIdx = uint3{(unsigned)tid % block.x, ((unsigned)tid / block.x) % block.y, (unsigned)tid / (block.x * block.y)};
                        block_bar = &bar;
                        warp = &warps[tid / 32];
                        lane = tid % 32;
                        kernel(args...);
                        warps[tid / 32].bar->arrive_and_drop();
                        bar.arrive_and_drop();
                    });
                for (auto& t : pool) t.join();
            }
}

}  // namespace emu

#ifndef RB_EMU_CUDA_HEADERS
// the handful of runtime calls the host code makes: device memory is host memory here, every stream is synchronous
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = (T*)std::malloc(bytes);
    return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t bytes) {
    *p = (T*)std::malloc(bytes);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMallocAsync(int** p, size_t bytes, cudaStream_t) {
    *p = (int*)std::malloc(bytes);
    return cudaSuccess;
}
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) {
    std::memset(p, v, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
#define RB_CUDA(x) (void)(x)
namespace rb {
inline void count_launch(int = 1) {}
}  // namespace rb
#endif  // RB_EMU_CUDA_HEADERS
