// lu_kernels.cu -- blocked right-looking LU with partial pivoting for ONE right-hand side, column-major, in place: the dense solve
// of the path wherever a matrix is really materialised (the 6N x 6N Newton matrix of the Gauss-Legendre integrator, the N x N
// system M a = Re(Phi') in the dense validation mode).  The trailing update A22 -= L21 U12 (K = 32) is the one true dense
// contraction on this path and runs on the FP64 tensor path (mma.sync m8n8k4.f64, "DMMA"); everything else is latency work.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/): MatrixSolver<N,1>::solve = cusolverDnDgetrf + cusolverDnDgetrs,
// L/MatrixSolver.cuh:114-125.
//
// Per panel of 32 columns, four launches:
//   lu_panel_kernel      one CTA: pivot search (warp shuffles), row swap inside the panel, scale, rank-1 update of the panel
//   lu_swap_trsm_kernel  one thread per column right of the panel and for b: the panel's row interchanges (the left part, L, is
//                        never read again because b is eliminated on the fly, exactly as the unblocked kernels of dense_kernels.cu
//                        treat it), then U12 = L11^{-1} A12 with L11 in shared memory
//   lu_gemm_kernel       A22 -= L21 U12: 64 x 64 tiles, 8 warps, 2 x 4 DMMA tiles of 8 x 8 per warp, 8 k-steps of 4
//   lu_gemv_kernel       b2  -= L21 b1
// then lu_backsolve_blocked_kernel: the back substitution with U, one CTA, 32 columns per pair of barriers.
// STATUS: written after round 1's GPU minutes were spent; selected only when asked for (RB_LU_BLOCKED=1 or
// rb_lu_solve(..., blocked = 1)) until it has passed tests/test_zz_gpu_implicit.py on hardware.
// The kernels and the launch sequence below also compile under g++ with tests/cpp/cuda_emu.h standing in for the CUDA headers
// (RB_EMULATE): the CPU test tier runs them thread for thread against LAPACK (tests/test_kernel_emulation.py).
#ifdef RB_EMULATE
#include "cuda_emu.h"
#else
#include "internal.cuh"
#endif
#include "launch.cuh"

namespace rb {

namespace {

constexpr int kNB = 32;              // panel width == K of the trailing update
constexpr int kPanelThreads = 1024;

__global__ void __launch_bounds__(kPanelThreads) lu_panel_kernel(double* A, int n, int k0, int kb, int* __restrict__ piv,
                                                                  int* __restrict__ info) {
    // Three block-wide barriers per column: (1) after the per-warp pivot candidates are in shared memory -- every warp then reduces
    // the 32 candidates itself, so the winner needs no broadcast (the candidate arrays are double-buffered against the next column);
    // (2) after the row interchange, which also leaves the pivot row in shared memory; (3) after the rank-1 update of the panel.
    __shared__ double s_val[2][32];
    __shared__ int s_idx[2][32];
    __shared__ double s_row[kNB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < kb; ++j) {
        const int col = k0 + j, buf = j & 1;
        double* cptr = A + (size_t)col * n;
        // pivot: largest |A[i, col]|, i >= col; ties to the lowest row (the choice of the unblocked kernel)
        double best = -1.0;
        int bi = col;
        for (int i = col + tid; i < n; i += kPanelThreads) {
            const double v = fabs(cptr[i]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[buf][warp] = best; s_idx[buf][warp] = bi; }
        __syncthreads();
        best = s_val[buf][lane];
        bi = s_idx[buf][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        const int p = __shfl_sync(0xffffffffu, bi, 0);
        const bool usable = best > 0.0;
        if (tid == 0) {
            piv[col] = p;
            if (!usable && *info == 0) *info = col + 1;   // zero (or NaN) column: singular, as getrf's info
        }
        if (tid < kb) {   // row interchange inside the panel; the thread that moves column k0 + tid also knows the new pivot-row entry
            double* q = A + (size_t)(k0 + tid) * n;
            const double top = q[col], low = q[p];
            if (p != col) {
                q[col] = low;
                q[p] = top;
            }
            s_row[tid] = low;
        }
        __syncthreads();
        if (usable) {
            const double pv = s_row[j];
            for (int i = col + 1 + tid; i < n; i += kPanelThreads) {
                const double l = cptr[i] / pv;
                cptr[i] = l;
                for (int c = j + 1; c < kb; ++c) A[(size_t)(k0 + c) * n + i] -= l * s_row[c];
            }
        }
        __syncthreads();
    }
}

// column c of the augmented matrix [A | b]: c == n addresses b
__device__ __forceinline__ double* aug_column(double* A, double* b, int n, int c) { return c < n ? A + (size_t)c * n : b; }

// one thread per column right of the panel (b included): the panel's row interchanges, then U12 = L11^{-1} A12 with L11 (unit lower
// triangular) in shared memory and the 32 entries of the column in registers
__global__ void __launch_bounds__(128) lu_swap_trsm_kernel(double* __restrict__ A, double* __restrict__ b, int n, int k0, int kb,
                                                           const int* __restrict__ piv) {
    __shared__ double L[kNB][kNB + 1];
    __shared__ int s_piv[kNB];
    for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) {
        const int i = idx % kb, j = idx / kb;
        L[i][j] = A[(size_t)(k0 + j) * n + k0 + i];
    }
    if ((int)threadIdx.x < kb) s_piv[threadIdx.x] = piv[k0 + threadIdx.x];
    __syncthreads();
    const int c = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n) return;
    double* colp = aug_column(A, b, n, c);
    for (int j = 0; j < kb; ++j) {
        const int r = k0 + j, p = s_piv[j];
        if (p != r) {
            const double t = colp[r];
            colp[r] = colp[p];
            colp[p] = t;
        }
    }
    colp += k0;
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = i < kb ? colp[i] : 0.0;
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
#pragma unroll
        for (int i = j + 1; i < kNB; ++i)
            if (i < kb) x[i] -= L[i][j] * x[j];   // unit lower triangular
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i)
        if (i < kb) colp[i] = x[i];
}

#ifndef RB_EMULATE
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

constexpr int kTile = 64;
constexpr int kGemmThreads = 256;

// C[r0.., c0..] -= L21[r0.., 0..kb) * U12[0..kb, c0..), the trailing block starting at row/column t0 = k0 + kb of the n x n matrix
__global__ void __launch_bounds__(kGemmThreads) lu_gemm_kernel(double* __restrict__ A, int n, int k0, int kb) {
    __shared__ double As[kTile][kNB + 1];   // -L21 tile, As[row][k]
    __shared__ double Bs[kNB][kTile + 1];   //  U12 tile, Bs[k][col]
    const int t0 = k0 + kb;
    const int r0 = t0 + blockIdx.x * kTile, c0 = t0 + blockIdx.y * kTile;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < kTile * kNB; idx += kGemmThreads) {
        const int i = idx % kTile, j = idx / kTile;   // consecutive threads: consecutive rows of one column of L21
        const int r = r0 + i;
        As[i][j] = (r < n && j < kb) ? -A[(size_t)(k0 + j) * n + r] : 0.0;
    }
    for (int idx = tid; idx < kNB * kTile; idx += kGemmThreads) {
        const int j = idx % kNB, c = idx / kNB;       // consecutive threads: consecutive rows of one column of U12
        const int cc = c0 + c;
        Bs[j][c] = (cc < n && j < kb) ? A[(size_t)cc * n + k0 + j] : 0.0;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;            // fragment coordinates of mma.m8n8k4.f64
    const int rb = (warp & 3) * 16, cb = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int r = r0 + rb + rt * 8 + g, c = c0 + cb + ct * 8 + 2 * t;
            acc[rt][ct][0] = (r < n && c < n) ? A[(size_t)c * n + r] : 0.0;
            acc[rt][ct][1] = (r < n && c + 1 < n) ? A[(size_t)(c + 1) * n + r] : 0.0;
        }
#pragma unroll
    for (int kk = 0; kk < kNB / 4; ++kk) {
        double a[2], bf[4];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) a[rt] = As[rb + rt * 8 + g][kk * 4 + t];
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) bf[ct] = Bs[kk * 4 + t][cb + ct * 8 + g];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) dmma_m8n8k4(acc[rt][ct][0], acc[rt][ct][1], a[rt], bf[ct]);
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int r = r0 + rb + rt * 8 + g, c = c0 + cb + ct * 8 + 2 * t;
            if (r < n && c < n) A[(size_t)c * n + r] = acc[rt][ct][0];
            if (r < n && c + 1 < n) A[(size_t)(c + 1) * n + r] = acc[rt][ct][1];
        }
}

__global__ void lu_gemv_kernel(const double* __restrict__ A, double* __restrict__ b, int n, int k0, int kb) {
    __shared__ double bs[kNB];
    if ((int)threadIdx.x < kb) bs[threadIdx.x] = b[k0 + threadIdx.x];
    __syncthreads();
    const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = b[i];
    for (int j = 0; j < kb; ++j) acc -= A[(size_t)(k0 + j) * n + i] * bs[j];
    b[i] = acc;
}

// U x = b in place, 32 columns at a time from the bottom: warp 0 solves the 32 x 32 triangle in shared memory, then every thread
// subtracts that block's contribution from the rows above it (coalesced down the columns).  Two block-wide barriers per 32 columns
// instead of two per column.
constexpr int kBackThreads = 1024;
__global__ void __launch_bounds__(kBackThreads) lu_backsolve_blocked_kernel(const double* __restrict__ A, double* __restrict__ b, int n) {
    __shared__ double Us[kNB][kNB + 1];
    __shared__ double xs[kNB];
    const int tid = threadIdx.x;
    for (int hi = n; hi > 0; hi -= kNB) {
        const int k0 = hi > kNB ? hi - kNB : 0, w = hi - k0;
        for (int idx = tid; idx < w * w; idx += kBackThreads) {
            const int i = idx % w, j = idx / w;
            Us[i][j] = A[(size_t)(k0 + j) * n + k0 + i];
        }
        if (tid < w) xs[tid] = b[k0 + tid];
        __syncthreads();
        if (tid < 32) {
            for (int j = w - 1; j >= 0; --j) {
                if (tid == j) xs[j] /= Us[j][j];
                __syncwarp();
                if (tid < j) xs[tid] -= Us[tid][j] * xs[j];
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += kBackThreads) {
            double acc = b[i];
            for (int j = 0; j < w; ++j) acc -= A[(size_t)(k0 + j) * n + i] * xs[j];
            b[i] = acc;
        }
        if (tid < w) b[k0 + tid] = xs[tid];
        __syncthreads();
    }
}

}  // namespace

// the factorisation with b eliminated on the fly: on return A holds U in its upper triangle and b holds L^{-1} P b
void lu_factor_blocked(double* A, double* b, int n, int* info, cudaStream_t st) {
    int* piv = nullptr;   // pivot rows, stream-ordered allocation: no synchronisation, no state shared between streams
    RB_CUDA(cudaMallocAsync(&piv, (size_t)std::max(n, 1) * sizeof(int), st));
    RB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    int launches = 0;
    for (int k0 = 0; k0 < n; k0 += kNB) {
        const int kb = std::min(kNB, n - k0);
        RB_LAUNCH(lu_panel_kernel, 1, kPanelThreads, st, A, n, k0, kb, piv, info);
        const int right = n - k0 - kb + 1;   // columns right of the panel, b included
        RB_LAUNCH(lu_swap_trsm_kernel, (right + 127) / 128, 128, st, A, b, n, k0, kb, (const int*)piv);
        launches += 2;
        const int m = n - k0 - kb;
        if (m > 0) {
            const int tiles = (m + kTile - 1) / kTile;
            RB_LAUNCH(lu_gemm_kernel, dim3(tiles, tiles), kGemmThreads, st, A, n, k0, kb);
            RB_LAUNCH(lu_gemv_kernel, (m + 255) / 256, 256, st, (const double*)A, b, n, k0, kb);
            launches += 2;
        }
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(piv, st);
    RB_CUDA(e);
    count_launch(launches);
}

void launch_lu_solve_blocked(double* A, double* b, int n, int* info, cudaStream_t st) {
    lu_factor_blocked(A, b, n, info, st);
    RB_LAUNCH(lu_backsolve_blocked_kernel, 1, kBackThreads, st, (const double*)A, b, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
