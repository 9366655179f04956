// implicit_kernels.cuh -- the element-wise kernels of the implicit side (implicit.cu): real <-> complex state, the 3N perturbed copies
// of a surface and the central differences that turn two batched RHS evaluations into the Jacobian, and the stage / residual / trial /
// Newton-matrix / update kernels of the Gauss-Legendre integrator.  Reference kernels replaced: see implicit.cu.
// Also compiles under g++ with tests/cpp/cuda_emu.h standing in for the CUDA headers (RB_EMULATE): the CPU test tier runs these
// kernels thread for thread (tests/test_kernel_emulation.py).
#pragma once
#ifdef RB_EMULATE
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace {

// ---- real <-> complex state ---------------------------------------------------------------------------------------------------
// [x | y | phi] -> [x + i y | phi + 0 i]
__global__ void real_to_complex_state_kernel(const double* __restrict__ y, double2* __restrict__ s, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    s[i] = make_double2(y[i], y[N + i]);
    s[N + i] = make_double2(y[2 * N + i], 0.0);
}

// [w | dPhi/dt] -> [Re w | Im w | Re dPhi/dt]
__global__ void complex_to_real_rhs_kernel(const double2* __restrict__ r, double* __restrict__ out, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double2 w = r[i];
    out[i] = w.x;
    out[N + i] = w.y;
    out[2 * N + i] = r[N + i].x;
}

// ---- finite-difference Jacobian -----------------------------------------------------------------------------------------------
// grid (ceil(N / 256), 3N, 1 or 2): member b = blockIdx.y = c N + j has coordinate c of point j moved by +eps (z == 0, into `pos`)
// or -eps (z == 1, into `neg`).  Batched layout: Z of member b at [b N, (b + 1) N), Phi of member b at 3 N^2 + [b N, (b + 1) N).
__global__ void perturbed_states_kernel(const double2* __restrict__ state, double2* __restrict__ pos, double2* __restrict__ neg,
                                        double eps, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.y;
    const int c = b / N, j = b - c * N;
    const double e = blockIdx.z ? -eps : eps;
    double2* __restrict__ out = blockIdx.z ? neg : pos;
    double2 z = state[i];
    double2 p = state[N + i];
    if (i == j) {
        if (c == 0) z.x += e;
        else if (c == 1) z.y += e;
        else p.x += e;
    }
    const size_t BN = (size_t)3 * N * N;
    out[(size_t)b * N + i] = z;
    out[BN + (size_t)b * N + i] = p;
}

// rhs of the batch: [w of member 0 .. w of member 3N-1 | dPhi/dt of member 0 ..]; C is 3N x 3N column-major, column = member
__global__ void jacobian_from_perturbed_kernel(const double2* __restrict__ pos, const double2* __restrict__ neg,
                                               double* __restrict__ C, int N, double eps) {
    const size_t total = (size_t)6 * N * N;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double2 a = pos[i], m = neg[i];
    const double d = 2.0 * eps;
    const double re = (a.x - m.x) / d, im = (a.y - m.y) / d;
    const size_t k = i / N, p = i - k * N;
    const size_t n3 = (size_t)3 * N;
    if (k < n3) {
        C[k * n3 + p] = re;
        C[k * n3 + p + N] = im;
    } else {
        C[(k - n3) * n3 + p + 2 * N] = re;
    }
}

// ---- Gauss-Legendre-2 (L/GLCoefficients.hpp) ----------------------------------------------------------------------------------
constexpr double kSqrt3 = 1.7320508075688772935;
constexpr double kA11 = 0.25, kA12 = 0.25 - kSqrt3 / 6.0, kA21 = 0.25 + kSqrt3 / 6.0, kA22 = 0.25;
constexpr double kB1 = 0.5, kB2 = 0.5;

// stage states y_i = y + h sum_j a_ij k_j, written once as the real vectors y1 | y2 (what the Jacobians are taken at) and once as the batch-2 complex
// state [Z_1 | Z_2 | Phi_1 | Phi_2] of the RHS assembler: both stages are then ONE batched RHS evaluation
__global__ void gl2_stage_states_batched_kernel(const double* __restrict__ y, double h, const double* __restrict__ k1,
                                                const double* __restrict__ k2, double* __restrict__ y1, double* __restrict__ y2,
                                                double2* __restrict__ cstate, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double s1[3], s2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t q = (size_t)c * N + i;
        const double a = k1[q], b = k2[q], v = y[q];
        s1[c] = v + h * (kA11 * a + kA12 * b);
        s2[c] = v + h * (kA21 * a + kA22 * b);
        y1[q] = s1[c];
        y2[q] = s2[c];
    }
    cstate[i] = make_double2(s1[0], s1[1]);
    cstate[N + i] = make_double2(s2[0], s2[1]);
    cstate[2 * N + i] = make_double2(s1[2], 0.0);
    cstate[3 * N + i] = make_double2(s2[2], 0.0);
}

// batched RHS [w_1 | w_2 | dPhi_1/dt | dPhi_2/dt] -> f(y1) | f(y2) as real vectors [Re w | Im w | Re dPhi/dt] each
__global__ void gl2_batched_rhs_to_real_kernel(const double2* __restrict__ crhs, double* __restrict__ fy, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t n = (size_t)3 * N;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const double2 w = crhs[m * N + i];
        fy[m * n + i] = w.x;
        fy[m * n + N + i] = w.y;
        fy[m * n + 2 * N + i] = crhs[2 * N + m * N + i].x;
    }
}

// R = k - f(y_stage) over both stages (2n entries) with sum R^2 and sum k^2 in the same pass: ONE CTA, fixed summation order
// (the Newton / Armijo decisions taken from these sums are then reproducible run to run)
constexpr int kNormThreads = 1024;
__global__ void __launch_bounds__(kNormThreads) gl2_residual_kernel(const double* __restrict__ fy, const double* __restrict__ k,
                                                                    double* __restrict__ R, size_t n2, double* __restrict__ sums) {
    __shared__ double sr[kNormThreads], sk[kNormThreads];
    double ar = 0.0, ak = 0.0;
    for (size_t i = threadIdx.x; i < n2; i += kNormThreads) {
        const double kv = k[i];
        const double r = kv - fy[i];
        R[i] = r;
        ar += r * r;
        ak += kv * kv;
    }
    sr[threadIdx.x] = ar;
    sk[threadIdx.x] = ak;
    __syncthreads();
    for (int w = kNormThreads / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            sr[threadIdx.x] += sr[threadIdx.x + w];
            sk[threadIdx.x] += sk[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[0] = sr[0];
        sums[1] = sk[0];
    }
}

// k_trial = k + alpha dK
__global__ void gl2_trial_kernel(const double* __restrict__ k, double alpha, const double* __restrict__ dK, double* __restrict__ kt,
                                 size_t n2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) kt[i] = k[i] + alpha * dK[i];
}

// dK <- -R (right-hand side of the Newton system; the LU solves in place)
__global__ void gl2_negate_kernel(const double* __restrict__ R, double* __restrict__ out, size_t n2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) out[i] = -R[i];
}

// Jacobian of the stage residual, 2n x 2n column-major: block (i, j) = delta_ij I - h a_ij J_i  (J_i n x n column-major)
__global__ void gl2_newton_matrix_kernel(const double* __restrict__ J1, const double* __restrict__ J2, double h,
                                         double* __restrict__ M, size_t n) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // row: consecutive threads -> consecutive addresses
    const size_t c = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (r >= n || c >= n) return;
    const double j1 = J1[r + c * n], j2 = J2[r + c * n];
    const double d = r == c ? 1.0 : 0.0;
    const size_t ld = 2 * n;
    M[r + c * ld] = d - h * kA11 * j1;
    M[r + (c + n) * ld] = -h * kA12 * j1;
    M[r + n + c * ld] = -h * kA21 * j2;
    M[r + n + (c + n) * ld] = d - h * kA22 * j2;
}

// y_next = y + h (b1 k1 + b2 k2)
__global__ void gl2_next_state_kernel(const double* __restrict__ y, double h, const double* __restrict__ k1,
                                      const double* __restrict__ k2, double* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = y[i] + h * (kB1 * k1[i] + kB2 * k2[i]);
}

}  // namespace
