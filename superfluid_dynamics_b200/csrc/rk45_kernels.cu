// rk45_kernels.cu -- element-wise kernels of the adaptive Runge-Kutta-Fehlberg 4(5) stepper.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   axpy / lincomb (2..5 slopes)         L/LinearAlgebra.cuh:10-112        (launched L/RK45.cuh:345-367)
//   rk45_error_and_y5                    L/RK45_Kernels.cuh:63-106         (launched L/RK45.cuh:385-399)
// The reference reduces the scaled error with one atomicAdd per block (order of additions not fixed); here the block sums are
// written out and added in index order by the last block, so repeated runs and replicas on different GPUs agree bit for bit.
#include "internal.cuh"

namespace rb {

// y_out = c1 k1 + c2 k2 + ... + y   (same left-to-right order as the reference's lincomb kernels)
__global__ void rk45_stage_kernel(const double2* __restrict__ y, const double2* __restrict__ k1, const double2* __restrict__ k2,
                                  const double2* __restrict__ k3, const double2* __restrict__ k4, const double2* __restrict__ k5,
                                  double2* __restrict__ out, double c1, double c2, double c3, double c4, double c5, int nk, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 a = k1[i];
    double sx = c1 * a.x, sy = c1 * a.y;
    if (nk >= 2) { a = k2[i]; sx = fma(c2, a.x, sx); sy = fma(c2, a.y, sy); }
    if (nk >= 3) { a = k3[i]; sx = fma(c3, a.x, sx); sy = fma(c3, a.y, sy); }
    if (nk >= 4) { a = k4[i]; sx = fma(c4, a.x, sx); sy = fma(c4, a.y, sy); }
    if (nk >= 5) { a = k5[i]; sx = fma(c5, a.x, sx); sy = fma(c5, a.y, sy); }
    const double2 yy = y[i];
    out[i] = make_double2(sx + yy.x, sy + yy.y);
}

void launch_rk45_stage(const double2* y, const double2* const k[5], double2* out, const double c[5], int nk, size_t n,
                       cudaStream_t st) {
    rk45_stage_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, k[0], k[1], k[2], k[3], k[4], out, c[0], c[1], c[2], c[3], c[4],
                                                                    nk, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// Fehlberg coefficients (L/RK45_Kernels.cuh:15-57): fifth-order weights b, fourth-order weights b*, d = b - b*
namespace fehlberg {
constexpr double b1 = 16.0 / 135.0, b3 = 6656.0 / 12825.0, b4 = 28561.0 / 56430.0, b5 = -9.0 / 50.0, b6 = 2.0 / 55.0;
constexpr double s1 = 25.0 / 216.0, s3 = 1408.0 / 2565.0, s4 = 2197.0 / 4104.0, s5 = -1.0 / 5.0;
constexpr double d1 = b1 - s1, d3 = b3 - s3, d4 = b4 - s4, d5 = b5 - s5, d6 = b6;
}  // namespace fehlberg

// y5 = y + h sum b_j k_j ;  e = h sum d_j k_j ;  sum over i of (|e_i| / (atol + rtol max(|y_i|, |y5_i|)))^2
__global__ void __launch_bounds__(256) rk45_error_y5_kernel(const double2* __restrict__ y, const double2* __restrict__ k1,
                                                             const double2* __restrict__ k3, const double2* __restrict__ k4,
                                                             const double2* __restrict__ k5, const double2* __restrict__ k6,
                                                             double2* __restrict__ y5_out, double h, double atol, double rtol,
                                                             double* __restrict__ partial, unsigned int* ticket,
                                                             double* __restrict__ sumsq, size_t n) {
    __shared__ double sred[256];
    __shared__ unsigned int s_last;
    using namespace fehlberg;
    double local = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)blockDim.x * gridDim.x) {
        const double2 yy = y[i], a1 = k1[i], a3 = k3[i], a4 = k4[i], a5 = k5[i], a6 = k6[i];
        // b2 = d2 = 0: the k2 terms of the reference's expression add exact zeros
        const double yx = yy.x + (h * b1) * a1.x + (h * b3) * a3.x + (h * b4) * a4.x + (h * b5) * a5.x + (h * b6) * a6.x;
        const double yi = yy.y + (h * b1) * a1.y + (h * b3) * a3.y + (h * b4) * a4.y + (h * b5) * a5.y + (h * b6) * a6.y;
        y5_out[i] = make_double2(yx, yi);
        const double ex = (h * d1) * a1.x + (h * d3) * a3.x + (h * d4) * a4.x + (h * d5) * a5.x + (h * d6) * a6.x;
        const double ei = (h * d1) * a1.y + (h * d3) * a3.y + (h * d4) * a4.y + (h * d5) * a5.y + (h * d6) * a6.y;
        double sc = atol + rtol * fmax(hypot(yy.x, yy.y), hypot(yx, yi));
        sc = fmax(sc, 1e-300);
        const double z = hypot(ex, ei) / sc;
        local += z * z;
    }
    sred[threadIdx.x] = local;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sred[threadIdx.x] += sred[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sred[0];
        __threadfence();
        s_last = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    if (s_last != gridDim.x - 1) return;
    __threadfence();
    double s = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(partial + b);
    sred[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sred[threadIdx.x] += sred[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *sumsq = sred[0];
        *ticket = 0u;
    }
}

int rk45_error_blocks(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 1024); }

void launch_rk45_error_y5(const double2* y, const double2* k1, const double2* k3, const double2* k4, const double2* k5,
                          const double2* k6, double2* y5_out, double h, double atol, double rtol, double* partial,
                          unsigned int* ticket, double* sumsq, size_t n, cudaStream_t st) {
    rk45_error_y5_kernel<<<rk45_error_blocks(n), 256, 0, st>>>(y, k1, k3, k4, k5, k6, y5_out, h, atol, rtol, partial, ticket, sumsq, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
