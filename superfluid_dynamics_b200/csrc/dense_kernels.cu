// dense_kernels.cu -- materialised operators (the entry points the reference's tests launch by name), the dPhi/dt
// expressions, the energy sums and a small in-place LU used as the validation solve.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   createMKernel, createFiniteDepthMKernel                       L/createM.cuh:43-92
//   createVelocityMatrices, createHeliumVelocityMatrices          L/WaterVelocities.cuh:38-107
//   compute_rhs_phi_expression, compute_rhs_helium_phi_expression{,_with_surface_tension,_expansion_terms}
//                                                                 L/createM.cuh:96-117, 171-211
//   KineticEnergy / GravitationalEnergy / VanDerWaalsEnergy / SurfaceEnergy / VolumeFlux (5 CUB reductions with a
//   cudaMalloc + cudaFreeAsync each)                              L/Energies.cuh:136-327
//   MatrixSolver<N,1>::solve (cusolverDnDgetrf + Dgetrs)          L/MatrixSolver.cuh:114-125
// Matrices are column-major A[k + j*n + b*n*n] = entry (row k, col j), exactly as the reference stores them.
#include <cstdlib>

#include "internal.cuh"
#include "../../include/roberts_b200_device.cuh"

namespace rb {

using rb_dev::cdivd;
using rb_dev::cot_half;

// ---- M -----------------------------------------------------------------------------------------
template <bool FINITE>
__global__ void create_M_kernel(double* __restrict__ A, const double2* __restrict__ Z, const double2* __restrict__ Zp,
                                const double2* __restrict__ Zpp, double rho, double h, int n, bool infinite_depth) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;   // row
    int j = blockIdx.y * blockDim.y + threadIdx.y;   // column
    size_t b = blockIdx.z;
    if (k >= n || j >= n) return;
    A[(size_t)k + (size_t)j * n + b * (size_t)n * n] =
        rb_dev::M_entry<FINITE>(k, j, Z + b * n, Zp + b * n, Zpp + b * n, rho, h, infinite_depth);
}

void launch_create_M(double* A, const double2* Z, const double2* Zp, const double2* Zpp, double rho, int n, size_t batch,
                     cudaStream_t st) {
    dim3 th(16, 16), bl((n + 15) / 16, (n + 15) / 16, (unsigned)batch);
    create_M_kernel<false><<<bl, th, 0, st>>>(A, Z, Zp, Zpp, rho, 0.0, n, true);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_create_finite_depth_M(double* A, const double2* Z, const double2* Zp, const double2* Zpp, double h, int n,
                                  size_t batch, bool infinite_depth, cudaStream_t st) {
    dim3 th(16, 16), bl((n + 15) / 16, (n + 15) / 16, (unsigned)batch);
    create_M_kernel<true><<<bl, th, 0, st>>>(A, Z, Zp, Zpp, 0.0, h, n, infinite_depth);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ---- V1, V2 ------------------------------------------------------------------------------------
__global__ void velocity_matrices_kernel(const double2* __restrict__ Z, const double2* __restrict__ Zp,
                                         const double2* __restrict__ Zpp, int n, double2* __restrict__ V1,
                                         double2* __restrict__ V2, bool lower, bool helium, double h, bool infinite_depth) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    size_t b = blockIdx.z;
    if (k >= n || j >= n) return;
    V1[(size_t)k + (size_t)j * n + b * (size_t)n * n] =
        rb_dev::V1_entry(k, j, Z + b * n, Zp + b * n, Zpp + b * n, V2 + b * n, lower, helium, h, infinite_depth);
}

void launch_velocity_matrices(const double2* Z, const double2* Zp, const double2* Zpp, int n, double2* V1, double2* V2,
                              bool lower, size_t batch, bool helium, double h, bool infinite_depth, cudaStream_t st) {
    dim3 th(16, 16), bl((n + 15) / 16, (n + 15) / 16, (unsigned)batch);
    velocity_matrices_kernel<<<bl, th, 0, st>>>(Z, Zp, Zpp, n, V1, V2, lower, helium, h, infinite_depth);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ---- dPhi/dt -----------------------------------------------------------------------------------
__global__ void rhs_phi_water_kernel(const double2* __restrict__ Z, const double2* __restrict__ V1,
                                     const double2* __restrict__ V2, double2* __restrict__ result, double rho, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double zi = Z[i].y;
    double2 v1 = V1[i], v2 = V2[i];
    double a1 = v1.x * v1.x + v1.y * v1.y;
    double a2 = v2.x * v2.x + v2.y * v2.y;
    double dot = V1[1].x * v2.x + v1.y * v2.y;   // the reference reads V1[1] here (L/createM.cuh:104); kept for parity
    result[i] = make_double2(-(1 + rho) * zi + 0.5 * a1 + 0.5 * rho * a2 - rho * dot, 0.0);
}

void launch_rhs_phi_water(const double2* Z, const double2* V1, const double2* V2, double2* result, double rho, int n,
                          cudaStream_t st) {
    rhs_phi_water_kernel<<<(n + 255) / 256, 256, 0, st>>>(Z, V1, V2, result, rho, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void rhs_phi_helium_kernel(const double2* __restrict__ Z, const double2* __restrict__ V1,
                                      double2* __restrict__ result, double h, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double vdw = h / 3.0;
    double2 v = V1[i];
    result[i] = make_double2(vdw * pow(1.0 + Z[i].y / h, -3.0) - vdw + 0.5 * v.x * v.x + 0.5 * v.y * v.y, 0.0);
}

void launch_rhs_phi_helium(const double2* Z, const double2* V1, double2* result, double h, int n, cudaStream_t st) {
    rhs_phi_helium_kernel<<<(n + 255) / 256, 256, 0, st>>>(Z, V1, result, h, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void rhs_phi_helium_st_kernel(const double2* __restrict__ Z, const double2* __restrict__ Zp,
                                         const double2* __restrict__ Zpp, const double2* __restrict__ V1,
                                         double2* __restrict__ result, double h, double kappa, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double zi = Z[i].y;
    double2 v = V1[i], zp = Zp[i], zpp = Zpp[i];
    double a1 = v.x * v.x + v.y * v.y;
    double curv = (zp.x * zpp.y - zp.y * zpp.x) / pow(zp.x * zp.x + zp.y * zp.y, 1.5);
    result[i] = make_double2(20.447761896665416 * h / 3.0 * (1.0 / pow(1.0 + zi / h, 3.0) - 1) + 0.5 * a1 + kappa * curv, 0.0);
}

void launch_rhs_phi_helium_st(const double2* Z, const double2* Zp, const double2* Zpp, const double2* V1, double2* result,
                              double h, double kappa, int n, cudaStream_t st) {
    rhs_phi_helium_st_kernel<<<(n + 255) / 256, 256, 0, st>>>(Z, Zp, Zpp, V1, result, h, kappa, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void rhs_phi_helium_exp_kernel(const double2* __restrict__ Z, const double2* __restrict__ V1,
                                          double2* __restrict__ result, double h, int N, int order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double2 v = V1[i];
    double y = Z[i].y;
    double kin = 0.5 * v.x * v.x + 0.5 * v.y * v.y;
    double vdw = 0.0;
    switch (order) {   // fall-through on purpose, as in the reference
        case 3: vdw += -10.0 / 3.0 * pow(y, 3.0) / (h * h);
        case 2: vdw += 2.0 * pow(y, 2.0) / h;
        case 1: vdw += -y;
        default: break;
    }
    result[i] = make_double2(vdw + kin, 0.0);
}

void launch_rhs_phi_helium_exp(const double2* Z, const double2* V1, double2* result, double h, int n, int order,
                               cudaStream_t st) {
    rhs_phi_helium_exp_kernel<<<(n + 255) / 256, 256, 0, st>>>(Z, V1, result, h, n, order);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ---- energies: the five sums of Energies.cuh in one launch, batch member 0 only (as the reference) -------------
constexpr int kEnergyThreads = 512;

__global__ void __launch_bounds__(kEnergyThreads) energies_kernel(const double2* __restrict__ Z, const double2* __restrict__ Zp,
                                                                   const double2* __restrict__ Phi,
                                                                   const double2* __restrict__ vel, double* __restrict__ out5,
                                                                   int N, int physics, double rho, double U, double depth,
                                                                   double kappa) {
    __shared__ double sred[5][kEnergyThreads];
    double s[5] = {0, 0, 0, 0, 0};
    for (int k = threadIdx.x; k < N; k += kEnergyThreads) {
        double2 z = Z[k], zp = Zp[k], ph = Phi[k], v = vel[k];
        // KineticEnergyCombination (velocitiesUpper == velocitiesLower in the reference's call, L/Energies.cuh:243)
        s[0] += (ph.x + 0.5 * U * (1.0 + rho) * z.x) * (-1.0 * zp.y * v.x + zp.x * v.y) -
                0.5 * U * ((v.x + rho * v.x) * zp.x + (v.y + rho * v.y) * zp.y + 0.5 * U * (1.0 - rho) * zp.x) * z.y;
        if (physics == 0) s[1] += z.y * z.y * zp.x;                                 // GravitationalEnergyCombination
        else s[1] += (1 / pow(1 + z.y / depth, 2.0) - 1.0);                          // VanDerWaalsEnergyCombination
        s[2] += sqrt(zp.x * zp.x + zp.y * zp.y);                                     // SurfaceEnergyCombination
        s[3] += (v.y * zp.x + v.x * zp.y);                                           // VolumeFluxCombination
        s[4] += z.y * zp.x;                                                          // volume
    }
    for (int q = 0; q < 5; ++q) sred[q][threadIdx.x] = s[q];
    __syncthreads();
    for (int w = kEnergyThreads / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int q = 0; q < 5; ++q) sred[q][threadIdx.x] += sred[q][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out5[0] = sred[0][0] * 0.25 / kPi;                                                         // :61-64
        out5[1] = physics == 0 ? sred[1][0] * 0.25 * (1.0 + rho) / kPi : sred[1][0] * depth * depth / 6.0;   // :77-80, :93-96
        out5[2] = (sred[2][0] - 2.0 * kPi) * kappa / (2.0 * kPi);                                  // :109-112
        out5[3] = sred[3][0] * 0.5 / kPi;                                                          // :125-128
        out5[4] = sred[4][0];
    }
}

void launch_energies(const double2* Z, const double2* Zp, const double2* Phi, const double2* vel, double* out5, int N,
                     int physics, double rho, double U, double depth, double kappa, cudaStream_t st) {
    energies_kernel<<<1, kEnergyThreads, 0, st>>>(Z, Zp, Phi, vel, out5, N, physics, rho, U, depth, kappa);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ---- validation solve: unblocked right-looking LU with partial pivoting, column-major, one right-hand side ------
constexpr int kLuThreads = 1024;

__global__ void __launch_bounds__(kLuThreads) lu_pivot_kernel(double* __restrict__ A, double* __restrict__ b, int n, int k,
                                                               int* __restrict__ info) {
    __shared__ double sval[kLuThreads];
    __shared__ int sidx[kLuThreads];
    double best = -1.0;
    int bi = k;
    for (int i = k + threadIdx.x; i < n; i += kLuThreads) {
        double v = fabs(A[(size_t)i + (size_t)k * n]);
        if (v > best) { best = v; bi = i; }
    }
    sval[threadIdx.x] = best;
    sidx[threadIdx.x] = bi;
    __syncthreads();
    for (int w = kLuThreads / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            double o = sval[threadIdx.x + w];
            int oi = sidx[threadIdx.x + w];
            if (o > sval[threadIdx.x] || (o == sval[threadIdx.x] && oi < sidx[threadIdx.x])) {
                sval[threadIdx.x] = o;
                sidx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    const int p = sidx[0];
    if (!(sval[0] > 0.0)) {   // zero column, or nothing but NaN in it (best stays at its -1 start): same verdict as the blocked panel
        if (threadIdx.x == 0 && *info == 0) *info = k + 1;
        return;
    }
    if (p != k) {
        for (int j = threadIdx.x; j < n; j += kLuThreads) {
            double t = A[(size_t)k + (size_t)j * n];
            A[(size_t)k + (size_t)j * n] = A[(size_t)p + (size_t)j * n];
            A[(size_t)p + (size_t)j * n] = t;
        }
        if (threadIdx.x == 0) { double t = b[k]; b[k] = b[p]; b[p] = t; }
    }
    __syncthreads();
    const double piv = A[(size_t)k + (size_t)k * n];
    for (int i = k + 1 + threadIdx.x; i < n; i += kLuThreads) A[(size_t)i + (size_t)k * n] /= piv;
}

__global__ void lu_update_kernel(double* __restrict__ A, double* __restrict__ b, int n, int k) {
    int i = k + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = k + 1 + blockIdx.y * blockDim.y + threadIdx.y;   // j == n addresses the right-hand side
    if (i >= n || j > n) return;
    double l = A[(size_t)i + (size_t)k * n];
    if (j < n) A[(size_t)i + (size_t)j * n] -= l * A[(size_t)k + (size_t)j * n];
    else b[i] -= l * b[k];
}

__global__ void __launch_bounds__(kLuThreads) lu_backsolve_kernel(const double* __restrict__ A, double* __restrict__ b, int n) {
    for (int k = n - 1; k >= 0; --k) {
        __syncthreads();
        if (threadIdx.x == 0) b[k] /= A[(size_t)k + (size_t)k * n];
        __syncthreads();
        double xk = b[k];
        for (int i = threadIdx.x; i < k; i += kLuThreads) b[i] -= A[(size_t)i + (size_t)k * n] * xk;
    }
}

void launch_lu_backsolve(const double* A, double* b, int n, cudaStream_t st) {
    lu_backsolve_kernel<<<1, kLuThreads, 0, st>>>(A, b, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_lu_solve_unblocked(double* A, double* b, int n, int* info, cudaStream_t st) {
    RB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    for (int k = 0; k < n; ++k) {
        lu_pivot_kernel<<<1, kLuThreads, 0, st>>>(A, b, n, k, info);
        int rem = n - k - 1;
        if (rem > 0) {
            dim3 th(32, 8), bl((rem + 31) / 32, (rem + 1 + 7) / 8);
            lu_update_kernel<<<bl, th, 0, st>>>(A, b, n, k);
        }
    }
    lu_backsolve_kernel<<<1, kLuThreads, 0, st>>>(A, b, n);
    RB_CUDA(cudaGetLastError());
    count_launch(2 * n);
}

// the blocked factorisation (lu_kernels.cu: panels of 32 columns, FP64 tensor-core trailing update) is the default from n = 64 up:
// measured faster than the two-launches-per-column factorisation above at every size tried (n = 384 ... 4096, see lu_kernels.cu);
// RB_LU_BLOCKED=0 selects the unblocked one
void launch_lu_solve(double* A, double* b, int n, int* info, cudaStream_t st) {
    static const bool blocked = [] {
        const char* v = std::getenv("RB_LU_BLOCKED");
        return !v || std::atoi(v) != 0;
    }();
    if (blocked && n >= 64) launch_lu_solve_blocked(A, b, n, info, st);
    else launch_lu_solve_unblocked(A, b, n, info, st);
}

}  // namespace rb
