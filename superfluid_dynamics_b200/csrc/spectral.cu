// spectral.cu -- element-wise parts of the FFT derivatives (the transforms themselves are issued by solver.cu).
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   batched_vector_subtract_singletime_complex_real   L/utilities.cuh:231-239   (launched L/Derivatives.cuh:317-318)
//   first_derivative_multiplication                   L/utilities.cuh:106-148
//   second_derivative_fft                             L/utilities.cuh:178-200
//   vector_mutiply_scalar / vector_scalar_add_complex_real   L/utilities.cuh:249-263 (L/Derivatives.cuh:254, 374, 382)
// The Nyquist conventions of the reference are reproduced exactly:
//   D1: modes 0..n/2-1 -> i k; mode n/2 -> i * pi * (n/2); mode n/2+1 -> 0; modes above -> i (k - n)
//   D2: -(k^2) for every mode, nothing zeroed.  The 1/n of the unnormalised inverse is folded into the multiply.
#include "internal.cuh"

namespace rb {

__device__ __forceinline__ double2 d1_coeff(double2 c, int i, int n) {
    const double dn = static_cast<double>(n);
    if (i < n / 2) {
        return make_double2(-i * c.y / dn, i * c.x / dn);
    } else if (i == n / 2) {
        return make_double2(-kPi * i * c.y / dn, kPi * i * c.x / dn);
    } else if (i == n / 2 + 1) {
        return make_double2(0.0, 0.0);
    }
    return make_double2(-(i - n) * c.y / dn, (i - n) * c.x / dn);
}

__device__ __forceinline__ double2 d2_coeff(double2 c, int i, int n) {
    if (i <= n / 2) {
        return make_double2(-i * i * c.x / (n), -i * i * c.y / (n));
    }
    return make_double2(-(i - n) * (i - n) * c.x / (n), -(i - n) * (i - n) * c.y / (n));
}

// Z - 2 pi j / N  and  Phi - (-(1+rho) pi U / N) j      (ZPhiDerivative ctor tables, L/Derivatives.cuh:282-287)
__global__ void sub_linear_kernel(const double2* __restrict__ Z, const double2* __restrict__ Phi, double2* __restrict__ zper,
                                  double2* __restrict__ phiper, int N, size_t total, double rho, double U) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double zl = 2 * kPi * (double)i / N;
    double2 z = Z[tid];
    zper[tid] = make_double2(z.x - zl, z.y);
    if (Phi) {
        double pl = -(1 + rho) * kPi * U / N * (double)i;
        double2 p = Phi[tid];
        phiper[tid] = make_double2(p.x - pl, p.y);
    }
}

void launch_sub_linear(const double2* Z, const double2* Phi, double2* out_zper, double2* out_phiper, int N, int batch,
                       double rho, double U, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    sub_linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, Phi, out_zper, out_phiper, N, total, rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// one pass over the spectra of Z_per and Phi_per: i k Z^, -k^2 Z^, i k Phi^
__global__ void spectral_multiply_zphi_kernel(const double2* __restrict__ hatZ, const double2* __restrict__ hatPhi,
                                              double2* __restrict__ d1z, double2* __restrict__ d2z, double2* __restrict__ d1phi,
                                              int N, size_t total) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double2 cz = hatZ[tid];
    d1z[tid] = d1_coeff(cz, i, N);
    d2z[tid] = d2_coeff(cz, i, N);
    if (hatPhi) d1phi[tid] = d1_coeff(hatPhi[tid], i, N);
}

void launch_spectral_multiply_zphi(const double2* hatZ, const double2* hatPhi, double2* out_d1z, double2* out_d2z,
                                   double2* out_d1phi, int N, int batch, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_zphi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(hatZ, hatPhi, out_d1z, out_d2z, out_d1phi, N,
                                                                                   total);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void spectral_multiply_kernel(const double2* __restrict__ hat, double2* __restrict__ out, int N, size_t total,
                                         int second) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double2 c = hat[tid];
    out[tid] = second ? d2_coeff(c, i, N) : d1_coeff(c, i, N);
}

void launch_spectral_multiply(const double2* hat, double2* out, int N, int batch, int second, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(hat, out, N, total, second);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// first derivative of a REAL sequence from its half spectrum (D2Z transform): the full coefficient array of the reference's
// Z2Z path is rebuilt with c[N-i] = conj(c[i]); `scale` (2 pi / N) is folded into the coefficients
__global__ void spectral_multiply_real_kernel(const double2* __restrict__ half, double2* __restrict__ out, int N, size_t total,
                                              double scale) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int nh = N / 2 + 1;
    size_t b = tid / N;
    int i = (int)(tid - b * N);
    double2 c;
    if (i < nh) {
        c = half[b * nh + i];
    } else {
        c = half[b * nh + (N - i)];
        c.y = -c.y;
    }
    double2 r = d1_coeff(c, i, N);
    out[tid] = make_double2(r.x * scale, r.y * scale);
}

void launch_spectral_multiply_real(const double2* half, double2* out, int N, int batch, double scale, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_real_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(half, out, N, total, scale);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// scaling by 2 pi / N (resp. its square) and the linear parts put back (L/Derivatives.cuh:321-324, 374, 380-383)
__global__ void finish_zphi_kernel(double2* __restrict__ Zp, double2* __restrict__ Zpp, double2* __restrict__ PhiP, int N,
                                   size_t total, double rho, double U) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const double s1 = 2.0 * kPi / N;
    const double s2 = 4.0 * kPi * kPi / ((double)N * N);
    double2 a = Zp[tid];
    Zp[tid] = make_double2(a.x * s1 + 2.0 * kPi / N, a.y * s1);
    double2 c = Zpp[tid];
    Zpp[tid] = make_double2(c.x * s2, c.y * s2);
    if (PhiP) {
        double2 p = PhiP[tid];
        p.x *= s1;
        p.y *= s1;
        if (U != 0) p.x += -(1 + rho) * kPi * U / N;
        PhiP[tid] = p;
    }
}

void launch_finish_zphi(double2* Zp, double2* Zpp, double2* PhiP, int N, int batch, double rho, double U, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    finish_zphi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Zp, Zpp, PhiP, N, total, rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void scale_kernel(double2* __restrict__ v, double s, size_t n) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    double2 a = v[tid];
    v[tid] = make_double2(a.x * s, a.y * s);
}

void launch_scale(double2* v, double s, size_t n, cudaStream_t st) {
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, s, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
