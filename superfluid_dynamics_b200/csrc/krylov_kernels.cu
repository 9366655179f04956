// krylov_kernels.cu -- small dense-vector kernels for the restarted GMRES used on the finite-depth helium operator
// (the reference solves that system with cuSOLVER LU, L/MatrixSolver.cuh:114-125; its image term, L/createM.cuh:87-88, puts the
// spectrum of M between 1/2 and N/(4 pi), where the plain Neumann iteration of the water operator does not converge).
// All reductions use a fixed thread count and a fixed tree: results are deterministic and identical on every rank.
#include "internal.cuh"

namespace rb {

constexpr int kDotThreads = 1024;

// out[j] = V[j] . w  for j < nvec, and out[nvec] = w . w
__global__ void __launch_bounds__(kDotThreads) multi_dot_kernel(const double* __restrict__ V, size_t ldv, int nvec,
                                                                 const double* __restrict__ w, double* __restrict__ out, int n) {
    __shared__ double sred[kDotThreads];
    const int j = blockIdx.x;
    const double* v = j < nvec ? V + (size_t)j * ldv : w;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kDotThreads) s = fma(v[i], w[i], s);
    sred[threadIdx.x] = s;
    __syncthreads();
    for (int k = kDotThreads / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sred[threadIdx.x] += sred[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = sred[0];
}

void launch_multi_dot(const double* V, size_t ldv, int nvec, const double* w, double* out, int n, cudaStream_t st) {
    multi_dot_kernel<<<nvec + 1, kDotThreads, 0, st>>>(V, ldv, nvec, w, out, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// w -= sum_j h[j] V[j]   (h on the device: no host round trip between the projection and the update)
__global__ void multi_axpy_kernel(double* __restrict__ w, const double* __restrict__ V, size_t ldv, int nvec,
                                  const double* __restrict__ h, double sign, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = w[i];
    for (int j = 0; j < nvec; ++j) acc = fma(sign * h[j], V[(size_t)j * ldv + i], acc);
    w[i] = acc;
}

void launch_multi_axpy(double* w, const double* V, size_t ldv, int nvec, const double* h, double sign, int n, cudaStream_t st) {
    multi_axpy_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, V, ldv, nvec, h, sign, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// out = sum_j y[j] V[j]
__global__ void combine_kernel(double* __restrict__ out, const double* __restrict__ V, size_t ldv, int nvec,
                               const double* __restrict__ y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int j = 0; j < nvec; ++j) acc = fma(y[j], V[(size_t)j * ldv + i], acc);
    out[i] = acc;
}

void launch_combine(double* out, const double* V, size_t ldv, int nvec, const double* y, int n, cudaStream_t st) {
    combine_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, V, ldv, nvec, y, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// vout = w / sqrt(*nrm2)
__global__ void normalize_kernel(double* __restrict__ vout, const double* __restrict__ w, const double* __restrict__ nrm2, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = *nrm2;
    vout[i] = s > 0.0 ? w[i] * rsqrt(s) : 0.0;
}

void launch_normalize(double* vout, const double* w, const double* nrm2, int n, cudaStream_t st) {
    normalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(vout, w, nrm2, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// out = a + alpha * b   (alpha may be 0 to copy)
__global__ void axpby_kernel(double* __restrict__ out, const double* __restrict__ a, double alpha, const double* __restrict__ b,
                             int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fma(alpha, b[i], a[i]);
}

void launch_axpby(double* out, const double* a, double alpha, const double* b, int n, cudaStream_t st) {
    axpby_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, a, alpha, b, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// spectral preconditioner: hat[m] *= invP[m]  (flat-film symbol of the image operator), and the real <-> complex glue
__global__ void precond_scale_kernel(double2* __restrict__ hat, const double* __restrict__ invP, int N, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = invP[i % N];
    double2 c = hat[i];
    hat[i] = make_double2(c.x * s, c.y * s);
}

void launch_precond_scale(double2* hat, const double* invP, int N, int n, cudaStream_t st) {
    precond_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(hat, invP, N, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void real_to_complex_kernel(const double* __restrict__ x, double2* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(x[i], 0.0);
}

void launch_real_to_complex(const double* x, double2* out, int n, cudaStream_t st) {
    real_to_complex_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, out, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void complex_to_real_kernel(const double2* __restrict__ c, double* __restrict__ out, double scale, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = c[i].x * scale;
}

void launch_complex_to_real(const double2* c, double* out, double scale, int n, cudaStream_t st) {
    complex_to_real_kernel<<<(n + 255) / 256, 256, 0, st>>>(c, out, scale, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
