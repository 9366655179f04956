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

// the same on the half spectrum of a real transform (D2Z -> scale -> Z2D): hat[b][m] *= invP[m] * norm, m <= N/2 (the symbol is
// symmetric, invP[m] == invP[N - m], so the half spectrum is all there is to scale); norm = 1/N folds the transform pair's
// normalisation in
__global__ void precond_scale_half_kernel(double2* __restrict__ half, const double* __restrict__ invP, int nh, int n, double norm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = invP[i % nh] * norm;
    double2 c = half[i];
    half[i] = make_double2(c.x * s, c.y * s);
}

void launch_precond_scale_half(double2* half, const double* invP, int N, int batch, cudaStream_t st) {
    const int nh = N / 2 + 1, n = nh * batch;
    precond_scale_half_kernel<<<(n + 255) / 256, 256, 0, st>>>(half, invP, nh, n, 1.0 / N);
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


// ------------------------------------------------------------------------------------------------
// Device-driven GMRES cycle (recorded RK4 steps of the finite-depth helium operator): the Arnoldi process, the Givens rotations of
// the least-squares problem and the convergence decision of every ensemble member live on the device, one CTA per member, so that a
// fixed sequence  [P^-1 v_k | w = M (.) | arnoldi(k)] x K  can be recorded into a CUDA graph: every kernel of the sequence returns
// at once when the cycle has ended (GmCtrl::done), exactly as the Richardson sweeps do.  Members converge individually (their own
// residual against their own ||b||); the cycle ends when the last one has.
// Every reduction is a fixed tree over a fixed thread count: the same numbers on every rank of a row-sharded run.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kGmThreads = 1024;

// sums of up to kGmMax + 1 per-thread values over the block, deterministic: warp shuffles, then warp 0 over the 32 warp sums
template <int NV>
__device__ void block_sums(double (&v)[NV], int nv, double* sred /* [32][NV] */, double* out /* [NV] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        if (j < nv) {
            double x = v[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) sred[warp * NV + j] = x;
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (j < nv) {
                double x = sred[lane * NV + j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                if (lane == 0) out[j] = x;
            }
        }
    }
    __syncthreads();
}

// the member-level ticket: the last CTA of the launch closes the kernel for the whole ensemble
__device__ bool last_member(GmCtrl* gc, int batch, int still_running) {
    __shared__ unsigned int s_t;
    if (threadIdx.x == 0) {
        if (still_running) atomicAdd(&gc->running, 1);
        __threadfence();
        s_t = atomicAdd(&gc->ticket, 1u);
    }
    __syncthreads();
    return s_t == (unsigned)(batch - 1);
}

}  // namespace

// r = b - w (w = M x0), v_0 = r / ||r||, g = (||r||, 0, ...); members whose guess already meets the tolerance are finished
__global__ void __launch_bounds__(kGmThreads) gm_start_kernel(const double* __restrict__ b, const double* __restrict__ w,
                                                               double* __restrict__ V0, GmMember* members, GmCtrl* gc,
                                                               SolveCtrl* ctrl, int N, int batch, double tol) {
    __shared__ double sred[32 * 2];
    __shared__ double sums[2];
    const int m = blockIdx.x;
    const size_t off = (size_t)m * N;
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < N; i += kGmThreads) {
        const double bi = b[off + i];
        const double r = bi - w[off + i];
        V0[off + i] = r;
        acc[0] = fma(r, r, acc[0]);
        acc[1] = fma(bi, bi, acc[1]);
    }
    block_sums<2>(acc, 2, sred, sums);
    const double beta = sqrt(sums[0]), bnorm = sqrt(sums[1]);
    double rel = bnorm > 0.0 ? beta / bnorm : (beta == 0.0 ? 0.0 : 1e300);
    if (!(rel == rel)) rel = 1e300;
    const bool conv = rel <= tol;
    const double inv = (beta > 0.0 && !conv) ? 1.0 / beta : 0.0;
    for (int i = threadIdx.x; i < N; i += kGmThreads) V0[off + i] *= inv;
    GmMember* mem = members + m;
    if (threadIdx.x == 0) {
        mem->g[0] = beta;
        mem->bnorm = bnorm;
        mem->rel = rel;
        mem->first_rel = rel;
        mem->k_used = 0;
        mem->running = conv ? 0 : 1;
    }
    if (last_member(gc, batch, conv ? 0 : 1) && threadIdx.x == 0) {
        const int running = atomicAdd(&gc->running, 0);
        gc->done = running == 0 ? 1 : 0;
        gc->k = 0;
        gc->running = 0;
        gc->ticket = 0u;
        ctrl->iters += 1;
        __threadfence();
    }
}

// Arnoldi step k of every member still running: classical Gram-Schmidt twice (CGS2) of w = M P^-1 v_k against v_0..v_k, the new
// Hessenberg column through the stored and one new Givens rotation, the residual estimate |g_{k+1}| / ||b||
__global__ void __launch_bounds__(kGmThreads) gm_arnoldi_kernel(double* __restrict__ V, size_t ldv, double* __restrict__ w,
                                                                 GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, int N, int batch,
                                                                 int k, int last_k, double tol) {
    if (*reinterpret_cast<volatile int*>(&gc->done)) return;
    __shared__ double sred[32 * (kGmMax + 1)];
    __shared__ double h[kGmMax + 1], hsum[kGmMax + 1];
    const int m = blockIdx.x;
    const size_t off = (size_t)m * N;
    GmMember* mem = members + m;
    const bool running = mem->running != 0;
    if (running) {
        const int nv = k + 1;
        if (threadIdx.x <= kGmMax) hsum[threadIdx.x] = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double acc[kGmMax + 1];
#pragma unroll
            for (int j = 0; j <= kGmMax; ++j) acc[j] = 0.0;
            for (int i = threadIdx.x; i < N; i += kGmThreads) {
                const double wi = w[off + i];
#pragma unroll
                for (int j = 0; j <= kGmMax; ++j)
                    if (j < nv) acc[j] = fma(V[(size_t)j * ldv + off + i], wi, acc[j]);
            }
            block_sums<kGmMax + 1>(acc, nv, sred, h);
            for (int i = threadIdx.x; i < N; i += kGmThreads) {
                double wi = w[off + i];
#pragma unroll
                for (int j = 0; j <= kGmMax; ++j)
                    if (j < nv) wi = fma(-h[j], V[(size_t)j * ldv + off + i], wi);
                w[off + i] = wi;
            }
            if ((int)threadIdx.x < nv) hsum[threadIdx.x] += h[threadIdx.x];
            __syncthreads();
        }
        double nn[1] = {0.0};
        for (int i = threadIdx.x; i < N; i += kGmThreads) {
            const double wi = w[off + i];
            nn[0] = fma(wi, wi, nn[0]);
        }
        block_sums<1>(nn, 1, sred, h);
        const double hk1 = sqrt(h[0]);
        const double inv = hk1 > 0.0 ? 1.0 / hk1 : 0.0;
        double* vn = V + (size_t)(k + 1) * ldv + off;
        for (int i = threadIdx.x; i < N; i += kGmThreads) vn[i] = w[off + i] * inv;
        if (threadIdx.x == 0) {
            double col[kGmMax + 2];
            for (int j = 0; j <= k; ++j) col[j] = hsum[j];
            col[k + 1] = hk1;
            for (int j = 0; j < k; ++j) {   // the rotations of the earlier columns
                const double a0 = col[j], a1 = col[j + 1];
                col[j] = mem->cs[j] * a0 + mem->sn[j] * a1;
                col[j + 1] = -mem->sn[j] * a0 + mem->cs[j] * a1;
            }
            const double a0 = col[k], a1 = col[k + 1];
            const double d = hypot(a0, a1);
            const double c = d > 0.0 ? a0 / d : 1.0, sgn = d > 0.0 ? a1 / d : 0.0;
            mem->cs[k] = c;
            mem->sn[k] = sgn;
            col[k] = d;
            for (int j = 0; j <= k; ++j) mem->H[j * kGmMax + k] = col[j];
            const double gk = mem->g[k];
            mem->g[k + 1] = -sgn * gk;
            mem->g[k] = c * gk;
            double rel = mem->bnorm > 0.0 ? fabs(mem->g[k + 1]) / mem->bnorm : 0.0;
            if (!(rel == rel)) rel = 1e300;
            mem->rel = rel;
            mem->k_used = k + 1;
            if (rel <= tol || rel >= 1e300) mem->running = 0;
        }
        __syncthreads();
    }
    const int still = running ? (members[m].running != 0 ? 1 : 0) : 0;
    if (last_member(gc, batch, still) && threadIdx.x == 0) {
        const int r = atomicAdd(&gc->running, 0);
        gc->k = k + 1;
        if (r == 0 || k == last_k) gc->done = 1;
        gc->running = 0;
        gc->ticket = 0u;
        ctrl->iters += 1;
        __threadfence();
    }
}

// end of the cycle: y from the triangular system of each member's k_used columns, t = sum_j y_j v_j (the correction before P^-1)
__global__ void __launch_bounds__(kGmThreads) gm_correction_kernel(const double* __restrict__ V, size_t ldv, double* __restrict__ t,
                                                                    GmMember* members, int N) {
    __shared__ double y[kGmMax];
    const int m = blockIdx.x;
    const size_t off = (size_t)m * N;
    GmMember* mem = members + m;
    const int ku = mem->k_used;
    if (threadIdx.x == 0) {
        for (int i = ku - 1; i >= 0; --i) {
            double acc = mem->g[i];
            for (int j = i + 1; j < ku; ++j) acc -= mem->H[i * kGmMax + j] * y[j];
            const double d = mem->H[i * kGmMax + i];
            y[i] = d != 0.0 ? acc / d : 0.0;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += kGmThreads) {
        double acc = 0.0;
        for (int j = 0; j < ku; ++j) acc = fma(y[j], V[(size_t)j * ldv + off + i], acc);
        t[off + i] = acc;
    }
}

void launch_gm_start(const double* b, const double* w, double* V0, GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, int N, int batch,
                     double tol, cudaStream_t st) {
    gm_start_kernel<<<batch, kGmThreads, 0, st>>>(b, w, V0, members, gc, ctrl, N, batch, tol);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_gm_arnoldi(double* V, size_t ldv, double* w, GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, int N, int batch, int k,
                       int last_k, double tol, cudaStream_t st) {
    gm_arnoldi_kernel<<<batch, kGmThreads, 0, st>>>(V, ldv, w, members, gc, ctrl, N, batch, k, last_k, tol);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_gm_correction(const double* V, size_t ldv, double* t, GmMember* members, int N, int batch, cudaStream_t st) {
    gm_correction_kernel<<<batch, kGmThreads, 0, st>>>(V, ldv, t, members, N);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
