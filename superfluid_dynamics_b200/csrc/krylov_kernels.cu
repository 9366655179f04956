// krylov_kernels.cu -- the restarted GMRES used on the finite-depth helium operator, device-driven
// (the reference solves that system with cuSOLVER LU, L/MatrixSolver.cuh:114-125; its image term, L/createM.cuh:87-88, puts the
// spectrum of M between 1/2 and N/(4 pi), where the plain Neumann iteration of the water operator does not converge).
// All reductions use a fixed thread count and a fixed tree: results are deterministic and identical on every rank.
#include "internal.cuh"

namespace rb {

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

// spectral preconditioner on the half spectrum of a real transform (D2Z -> scale -> Z2D): hat[b][m] *= invP[m] * norm, m <= N/2 (the symbol is
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
                double x = lane < (int)(blockDim.x >> 5) ? sred[lane * NV + j] : 0.0;   // (blocks of fewer than 32 warps)
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
        atomicMax(&gc->worst_bits, (unsigned long long)__double_as_longlong(rel));
    }
    if (last_member(gc, batch, conv ? 0 : 1) && threadIdx.x == 0) {
        const int running = atomicAdd(&gc->running, 0);
        gc->done = running == 0 ? 1 : 0;
        gc->k = 0;
        gc->worst_rel = __longlong_as_double((long long)atomicAdd(&gc->worst_bits, 0ull));
        gc->worst_bits = 0ull;
        gc->running = 0;
        gc->ticket = 0u;
        ctrl->iters += 1;
        __threadfence();
    }
}

// Arnoldi step k of every member still running: classical Gram-Schmidt twice (CGS2) of w = M P^-1 v_k against v_0..v_k, the new
// Hessenberg column through the stored and one new Givens rotation, the residual estimate |g_{k+1}| / ||b||.
// Grid = (C, batch): C CTAs share one member, each owning a contiguous slice of the vectors (its slice of w stays in registers);
// the three reductions of the step (two projections, one norm) are slice partials in global memory, one in-kernel barrier over the
// member's C CTAs each, and then EVERY CTA adds the C partials in slice order -- the same numbers everywhere, on every rank.
// (Round-2 profile, one CTA per member: 150 us per step at N = 16384 -- replicated on every rank of a sharded run, where four sharded
// image sweeps take 330 us: the step, not the sweep, bounded the helium film on 8 GPUs.)  The host picks C so that all C x batch CTAs
// are co-resident (C = 1 for ensembles).
constexpr int kGmSliceThreads = 256;
constexpr int kGmSlice = 1024;                              // preferred elements per CTA (4 per thread)
constexpr int kGmPartStride = 2 * (kGmMax + 1) + 2;          // per CTA: projections of pass 0 | of pass 1 | norm (separate slots: a fast CTA
                                                            // may publish its next partial while a slow one still reads the previous ones)

namespace {
// barrier over the C CTAs of one member: monotonic arrival counter, target = arrivals expected so far
__device__ __forceinline__ void member_barrier(unsigned int* bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (atomicAdd(bar, 0u) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}
}  // namespace

__global__ void __launch_bounds__(kGmSliceThreads) gm_arnoldi_kernel(double* __restrict__ V, size_t ldv, double* __restrict__ w,
                                                                      GmMember* members, GmCtrl* gc, SolveCtrl* ctrl,
                                                                      double* __restrict__ part, int N, int batch, int k, int last_k,
                                                                      double tol) {
    if (*reinterpret_cast<volatile int*>(&gc->done)) return;
    __shared__ double sred[32 * (kGmMax + 1)];
    __shared__ double h[kGmMax + 1], hsum[kGmMax + 1];
    const int C = gridDim.x, c = blockIdx.x, m = blockIdx.y;
    const size_t off = (size_t)m * N;
    GmMember* mem = members + m;
    const bool running = mem->running != 0;
    const unsigned int gen = mem->gen;                      // Arnoldi steps this member has run so far: base of its barrier targets (read by every
                                                            // CTA of the member before its first arrival; advanced by CTA 0 after the last barrier)
    if (running) {
        const int nv = k + 1;
        const int S = (N + C - 1) / C;                      // this CTA's slice [lo, hi)
        const int lo = c * S, hi = min(N, lo + S);
        if (threadIdx.x <= kGmMax) hsum[threadIdx.x] = 0.0;
        double* mypart = part + ((size_t)m * C + c) * kGmPartStride;
        const double* allpart = part + (size_t)m * C * kGmPartStride;
        for (int pass = 0; pass < 2; ++pass) {
            double acc[kGmMax + 1];
#pragma unroll
            for (int j = 0; j <= kGmMax; ++j) acc[j] = 0.0;
            for (int i = lo + threadIdx.x; i < hi; i += kGmSliceThreads) {
                const double wi = w[off + i];
#pragma unroll
                for (int j = 0; j <= kGmMax; ++j)
                    if (j < nv) acc[j] = fma(V[(size_t)j * ldv + off + i], wi, acc[j]);
            }
            block_sums<kGmMax + 1>(acc, nv, sred, h);
            if (C > 1) {
                if ((int)threadIdx.x < nv) mypart[pass * (kGmMax + 1) + threadIdx.x] = h[threadIdx.x];
                member_barrier(&mem->bar, (3u * gen + (unsigned)pass + 1u) * (unsigned)C);
                if ((int)threadIdx.x < nv) {
                    double t = 0.0;
                    for (int cc = 0; cc < C; ++cc) t += __ldcg(allpart + (size_t)cc * kGmPartStride + pass * (kGmMax + 1) + threadIdx.x);
                    h[threadIdx.x] = t;
                }
                __syncthreads();
            }
            for (int i = lo + threadIdx.x; i < hi; i += kGmSliceThreads) {
                double wi = w[off + i];
#pragma unroll
                for (int j = 0; j <= kGmMax; ++j)
                    if (j < nv) wi = fma(-h[j], V[(size_t)j * ldv + off + i], wi);
                w[off + i] = wi;                            // (each thread re-reads only what it wrote itself)
            }
            if ((int)threadIdx.x < nv) hsum[threadIdx.x] += h[threadIdx.x];
            __syncthreads();
        }
        double nn[1] = {0.0};
        for (int i = lo + threadIdx.x; i < hi; i += kGmSliceThreads) {
            const double wi = w[off + i];
            nn[0] = fma(wi, wi, nn[0]);
        }
        block_sums<1>(nn, 1, sred, h);
        if (C > 1) {
            if (threadIdx.x == 0) mypart[2 * (kGmMax + 1)] = h[0];
            member_barrier(&mem->bar, (3u * gen + 3u) * (unsigned)C);
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int cc = 0; cc < C; ++cc) t += __ldcg(allpart + (size_t)cc * kGmPartStride + 2 * (kGmMax + 1));
                h[0] = t;
            }
            __syncthreads();
        }
        const double hk1 = sqrt(h[0]);
        const double inv = hk1 > 0.0 ? 1.0 / hk1 : 0.0;
        double* vn = V + (size_t)(k + 1) * ldv + off;
        for (int i = lo + threadIdx.x; i < hi; i += kGmSliceThreads) vn[i] = w[off + i] * inv;
        if (c == 0 && threadIdx.x == 0) {
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
            const double cr = d > 0.0 ? a0 / d : 1.0, sgn = d > 0.0 ? a1 / d : 0.0;
            mem->cs[k] = cr;
            mem->sn[k] = sgn;
            col[k] = d;
            for (int j = 0; j <= k; ++j) mem->H[j * kGmMax + k] = col[j];
            const double gk = mem->g[k];
            mem->g[k + 1] = -sgn * gk;
            mem->g[k] = cr * gk;
            double rel = mem->bnorm > 0.0 ? fabs(mem->g[k + 1]) / mem->bnorm : 0.0;
            if (!(rel == rel)) rel = 1e300;
            mem->rel = rel;
            mem->k_used = k + 1;
            if (rel <= tol || rel >= 1e300) mem->running = 0;
            mem->gen = gen + 1u;
        }
        __syncthreads();
    }
    // the launch-level ticket counts every CTA (C x batch); a member is still running if its CTA 0 says so
    const int still = (running && c == 0) ? (members[m].running != 0 ? 1 : 0) : 0;
    if (last_member(gc, batch * C, still) && threadIdx.x == 0) {
        const int r = atomicAdd(&gc->running, 0);
        gc->k = k + 1;
        gc->k_total += 1;
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

int gm_arnoldi_slices(int N, int batch) {
    // CTAs per member: all C x batch CTAs of a launch meet at in-kernel barriers (per member), so those of a member must be resident
    // together: keep the whole launch within one CTA per SM
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (N + kGmSlice - 1) / kGmSlice;
    return std::max(1, std::min(want, sms / std::max(1, batch)));
}

void launch_gm_arnoldi(double* V, size_t ldv, double* w, GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, double* part, int N, int batch,
                       int k, int last_k, double tol, cudaStream_t st) {
    const int C = gm_arnoldi_slices(N, batch);
    gm_arnoldi_kernel<<<dim3(C, batch), kGmSliceThreads, 0, st>>>(V, ldv, w, members, gc, ctrl, part, N, batch, k, last_k, tol);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_gm_correction(const double* V, size_t ldv, double* t, GmMember* members, int N, int batch, cudaStream_t st) {
    gm_correction_kernel<<<batch, kGmThreads, 0, st>>>(V, ldv, t, members, N);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
