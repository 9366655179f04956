// pair_kernels.cu -- the O(N^2) cotangent-kernel sums of the Roberts boundary-integral method, matrix-free.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   createMKernel / createFiniteDepthMKernel + cusolverDnDgetrf/Dgetrs   (L/createM.cuh:43-92, L/MatrixSolver.cuh:114-125)
//   createVelocityMatrices / createHeliumVelocityMatrices + cublasZgemv  (L/WaterVelocities.cuh:38-107, 206-242)
// Both the application of M to a vector and the velocity summation reduce to the same complex row sum
//     S_k = sum_{j != k} cot((z_k - z_j)/2) x_j ,  x real,
// which is evaluated here without transcendental functions in the inner loop:
//     cot((z_k - z_j)/2) = i (E_k + E_j)/(E_k - E_j) = i [ 1 + 2 E_j/(E_k - E_j) ],   E = exp(i z)
//     S_k = i [ (sum_j x_j - x_k) + 2 T_k ],   T_k = sum_{j != k} F_j / (E_k - E_j),   F_j = x_j E_j .
// Near the diagonal E_k - E_j cancels; there the same formula is used with cell-local exponentials
//     P = expm1(i (z - z_c)),  E_k - E_j = E_c (P_k - P_j),  F_j = x_j (1 + P_j),
// z_c the centre of the source cell, which keeps every matrix entry accurate to ~1e-14 relative.
// Inner loop: 13 FP64-pipe instructions per pair (2 DADD, 3 DMUL, 8 DFMA) + one MUFU.RCP64H.
#include "internal.cuh"

namespace rb {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {
    // MUFU.RCP64H seed (>= 20 bits) + one cubically convergent step: relative error <= seed^3 < 1e-18
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    double t = fma(e, e, e);
    return fma(r, t, r);
}

__device__ __forceinline__ double ldcg_d(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg_d2(const double2* p) { return __ldcg(p); }

// deterministic block reduction of one register value: fixed tree -- shuffles inside a warp, then every thread adds the warp sums
// in warp order (two block barriers instead of one per tree level; the barriers of these reductions sit in the serial tail of a sweep)
template <int THREADS>
__device__ __forceinline__ double block_reduce_fixed(double v, double* sred) {
    static_assert(THREADS % 32 == 0 && THREADS <= 1024, "whole warps");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) r += sred[w];
    __syncthreads();
    return r;
}

// deterministic block sum of n values from global memory (fixed thread count, fixed tree)
template <int THREADS>
__device__ __forceinline__ double block_sum_fixed(const double* p, int n, double* sred) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += THREADS) s += ldcg_d(p + i);
    return block_reduce_fixed<THREADS>(s, sred);
}

// store a result at the same arena offset on every rank (single GPU: a plain store)
template <typename T>
__device__ __forceinline__ void mirror_store(const CommView& c, T* local_ptr, T v) {
    if (c.nranks <= 1) {
        *local_ptr = v;
        return;
    }
    const size_t off = (size_t)(reinterpret_cast<char*>(local_ptr) - c.my_base);
    for (int r = 0; r < c.nranks; ++r) *reinterpret_cast<T*>(c.peer_base[r] + off) = v;
}

// tell every rank that this rank's rows of the current exchange are in place (call after a system-scope fence)
__device__ __forceinline__ void comm_signal(const CommView& c) {
    const unsigned long long e = *c.signal_epoch + 1ull;
    *c.signal_epoch = e;
    for (int r = 0; r < c.nranks; ++r) {
        volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(c.peer_base[r] + c.off_flags) + c.rank;
        *f = e;
    }
    __threadfence_system();
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    double inv = 1.0 / (b.x * b.x + b.y * b.y);
    return make_double2((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}

// expm1(i (dx + i dy)) = exp(-dy) (cos dx + i sin dx) - 1, accurate for small |d|
__device__ __forceinline__ double2 cexpm1_i(double dx, double dy) {
    double s, c;
    sincos(dx, &s, &c);
    double em = expm1(-dy);
    double sh = sin(0.5 * dx);
    // (1+em) c - 1 = em c + (c - 1) = em c - 2 sin^2(dx/2)
    return make_double2(fma(em, c, -2.0 * sh * sh), (1.0 + em) * s);
}

__device__ __forceinline__ int centre_index(int cell, int N) {
    int i = cell * kCell + kCell / 2;
    return i < N ? i : N - 1;
}

// local coordinate of point z relative to the centre zc, wrapped to the nearest period in x with a two-word 2 pi
__device__ __forceinline__ double2 local_coord(double2 z, double2 zc) {
    // exact difference hi + lo (TwoSum)
    double hi = z.x - zc.x;
    double bb = hi - z.x;
    double lo = (z.x - (hi - bb)) + (-zc.x - bb);
    double m = rint(hi * (1.0 / kTwoPiHi));
    double dx = ((hi - m * kTwoPiHi) + lo) - m * kTwoPiLo;   // m in {-1,0,1}: m*kTwoPiHi exact, hi - m*2pi exact (Sterbenz) when wrapped
    return cexpm1_i(dx, z.y - zc.y);
}

// ------------------------------------------------------------------------------------------------
// geometry: everything that depends on the surface only (once per RHS, reused by every sweep)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void geometry_point(const Geometry& g, double2* __restrict__ phiprime_c, int N, int ncell, double rhoM,
                                               double depth, int finite_image, int use_local, int raw_derivs, double rho, double U,
                                               size_t tid, int i, int b) {
    const double2* Zb = g.Z + (size_t)b * N;
    double2 z = Zb[i];
    double2 zp = g.Zp[tid];
    double2 zpp = g.Zpp[tid];
    double2 php = phiprime_c ? phiprime_c[tid] : make_double2(0.0, 0.0);
    if (raw_derivs) {
        // the arrays hold the unscaled inverse transforms: apply 2 pi / N (squared for Z'') and put the linear parts back
        // (finish_zphi_kernel of spectral.cu folded in; L/Derivatives.cuh:321-324, 374, 380-383)
        const double s1 = 2.0 * kPi / N;
        const double s2 = 4.0 * kPi * kPi / ((double)N * N);
        zp = make_double2(zp.x * s1 + 2.0 * kPi / N, zp.y * s1);
        zpp = make_double2(zpp.x * s2, zpp.y * s2);
        g.Zp[tid] = zp;
        g.Zpp[tid] = zpp;
        if (phiprime_c) {
            php.x *= s1;
            php.y *= s1;
            if (U != 0) php.x += -(1 + rho) * kPi * U / N;
            phiprime_c[tid] = php;
        }
    }

    double s, c;
    sincos(z.x, &s, &c);
    double em = exp(-z.y);
    g.EG[tid] = make_double2(em * c, em * s);
    if (finite_image) {
        double ep = exp(z.y + 2.0 * depth);
        g.EI[tid] = make_double2(ep * c, ep * s);
    }
    if (use_local) {
        int cell = i / kCell;
        int cm = cell == 0 ? ncell - 1 : cell - 1;
        int cp = cell == ncell - 1 ? 0 : cell + 1;
        g.P0[tid] = local_coord(z, Zb[centre_index(cell, N)]);
        g.Pm[tid] = local_coord(z, Zb[centre_index(cm, N)]);
        g.Pp[tid] = local_coord(z, Zb[centre_index(cp, N)]);
    }
    // diagonal terms (without the image contribution: the image sum runs over every j including j == k)
    double2 q = cdiv(zpp, zp);                       // Zpp/Zp
    double cK = 0.25 * (1.0 - rhoM) / kPi;
    g.Mdiag[tid] = 0.5 * (1.0 + rhoM) + cK * q.y;    // L/createM.cuh:56, :79
    double2 q2 = cdiv(q, zp);                        // Zpp/Zp^2
    double2 hz = cdiv(make_double2(0.5, 0.0), zp);   // 1/(2 Zp)
    // -i/(4 pi) * q2 + 1/(2 Zp)   (L/WaterVelocities.cuh:55-59), multiply_by_i(-(1/4pi) q2) = (q2.y/(4pi), -q2.x/(4pi))
    g.V1diag[tid] = make_double2(q2.y * (0.25 / kPi) + hz.x, -q2.x * (0.25 / kPi) + hz.y);
    double2 iz = cdiv(make_double2(1.0 / (2.0 * kPi), 0.0), zp);
    g.V2[tid] = make_double2(-iz.y, iz.x);           // i/(2 pi Zp)   (:66)
    if (phiprime_c) g.b[tid] = php.x;    // complex_to_real, L/BaseBoundaryIntegrator.cuh:299
}

__global__ void geometry_kernel(Geometry g, double2* __restrict__ phiprime_c, int N, int batch, int ncell, int physics,
                                double rhoM, double depth, int finite_image, int use_local, int raw_derivs, double rho,
                                double U) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)N * batch) return;
    int b = (int)(tid / N);
    int i = (int)(tid - (size_t)b * N);
    geometry_point(g, phiprime_c, N, ncell, rhoM, depth, finite_image, use_local, raw_derivs, rho, U, tid, i, b);
}

void launch_geometry(const Geometry& g, double2* phiprime_c, int N, int batch, int ncell, int physics, double rhoM,
                     double depth, int finite_image, int use_local, int raw_derivs, double rho, double U, cudaStream_t st) {
    size_t n = (size_t)N * batch;
    int threads = 128;
    geometry_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(g, phiprime_c, N, batch, ncell, physics, rhoM,
                                                                                   depth, finite_image, use_local, raw_derivs,
                                                                                   rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// start of a solve: initial iterate, its per-cell sums, per-cell ||b||^2, control block reset.
// The iterate is, in this order of preference,
//   (a) the polynomial extrapolation of the solutions found at the same RK stage of the previous `order` steps
//       (history ring in device memory, position read from a device counter so that one CUDA graph serves every step),
//   (b) a caller-supplied warm vector,  (c) the first Neumann term omega * b.
// ------------------------------------------------------------------------------------------------
// (b, Zp, Mdiag deliberately not __restrict__: the fused geometry + guess kernel writes them through the Geometry pointers first)
__device__ __forceinline__ void guess_cell(const double* b, const double* __restrict__ warm, const HistoryRing& hist,
                                           double* __restrict__ x0, double* __restrict__ xsum_part,
                                           double* __restrict__ bnorm_part, SolveCtrl* ctrl, double omega, int N, int ncell,
                                           const double2* Zp, const double* Mdiag, double cK, double* sred) {
    int cell = blockIdx.x, bm = blockIdx.y;
    int i = cell * kCell + threadIdx.x;
    double xv = 0.0, bv = 0.0;
    int cnt = 0, m = 0;
    if (hist.base) {
        cnt = *hist.counter;
        m = cnt < hist.order ? cnt : hist.order;
    }
    if (i < N) {
        size_t o = (size_t)bm * N + i;
        bv = b[o];
        if (m > 0) {
            // Lagrange extrapolation to the next equally spaced point: c_i = (-1)^(i+1) C(m, i), i = 1..m
            // (m=1: 1 | 2: 2,-1 | 3: 3,-3,1 | 4: 4,-6,4,-1 | 5: 5,-10,10,-5,1 | 6: 6,-15,20,-15,6,-1)
            const bool pr = hist.predict && hist.Abase && Zp;
            double cf = 1.0, Aer = 0.0, Aei = 0.0;
            xv = 0.0;
            for (int back = 1; back <= m; ++back) {
                cf = -cf * (double)(m - back + 1) / (double)back;          // -> (-1)^back C(m, back)
                const size_t so = (size_t)((cnt - back) % hist.ring) * hist.stride + o;
                xv = fma(-cf, hist.base[so], xv);
                if (pr) {
                    const double2 A = hist.Abase[so];
                    Aer = fma(-cf, A.x, Aer);
                    Aei = fma(-cf, A.y, Aei);
                }
            }
            if (pr) {
                // predicted sweep: the row sums of the extrapolated iterate are the extrapolated row sums (linear in x, smooth in
                // time); today's Zp_k, Mdiag_k, b_k carry today's round-off noise, which is what the extrapolation cannot know
                const double2 zp = Zp[o];
                const double Mx = fma(Mdiag[o], xv, cK * (zp.x * Aer - zp.y * Aei));
                xv = fma(omega, bv - Mx, xv);
            }
        } else {
            xv = warm ? warm[o] : omega * bv;
        }
        x0[o] = xv;
    }
    double sx = block_reduce_fixed<kCell>(xv, sred);
    double sb = block_reduce_fixed<kCell>(bv * bv, sred);
    if (threadIdx.x == 0) {
        xsum_part[(size_t)bm * ncell + cell] = sx;
        bnorm_part[(size_t)bm * ncell + cell] = sb;
        if (cell == 0 && bm == 0) {
            ctrl->done = 0;
            ctrl->iters = 0;
            ctrl->final_buf = 0;
            ctrl->converged = 0;
            ctrl->stagnated = 0;
            ctrl->rel2 = 0.0;
            ctrl->prev_rel2 = 1e300;
            ctrl->first_rel2 = 1e300;
            ctrl->max_rel2_bits = 0ull;
            ctrl->members_done = 0u;
        }
    }
}

__global__ void __launch_bounds__(kCell) guess_kernel(const double* __restrict__ b, const double* __restrict__ warm,
                                                       HistoryRing hist, double* __restrict__ x0,
                                                       double* __restrict__ xsum_part, double* __restrict__ bnorm_part,
                                                       SolveCtrl* ctrl, double omega, int N, int ncell,
                                                       const double2* __restrict__ Zp, const double* __restrict__ Mdiag, double cK) {
    __shared__ double sred[kCell];
    guess_cell(b, warm, hist, x0, xsum_part, bnorm_part, ctrl, omega, N, ncell, Zp, Mdiag, cK, sred);
}

// geometry of a cell's points followed by the start of the solve for that cell: one launch instead of two (every thread reads back
// only what it wrote itself: b, Zp, Mdiag of its own point)
__global__ void __launch_bounds__(kCell) geometry_guess_kernel(Geometry g, double2* __restrict__ phiprime_c, int N, int ncell,
                                                                double rhoM, double depth, int finite_image, int use_local,
                                                                int raw_derivs, double rho, double U,
                                                                const double* __restrict__ warm, HistoryRing hist,
                                                                double* __restrict__ x0, double* __restrict__ xsum_part,
                                                                double* __restrict__ bnorm_part, SolveCtrl* ctrl, double omega,
                                                                double cK) {
    __shared__ double sred[kCell];
    const int i = blockIdx.x * kCell + threadIdx.x;
    const int bm = blockIdx.y;
    if (i < N)
        geometry_point(g, phiprime_c, N, ncell, rhoM, depth, finite_image, use_local, raw_derivs, rho, U, (size_t)bm * N + i, i, bm);
    guess_cell(g.b, warm, hist, x0, xsum_part, bnorm_part, ctrl, omega, N, ncell, g.Zp, g.Mdiag, cK, sred);
}

void launch_geometry_guess(const Geometry& g, double2* phiprime_c, int N, int batch, int ncell, double rhoM, double depth,
                           int finite_image, int use_local, double rho, double U, const double* warm, const HistoryRing& hist,
                           double* x0, double* xsum_part, double* bnorm_part, SolveCtrl* ctrl, double omega, double cK,
                           cudaStream_t st) {
    geometry_guess_kernel<<<dim3(ncell, batch), kCell, 0, st>>>(g, phiprime_c, N, ncell, rhoM, depth, finite_image, use_local, 1, rho,
                                                                U, warm, hist, x0, xsum_part, bnorm_part, ctrl, omega, cK);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_guess(const double* b, const double* warm, const HistoryRing& hist, double* x0, double* xsum_part,
                  double* bnorm_part, SolveCtrl* ctrl, double omega, int N, int batch, int ncell, cudaStream_t st,
                  const double2* Zp, const double* Mdiag, double cK) {
    guess_kernel<<<dim3(ncell, batch), kCell, 0, st>>>(b, warm, hist, x0, xsum_part, bnorm_part, ctrl, omega, N, ncell, Zp, Mdiag,
                                                       cK);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// end of a solve: pick the iterate buffer the control block names, publish a (real + complex), its per-cell sums and,
// when a history ring is attached, the ring slot of this step (real_to_complex of L/BaseBoundaryIntegrator.cuh:201 folded in)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCell) finish_solve_kernel(const double* __restrict__ buf0, const double* __restrict__ buf1,
                                                              const double2* __restrict__ A0, const double2* __restrict__ A1,
                                                              const SolveCtrl* ctrl, double* __restrict__ a_out,
                                                              double2* __restrict__ a_complex, double* __restrict__ xsum_part,
                                                              HistoryRing hist, int N, int ncell, FinishPost post) {
    __shared__ double sred[kCell];
    const int fb = (ctrl && ctrl->final_buf) ? 1 : 0;
    const double* src = fb ? buf1 : buf0;
    const double* nxt = fb ? buf0 : buf1;   // combined sweeps: x + omega r of the verified iterate
    int cell = blockIdx.x, bm = blockIdx.y;
    int i = cell * kCell + threadIdx.x;
    double v = 0.0;
    if (i < N) {
        size_t o = (size_t)bm * N + i;
        v = src[o];
        if (a_out) a_out[o] = v;
        if (a_complex) a_complex[o] = make_double2(v, 0.0);
        if (hist.base) {
            const size_t so = (size_t)(*hist.counter % hist.ring) * hist.stride + o;
            const bool keepA = hist.Abase && A0;
            hist.base[so] = (hist.store_next && !keepA) ? nxt[o] : v;   // with row sums the ring must hold the iterate they belong to
            if (keepA) hist.Abase[so] = (fb ? A1 : A0)[o];
        }
        if (post.vel) {
            // ---- what the sweep epilogue left open: the V2 a' term (its transform ran beside the sweep), dPhi/dt, the RK update ----
            double2 u = post.vel[o];                  // conj(w) so far
            double dphi = 0.0;
            if (post.aprime0) {
                const double2 ap = (fb ? post.aprime1 : post.aprime0)[o];
                const double2 v2 = post.V2[o];
                const double tr = v2.x * ap.x - v2.y * ap.y;
                const double ti = v2.x * ap.y + v2.y * ap.x;
                u = make_double2(u.x + tr, u.y - ti);
                post.vel[o] = u;
                double2 uu = post.vel_upper[o];
                post.vel_upper[o] = make_double2(uu.x + tr, uu.y - ti);
                if (post.dphi) {
                    const double wr = u.x, wi = -u.y;
                    const double y = post.Z[o].y;
                    if (post.rhs_phi_kind == 1) {
                        dphi = -y + 0.5 * (wr * wr + wi * wi);        // L/createM.cuh:105 at rho = 0
                    } else {
                        const double vdw = post.depth / 3.0;           // L/createM.cuh:113-115
                        dphi = vdw * pow(1.0 + y / post.depth, -3.0) - vdw + (0.5 * wr * wr + 0.5 * wi * wi);
                    }
                    post.dphi[o] = make_double2(dphi, 0.0);
                }
            } else if (post.dphi) {
                dphi = post.dphi[o].x;
            }
            if (post.update == 1) {          // y_i = y0 + c k   (cublasZaxpy, L/AutonomousRungeKuttaStepper.cuh:349-357)
                const double2 z0 = post.y0[o], p0 = post.y0[post.BN + o];
                post.y_out[o] = make_double2(fma(post.c, u.x, z0.x), fma(post.c, u.y, z0.y));
                post.y_out[post.BN + o] = make_double2(fma(post.c, dphi, p0.x), fma(post.c, 0.0, p0.y));
            } else if (post.update == 2) {   // y0 += h/6 (k1 + 2 k2 + 2 k3 + k4)   (add_k_vectors + Zaxpy, :243, :361)
                const double2 a1 = post.k1[o], a2 = post.k2[o], a3 = post.k3[o];
                const double2 b1 = post.k1[post.BN + o], b2 = post.k2[post.BN + o], b3 = post.k3[post.BN + o];
                const double2 z0 = post.y_out[o], p0 = post.y_out[post.BN + o];
                const double sx = a1.x + 2.0 * a2.x + 2.0 * a3.x + u.x, sy = a1.y + 2.0 * a2.y + 2.0 * a3.y + u.y;
                const double px = b1.x + 2.0 * b2.x + 2.0 * b3.x + dphi, py = b1.y + 2.0 * b2.y + 2.0 * b3.y + 0.0;
                post.y_out[o] = make_double2(fma(post.c, sx, z0.x), fma(post.c, sy, z0.y));
                post.y_out[post.BN + o] = make_double2(fma(post.c, px, p0.x), fma(post.c, py, p0.y));
            }
        }
    }
    double sx = block_reduce_fixed<kCell>(v, sred);
    if (threadIdx.x == 0 && xsum_part) xsum_part[(size_t)bm * ncell + cell] = sx;
}

void launch_finish_solve(const double* buf0, const double* buf1, const SolveCtrl* ctrl, double* a_out, double2* a_complex,
                         double* xsum_part, const HistoryRing& hist, int N, int batch, int ncell, cudaStream_t st,
                         const double2* A0, const double2* A1, const FinishPost* post) {
    FinishPost p;
    if (post) p = *post;
    finish_solve_kernel<<<dim3(ncell, batch), kCell, 0, st>>>(buf0, buf1, A0, A1, ctrl, a_out, a_complex, xsum_part, hist, N,
                                                              ncell, p);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void advance_counter_kernel(int* counter) { *counter += 1; }

// last kernel of a recorded RK4 step: advance the history counter and fold the four stage solves' control blocks into the chunk
// aggregate (so that the host can launch several recorded steps back to back and look once)
__global__ void step_end_kernel(int* counter, const SolveCtrl* __restrict__ ctrl_all, StepAgg* agg, int opt_mask) {
    if (counter) *counter += 1;
    if (!agg) return;
    bool all_done = true, failed = false;
    for (int i = 0; i < 4; ++i) {
        const SolveCtrl c = ctrl_all[i];
        all_done = all_done && c.done;
        failed = failed || (c.done && !c.converged && !c.stagnated);
        if (c.stagnated && !c.converged) agg->stagnated += 1;
        const int occ = c.iters + ((opt_mask >> i) & 1);
        if (occ > agg->max_occupied) agg->max_occupied = occ;
        agg->sum_iters += c.iters;
        if (!(c.rel2 <= agg->worst_rel2)) agg->worst_rel2 = c.rel2;   // NaN sticks
        agg->first_rel2[i] = c.first_rel2;
        agg->rel2_last[i] = c.rel2;
        agg->iters_last[i] = c.iters;
        agg->conv_last[i] = c.converged;
        agg->stag_last[i] = c.stagnated;
    }
    agg->steps += 1;
    if (!all_done) agg->not_done += 1;
    if (failed) agg->failed += 1;
}

void launch_step_end(int* counter, const SolveCtrl* ctrl_all, StepAgg* agg, int opt_mask, cudaStream_t st) {
    step_end_kernel<<<1, 1, 0, st>>>(counter, ctrl_all, agg, opt_mask);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_advance_counter(int* counter, cudaStream_t st) {
    advance_counter_kernel<<<1, 1, 0, st>>>(counter);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// the sweep
// ------------------------------------------------------------------------------------------------
struct __align__(16) SrcEntry {   // one source point in shared memory: 2 x LDS.128, broadcast to the warp
    double p, q;                  // e_j
    double fr, fi;                // F_j = x_j E_j
};

template <bool DIAG, int RPT>
__device__ __forceinline__ void tile_accumulate(const SrcEntry* __restrict__ sh, int tile, const double2 (&ek)[RPT],
                                                const int (&sd)[RPT], double2 (&acc)[RPT]) {
#pragma unroll 4
    for (int s = 0; s < tile; ++s) {
        const double2 e = *reinterpret_cast<const double2*>(&sh[s].p);
        const double2 f = *reinterpret_cast<const double2*>(&sh[s].fr);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            double dr = ek[r].x - e.x;
            double di = ek[r].y - e.y;
            double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp(n2);
            if (DIAG) inv = (s == sd[r]) ? 0.0 : inv;
            double tr = fma(f.y, di, f.x * dr);        // Re(F conj(d))
            double ti = fma(f.y, dr, -(f.x * di));     // Im(F conj(d))
            acc[r].x = fma(tr, inv, acc[r].x);
            acc[r].y = fma(ti, inv, acc[r].y);
        }
    }
}

// closing of a solver sweep (MV, or VEL in combined mode): per-cell sums, tickets, and the convergence decision by the last CTA
template <int THREADS>
__device__ __forceinline__ void solver_sweep_close(const SweepArgs& a, double sx, double sr, int bm, int cellK, double* sred,
                                                   unsigned int* s_ticket) {
    const int t = threadIdx.x;
    sx = block_reduce_fixed<THREADS>(sx, sred);
    sr = block_reduce_fixed<THREADS>(sr, sred);
    if (a.comm.nranks > 1) __threadfence_system();   // this CTA's remote row stores before its ticket
    __syncthreads();
    if (t == 0) {
        mirror_store(a.comm, a.xsum_part_out + (size_t)bm * a.ncell + cellK, sx);
        a.rnorm_part[(size_t)bm * a.ncell + cellK] = sr;
        if (a.comm.nranks > 1) __threadfence_system(); else __threadfence();
        *s_ticket = atomicAdd(a.member_tickets + bm, 1u);
    }
    __syncthreads();
    if (*s_ticket != (unsigned)(a.row_cells - 1)) return;
    // ---- level 2: last row cell of this batch member ---------------------------------------
    __threadfence();
    double rn = block_sum_fixed<THREADS>(a.rnorm_part + (size_t)bm * a.ncell + a.row_cell0, a.row_cells, sred);
    if (a.comm.nranks > 1) {
        // row-sharded: publish this rank's residual sum and signal; the decision is taken by comm_wait_kernel on every
        // rank from the same numbers in the same order
        if (t == 0) {
            a.member_tickets[bm] = 0u;
            double* slot = reinterpret_cast<double*>(a.comm.my_base + a.comm.off_rn) + a.out_buf * kMaxRanks + a.comm.rank;
            mirror_store(a.comm, slot, rn);
            __threadfence_system();
            comm_signal(a.comm);
        }
        return;
    }
    double bn = block_sum_fixed<THREADS>(a.bnorm_part + (size_t)bm * a.ncell, a.ncell, sred);
    if (t == 0) {
        a.member_tickets[bm] = 0u;
        double rel2 = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
        if (!(rel2 == rel2)) rel2 = 1e300;   // NaN -> "not converged"
        atomicMax(&a.ctrl->max_rel2_bits, (unsigned long long)__double_as_longlong(rel2));
        __threadfence();
        unsigned int m = atomicAdd(&a.ctrl->members_done, 1u);
        if (m == (unsigned)(a.batch - 1)) {
            // ---- level 3: last member -> decide ------------------------------------------------
            __threadfence();
            unsigned long long bits = atomicAdd(&a.ctrl->max_rel2_bits, 0ull);
            double worst = __longlong_as_double((long long)bits);
            volatile SolveCtrl* c = a.ctrl;
            solve_decide(c, worst, a.tol2, a.max_iters, a.final_buf_on_done);
            c->max_rel2_bits = 0ull;
            c->members_done = 0u;
            __threadfence();
        }
    }
}

// Far tiles (no cell-local coordinates, no j == k) and the image sum, every mode: with conj(d) = conj(E_k) - conj(E_j),
//   T_k = sum_j F_j conj(d)/|d|^2 = conj(E_k) U_k - V_k ,   U_k = sum_j F_j / |d|^2 (complex),  V_k = sum_j g_j / |d|^2 (real),
//   g_j = x_j |E_j|^2: 2 DADD + DMUL + DFMA + (MUFU + 3 DFMA) + 3 DFMA = 10 FP64-pipe instructions per pair for the 20 algorithmic
//   flops; the target-side product is applied once per row after the loop.  The two sums cancel by at most |E|/|E_k - E_j|
//   <= 1/(cell width) ~ 40 in far tiles (<= 1/(1 - e^{-2h}) for the image sources), i.e. ~1e-15 relative in the row sum.
template <int RPT>
__device__ __forceinline__ void tile_accumulate_far(const SrcEntry* __restrict__ sh, const double* __restrict__ sg, int tile,
                                                    const double2 (&ek)[RPT], double2 (&U)[RPT],
                                                    double (&V)[RPT]) {
#pragma unroll 4
    for (int s = 0; s < tile; ++s) {
        const double2 e = *reinterpret_cast<const double2*>(&sh[s].p);
        const double2 f = *reinterpret_cast<const double2*>(&sh[s].fr);
        const double gj = sg[s];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            double dr = ek[r].x - e.x;
            double di = ek[r].y - e.y;
            double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp(n2);
            U[r].x = fma(f.x, inv, U[r].x);
            U[r].y = fma(f.y, inv, U[r].y);
            V[r] = fma(gj, inv, V[r]);
        }
    }
}

template <int MODE, bool IMAGE, int RPT>
__global__ void __launch_bounds__(kCell / RPT, RPT == 4 ? 8 : (IMAGE ? 5 : 7)) sweep_kernel(const SweepArgs a) {
    constexpr int THREADS = kCell / RPT;   // one CTA = one 256-row cell: THREADS threads x RPT rows
    __shared__ SrcEntry sh[kCell];
    __shared__ SrcEntry shI[IMAGE ? kCell : 1];
    __shared__ double shg[kCell];
    __shared__ double shgI[IMAGE ? kCell : 1];
    constexpr bool REALPATH = (MODE == kSweepMV);   // solver sweeps publish Re(Zp T) only
    __shared__ double sred[THREADS];
    __shared__ unsigned int s_ticket;

    if ((MODE == kSweepMV || MODE == kSweepVEL) && a.skip_if_done) {
        if (*reinterpret_cast<volatile int*>(&a.ctrl->done)) return;
    }
    const int t = threadIdx.x;
    const int cellK = a.row_cell0 + blockIdx.x;
    const int chunk = blockIdx.y;
    const int bm = blockIdx.z;
    const int N = a.N;
    const size_t boff = (size_t)bm * N;
    const double* __restrict__ x = a.x + boff;
    const double2* __restrict__ EG = a.g.EG + boff;
    const double2* __restrict__ P0 = a.g.P0 + boff;

    int krow[RPT];
    int lrow[RPT];
    double2 acc[RPT], ekG[RPT], zpk[RPT], U[RPT], UI[RPT];
    double V[RPT], VI[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        lrow[r] = t + r * THREADS;
        krow[r] = cellK * kCell + lrow[r];
        acc[r] = make_double2(0.0, 0.0);
        U[r] = make_double2(0.0, 0.0);
        UI[r] = make_double2(0.0, 0.0);
        V[r] = 0.0;
        VI[r] = 0.0;
        ekG[r] = krow[r] < N ? EG[krow[r]] : make_double2(3.0e150, 0.0);
        zpk[r] = make_double2(0.0, 0.0);
        if (REALPATH && krow[r] < N) zpk[r] = a.g.Zp[boff + krow[r]];
    }

    const int tile = a.tile;
    const int tile0 = chunk * a.tiles_per_chunk;
    for (int it = 0; it < a.tiles_per_chunk; ++it) {
        const int j0 = (tile0 + it) * tile;
        if (j0 >= N) break;
        const int cellJ = j0 / kCell;
        int dist = cellJ - cellK;
        if (dist < 0) dist += a.ncell;
        const bool near = a.use_local && (dist == 0 || dist == 1 || dist == a.ncell - 1);
        // ---- stage the source tile -------------------------------------------------------
        __syncthreads();
        for (int s = t; s < tile; s += THREADS) {
            int j = j0 + s;
            SrcEntry e;
            if (j < N) {
                double xj = x[j];
                if (near) {
                    double2 p = P0[j];
                    e.p = p.x; e.q = p.y;
                    e.fr = xj * (1.0 + p.x); e.fi = xj * p.y;
                } else {
                    double2 g = EG[j];
                    e.p = g.x; e.q = g.y;
                    e.fr = xj * g.x; e.fi = xj * g.y;
                    shg[s] = xj * (g.x * g.x + g.y * g.y);
                }
                sh[s] = e;
                if (IMAGE) {
                    double2 gi = a.g.EI[boff + j];
                    SrcEntry ei;
                    ei.p = gi.x; ei.q = gi.y; ei.fr = xj * gi.x; ei.fi = xj * gi.y;
                    shI[s] = ei;
                    shgI[s] = xj * (gi.x * gi.x + gi.y * gi.y);
                }
            } else {
                e.p = 1.0e150; e.q = 0.0; e.fr = 0.0; e.fi = 0.0;   // contributes exactly 0
                sh[s] = e;
                shg[s] = 0.0;
                if (IMAGE) {
                    shI[s] = e;
                    shgI[s] = 0.0;
                }
            }
        }
        // ---- this tile's view of the targets ---------------------------------------------
        double2 ek[RPT];
        int sd[RPT];
        const double2* __restrict__ tk = EG;
        if (near) tk = (dist == 0) ? P0 : (dist == 1 ? a.g.Pp + boff : a.g.Pm + boff);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            ek[r] = krow[r] < N ? tk[krow[r]] : make_double2(3.0e150, 0.0);
            sd[r] = lrow[r] - (j0 - cellJ * kCell);
        }
        __syncthreads();
        if (dist == 0) tile_accumulate<true, RPT>(sh, tile, ek, sd, acc);
        else if (a.use_local && !near) tile_accumulate_far<RPT>(sh, shg, tile, ekG, U, V);
        else           tile_accumulate<false, RPT>(sh, tile, ek, sd, acc);
        if (IMAGE) tile_accumulate_far<RPT>(shI, shgI, tile, ekG, UI, VI);
    }

    // ---- publish the partial sums; the last CTA(s) of this row cell add them up ----------------------
    const size_t pbase = ((size_t)bm * a.nchunks + chunk) * N;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        if (krow[r] < N) {
            // far part of this chunk: T += conj(E_k) U - V
            acc[r].x += fma(ekG[r].x, U[r].x, ekG[r].y * U[r].y) - V[r];
            acc[r].y += fma(ekG[r].x, U[r].y, -(ekG[r].y * U[r].x));
            if (REALPATH) acc[r] = make_double2(zpk[r].x * acc[r].x - zpk[r].y * acc[r].y, 0.0);   // Re(Zp T) of this chunk
            a.partial[pbase + krow[r]] = acc[r];
            if (IMAGE)
                a.partial_img[pbase + krow[r]] = make_double2(fma(ekG[r].x, UI[r].x, ekG[r].y * UI[r].y) - VI[r],
                                                               fma(ekG[r].x, UI[r].y, -(ekG[r].y * UI[r].x)));
        }
    }
    __threadfence();
    __syncthreads();

    // sum of `count` consecutive slices (stride N) of a partial-sum workspace for this thread's rows: independent loads are issued in
    // batches (memory-level parallelism), the additions stay in slice order (deterministic)
    auto ordered_sum = [&](const double2* __restrict__ base, const double2* __restrict__ ibase, int count, double2 (&T)[RPT],
                           double2 (&TI)[RPT]) {
        constexpr int kBatch = (IMAGE ? 4 : 8) * 2 / RPT;
        const size_t cstride = (size_t)N;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            T[r] = make_double2(0.0, 0.0);
            TI[r] = make_double2(0.0, 0.0);
        }
        for (int c0 = 0; c0 < count; c0 += kBatch) {
            double2 v[RPT][kBatch], vi[RPT][kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const bool ok = (c0 + u) < count && krow[r] < N;
                    v[r][u] = ok ? ldcg_d2(base + (size_t)(c0 + u) * cstride + krow[r]) : make_double2(0.0, 0.0);
                    if (IMAGE) vi[r][u] = ok ? ldcg_d2(ibase + (size_t)(c0 + u) * cstride + krow[r]) : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    T[r].x += v[r][u].x; T[r].y += v[r][u].y;
                    if (IMAGE) { TI[r].x += vi[r][u].x; TI[r].y += vi[r][u].y; }
                }
            }
        }
    };

    double2 T[RPT], TI[RPT];
    unsigned int* cticket = a.cell_tickets + (size_t)bm * a.ncell + cellK;
    // Two levels when a.chunk_group > 0: the chunks of a row cell are reduced in groups as the groups complete (overlapped with the
    // CTAs still computing), and the last group to complete adds the group sums.  The serial tail of a launch shrinks from nchunks
    // to ~2 sqrt(nchunks) slices; the summation order (chunks in order within a group, groups in order) stays fixed.  One loop
    // with a single ordered_sum site keeps the register budget of the main loop what it was.
    const int grp = a.chunk_group;
    const int g = grp > 0 ? chunk / grp : 0;
    bool final_level = grp <= 0;
    for (;;) {
        unsigned int* ticket = final_level ? cticket : a.group_tickets + ((size_t)bm * a.ncell + cellK) * a.ngroups + g;
        const int first = final_level ? 0 : g * grp;
        const bool from_groups = final_level && grp > 0;
        const int count = final_level ? (grp > 0 ? a.ngroups : a.nchunks) : min(grp, a.nchunks - first);
        if (t == 0) s_ticket = atomicAdd(ticket, 1u);
        __syncthreads();
        if (s_ticket != (unsigned)(count - 1)) return;
        __threadfence();
        if (t == 0) *ticket = 0u;   // ready for the next launch
        const size_t off = from_groups ? (size_t)bm * a.ngroups * N : ((size_t)bm * a.nchunks + first) * N;
        const double2* src = (from_groups ? a.gpartial : a.partial) + off;
        const double2* srcI = IMAGE ? (from_groups ? a.gpartial_img : a.partial_img) + off : nullptr;
        ordered_sum(src, srcI, count, T, TI);
        if (final_level) break;
        const size_t gbase = ((size_t)bm * a.ngroups + g) * N;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            if (krow[r] < N) {
                a.gpartial[gbase + krow[r]] = T[r];
                if (IMAGE) a.gpartial_img[gbase + krow[r]] = TI[r];
            }
        }
        __threadfence();
        __syncthreads();
        final_level = true;
    }
    // ---- epilogue (one CTA per row cell gets here) ---------------------------------------------------
    const double sumx = block_sum_fixed<THREADS>(a.xsum_part + (size_t)bm * a.ncell, a.ncell, sred);
    const double inv4pi = 0.25 / kPi;

    if (MODE == kSweepMV) {
        double sx = 0.0, sr = 0.0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            if (krow[r] < N) {
                const size_t o = boff + krow[r];
                double xk = x[krow[r]];
                double2 zp = a.g.Zp[o];
                double Ar = (sumx - xk) + 2.0 * T[r].x;
                double Ai = 2.0 * T[r].y;
                // Im(Zp * S) with S = i A  ->  Re(Zp A);  the real-part path delivers T[r].x = Re(Zp T) directly
                double Kx = REALPATH ? a.cK * fma(zp.x, sumx - xk, 2.0 * T[r].x) : a.cK * (zp.x * Ar - zp.y * Ai);
                if (IMAGE) Kx -= inv4pi * (sumx + 2.0 * TI[r].x);   // -(1/4pi) Im(S_img), S_img = i (sumx + 2 T_img)
                double Mx = fma(a.g.Mdiag[o], xk, Kx);
                if (a.apply_only) {   // operator application for the Krylov solver: y = M x
                    mirror_store(a.comm, a.x_out + o, Mx);
                    continue;
                }
                double res = a.g.b[o] - Mx;
                double xn = fma(a.omega, res, xk);
                mirror_store(a.comm, a.x_out + o, xn);
                sx += xn;
                sr += res * res;
            }
        }
        if (a.apply_only) {
            if (a.comm.nranks > 1) {
                __threadfence_system();
                __syncthreads();
                if (t == 0) s_ticket = atomicAdd(a.member_tickets + bm, 1u);
                __syncthreads();
                if (s_ticket == (unsigned)(a.row_cells - 1) && t == 0) {
                    a.member_tickets[bm] = 0u;
                    __threadfence_system();
                    comm_signal(a.comm);
                }
            }
            return;
        }
        solver_sweep_close<THREADS>(a, sx, sr, bm, cellK, sred, &s_ticket);
        return;
    }

    if (MODE == kSweepVEL) {
        double sx = 0.0, sr = 0.0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            if (krow[r] < N) {
                const size_t o = boff + krow[r];
                double ak = x[krow[r]];
                double2 zp = a.g.Zp[o];
                double2 v1d = a.g.V1diag[o];
                double2 v2 = a.g.V2[o];
                const double2 ap = a.defer_aprime ? make_double2(0.0, 0.0) : a.aprime[o];   // deferred: finish_solve adds V2 a'
                // w = (-i/4pi) S + V1diag a + V2 a',  S = i A  ->  A/(4 pi)
                double wr = inv4pi * ((sumx - ak) + 2.0 * T[r].x) + v1d.x * ak + (v2.x * ap.x - v2.y * ap.y);
                double wi = inv4pi * (2.0 * T[r].y) + v1d.y * ak + (v2.x * ap.y + v2.y * ap.x);
                if (IMAGE) {   // + (i/4pi) S_img = -(sumx + 2 T_img)/(4 pi)
                    wr -= inv4pi * (sumx + 2.0 * TI[r].x);
                    wi -= inv4pi * (2.0 * TI[r].y);
                }
                mirror_store(a.comm, a.vel_lower + o, make_double2(wr, -wi));   // conj, L/WaterVelocities.cuh:241
                double2 az = cdiv(make_double2(ak, 0.0), zp);       // upper fluid: diagonal -1/(2Zp) instead of +1/(2Zp)
                a.vel_upper[o] = make_double2(wr - az.x, -(wi - az.y));
                if (a.dphi && !a.defer_aprime) {
                    double y = a.g.Z[o].y;
                    double kin = 0.5 * wr * wr + 0.5 * wi * wi;
                    double d;
                    if (a.rhs_phi_kind == 1) {
                        d = -y + 0.5 * (wr * wr + wi * wi);        // L/createM.cuh:105 at rho = 0
                    } else {
                        double vdw = a.depth / 3.0;                // L/createM.cuh:113-115
                        d = vdw * pow(1.0 + y / a.depth, -3.0) - vdw + kin;
                    }
                    mirror_store(a.comm, a.dphi + o, make_double2(d, 0.0));
                }
                if (a.combined) {
                    // the same row sum verifies the iterate: r = b - M a, and prepares the next one in case it is needed
                    double Ar = (sumx - ak) + 2.0 * T[r].x;
                    double Ai = 2.0 * T[r].y;
                    if (a.A_out) mirror_store(a.comm, a.A_out + o, make_double2(Ar, Ai));
                    double Kx = a.cK * (zp.x * Ar - zp.y * Ai);
                    if (IMAGE) Kx -= inv4pi * (sumx + 2.0 * TI[r].x);
                    double res = a.g.b[o] - fma(a.g.Mdiag[o], ak, Kx);
                    double xn = fma(a.omega, res, ak);
                    mirror_store(a.comm, a.x_out + o, xn);
                    sx += xn;
                    sr += res * res;
                }
            }
        }
        if (a.combined) {
            solver_sweep_close<THREADS>(a, sx, sr, bm, cellK, sred, &s_ticket);
            return;
        }
        if (a.comm.nranks > 1) {
            // the last row cell of this rank signals that its rows of k are in every arena
            __threadfence_system();
            __syncthreads();
            if (t == 0) s_ticket = atomicAdd(a.member_tickets + bm, 1u);
            __syncthreads();
            if (s_ticket == (unsigned)(a.row_cells - 1) && t == 0) {
                a.member_tickets[bm] = 0u;
                __threadfence_system();
                comm_signal(a.comm);
            }
        }
        return;
    }

    // RAW: S_k = i A_k
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        if (krow[r] < N) {
            double xk = x[krow[r]];
            double Ar = (sumx - xk) + 2.0 * T[r].x;
            double Ai = 2.0 * T[r].y;
            a.raw_out[boff + krow[r]] = make_double2(-Ai, Ar);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// row-sharded runs: wait until every rank has signalled the current exchange; for a solver sweep also take the convergence
// decision (identical on every rank: same residual sums, same order)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) comm_wait_kernel(CommView c, SolveCtrl* ctrl, int decide, int parity, int final_buf,
                                                        const double* __restrict__ bnorm_part, int ncell, double tol2,
                                                        int max_iters) {
    const int lane = threadIdx.x;
    // decide == 1: the sweep before was launched with skip_if_done and has skipped itself (no signal was sent): nothing to wait
    // for.  decide == 2: that sweep always runs and signals (rb_bench_sweep), so its epoch must always be consumed -- returning
    // here would leave the peers' flags ahead of wait_epoch and let every later wait pass on stale flags.
    // decide == 3: as 1 (the sweep before skipped itself when ctrl->done), but nothing to decide after the wait (operator application
    // of the recorded GMRES cycle)
    if ((decide == 1 || decide == 3) && *reinterpret_cast<volatile int*>(&ctrl->done)) return;
    const unsigned long long expected = *c.wait_epoch + 1ull;
    bool timed_out = false;
    if (lane < c.nranks) {
        volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(c.my_base + c.off_flags) + lane;
        const long long t0 = clock64();
        while (*f < expected) {
            if (clock64() - t0 > 20000000000ll) {   // ~10 s: a peer died; give up instead of hanging the device
                timed_out = true;
                break;
            }
        }
    }
    timed_out = __any_sync(0xffffffffu, timed_out);
    __threadfence_system();
    __syncwarp();
    if (lane == 0) *c.wait_epoch = expected;
    if (timed_out) {
        if (lane == 0) {
            *c.error_flag = 1;
            ctrl->done = 1;
            ctrl->converged = 0;
            ctrl->stagnated = 0;
        }
        return;
    }
    if (!decide || decide == 3) return;
    // ||b||^2 over all cells: lanes stride, then a fixed-order combine
    double bl = 0.0;
    for (int i = lane; i < ncell; i += 32) bl += __ldcg(bnorm_part + i);
    double bn = 0.0;
    for (int l = 0; l < 32; ++l) bn += __shfl_sync(0xffffffffu, bl, l);
    if (lane == 0) {
        volatile double* rnp = reinterpret_cast<double*>(c.my_base + c.off_rn) + parity * kMaxRanks;
        double rn = 0.0;
        for (int r = 0; r < c.nranks; ++r) rn += rnp[r];
        double worst = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
        solve_decide(ctrl, worst, tol2, max_iters, final_buf);
    }
}

void launch_comm_wait(const CommView& c, SolveCtrl* ctrl, int decide, int parity, int final_buf, const double* bnorm_part,
                      int ncell, double tol2, int max_iters, cudaStream_t st) {
    comm_wait_kernel<<<1, 32, 0, st>>>(c, ctrl, decide, parity, final_buf, bnorm_part, ncell, tol2, max_iters);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

template <int RPT>
static void launch_sweep_rpt(const SweepArgs& a, int mode, cudaStream_t st) {
    dim3 grid(a.row_cells, a.nchunks, a.batch);
    dim3 block(kCell / RPT);
    if (a.has_image) {
        if (mode == kSweepMV) sweep_kernel<kSweepMV, true, RPT><<<grid, block, 0, st>>>(a);
        else if (mode == kSweepVEL) sweep_kernel<kSweepVEL, true, RPT><<<grid, block, 0, st>>>(a);
        else throw std::runtime_error("raw cotangent sum with image term is not defined");
    } else {
        if (mode == kSweepMV) sweep_kernel<kSweepMV, false, RPT><<<grid, block, 0, st>>>(a);
        else if (mode == kSweepVEL) sweep_kernel<kSweepVEL, false, RPT><<<grid, block, 0, st>>>(a);
        else sweep_kernel<kSweepRAW, false, RPT><<<grid, block, 0, st>>>(a);
    }
}

void launch_sweep(const SweepArgs& a, int mode, cudaStream_t st) {
    if (a.rows_per_thread == 4) launch_sweep_rpt<4>(a, mode, st);
    else launch_sweep_rpt<2>(a, mode, st);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
