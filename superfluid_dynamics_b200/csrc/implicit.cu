// implicit.cu -- the implicit side of the path: real-state wrapper of the RHS, finite-difference Jacobian from ONE pair of batched
// RHS evaluations (batch = 3N perturbed surfaces through the same matrix-free sweeps as every other RHS), and the two-stage
// Gauss-Legendre integrator (order 4, symplectic) with a damped Newton iteration on the stage slopes.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/, P/ = CuSuperHelium/Python/):
//   convertToComplexStateKernel, convertToRealRhsKernel, RealBoundaryItegralCalculator<N>::run   L/RealBoundaryIntegralCalculator.cuh:4-70
//   createInitialState, createInitialBatchedZ (x2), createJacobianMatrixFromPerturbedRhs,
//   JacobianCalculator<N>::calculateJacobian                                                     L/JacobianCalculator.cuh:11-284
//   stageStatesKernel, computeResidualsKernel + 2 cublasDdot, createKTrial, fillMMatrix, multiplyVector, calculateNextStateKernel,
//   GaussLegendre2<N>::gaussLegendreS2Step / residualAndPhi / runEvolution / copy*ToHost          L/GaussLegendre.cuh:16-612
//   calculateJacobian, calculatePerturbedStates256, integrateSimulationGL2 (+ _freeMemory)        L/Export.cu:392-518, 600-739
// The Newton iteration follows the reference's own Python statement of the integrator (P/integration/gauss_legendre.py:55-267),
// which is the algorithm the CUDA class transcribes; where the CUDA transcription departs from it (see gl2_step below) the
// Python statement is kept.  Everything here is issued on the stream of the batch-1 solver; the only host synchronisations are
// the ones the algorithm needs (one per residual evaluation: the Newton and Armijo decisions are taken on the host).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/roberts_b200.h"
// The host logic below also compiles under g++ with tests/cpp/cuda_emu.h standing in for the CUDA runtime and a mock of the RHS
// assembler behind it (RB_EMULATE, tests/cpp/emu_implicit.cpp): the CPU test tier drives rb_jacobian_* / rb_gl2_* and the legacy
// exports through the reference's trajectories without a GPU (tests/test_kernel_emulation.py).
#ifdef RB_EMULATE
#include "cuda_emu.h"
#else
#include "internal.cuh"
#endif
#include "implicit_kernels.cuh"
#include "launch.cuh"

using namespace rb;

#define RB_TRY try {
#define RB_CATCH                      \
    }                                 \
    catch (const std::exception& e) { \
        return report_error(e);       \
    }                                 \
    return 0;

namespace {

template <typename T>
T* dmalloc(size_t n) {
    T* p = nullptr;
    RB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

void check_rc(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + rb_last_error());
}

inline unsigned blocks_for(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

void launched() {
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// JacobianCalculator<N>
// ------------------------------------------------------------------------------------------------
struct rb_jacobian {
    int N = 0;
    double eps = 1e-6;                 // L/JacobianCalculator.cuh:171
    rb_solver* batched = nullptr;      // BaseBoundaryIntegralCalculator<N, 3N>
    double2* state = nullptr;          // 2N: the complex state
    double2 *zpos = nullptr, *zneg = nullptr, *rpos = nullptr, *rneg = nullptr;   // 6 N^2 each
    size_t calls = 0;
};

static void jacobian_free(rb_jacobian* j) {
    if (!j) return;
    if (j->batched) rb_destroy(j->batched);
    for (void* p : {(void*)j->state, (void*)j->zpos, (void*)j->zneg, (void*)j->rpos, (void*)j->rneg})
        if (p) cudaFree(p);
    delete j;
}

static cudaStream_t jacobian_stream(rb_jacobian* j) { return (cudaStream_t)rb_get_stream(j->batched); }

static void jacobian_calculate(rb_jacobian* j, const double* state_real, double* jac) {
    const int N = j->N;
    cudaStream_t st = jacobian_stream(j);
    RB_LAUNCH_EW(real_to_complex_state_kernel, blocks_for(N), 256, st, state_real, j->state, N);
    launched();
    RB_LAUNCH_EW(perturbed_states_kernel, dim3(blocks_for(N), 3 * N, 2), 256, st, j->state, j->zpos, j->zneg, j->eps, N);
    launched();
    check_rc(rb_rhs(j->batched, (const rb_complex*)j->zpos, (rb_complex*)j->rpos), "rb_rhs (Jacobian, +eps)");
    check_rc(rb_rhs(j->batched, (const rb_complex*)j->zneg, (rb_complex*)j->rneg), "rb_rhs (Jacobian, -eps)");
    RB_LAUNCH_EW(jacobian_from_perturbed_kernel, blocks_for((size_t)6 * N * N), 256, st, j->rpos, j->rneg, jac, N, j->eps);
    launched();
    ++j->calls;
}

// ------------------------------------------------------------------------------------------------
// GaussLegendre2<N>
// ------------------------------------------------------------------------------------------------
struct rb_gl2 {
    rb_solver* s = nullptr;        // batch-1 RHS assembler (the problem behind RealBoundaryItegralCalculator)
    rb_jacobian* jac = nullptr;
    rb_gl2_options opt;
    int N = 0;
    size_t n = 0;                  // 3N
    double* y = nullptr;           // devState
    bool owns_y = false;
    double *ynext = nullptr, *k = nullptr, *ktrial = nullptr, *ystage = nullptr, *fy = nullptr, *R = nullptr, *dK = nullptr;
    double *J1 = nullptr, *J2 = nullptr, *Jf = nullptr, *M = nullptr;
    double2 *cstate = nullptr, *crhs = nullptr;   // RealBoundaryItegralCalculator's devComplexState / devComplexRHS
    rb_solver* s2 = nullptr;       // owned batch-2 assembler: both stage states in one RHS evaluation
    double2 *cstate2 = nullptr, *crhs2 = nullptr;   // its [Z_1 | Z_2 | Phi_1 | Phi_2] state and RHS (4N each)
    double* sums = nullptr;        // device [2]
    double* h_sums = nullptr;      // pinned [2]
    int* lu_info = nullptr;
    int* h_info = nullptr;         // pinned
    double hmin = -1.0;
    std::vector<double> times;     // devTimes
    double* log = nullptr;         // devYs
    size_t log_count = 0, log_cap = 0;
    rb_gl2_stats stats{};
};

// options the Armijo search cannot terminate with are refused up front (the reference would loop forever on them)
static void gl2_check_options(const rb_gl2_options& o) {
    if (!(o.backtrack > 0.0 && o.backtrack < 1.0)) throw std::invalid_argument("Gauss-Legendre options: backtrack must lie in (0, 1)");
    if (!(o.minAlpha > 0.0)) throw std::invalid_argument("Gauss-Legendre options: minAlpha must be positive");
    if (!(o.newtonTolerance >= 0.0)) throw std::invalid_argument("Gauss-Legendre options: newtonTolerance must not be negative");
}

static cudaStream_t gl2_stream(rb_gl2* g) { return (cudaStream_t)rb_get_stream(g->s); }

static void gl2_free(rb_gl2* g) {
    if (!g) return;
    if (g->owns_y && g->y) cudaFree(g->y);
    for (void* p : {(void*)g->ynext, (void*)g->k, (void*)g->ktrial, (void*)g->ystage, (void*)g->fy, (void*)g->R, (void*)g->dK, (void*)g->J1,
                    (void*)g->J2, (void*)g->Jf, (void*)g->M, (void*)g->cstate, (void*)g->crhs, (void*)g->sums, (void*)g->lu_info,
                    (void*)g->log})
        if (p) cudaFree(p);
    if (g->s2) rb_destroy(g->s2);
    for (void* p : {(void*)g->cstate2, (void*)g->crhs2})
        if (p) cudaFree(p);
    if (g->h_sums) cudaFreeHost(g->h_sums);
    if (g->h_info) cudaFreeHost(g->h_info);
    delete g;
}

// RealBoundaryItegralCalculator<N>::run, L/RealBoundaryIntegralCalculator.cuh:59-70
static void real_rhs(rb_solver* s, double2* cstate, double2* crhs, const double* y, double* out, int N) {
    cudaStream_t st = (cudaStream_t)rb_get_stream(s);
    RB_LAUNCH_EW(real_to_complex_state_kernel, blocks_for(N), 256, st, y, cstate, N);
    launched();
    check_rc(rb_rhs(s, (const rb_complex*)cstate, (rb_complex*)crhs), "rb_rhs");
    RB_LAUNCH_EW(complex_to_real_rhs_kernel, blocks_for(N), 256, st, crhs, out, N);
    launched();
}

static void gl2_rhs(rb_gl2* g, const double* y, double* out) {
    real_rhs(g->s, g->cstate, g->crhs, y, out, g->N);
    ++g->stats.rhs_evaluations;
}

struct Staging {
    double residualNorm, phi, normK;
};

// residualAndPhi, L/GaussLegendre.cuh:589-611: stage states of the slopes k (2n: k1 | k2), f at both, R = k - f, phi = |R|^2 / 2.
// Both stages go through the batch-2 assembler as one RHS evaluation; one host synchronisation.
static Staging gl2_residual_and_phi(rb_gl2* g, const double* y, const double* k, double h) {
    const size_t n = g->n;
    cudaStream_t st = gl2_stream(g);
    RB_LAUNCH_EW(gl2_stage_states_batched_kernel, blocks_for(g->N), 256, st, y, h, k, k + n, g->ystage, g->ystage + n, g->cstate2,
                 g->N);
    launched();
    check_rc(rb_rhs(g->s2, (const rb_complex*)g->cstate2, (rb_complex*)g->crhs2), "rb_rhs (both stages)");
    RB_LAUNCH_EW(gl2_batched_rhs_to_real_kernel, blocks_for(g->N), 256, st, g->crhs2, g->fy, g->N);
    launched();
    g->stats.rhs_evaluations += 2;
    RB_LAUNCH(gl2_residual_kernel, 1, kNormThreads, st, g->fy, k, g->R, 2 * n, g->sums);
    launched();
    RB_CUDA(cudaMemcpyAsync(g->h_sums, g->sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    Staging r;
    r.phi = 0.5 * g->h_sums[0];
    r.residualNorm = std::sqrt(g->h_sums[0]);
    r.normK = std::sqrt(g->h_sums[1]);
    // slopes that throw the stage states far off the surface can leave the inner solve (or the sums) without an answer: such
    // slopes are simply not acceptable to the line search
    double status[8];   // [0] converged, [1] stagnated on the round-off floor (accepted); neither: the solve failed
    check_rc(rb_solve_status(g->s2, status), "rb_solve_status");
    if ((status[0] == 0.0 && status[1] == 0.0) || !(r.phi == r.phi)) {
        r.phi = HUGE_VAL;
        r.residualNorm = HUGE_VAL;
    }
    return r;
}

static void gl2_jacobian(rb_gl2* g, const double* y, double* J) {
    jacobian_calculate(g->jac, y, J);
    ++g->stats.jacobians;
    double status[8];
    check_rc(rb_solve_status(g->jac->batched, status), "rb_solve_status (Jacobian)");
    if (status[0] == 0.0 && status[1] == 0.0) throw std::runtime_error("Gauss-Legendre: the batched solve behind the Jacobian did not converge");
}

// gaussLegendreS2Step, L/GaussLegendre.cuh:441-560 == gauss_legendre_s2_step, P/integration/gauss_legendre.py:55-170.
// Kept from the Python statement where the CUDA transcription departs from it:
//   * Newton matrix blocks (1,2) = -h a12 J1 and (2,1) = -h a21 J2 (fillMMatrix exchanges the two coefficients, :45-48, and is
//     launched with grid and block exchanged, :483, which cannot run for 3N / 16 > 32);
//   * residual norm and phi over both stages (cublasDdot over 3N of the 6N entries, :606), tolerance relative to 1 + |k|
//     (1 + |k|^2 there, :607 with :469);
//   * stage states from the slopes under test (stageStates reads the member buffers whatever was passed, :580-586).
// Returns false when the line search fails without a fallback left (the state is not advanced).
static bool gl2_step(rb_gl2* g, const double* ycur, double* dest, double h) {
    const size_t n = g->n;
    const rb_gl2_options& o = g->opt;
    cudaStream_t st = gl2_stream(g);
    rb_gl2_stats& res = g->stats;
    res.converged = 0;
    res.simplifiedFallbackUsed = 0;
    res.numberIterations = 0;
    bool freeze = false;
    double* k = g->k;
    double* ktrial = g->ktrial;

    // predictor: both stages start from f(y)
    gl2_rhs(g, ycur, k);
    RB_CUDA(cudaMemcpyAsync(k + n, k, n * sizeof(double), cudaMemcpyDeviceToDevice, st));

    Staging sr{};
    bool have_sr = false;   // an accepted trial leaves its residual, stage states and norms in place: not evaluated twice
    size_t it = 0;
    for (; it < o.maxNewtonIterations; ++it) {
        if (!have_sr) sr = gl2_residual_and_phi(g, ycur, k, h);
        have_sr = false;
        if (sr.residualNorm <= o.newtonTolerance * (1.0 + sr.normK)) {
            res.converged = 1;
            break;
        }
        const double *J1 = g->Jf, *J2 = g->Jf;
        if (!freeze) {
            gl2_jacobian(g, g->ystage, g->J1);
            gl2_jacobian(g, g->ystage + n, g->J2);
            J1 = g->J1;
            J2 = g->J2;
        }
        {
            dim3 th(32, 8), bl(blocks_for(n, 32), blocks_for(n, 8));
            RB_LAUNCH_EW(gl2_newton_matrix_kernel, bl, th, st, J1, J2, h, g->M, n);
            launched();
        }
        RB_LAUNCH_EW(gl2_negate_kernel, blocks_for(2 * n), 256, st, g->R, g->dK, 2 * n);
        launched();
        launch_lu_solve(g->M, g->dK, (int)(2 * n), g->lu_info, st);   // MatrixSolver<6N,1>::solve, :487
        RB_CUDA(cudaMemcpyAsync(g->h_info, g->lu_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        ++res.linear_solves;
        ++res.newton_iterations;

        double alpha = 1.0;
        const double r2 = sr.residualNorm * sr.residualNorm;
        double target = sr.phi - o.armijo_c * alpha * r2;
        bool first_trial = true;
        while (true) {
            RB_LAUNCH_EW(gl2_trial_kernel, blocks_for(2 * n), 256, st, k, alpha, g->dK, ktrial, 2 * n);
            launched();
            Staging tr = gl2_residual_and_phi(g, ycur, ktrial, h);   // synchronises: h_info is valid from here on
            if (first_trial && *g->h_info != 0)
                throw std::runtime_error("Gauss-Legendre Newton matrix is singular (pivot " + std::to_string(*g->h_info) + ")");
            first_trial = false;
            if (tr.phi <= target) {
                std::swap(k, ktrial);
                sr = tr;
                have_sr = true;
                break;
            }
            alpha *= o.backtrack;
            target = sr.phi - o.armijo_c * alpha * r2;
            if (alpha < o.minAlpha) {
                if (!freeze && o.allowSimplifiedFallback) {
                    freeze = true;
                    res.simplifiedFallbackUsed = 1;
                    gl2_jacobian(g, ycur, g->Jf);   // chord Newton about the base point of the step
                    break;
                }
                res.converged = 0;
                res.numberIterations = it;
                res.residualNorm = sr.residualNorm;
                return false;
            }
        }
    }
    if (!res.converged && !have_sr) sr = gl2_residual_and_phi(g, ycur, k, h);
    RB_LAUNCH_EW(gl2_next_state_kernel, blocks_for(n), 256, st, ycur, h, k, k + n, dest, n);
    launched();
    res.numberIterations = it;
    res.residualNorm = sr.residualNorm;
    return true;
}

static void gl2_log_append(rb_gl2* g, double t) {
    const size_t n = g->n;
    cudaStream_t st = gl2_stream(g);
    if (g->log_count == g->log_cap) {
        const size_t cap = std::max<size_t>(64, 2 * g->log_cap);
        double* grown = dmalloc<double>(cap * n);
        if (g->log_count) RB_CUDA(cudaMemcpyAsync(grown, g->log, g->log_count * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (g->log) {
            RB_CUDA(cudaStreamSynchronize(st));
            cudaFree(g->log);
        }
        g->log = grown;
        g->log_cap = cap;
    }
    RB_CUDA(cudaMemcpyAsync(g->log + g->log_count * n, g->y, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    ++g->log_count;
    g->times.push_back(t);
}

// runEvolution, L/GaussLegendre.cuh:216-299 == integrate_gl2, P/integration/gauss_legendre.py:173-267.  As in the Python statement the
// running step size and the lower bound |t1 - t0| / 2^20 belong to the call: the CUDA class writes the (possibly reduced, possibly
// end-truncated) size back into its options (:290) and keeps the first call's bound (:226-228), so that a second runEvolution on
// the same object would start from whatever sliver the first one ended with.
static void gl2_evolve(rb_gl2* g, double t0, double t1) {
    const rb_gl2_options& o = g->opt;
    if (!(o.stepSize > 0.0)) throw std::invalid_argument("Step size must be positive.");   // the reference tests < 0 only: 0 (or NaN) never ends
    if (!g->y) throw std::runtime_error("Initial state not set. Call initialize() before running evolution.");
    const size_t n = g->n;
    cudaStream_t st = gl2_stream(g);
    const double total = std::fabs(t1 - t0);
    g->hmin = total / std::pow(2.0, 20);
    double h = o.stepSize;
    if (o.returnTrajectory) gl2_log_append(g, t0);
    double t = t0;
    const double forward = t1 >= t0 ? 1.0 : -1.0;
    while ((t - t1) * forward < 0.0) {
        double htry = std::min(h, std::fabs(t1 - t)) * forward;
        bool ok = false;
        for (size_t i = 0; i < o.maxStepsHalves + 1; ++i) {
            if (gl2_step(g, g->y, g->ynext, htry) && g->stats.converged) {
                ok = true;
                break;
            }
            if (std::fabs(htry) <= g->hmin) break;
            htry *= 0.5;
            ++g->stats.steps_halved;
        }
        if (!ok) {
            char msg[256];
            std::snprintf(msg, sizeof(msg), "Gauss-Legendre 2nd Order method failed to converge at t~%g; residual=%.3e; last htry=%.3e", t,
                          g->stats.residualNorm, htry);
            throw std::runtime_error(msg);
        }
        RB_CUDA(cudaMemcpyAsync(g->y, g->ynext, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        t += htry;
        h = std::fabs(htry);   // the (possibly reduced) size is kept for the following steps of this call
        ++g->stats.steps_accepted;
        if (o.returnTrajectory) gl2_log_append(g, t);
    }
    RB_CUDA(cudaStreamSynchronize(st));
}

static void malloc_copy_out(const std::vector<double>& v, double** out) {
    double* p = (double*)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(double));
    if (!p) throw std::runtime_error("out of host memory");
    std::memcpy(p, v.data(), v.size() * sizeof(double));
    *out = p;
}

extern "C" {

int rb_real_rhs(rb_solver* s, const double* state_real_dev, double* rhs_real_dev) {
    RB_TRY
    if (!s || !state_real_dev || !rhs_real_dev) throw std::runtime_error("rb_real_rhs: null argument");
    int N = 0, B = 0;
    check_rc(rb_get_props(s, nullptr, &N, &B), "rb_get_props");
    if (B != 1) throw std::runtime_error("rb_real_rhs: the solver must have batch 1");
    double2* tmp = dmalloc<double2>((size_t)4 * N);
    try {
        real_rhs(s, tmp, tmp + 2 * N, state_real_dev, rhs_real_dev, N);
        RB_CUDA(cudaStreamSynchronize((cudaStream_t)rb_get_stream(s)));
    } catch (...) {
        cudaFree(tmp);
        throw;
    }
    cudaFree(tmp);
    RB_CATCH
}

int rb_perturbed_states(const rb_complex* state_dev, rb_complex* batched_dev, double eps, int N, void* cuda_stream) {
    RB_TRY
    if (!state_dev || !batched_dev || N < 1) throw std::runtime_error("rb_perturbed_states: bad argument");
    RB_LAUNCH_EW(perturbed_states_kernel, dim3(blocks_for(N), 3 * N, 1), 256, (cudaStream_t)cuda_stream, (const double2*)state_dev,
                 (double2*)batched_dev, nullptr, eps, N);
    launched();
    RB_CATCH
}

rb_jacobian* rb_jacobian_create(int N, const rb_props* props) {
    try {
        if (N < 2) throw std::runtime_error("rb_jacobian_create: N must be >= 2");
        std::unique_ptr<rb_jacobian, void (*)(rb_jacobian*)> j(new rb_jacobian, jacobian_free);
        j->N = N;
        rb_props p;
        if (props) p = *props; else rb_default_props(&p);
        p.guess_mode = RB_GUESS_WARM;   // the -eps batch starts from the +eps solutions, the next Jacobian from this one's
        p.compute_energies = 0;
        j->batched = rb_create(N, 3 * N, &p);
        if (!j->batched) throw std::runtime_error(std::string("rb_create (batch 3N): ") + rb_last_error());
        const size_t big = (size_t)6 * N * N;
        j->state = dmalloc<double2>((size_t)2 * N);
        j->zpos = dmalloc<double2>(big);
        j->zneg = dmalloc<double2>(big);
        j->rpos = dmalloc<double2>(big);
        j->rneg = dmalloc<double2>(big);
        return j.release();
    } catch (const std::exception& e) {
        report_error(e);
        return nullptr;
    }
}

int rb_jacobian_destroy(rb_jacobian* j) {
    RB_TRY
    if (j) {
        cudaDeviceSynchronize();
        jacobian_free(j);
    }
    RB_CATCH
}

int rb_jacobian_set_epsilon(rb_jacobian* j, double eps) {
    RB_TRY
    if (!j) throw std::runtime_error("rb_jacobian_set_epsilon: null calculator");
    j->eps = eps;
    RB_CATCH
}

int rb_jacobian_set_stream(rb_jacobian* j, void* cuda_stream) {
    RB_TRY
    if (!j) throw std::runtime_error("rb_jacobian_set_stream: null calculator");
    check_rc(rb_set_stream(j->batched, cuda_stream), "rb_set_stream");
    RB_CATCH
}

rb_solver* rb_jacobian_solver(rb_jacobian* j) { return j ? j->batched : nullptr; }

int rb_jacobian_calculate(rb_jacobian* j, const double* state_real_dev, double* jac_dev) {
    RB_TRY
    if (!j || !state_real_dev || !jac_dev) throw std::runtime_error("rb_jacobian_calculate: null argument");
    jacobian_calculate(j, state_real_dev, jac_dev);
    RB_CATCH
}

int rb_lu_solve(double* A_dev, double* b_dev, int n, int blocked, int* info_host, void* cuda_stream) {
    RB_TRY
    if (!A_dev || !b_dev || n < 1) throw std::runtime_error("rb_lu_solve: bad argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int* info = dmalloc<int>(1);
    int h = 0;
    try {
        if (blocked > 0) launch_lu_solve_blocked(A_dev, b_dev, n, info, st);
        else if (blocked == 0) launch_lu_solve_unblocked(A_dev, b_dev, n, info, st);
        else launch_lu_solve(A_dev, b_dev, n, info, st);
        RB_CUDA(cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFree(info);
        throw;
    }
    cudaFree(info);
    if (info_host) *info_host = h;
    RB_CATCH
}

void rb_gl2_default_options(rb_gl2_options* o) {
    o->stepSize = 0.01;
    o->newtonTolerance = 1e-10;
    o->maxNewtonIterations = 20;
    o->allowSimplifiedFallback = 0;
    o->returnTrajectory = 1;
    o->armijo_c = 1e-4;
    o->backtrack = 0.5;
    o->minAlpha = 1e-6;
    o->maxStepsHalves = 6;
}

rb_gl2* rb_gl2_create(rb_solver* s, rb_jacobian* j, const rb_gl2_options* options) {
    try {
        if (!s || !j) throw std::runtime_error("rb_gl2_create: null argument");
        int N = 0, B = 0;
        check_rc(rb_get_props(s, nullptr, &N, &B), "rb_get_props");
        if (B != 1) throw std::runtime_error("rb_gl2_create: the RHS assembler must have batch 1");
        if (j->N != N) throw std::runtime_error("rb_gl2_create: the Jacobian calculator was built for a different N");
        std::unique_ptr<rb_gl2, void (*)(rb_gl2*)> g(new rb_gl2, gl2_free);
        g->s = s;
        g->jac = j;
        if (options) g->opt = *options; else rb_gl2_default_options(&g->opt);
        gl2_check_options(g->opt);
        g->N = N;
        const size_t n = g->n = (size_t)3 * N;
        check_rc(rb_set_stream(j->batched, rb_get_stream(s)), "rb_set_stream (Jacobian)");   // one stream: plain program order
        g->ynext = dmalloc<double>(n);
        g->k = dmalloc<double>(2 * n);
        g->ktrial = dmalloc<double>(2 * n);
        g->ystage = dmalloc<double>(2 * n);
        g->fy = dmalloc<double>(2 * n);
        g->R = dmalloc<double>(2 * n);
        g->dK = dmalloc<double>(2 * n);
        g->J1 = dmalloc<double>(n * n);
        g->J2 = dmalloc<double>(n * n);
        g->Jf = dmalloc<double>(n * n);
        g->M = dmalloc<double>(4 * n * n);
        {
            rb_props p2;
            check_rc(rb_get_props(s, &p2, nullptr, nullptr), "rb_get_props");
            p2.compute_energies = 0;
            g->s2 = rb_create(N, 2, &p2);
            if (!g->s2) throw std::runtime_error(std::string("rb_create (batch 2): ") + rb_last_error());
            // trial slopes of a diverging Newton iteration can describe surfaces that have no finite RHS: the inner solve then fails,
            // which the line search must see as "not acceptable" (status below), not as an error of the call
            check_rc(rb_set_strict(g->s2, 0), "rb_set_strict (stage assembler)");
            check_rc(rb_set_stream(g->s2, rb_get_stream(s)), "rb_set_stream (stage assembler)");
            g->cstate2 = dmalloc<double2>((size_t)4 * N);
            g->crhs2 = dmalloc<double2>((size_t)4 * N);
        }
        g->cstate = dmalloc<double2>((size_t)2 * N);
        g->crhs = dmalloc<double2>((size_t)2 * N);
        g->sums = dmalloc<double>(2);
        g->lu_info = dmalloc<int>(1);
        RB_CUDA(cudaMallocHost(&g->h_sums, 2 * sizeof(double)));
        RB_CUDA(cudaMallocHost(&g->h_info, sizeof(int)));
        *g->h_info = 0;
        return g.release();
    } catch (const std::exception& e) {
        report_error(e);
        return nullptr;
    }
}

int rb_gl2_destroy(rb_gl2* g) {
    RB_TRY
    if (g) {
        cudaDeviceSynchronize();
        gl2_free(g);
    }
    RB_CATCH
}

int rb_gl2_set_options(rb_gl2* g, const rb_gl2_options* options) {
    RB_TRY
    if (!g || !options) throw std::runtime_error("rb_gl2_set_options: null argument");
    gl2_check_options(*options);
    g->opt = *options;
    RB_CATCH
}

int rb_gl2_initialize(rb_gl2* g, double* state, int on_device) {
    RB_TRY
    if (!g || !state) throw std::runtime_error("rb_gl2_initialize: null argument");
    cudaStream_t st = gl2_stream(g);
    if (on_device) {
        if (g->owns_y && g->y) cudaFree(g->y);
        g->y = state;   // caller keeps ownership, L/GaussLegendre.cuh:306-308
        g->owns_y = false;
    } else {
        if (!g->owns_y || !g->y) g->y = dmalloc<double>(g->n);
        g->owns_y = true;
        RB_CUDA(cudaMemcpyAsync(g->y, state, g->n * sizeof(double), cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaStreamSynchronize(st));
    }
    RB_CATCH
}

int rb_gl2_step(rb_gl2* g, double h, int* converged) {
    RB_TRY
    if (!g || !g->y) throw std::runtime_error("rb_gl2_step: initialize() has not been called");
    const bool ok = gl2_step(g, g->y, g->ynext, h) && g->stats.converged;
    if (ok) {
        RB_CUDA(cudaMemcpyAsync(g->y, g->ynext, g->n * sizeof(double), cudaMemcpyDeviceToDevice, gl2_stream(g)));
        ++g->stats.steps_accepted;
    }
    RB_CUDA(cudaStreamSynchronize(gl2_stream(g)));
    if (converged) *converged = ok ? 1 : 0;
    RB_CATCH
}

int rb_gl2_evolve(rb_gl2* g, double t0, double t1) {
    RB_TRY
    if (!g) throw std::runtime_error("rb_gl2_evolve: null integrator");
    gl2_evolve(g, t0, t1);
    RB_CATCH
}

int rb_gl2_copy_trajectory(rb_gl2* g, double** times_out, size_t* times_count, double** states_out, size_t* states_count) {
    RB_TRY
    if (!g || !states_out || !states_count) throw std::runtime_error("rb_gl2_copy_trajectory: null argument");
    const bool traj = g->opt.returnTrajectory != 0;
    if (!traj && !g->y) throw std::runtime_error("rb_gl2_copy_trajectory: initialize() has not been called");
    if (times_out) {
        *times_out = nullptr;
        if (traj) malloc_copy_out(g->times, times_out);
    }
    if (times_count) *times_count = traj ? g->times.size() : 0;
    const size_t count = traj ? g->log_count : 1;
    double* h = (double*)std::malloc(std::max<size_t>(count, 1) * g->n * sizeof(double));
    if (!h) throw std::runtime_error("rb_gl2_copy_trajectory: out of host memory");
    cudaStream_t st = gl2_stream(g);
    if (count) RB_CUDA(cudaMemcpyAsync(h, traj ? g->log : g->y, count * g->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    *states_out = h;
    *states_count = count;
    RB_CATCH
}

double* rb_gl2_dev_state(rb_gl2* g) { return g ? g->y : nullptr; }

int rb_gl2_get_state(rb_gl2* g, double* state_host) {
    RB_TRY
    if (!g || !g->y || !state_host) throw std::runtime_error("rb_gl2_get_state: initialize() has not been called");
    cudaStream_t st = gl2_stream(g);
    RB_CUDA(cudaMemcpyAsync(state_host, g->y, g->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    RB_CATCH
}

int rb_gl2_get_stats(rb_gl2* g, rb_gl2_stats* out) {
    RB_TRY
    if (!g || !out) throw std::runtime_error("rb_gl2_get_stats: null argument");
    *out = g->stats;
    RB_CATCH
}

// ---- legacy exports ---------------------------------------------------------------------------------------------------------------
// L/Export.cu:392-518
int calculateJacobian(const double* state, double* jac, double L, double rho, double kappa, double depth, double epsilon, size_t N) {
    RB_TRY
    if (!state || !jac) throw std::runtime_error("calculateJacobian: null argument");
    rb_props p = helium_props_from_si(L, rho, kappa, depth, false, 1, false);
    std::unique_ptr<rb_jacobian, void (*)(rb_jacobian*)> j(rb_jacobian_create((int)N, &p), jacobian_free);
    if (!j) throw std::runtime_error(rb_last_error());
    j->eps = epsilon;
    const size_t n = 3 * N;
    double* d = dmalloc<double>(n + n * n);
    try {
        cudaStream_t st = jacobian_stream(j.get());
        RB_CUDA(cudaMemcpyAsync(d, state, n * sizeof(double), cudaMemcpyHostToDevice, st));
        jacobian_calculate(j.get(), d, d + n);
        RB_CUDA(cudaMemcpyAsync(jac, d + n, n * n * sizeof(double), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
    RB_CATCH
}

// L/Export.cu:600-663 (the physical arguments are accepted and unused there as well)
int calculatePerturbedStates256(const double* x, const double* y, const double* phi, rb_complex* Zperturbed, double, double, double,
                                double, double epsilon) {
    RB_TRY
    if (!x || !y || !phi || !Zperturbed) throw std::runtime_error("calculatePerturbedStates256: null argument");
    const int N = 256;
    std::vector<double2> host(2 * N);
    for (int i = 0; i < N; ++i) {
        host[i] = make_double2(x[i], y[i]);
        host[N + i] = make_double2(phi[i], 0.0);
    }
    const size_t big = (size_t)6 * N * N;
    double2* d = dmalloc<double2>(2 * N + big);
    try {
        RB_CUDA(cudaMemcpy(d, host.data(), 2 * N * sizeof(double2), cudaMemcpyHostToDevice));
        RB_LAUNCH_EW(perturbed_states_kernel, dim3(blocks_for(N), 3 * N, 1), 256, nullptr, d, d + 2 * N, nullptr, epsilon, N);
        launched();
        RB_CUDA(cudaMemcpy(Zperturbed, d + 2 * N, big * sizeof(double2), cudaMemcpyDeviceToHost));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
    RB_CATCH
}

// L/Export.cu:665-739
int integrateSimulationGL2(double* initialState, double** statesOut, size_t* statesCount, double** timesOut, size_t* timesCount,
                           SimProperties* simProperties, GaussLegendreOptions* glCOptions, size_t N) {
    RB_TRY
    if (!initialState || !statesOut || !statesCount || !simProperties || !glCOptions)
        throw std::runtime_error("integrateSimulationGL2: null argument");
    rb_props p = helium_props_from_si(simProperties->L, simProperties->rho, simProperties->kappa, simProperties->depth,
                                      simProperties->use_expansions, simProperties->expansion_order, simProperties->infinite_depth);
    p.guess_mode = RB_GUESS_WARM;
    p.compute_energies = 0;
    rb_gl2_options o;   // createOptionsFromCOptions, L/GaussLegendre.cuh:92-105
    o.stepSize = glCOptions->stepSize;
    o.newtonTolerance = glCOptions->newtonTolerance;
    o.maxNewtonIterations = glCOptions->maxNewtonIterations;
    o.allowSimplifiedFallback = glCOptions->allowSimplifiedFallback;
    o.returnTrajectory = glCOptions->returnTrajectory;
    o.armijo_c = glCOptions->armijo_c;
    o.backtrack = glCOptions->backtrack;
    o.minAlpha = glCOptions->minAlpha;
    o.maxStepsHalves = glCOptions->maxStepsHalves;
    std::unique_ptr<rb_solver, int (*)(rb_solver*)> s(rb_create((int)N, 1, &p), rb_destroy);
    if (!s) throw std::runtime_error(rb_last_error());
    std::unique_ptr<rb_jacobian, void (*)(rb_jacobian*)> j(rb_jacobian_create((int)N, &p), jacobian_free);
    if (!j) throw std::runtime_error(rb_last_error());
    std::unique_ptr<rb_gl2, void (*)(rb_gl2*)> g(rb_gl2_create(s.get(), j.get(), &o), gl2_free);
    if (!g) throw std::runtime_error(rb_last_error());
    check_rc(rb_gl2_initialize(g.get(), initialState, 0), "rb_gl2_initialize");
    gl2_evolve(g.get(), glCOptions->t0, glCOptions->t1);
    double* times = nullptr;
    size_t tcount = 0;
    check_rc(rb_gl2_copy_trajectory(g.get(), &times, &tcount, statesOut, statesCount), "rb_gl2_copy_trajectory");
    if (timesOut) *timesOut = times; else std::free(times);
    if (timesCount) *timesCount = tcount;
    cudaDeviceSynchronize();   // g, j, s are released in this order by the unique_ptrs; nothing may still be in flight
    RB_CATCH
}

int integrateSimulationGL2_freeMemory(double* statesOut, double* timesOut) {
    std::free(statesOut);
    std::free(timesOut);
    return 0;
}

}  // extern "C"
