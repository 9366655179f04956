// host.cuh -- what the host-side translation units of libroberts_b200 share (not part of the ABI): the solver and stepper objects,
// error plumbing, and the handful of functions that cross file boundaries.
//   solver.cu   the RHS assembler (create / free, derivatives, solves, RHS variants) and its C ABI
//   stepper.cu  the RK4 stepper (recorded steps, rollback, asynchronous chunks) and its C ABI
//   comm.cu     row sharding over the GPUs of one node (IPC arenas), sweep-plan queries
//   probes.cu   measurement entry points (FP64 probes, rb_bench_sweep)
//   exports.cu  the legacy L/Export.cuh names (SI in, nondimensionalised inside)
//   drive.cu    the optomechanically driven film (augmented and explicitly time-dependent steppers)
//   rk45.cu     the adaptive RKF45 stepper
#pragma once
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/roberts_b200.h"
#include "internal.cuh"

using namespace rb;

// ---- error plumbing ------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
int fail(const std::exception& e);   // records the message for rb_last_error, prints it (L/Export.cu: std::cerr << "Error: " ...), returns -1
#define RB_TRY try {
#define RB_CATCH                      \
    }                                 \
    catch (const std::exception& e) { \
        return fail(e);               \
    }                                 \
    return 0;

inline void cufft_check(cufftResult r, const char* what) {
    if (r != CUFFT_SUCCESS) throw std::runtime_error(std::string(what) + " failed: cufft error " + std::to_string((int)r));
}

template <typename T>
inline T* dmalloc(size_t n) {
    T* p = nullptr;
    RB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

// a device allocation that is released on every path out of its scope (the legacy exports allocate temporaries and may throw)
struct DeviceFree {
    void operator()(void* p) const { cudaFree(p); }
};
template <typename T>
using device_ptr = std::unique_ptr<T, DeviceFree>;
template <typename T>
inline device_ptr<T> dmalloc_scoped(size_t n) { return device_ptr<T>(dmalloc<T>(n)); }

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// ---- the RHS assembler ----------------------------------------------------------------------------
struct rb_solver {
    int N = 0, batch = 0, ncell = 0;
    size_t BN = 0;
    rb_props props{};
    cudaStream_t stream = nullptr;       // the stream every kernel of this solver is issued on
    cudaStream_t own_stream = nullptr;   // blocking stream used when the caller hands over the legacy default stream
    int device = 0;

    // derived physics
    double rhoM = 0, cK = 0, omega = 0;
    int has_image = 0, use_local = 0, rhs_phi_kind = 0;
    bool matrix_free_solve = true;

    // chunking of the tiled sweep (pair_kernels.cu)
    int tile = 256, tiles_per_chunk = 1, nchunks = 1;
    int chunk_group = 0;           // two-level reduction of the chunk partials: chunks per group (0: single level)
    int v1_rows = 2;               // tiled kernel: rows per thread (RB_V1_ROWS)
    // schedule of the persistent sweep (pair_kernels2.cu); used whenever there is no image sum
    bool use_v2 = false;
    int v2_RB = 0, v2_R = 0, v2_groups = 0, v2_spg = 0, v2_TS = 0, v2_bpm = 0, v2_total_blocks = 0;
    Sweep2Launch v2l;
    bool use_v3 = false;             // warp-per-row-group kernel (one member, no image sum, small N)
    Sweep3Launch v3l;
    int v2_split = 1;
    double v2_eff = 0.0;
    double* v2_rnorm_part = nullptr;
    unsigned int* v2_ticket = nullptr;
    double2* v2_partial = nullptr;
    double* v2_xs_part = nullptr;
    unsigned int* v2_blk_tickets = nullptr;

    // device buffers
    double2* deriv = nullptr;      // [3][BN]: Zp | Zpp | PhiPrime(complex)
    double2* fwork = nullptr;      // [3][BN]: FFT work (periodic parts / spectra)
    double2 *EG = nullptr, *P0 = nullptr, *Pm = nullptr, *Pp = nullptr, *EI = nullptr, *V1diag = nullptr, *V2 = nullptr;
    double *Mdiag = nullptr, *b = nullptr, *a = nullptr;
    double* xbuf[2] = {nullptr, nullptr};
    double* xsum_part[2] = {nullptr, nullptr};
    double *xsum_a = nullptr, *rnorm_part = nullptr, *bnorm_part = nullptr, *energies = nullptr;
    double2 *ac = nullptr, *aprime = nullptr, *vel_upper = nullptr;
    double2 *partial = nullptr, *partial_img = nullptr, *gpartial = nullptr, *gpartial_img = nullptr;
    unsigned int* group_tickets = nullptr;
    int ngroups = 1;
    unsigned int *cell_tickets = nullptr, *member_tickets = nullptr;
    SolveCtrl* ctrl_all = nullptr;   // [4]: one control block per RK stage (standalone calls use block 0)
    SolveCtrl* ctrl = nullptr;       // the block the next solve uses
    SolveCtrl* h_ctrl = nullptr;     // pinned, [4]
    // arena: one allocation holding everything a peer rank may write (iterates, their per-cell sums, residual slots, flags,
    // the four RK stage slopes); identical layout on every rank
    char* arena = nullptr;
    size_t arena_bytes = 0;
    double2* kbuf[4] = {nullptr, nullptr, nullptr, nullptr};
    double2* Abuf[2] = {nullptr, nullptr};   // row sums A_k of the iterate a combined sweep verified, by iterate-buffer parity (arena)
    // launch-bound regime (N <= 4096): the a' transform of a round runs on a side stream beside that round's combined sweep (a fork
    // and a join inside the recorded step); the sweep leaves V2 a' and dPhi/dt to finish_solve, which also does the RK update
    bool overlap_ok = false;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cufftHandle plan_d2z_side = 0, plan1_side = 0;
    bool side_plans = false;
    double2* aprime2[2] = {nullptr, nullptr};   // a' per iterate-buffer parity
    double2* half_side = nullptr;               // D2Z half spectrum of the side stream
    FinishPost post_update;                     // set by the stepper before a stage: RK update to fold into finish_solve
    bool post_update_done = false;
    CommView comm;
    void* peer_mapped[kMaxRanks] = {};
    unsigned long long* epochs = nullptr;   // [2] signal / wait counters + error flag
    int row_cell0 = 0, row_cells = 0;
    // restarted GMRES for the finite-depth helium operator (device-driven cycles; the host looks once per cycle outside recorded steps)
    bool use_gmres = false;
    int gm_m = 0;                  // restart length
    size_t gm_ld = 0;              // stride of the Krylov vectors (BN rounded up to 32: the real transforms of the preconditioner want
                                   // 16-byte aligned vectors, also for odd N)
    double *gm_V = nullptr, *gm_x = nullptr, *gm_t = nullptr, *gm_dev = nullptr, *gm_invP = nullptr;
    double* gm_host = nullptr;     // pinned
    // the same solver driven from the device inside recorded RK4 steps (krylov_kernels.cu: gm_*_kernel)
    bool gm_device = false;        // recorded steps use the cycle too (RB_DEVICE_GMRES=0: helium steps are not recorded)
    GmMember* gm_members = nullptr;
    double* gm_part = nullptr;     // slice partials of the multi-CTA Arnoldi kernel
    GmCtrl* gm_ctrl = nullptr;     // viewed as a SolveCtrl by the sweeps that skip themselves once the cycle has ended (first member: done)
    double* Mdense = nullptr;      // dense validation path, allocated on demand
    int* lu_info = nullptr;
    double2* scratch_state = nullptr;   // legacy host-vector exports

    cufftHandle plan1 = 0, plan2 = 0, plan3 = 0, plan_d2z = 0, plan_z2d = 0;   // plan_z2d: helium preconditioner only
    // shared-memory FFT derivatives (small power-of-two N, launch-bound regime): twiddle table exp(-2 pi i k / N), k < N/2
    bool own_fft = false;
    bool own_fft_skippable = false;
    int logN = 0;
    double2* fft_tw = nullptr;
    bool plans = false;

    // warm start: stage-history ring attached by the stepper for the next solve (base == nullptr: none)
    HistoryRing hist;
    bool have_prev_a = false;

    // capture mode: fixed sweep count, no host synchronisation inside rb_rhs
    int fixed_sweeps = 0;

    // statistics of the last solve (inside the RK4 stepper: of the last step, aggregated over its four stage solves)
    int last_iters = 0, last_converged = 0, last_stagnated = 0;
    double last_rel = 0;
    // how solves ended since the solver was created: converged = tolerance met; stagnated = stopped on the round-off floor of the
    // residual (<= 1e-10) above the tolerance; failed = neither (iteration cap, NaN, a peer that never signalled)
    long long stagnated_solves = 0, failed_solves = 0;
    double worst_rel = 0;            // largest final relative residual of any solve that was accepted
    bool strict = true;              // a failed solve makes the call return -1 (rb_set_strict(s, 0): report through the statistics only)
    int kpred = 0;
    long long total_sweeps = 0;      // sweep kernels launched (including ones that skipped)
    long long sum_iters = 0;         // M*x applications actually performed, summed over solves
    long long num_solves = 0;
    long long vel_sweeps = 0;        // velocity-only sweeps (the combined verify+velocity sweeps are counted in sum_iters)
    bool combined_ok = true;         // RB_COMBINED=0 disables the combined sweep
    bool optimistic = false;         // recorded steps: the FIRST sweep is already a combined one (guess expected to verify as is)
    bool hist_store_next = true;     // RB_HIST_NEXT=0: history keeps the verified iterate instead of its successor

    const double2* cur_Z = nullptr;
    const double2* cur_Phi = nullptr;
    const double2* cur_vel = nullptr;

    double2* Zp() const { return deriv; }
    double2* Zpp() const { return deriv + BN; }
    double2* PhiPc() const { return deriv + 2 * BN; }
};

// solver.cu
rb_solver* solver_create(int N, int batch, const rb_props* pin);
void solver_free(rb_solver* s);
void choose_chunking(rb_solver* s);
void alloc_partials(rb_solver* s);
void plan_sweep2(rb_solver* s);
void choose_sweep_kernel(rb_solver* s);
void surface_stage(rb_solver* s, const double2* Z, const double2* Phi);
SweepArgs base_args(rb_solver* s, const double2* Z);
void launch_mv(rb_solver* s, const SweepArgs& base, int i, int skip);
void note_solve_end(rb_solver* s, int converged, int stagnated, double rel, int iters, const char* what);
void vorticities(rb_solver* s, const double2* state);
void fft_derivative(rb_solver* s, const double2* in, double2* out, int second, double scaling);
void rhs(rb_solver* s, const double2* state, double2* out);

// ---- the RK4 stepper ------------------------------------------------------------------------------
constexpr int kHistRing = 12;  // > max order (6) + the steps of an asynchronously launched chunk: a chunk that is rolled back and
                               // repeated never reads a slot the failed attempt has overwritten (kChunkMax + order <= kHistRing)
constexpr int kChunkMax = 6;   // recorded steps launched back to back between two host looks at the solve status

struct rb_stepper {
    rb_solver* s = nullptr;
    double dt = 1e-2;
    double t = 0.0;
    double2* y0 = nullptr;
    bool owns_y0 = false;
    double2 *k[4] = {nullptr, nullptr, nullptr, nullptr}, *ytmp = nullptr, *ybackup = nullptr;
    // stage history of the vortex-sheet strengths for the warm start: hist[stage] = [kHistRing][BN]
    double* hist[4] = {nullptr, nullptr, nullptr, nullptr};
    int* d_counter = nullptr;     // device: completed steps since the history was reset
    int h_counter = 0;            // host mirror
    int order = 4;                // extrapolation order (RB_GUESS_ORDER)
    int predict = 0;              // guess = one Richardson sweep whose row sums are extrapolated in time (RB_GUESS_PREDICT; auto: on
                                  // for tolerances above the round-off floor of that extrapolation)
    // CUDA graph of one step (fixed number of self-skipping sweeps per solve)
    bool use_graph = true;
    cudaGraphExec_t graph_exec = nullptr;      // the graph in use (owned by graph_cache)
    cudaGraphExec_t graph_cache[16] = {};      // one recorded step per mask of optimistic stages
    int graph_kernels[16] = {};                // kernels of this library recorded in each (cuFFT's own are not counted)
    int opt_mask = 0;                          // bit i: stage i starts with a combined sweep (its guess verified as it stood lately)
    int graph_mask = 0;                        // mask graph_exec was recorded with
    int opt_policy = 1;                        // RB_OPTIMISTIC: 0 never, 1 adaptive per stage, 2 always
    long long opt_stage_solves = 0, one_sweep_solves = 0;
    double first_rel[4] = {0, 0, 0, 0};        // residual of the initial iterate of each stage in the last step
    int graph_sweeps = 0;
    double graph_dt = 0;
    double2* graph_y0 = nullptr;
    int graph_hits_below = 0;     // consecutive steps that needed far fewer sweeps than captured
    // "tight" recording: after 8 steps in a row that all needed the same number of sweeps, record exactly that many (no surplus,
    // self-skipping round per solve: at N <= 4096 such a round -- a skipped sweep, its a' transform, the fork / join around it -- is
    // ~8 % of a step); a step that then runs out of sweeps is rolled back and redone as always, and tight recording is banned for a while
    bool tight_ok = true;
    bool tight = false;
    int tight_hits = 0, tight_ban = 0, tight_max = 0;
    long long tight_failures = 0;
    long long graph_launches = 0, graph_captures = 0, fallback_steps = 0;
    cudaEvent_t ev = nullptr;
    // asynchronous chunks: several recorded steps launched back to back, one host synchronisation per chunk (launch-bound regime)
    StepAgg* d_agg = nullptr;
    StepAgg* h_agg = nullptr;     // pinned
    double2* ycheck = nullptr;    // state at the start of the chunk (rollback)
    int chunk = 0;                // steps per chunk (0: off)
    long long chunks_launched = 0, chunks_rolled_back = 0;
    // logging
    size_t log_every = 0, log_capacity = 0, log_count = 0, step_index = 0;
    double2* log_states = nullptr;
    std::vector<double> log_times;
};

// stepper.cu
void stepper_free(rb_stepper* st);
void stepper_step(rb_stepper* st);
void stepper_run(rb_stepper* st, size_t n);

// ---- nondimensionalisation of the legacy exports (exports.cu) -----------------------------------------
constexpr double kAlphaHamaker = 3.5e-24;    // L/constants.cuh:11
constexpr double kHbar = 1.054571817e-34;    // L/constants.cuh:10
struct Adim {
    double base_length, base_acceleration, base_time, base_energy, kappa, depth, rho;
};
Adim adimensionalize(double L, double rho, double kappa, double depth, double rhoHelium = 150.0);
rb_props helium_props(const Adim& ad, bool use_expansions, int expansion_order, bool infinite_depth);
