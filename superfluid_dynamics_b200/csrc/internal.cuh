// internal.cuh -- shared declarations of libroberts_b200 (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <stdexcept>
#include <string>

namespace rb {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPiHi = 6.283185307179586;            // fl(2 pi)
constexpr double kTwoPiLo = 2.4492935982947064e-16;       // 2 pi - fl(2 pi)

// ---- tiling constants of the pair-interaction (cotangent-sum) kernels -------------------
constexpr int kCell = 256;           // points per cell == targets (rows) per CTA == max source tile
// the tiled sweep kernel is templated on the rows per thread RPT (2 or 4): kCell / RPT threads per CTA
constexpr int kMinCellsForLocal = 4; // below this every tile uses the global exponentials

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        throw std::runtime_error(std::string(what) + " failed at " + file + ":" + std::to_string(line) + ": " +
                                 cudaGetErrorString(e));
    }
}
#define RB_CUDA(x) ::rb::cuda_check((x), #x, __FILE__, __LINE__)

// number of kernels of this library launched since load (cuFFT's own kernels are not counted)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

// per-solve control block living in device memory (read by every sweep CTA)
struct SolveCtrl {
    int done;            // 1: converged (or gave up) -> later sweeps of this solve return immediately
    int iters;           // M*x applications performed
    int final_buf;       // which of the two iterate buffers holds the answer
    int converged;       // 1 if the tolerance was met
    double rel2;         // last max_b ||r||^2/||b||^2
    double prev_rel2;
    unsigned long long max_rel2_bits;   // atomicMax accumulator over batch members (non-negative doubles order as uint64)
    unsigned int members_done;          // level-3 ticket
    int stagnated;       // 1 if the iteration stopped on the round-off floor of the residual (rel <= 1e-10, no longer decreasing)
                         // WITHOUT meeting the tolerance: reported as its own status, never as converged
    double first_rel2;   // residual of the initial iterate (set by the first sweep of the solve): quality of the guess
};

// How a solve ended, decided once per sweep from the worst member's ||r||^2/||b||^2 by whichever thread closes the sweep (the last
// CTA of a sweep kernel, or comm_wait_kernel on a row-sharded run; every rank takes it from the same numbers in the same order).
//   converged: rel <= tolerance, nothing else.
//   stagnated: rel <= 1e-10 and no longer contracting -- the residual shrank by less than a factor 0.8 in the last sweep (the
//              slowest contraction the Neumann iteration has on any surface it converges on is ~0.55 per sweep, trochoid
//              steepness 0.9, which must NOT count: round 1's threshold of 0.5 cut such solves short at 1e-10) -- i.e. the residual
//              sits on its round-off floor, which grows like N eps: the iteration cannot do better; the solve ends with
//              converged = 0, stagnated = 1 and the host counts it separately.
//   neither and iters >= max_iters: the solve failed (done = 1, converged = 0, stagnated = 0).
#ifdef __CUDACC__
// (iters_before / prev: the block's iters and prev_rel2 -- the second form lets a caller load them early, beside its other loads)
__device__ __forceinline__ void solve_decide(volatile SolveCtrl* c, double worst, double tol2, int max_iters, int final_buf, int iters_before,
                                             double prev) {
    if (!(worst == worst)) worst = 1e300;   // NaN -> "not converged"
    const int iters = iters_before + 1;
    const bool conv = worst <= tol2;
    const bool stagnated = !conv && iters >= 3 && worst <= 1e-20 && worst > 0.64 * prev;
    c->iters = iters;
    c->rel2 = worst;
    c->prev_rel2 = worst;
    if (iters == 1) c->first_rel2 = worst;
    c->final_buf = final_buf;
    if (conv || stagnated || iters >= max_iters) {
        c->converged = conv ? 1 : 0;
        c->stagnated = stagnated ? 1 : 0;
        c->done = 1;
    }
}
__device__ __forceinline__ void solve_decide(volatile SolveCtrl* c, double worst, double tol2, int max_iters, int final_buf) {
    const int iters_before = c->iters;
    const double prev = c->prev_rel2;
    solve_decide(c, worst, tol2, max_iters, final_buf, iters_before, prev);
}
#endif

// what the steps of one asynchronously launched chunk of recorded RK4 steps left behind (device memory, folded by the last kernel
// of every recorded step, read by the host once per chunk instead of once per step)
struct StepAgg {
    int steps;             // steps folded in since the host cleared the block
    int not_done;          // steps in which some stage's solve ran out of recorded sweeps
    int failed;            // steps in which some stage ended neither converged nor stagnated
    int stagnated;         // stage solves that ended stagnated (accepted, counted)
    int max_occupied;      // max over the steps and stages of the sweeps a solve occupied in the recorded sequence
    int pad;
    long long sum_iters;   // M*x applications over all solves
    double worst_rel2;     // largest final ||r||^2/||b||^2
    double first_rel2[4];  // of the last step: residual of each stage's initial iterate
    double rel2_last[4];   // of the last step: final residuals
    int iters_last[4];
    int conv_last[4];
    int stag_last[4];
};

// device-driven GMRES cycle (krylov_kernels.cu): per-member least-squares state and the cycle's control block
constexpr int kGmMax = 16;   // Krylov vectors per recorded cycle
struct GmMember {
    double H[(kGmMax + 1) * kGmMax];   // Hessenberg matrix after the rotations (upper triangular part used), H[i * kGmMax + j]
    double cs[kGmMax], sn[kGmMax];     // Givens rotations
    double g[kGmMax + 1];              // rotated right-hand side: |g[k]| is the residual norm after k steps
    double bnorm, rel, first_rel;
    int k_used;                        // columns this member's correction uses
    int running;                       // 0: this member's residual estimate met the tolerance (or broke down)
    unsigned int bar;                  // arrivals at the in-kernel barriers of the CTAs that share this member (monotonic, never reset)
    unsigned int gen;                  // Arnoldi steps this member has run since creation: base of the barrier targets (never reset)
};
struct GmCtrl {
    int done;                          // 1: the cycle has ended (every member finished, or the recorded iterations are used up)
    int k;                             // Arnoldi steps performed in this cycle
    int running;                       // scratch: members still running, counted during a kernel
    unsigned int ticket;               // scratch: member-level ticket
    int k_total;                       // Arnoldi steps performed since the host cleared the block (all cycles of one solve)
    int pad;
    unsigned long long worst_bits;     // scratch: atomicMax accumulator of the members' residuals (non-negative doubles order as uint64)
    double worst_rel;                  // largest true relative residual over the members at the start of the current cycle
};

// geometry of the surface, per point; all arrays are [batch][N]
struct Geometry {
    const double2* Z;      // surface points
    double2* Zp;           // dZ/dj
    double2* Zpp;          // d2Z/dj2
    double2* EG;           // exp(i z)                      global exponentials
    double2* P0;           // expm1(i (z - zc[cell(z)]))    local, own cell centre
    double2* Pm;           // ... relative to the previous cell's centre
    double2* Pp;           // ... relative to the next cell's centre
    double2* EI;           // exp(i (conj z - 2 i h))       image sources (finite-depth helium) or nullptr
    double* Mdiag;         // diagonal of M
    double2* V1diag;       // diagonal of V1 (lower fluid)
    double2* V2;           // i / (2 pi Zp)
    double* b;             // Re(Phi')
};

// solutions of previous steps at one RK stage, kept on the device: slot (counter % ring) is written by the current step
struct HistoryRing {
    double* base = nullptr;      // [ring][stride] or nullptr (no history)
    size_t stride = 0;           // batch * N
    int ring = 0;                // number of slots (> order so that a repeated step never reads what it overwrote)
    int order = 0;               // extrapolation order (1..4)
    const int* counter = nullptr;   // device: steps completed so far
    int store_next = 0;          // finish_solve: record the iterate AFTER the verified one (x + omega r, one contraction better)
    double2* Abase = nullptr;    // [ring][stride] row sums A_k = (sum x - x_k) + 2 T_k of the recorded solutions (S_k = i A_k)
    int predict = 0;             // guess: x0 = extrap(a) + omega (b - M~ extrap(a)),  M~ x = Mdiag x + cK Re(Zp extrap(A)):
                                 // one Richardson sweep whose O(N^2) row sums are replaced by their extrapolation in time
};

// multi-GPU view: every rank owns one "arena" allocation with the same layout; arenas of the peers are mapped into this
// process (CUDA IPC), so a result row is published by storing it at the same offset in every arena over NVLink.
constexpr int kMaxRanks = 8;
struct CommView {
    int nranks = 1;                      // 1: single GPU, plain local stores
    int rank = 0;
    char* my_base = nullptr;             // this rank's arena
    char* peer_base[kMaxRanks] = {};     // every rank's arena as mapped here (peer_base[rank] == my_base)
    size_t off_rn = 0;                   // double rn[2][kMaxRanks]: per-rank residual sums, double-buffered by iterate parity
    size_t off_flags = 0;                // unsigned long long flags[kMaxRanks]: epoch of the last signal received from each rank
    unsigned long long* signal_epoch = nullptr;   // local counters (device memory)
    unsigned long long* wait_epoch = nullptr;
    int* error_flag = nullptr;           // set when a wait timed out
};

// work folded into the kernel that closes a solve (finish_solve): the velocity terms a combined sweep left open because their
// inputs were still being transformed on the side stream, dPhi/dt, and the RK stage / final update of the stepper
struct FinishPost {
    double2* vel = nullptr;              // u + i v rows as the sweep left them (nullptr: nothing to do)
    double2* vel_upper = nullptr;
    double2* dphi = nullptr;             // dPhi/dt + 0 i (nullptr: computed by a separate kernel)
    const double2* aprime0 = nullptr;    // a' of the iterate in buffer 0 / 1 (nullptr: the sweep already added V2 a')
    const double2* aprime1 = nullptr;
    const double2* V2 = nullptr;
    const double2* Z = nullptr;
    int rhs_phi_kind = 0;
    double depth = 0.0;
    int update = 0;                      // 0 none, 1: y_out = y0 + c k, 2: y_out += c (k1 + 2 k2 + 2 k3 + k) with c = h/6
    double c = 0.0;
    const double2* y0 = nullptr;
    double2* y_out = nullptr;
    const double2 *k1 = nullptr, *k2 = nullptr, *k3 = nullptr;
    size_t BN = 0;
};

enum SweepMode { kSweepMV = 0, kSweepVEL = 1, kSweepRAW = 2 };

struct SweepArgs {
    // sizes
    int N, batch, ncell;             // ncell = ceil(N / kCell)
    int rows_per_thread;             // tiled kernel: register blocking (2 or 4 rows per thread, CTA = kCell / rows threads)
    int tile;                        // sources per smem tile (64, 128 or 256; divides kCell)
    int tiles_per_chunk;             // source tiles handled by one CTA
    int nchunks;                     // gridDim.y
    int row_cell0, row_cells;        // this rank's range of row cells (multi-GPU row sharding)
    int use_local;                   // near-field tiles use cell-local exponentials
    int has_image;                   // finite-depth helium image sum
    // inputs
    Geometry g;
    const double* x;                 // MV: iterate in; VEL/RAW: strengths a (real), [batch][N]
    double* x_out;                   // MV: iterate out
    const double* xsum_part;         // [batch][ncell] partial sums of x per cell (producer epilogue)
    double* xsum_part_out;           // MV: partial sums of the new iterate
    double* rnorm_part;              // MV: [batch][ncell] partial sums of r^2
    const double* bnorm_part;        // [batch][ncell] partial sums of b^2
    // workspaces
    double2* partial;                // [batch][nchunks][N] partial T sums
    double2* partial_img;            // same for the image sum
    unsigned int* cell_tickets;      // [batch][ncell]   level-1 counters (zero on entry, reset on exit)
    int chunk_group, ngroups;        // two-level reduction of the chunk partials: chunks per group (0: single level), number of groups
    double2* gpartial;               // [batch][ngroups][N] group sums
    double2* gpartial_img;
    unsigned int* group_tickets;     // [batch][ncell][ngroups] (zero on entry, reset on exit)
    unsigned int* member_tickets;    // [batch]          level-2 counters
    SolveCtrl* ctrl;
    // physics
    double cK;                       // (1-rho)/(4 pi)   (helium: 1/(4 pi))
    double omega;                    // Richardson relaxation 2/(1+rho)
    double rho;
    double depth;
    double tol2;                     // tolerance^2
    int max_iters;
    int rhs_phi_kind;                // VEL epilogue: 0 none, 1 water rho==0, 2 helium vdw, 3 helium expansion, 4 helium + surface tension
    int expansion_order;
    double kappa;
    int apply_only;                  // MV sweeps: x_out = M x (no update, no convergence logic) -- the Krylov solver's operator
    int skip_if_done;                // MV sweeps: return at once when ctrl->done
    int out_buf;                     // MV sweeps: index (0/1) of the iterate buffer x_out lives in
    int combined;                    // VEL sweeps: also evaluate r = b - M x from the same row sums, write x_out = x + omega r and
                                     // take the convergence decision: "verify the iterate and produce its velocities in one pass"
    int final_buf_on_done;           // buffer index recorded in ctrl->final_buf when this sweep declares convergence
    // VEL outputs
    const double2* aprime;           // da/dj (complex, imaginary part kept as the reference does)
    double2* vel_lower;              // u + i v   -> rhs[0 .. BN)
    double2* vel_upper;              // upper-fluid velocities
    double2* dphi;                   // dPhi/dt + 0 i -> rhs[BN .. 2BN)  (nullptr: separate kernel)
    double2* raw_out;                // RAW: S_k = sum_{j!=k} cot((z_k - z_j)/2) x_j
    double2* A_out;                  // combined VEL sweeps: A_k of the input iterate (kept for the time extrapolation of the row sums)
    int defer_aprime;                // VEL sweeps: leave V2 a' and dPhi/dt to finish_solve (a' is transformed beside this sweep)
    CommView comm;                   // row sharding over GPUs (nranks == 1: off)
    // persistent one-wave variant (sweep2_kernel): static schedule of row blocks, in-CTA source split, no global partials
    int v2_RB;                       // rows per row block (multiple of 32 * R)
    int v2_R;                        // rows per thread (1 or 2)
    int v2_groups;                   // source groups per CTA; threads = (RB / R) * groups
    int v2_spg;                      // sources per group per staged tile (power of two >= 32)
    int v2_TS;                       // staged tile = groups * spg sources
    int v2_bpm;                      // row blocks per batch member (of this rank's rows)
    int v2_total_blocks;             // batch * bpm
    int v2_row_begin, v2_row_end;    // this rank's rows [begin, end)
    int v2_split;                    // source parts per row block (work item = row block x part); 1 = no split
    double2* v2_partial;             // [batch][split][rows] row sums per part (split > 1)
    double* v2_xs_part;              // [total_blocks * split] sum of x over each part's sources
    unsigned int* v2_blk_tickets;    // [total_blocks] zero on entry, reset on exit
    double* v2_rnorm_part;           // [total_blocks] residual sums per row block
    unsigned int* v2_ticket;         // one counter, zero on entry, reset on exit
};

struct Sweep2Launch {
    int grid = 0, threads = 0;
    size_t smem = 0;
};

// warp-per-row-group kernel (pair_kernels3.cu): R rows per warp, TS staged sources per tile
struct Sweep3Launch {
    int grid = 0, threads = 0, R = 1, TS = 0, S = 1;   // S warps share a row group (cells split between them)
    bool batched = false;                              // ensembles: one member at a time per CTA (TS = padded sources per member)
    size_t smem = 0;
};

// ---- launch wrappers (each defined in the .cu named in the comment) -----------------------
// pair_kernels.cu
void launch_geometry(const Geometry& g, double2* phiprime_c, int N, int batch, int ncell, int physics,
                     double rhoM, double depth, int finite_image, int use_local, int raw_derivs, double rho, double U,
                     cudaStream_t st);
void launch_sweep(const SweepArgs& a, int mode, cudaStream_t st);
// pair_kernels2.cu
void launch_sweep2(const SweepArgs& a, const Sweep2Launch& l, int mode, cudaStream_t st);
// pair_kernels3.cu
void launch_sweep3(const SweepArgs& a, const Sweep3Launch& l, int mode, cudaStream_t st);
size_t sweep3_smem(int TS);
size_t sweep3b_smem(int NP);
void launch_guess(const double* b, const double* warm, const HistoryRing& hist, double* x0, double* xsum_part,
                  double* bnorm_part, SolveCtrl* ctrl, double omega, int N, int batch, int ncell, cudaStream_t st,
                  const double2* Zp = nullptr, const double* Mdiag = nullptr, double cK = 0.0);
void launch_advance_counter(int* counter, cudaStream_t st);
void launch_step_end(int* counter, const SolveCtrl* ctrl_all, StepAgg* agg, int opt_mask, cudaStream_t st);
void launch_comm_wait(const CommView& c, SolveCtrl* ctrl, int decide, int parity, int final_buf, const double* bnorm_part,
                      int ncell, double tol2, int max_iters, cudaStream_t st);
// spectral.cu
void launch_sub_linear(const double2* Z, const double2* Phi, double2* out_zper, double2* out_phiper, int N, int batch,
                       double rho, double U, cudaStream_t st);
void launch_spectral_multiply_zphi(const double2* hatZ, const double2* hatPhi, double2* out_d1z, double2* out_d2z,
                                   double2* out_d1phi, int N, int batch, cudaStream_t st);
void launch_spectral_multiply(const double2* hat, double2* out, int N, int batch, int second, cudaStream_t st);
void launch_spectral_multiply_real(const double2* half, double2* out, int N, int batch, double scale, cudaStream_t st);
void launch_fft_zphi(const double2* Z, const double2* Phi, double2* Zp, double2* Zpp, double2* PhiP, int N, int logN, int batch,
                     const double2* tw, double rho, double U, cudaStream_t st);
void launch_fft_real_derivative(const double* x, double2* out, int N, int logN, int batch, const double2* tw, double scale,
                                const SolveCtrl* ctrl, cudaStream_t st);
void launch_finish_zphi(double2* Zp, double2* Zpp, double2* PhiP, int N, int batch, double rho, double U, cudaStream_t st);
void launch_scale(double2* v, double s, size_t n, cudaStream_t st);
void launch_finish_solve(const double* buf0, const double* buf1, const SolveCtrl* ctrl, double* a_out, double2* a_complex,
                         double* xsum_part, const HistoryRing& hist, int N, int batch, int ncell, cudaStream_t st,
                         const double2* A0 = nullptr, const double2* A1 = nullptr, const FinishPost* post = nullptr);
void launch_geometry_guess(const Geometry& g, double2* phiprime_c, int N, int batch, int ncell, double rhoM, double depth,
                           int finite_image, int use_local, double rho, double U, const double* warm, const HistoryRing& hist,
                           double* x0, double* xsum_part, double* bnorm_part, SolveCtrl* ctrl, double omega, double cK,
                           cudaStream_t st);
// krylov_kernels.cu
void launch_axpby(double* out, const double* a, double alpha, const double* b, int n, cudaStream_t st);
void launch_precond_scale_half(double2* half, const double* invP, int N, int batch, cudaStream_t st);
void launch_gm_start(const double* b, const double* w, double* V0, GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, int N, int batch,
                     double tol, cudaStream_t st);
void launch_gm_arnoldi(double* V, size_t ldv, double* w, GmMember* members, GmCtrl* gc, SolveCtrl* ctrl, double* part, int N, int batch,
                       int k, int last_k, double tol, cudaStream_t st);
int gm_arnoldi_slices(int N, int batch);   // CTAs per member of the Arnoldi kernel
void launch_gm_correction(const double* V, size_t ldv, double* t, GmMember* members, int N, int batch, cudaStream_t st);
// dense_kernels.cu
void launch_create_M(double* A, const double2* Z, const double2* Zp, const double2* Zpp, double rho, int n, size_t batch,
                     cudaStream_t st);
void launch_create_finite_depth_M(double* A, const double2* Z, const double2* Zp, const double2* Zpp, double h, int n,
                                  size_t batch, bool infinite_depth, cudaStream_t st);
void launch_velocity_matrices(const double2* Z, const double2* Zp, const double2* Zpp, int n, double2* V1, double2* V2,
                              bool lower, size_t batch, bool helium, double h, bool infinite_depth, cudaStream_t st);
void launch_rhs_phi_water(const double2* Z, const double2* V1, const double2* V2, double2* result, double rho, int n,
                          cudaStream_t st);
void launch_rhs_phi_helium(const double2* Z, const double2* V1, double2* result, double h, int n, cudaStream_t st);
void launch_rhs_phi_helium_st(const double2* Z, const double2* Zp, const double2* Zpp, const double2* V1, double2* result,
                              double h, double kappa, int n, cudaStream_t st);
void launch_rhs_phi_helium_exp(const double2* Z, const double2* V1, double2* result, double h, int n, int order,
                               cudaStream_t st);
void launch_energies(const double2* Z, const double2* Zp, const double2* Phi, const double2* vel, double* out5, int N,
                     int physics, double rho, double U, double depth, double kappa, cudaStream_t st);
void launch_lu_solve(double* A, double* b, int n, int* info, cudaStream_t st);            // the solve every caller uses
void launch_lu_solve_unblocked(double* A, double* b, int n, int* info, cudaStream_t st);  // two launches per column
void launch_lu_backsolve(const double* A, double* b, int n, cudaStream_t st);
// lu_kernels.cu
void launch_lu_solve_blocked(double* A, double* b, int n, int* info, cudaStream_t st);    // panels of 32, DMMA trailing update
// stepper_kernels.cu
void launch_stage_update(double2* y_out, const double2* y0, const double2* k, double c, size_t n, cudaStream_t st);
void launch_final_update(double2* y0, const double2* k1, const double2* k2, const double2* k3, const double2* k4, double h,
                         size_t n, cudaStream_t st);
// drive_kernels.cu (rb_opto: include/roberts_b200.h, included before this header by the translation units that use these)
#ifdef ROBERTS_B200_H
void launch_light_intensity(const double2* Z, double* out, const rb_opto& v, size_t n, cudaStream_t st);
void launch_augmented_terms(const double2* state, double2* rhs, const rb_opto& v, size_t BN, cudaStream_t st);
void launch_timed_drive(double2* rhs_phi, const double2* Z, const double2* w, double* delayed, const rb_opto& v, double time,
                        double prev_time, int save, size_t BN, cudaStream_t st);
// solver.cu services for the other host-side translation units (implicit.cu)
int report_error(const std::exception& e);   // records the message for rb_last_error, prints it, returns -1
rb_props helium_props_from_si(double L, double rho, double kappa, double depth, bool use_expansions, int expansion_order,
                              bool infinite_depth);   // adimensionalizeProperties + HeliumBoundaryProblem, L/Export.cu:1222-1246
#endif
// rk45_kernels.cu
void launch_rk45_stage(const double2* y, const double2* const k[5], double2* out, const double c[5], int nk, size_t n,
                       cudaStream_t st);
int rk45_error_blocks(size_t n);
void launch_rk45_error_y5(const double2* y, const double2* k1, const double2* k3, const double2* k4, const double2* k5,
                          const double2* k6, double2* y5_out, double h, double atol, double rtol, double* partial,
                          unsigned int* ticket, double* sumsq, size_t n, cudaStream_t st);
void launch_fp64_peak(double* sink, int iters, int blocks, cudaStream_t st);
void launch_fp64_peak3(double* sink, int iters, int blocks, double seed, cudaStream_t st);
void launch_fp64_mix(double* sink, int iters, int blocks, int nm, int nf, cudaStream_t st);

}  // namespace rb
