// solver.cu -- host orchestration of one RHS evaluation, the RK4 stepper and the C ABI (include/roberts_b200.h).
//
// Mirrors (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   BaseBoundaryIntegralCalculator<N,B>::runTimeStep / calculateVorticities   L/BaseBoundaryIntegrator.cuh:138-306
//   ZPhiDerivative<N,B>::exec, FftDerivative<N,B>::exec                        L/Derivatives.cuh:190-257, 311-384
//   AutonomousRungeKuttaStepperBase::runStep / initialize / runEvolution       L/AutonomousRungeKuttaStepper.cuh:124-437
//   calculateRHSNFromVectors, adimensionalizeProperties                        L/Export.cu:194-265, 1213-1246
// N and the batch size are runtime values.  No CPU fallback: everything below needs a CUDA device.
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/roberts_b200.h"
#include "internal.cuh"

using namespace rb;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
namespace rb {
unsigned long long g_launch_count = 0;
}

static int fail(const std::exception& e) {
    g_last_error = e.what();
    std::fprintf(stderr, "Error: %s\n", e.what());   // L/Export.cu: std::cerr << "Error: " << e.what()
    return -1;
}
#define RB_TRY try {
#define RB_CATCH                      \
    }                                 \
    catch (const std::exception& e) { \
        return fail(e);               \
    }                                 \
    return 0;

static void cufft_check(cufftResult r, const char* what) {
    if (r != CUFFT_SUCCESS) throw std::runtime_error(std::string(what) + " failed: cufft error " + std::to_string((int)r));
}

template <typename T>
static T* dmalloc(size_t n) {
    T* p = nullptr;
    RB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// the RHS assembler
// ------------------------------------------------------------------------------------------------
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
    // restarted GMRES for the finite-depth helium operator (host-driven, one synchronisation per iteration)
    bool use_gmres = false;
    int gm_m = 0;                  // restart length
    size_t gm_ld = 0;              // stride of the Krylov vectors (BN rounded up to 32: the real transforms of the preconditioner want
                                   // 16-byte aligned vectors, also for odd N)
    double *gm_V = nullptr, *gm_x = nullptr, *gm_t = nullptr, *gm_dev = nullptr, *gm_invP = nullptr;
    double* gm_host = nullptr;     // pinned
    // the same solver driven from the device inside recorded RK4 steps (krylov_kernels.cu: gm_*_kernel)
    bool gm_device = false;
    GmMember* gm_members = nullptr;
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

static void solver_free(rb_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->plans) {
        cufftDestroy(s->plan1);
        cufftDestroy(s->plan2);
        cufftDestroy(s->plan3);
        cufftDestroy(s->plan_d2z);
        if (s->plan_z2d) cufftDestroy(s->plan_z2d);
    }
    for (int r = 0; r < kMaxRanks; ++r)
        if (s->peer_mapped[r]) cudaIpcCloseMemHandle(s->peer_mapped[r]);
    void* ptrs[] = {s->deriv, s->fwork, s->EG, s->P0, s->Pm, s->Pp, s->EI, s->V1diag, s->V2, s->Mdiag, s->b, s->a,
                    s->arena, s->epochs, s->xsum_a, s->rnorm_part, s->bnorm_part,
                    s->energies, s->ac, s->aprime, s->vel_upper, s->partial, s->partial_img, s->gpartial, s->gpartial_img, s->group_tickets, s->cell_tickets,
                    s->member_tickets, s->ctrl_all, s->Mdense, s->lu_info, s->scratch_state, s->gm_V, s->gm_x, s->gm_t,
                    s->gm_dev, s->gm_invP, s->gm_members, s->gm_ctrl, s->v2_rnorm_part, s->v2_ticket, s->fft_tw, s->v2_partial, s->v2_xs_part,
                    s->v2_blk_tickets};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
    if (s->side_plans) {
        cufftDestroy(s->plan_d2z_side);
        cufftDestroy(s->plan1_side);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    for (auto p : {s->aprime2[0], s->aprime2[1], s->half_side})
        if (p) cudaFree(p);
    if (s->gm_host) cudaFreeHost(s->gm_host);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

// Source chunking of the tiled sweep.  A CTA is one 256-row cell x one chunk of sources; the grid is (row cells, chunks, members).
// Measured on a B200 (solver sweep alone, us; profiles/r02a_shardtune_*.log, r02b_*), N = 65536, 4 rows per thread, per-rank share
// of a G-rank run timed on one GPU with rb_debug_set_row_range:
//   G = 1 (256 row cells): chunks 4 / 8 / 18 / 37 / 64 -> 3450 / 2995 / 2976 / 2887 / 2876          (ideal at the 1-GPU rate: 2876)
//   G = 2 (128): 9 / 18 / 37 / 128 -> 1600 / 1574 / 1516 / 1479                                      (1438)
//   G = 4 (64):  9 / 18 / 37 / 74 / 256 -> 884 / 855 / 774 / 736 / 751                               (719)
//   G = 8 (32):  9 / 18 / 37 / 74 / 148 / 512 -> 605 / 502 / 431 / 404 / 390 / 451                   (359)
// i.e. one balanced wave (32 x 37 = 1184 CTAs) is NOT the optimum: several waves of CTAs with 256-1024 sources each balance better,
// until the serial reduction of the chunk partials by the finishing CTA of each row cell (one batch of loads per 4-8 chunks, at the
// tail of the launch where nothing overlaps it) eats the gain -- which is what the two-level reduction in sweep_kernel removes
// (groups of ~sqrt(nchunks) chunks are reduced as they complete).  Fewer, longer chunks also cost accuracy: a chunk's sum is one
// serial accumulation, and at N = 65536 with 8 chunks the residual's round-off floor rises enough to need a third sweep per solve.
static void choose_chunking(rb_solver* s) {
    const int N = s->N;
    // round 1, whole surfaces: N = 65536: 2 rows per thread 3200 (target 1184) / 3108 (2368); 4 rows per thread 3391 / 3011 /
    // 2931 (4736) / 2893 (9472) / 2875 (18944); 8 rows per thread 3556 at best -- the persistent kernel: 3113;
    // N = 32768: 832 / 813 | 912 / 830 / 819; N = 16384: 248 / 234 | 314 / 265 / 235; N = 8192: 80.5 / 74.2 | 102 / 84; N = 4096: 32.8 / 32.9 | 41
    // round 2, with the two-level reduction (profiles/r02b_shardtune.log), chunks -> us:
    //   N = 65536: G = 1: 37 / 64 / 74 / 128 -> 2896 / 2873 / 2875 / 2846;  G = 2: 37 / 74 / 128 / 256 -> 1523 / 1483 / 1462 / 1452;
    //              G = 4: 74 / 148 / 256 / 512 -> 736 / 731 / 728 / 737;    G = 8: 74 / 148 / 256 / 512 -> 386 / 379 / 376 / 381
    //              (the same 8-GPU shard with the single-level reduction: 148 -> 400, 512 -> 488)
    //   N = 16384: G = 1: 32 / 64 / 128 -> 241 / 225 / 219, 4 rows per thread 64 / 128 -> 226 / 213;  G = 2: 32 / 64 / 128 -> 144 / 129 / 124;
    //              G = 4: 64 / 128 / 256 -> 78 / 72 / 73;  G = 8: 64 / 128 / 256 -> 57 / 49 / 47
    //   N = 4096:  16 / 32 / 64 -> 46 / 37 / 35 (single level at 64: 41)
    const bool big = N >= 49152;
    const bool mid = N >= 16384;
    // (with the image sum 4 rows per thread spill at 128 registers: helium N = 16384 sweep 416 us against 404 with 2 rows)
    s->v1_rows = env_int("RB_V1_ROWS", (big || (mid && !s->has_image)) ? 4 : 2) == 4 ? 4 : 2;
    const int target = env_int("RB_TARGET_CTAS", big ? 32768 : (mid ? 148 * 64 : (N >= 2048 ? 148 * 16 : 148 * 8)));
    const long rows = (long)(s->row_cells > 0 ? s->row_cells : s->ncell) * s->batch;
    int wanted = (int)std::max<long>(1, (target + rows - 1) / rows);
    // never below `min_srcs` sources per CTA: the fixed cost of a CTA (prologue, partial store, tickets) is ~3 us ~ 16 sources' worth
    const int min_srcs = env_int("RB_MIN_SRCS", big ? 256 : (mid ? 128 : 64));
    const int max_chunks = std::max(1, (N + min_srcs - 1) / min_srcs);
    wanted = std::min(wanted, max_chunks);
    wanted = env_int("RB_NCHUNKS", wanted);
    int srcs = (N + wanted - 1) / wanted;
    srcs = ((srcs + 63) / 64) * 64;
    // largest tile (fewest barriers) that divides the chunk; tiles never straddle a 256-point cell
    if (srcs % 256 == 0) s->tile = 256;
    else if (srcs % 128 == 0) s->tile = 128;
    else s->tile = 64;
    s->tiles_per_chunk = srcs / s->tile;
    int t = env_int("RB_TILE", 0);
    if (t == 64 || t == 128 || t == 256) {
        s->tile = t;
        s->tiles_per_chunk = std::max(1, (srcs + t - 1) / t);
    }
    int per = s->tile * s->tiles_per_chunk;
    s->nchunks = (N + per - 1) / per;
    // two-level reduction of the chunk partials: groups of ~sqrt(nchunks) chunks (a multiple of 4 = one batch of loads)
    int grp = 1;
    if (s->nchunks > 16) {
        grp = 4;
        while (grp * grp < s->nchunks) grp += 4;
    }
    grp = env_int("RB_CHUNK_GROUP", grp);
    s->chunk_group = grp >= 2 && grp < s->nchunks ? grp : 0;   // 0: single level
    if (env_int("RB_VERBOSE", 0))
        std::fprintf(stderr, "[roberts_b200] tiled sweep plan: N=%d rows/thread=%d row cells=%ld tile=%d tiles/chunk=%d nchunks=%d group=%d -> %ld CTAs\n",
                     N, s->v1_rows, rows, s->tile, s->tiles_per_chunk, s->nchunks, s->chunk_group, rows * s->nchunks);
}

static void set_stream(rb_solver* s, cudaStream_t st);

// workspaces of the tiled sweep that depend on the chunking (re-made whenever the plan changes)
static void alloc_partials(rb_solver* s) {
    for (void* p : {(void*)s->partial, (void*)s->partial_img, (void*)s->gpartial, (void*)s->gpartial_img, (void*)s->group_tickets})
        if (p) cudaFree(p);
    s->partial = s->partial_img = s->gpartial = s->gpartial_img = nullptr;
    s->group_tickets = nullptr;
    s->ngroups = s->chunk_group > 0 ? (s->nchunks + s->chunk_group - 1) / s->chunk_group : 1;
    s->partial = dmalloc<double2>((size_t)s->nchunks * s->BN);
    if (s->has_image) s->partial_img = dmalloc<double2>((size_t)s->nchunks * s->BN);
    if (s->chunk_group > 0) {
        s->gpartial = dmalloc<double2>((size_t)s->ngroups * s->BN);
        if (s->has_image) s->gpartial_img = dmalloc<double2>((size_t)s->ngroups * s->BN);
        const size_t nt = (size_t)s->batch * s->ncell * s->ngroups;
        s->group_tickets = dmalloc<unsigned int>(nt);
        RB_CUDA(cudaMemset(s->group_tickets, 0, nt * sizeof(unsigned int)));
    }
}

// static schedule of the persistent sweep: row blocks of RB rows, (RB/R) x groups threads, staged tiles of groups*spg sources
static void plan_sweep2(rb_solver* s) {
    int nSM = 148;
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, s->device);
    const int N = s->N, B = s->batch;
    const int row_begin = s->row_cell0 * kCell;
    const int row_end = std::min(N, (s->row_cell0 + s->row_cells) * kCell);
    const int rows = row_end - row_begin;
    const long units32 = (long)B * ((rows + 31) / 32);
    int R, RB;
    if (units32 <= 2L * nSM) {
        R = 1;
        RB = 32;
    } else if (B > 1 && rows >= 128 && env_int("RB_V2_R4", 1)) {
        // ensembles: whole members (or 128-row multiples of them) per row block, 4 rows per thread -- the register blocking that
        // carries the tiled kernel to 0.8 of the FP64 peak; round 1 ran them at 2 rows per thread and 0.39 of the peak
        R = 4;
        const long units128 = (long)B * ((rows + 127) / 128);
        long m = (units128 + nSM - 1) / nSM;
        m = std::max(1L, std::min(4L, m));
        m = std::min<long>(m, (rows + 127) / 128);
        RB = 128 * (int)m;
    } else {
        R = 2;
        const long units64 = (long)B * ((rows + 63) / 64);
        long m = (units64 + nSM - 1) / nSM;
        m = std::max(1L, std::min(8L, m));
        m = std::min<long>(m, (rows + 63) / 64);
        RB = 64 * (int)m;
    }
    RB = env_int("RB_V2_RB", RB);
    R = env_int("RB_V2_R", R);
    const int nrt = RB / R;
    const int max_threads = R == 4 ? (env_int("RB_V2_R4_THREADS", 512) <= 256 ? 256 : 512) : (R == 2 ? 896 : 1024);   // launch bounds of sweep2_kernel<., R>
    int G = 1;
    while (nrt * G * 2 <= max_threads && G * 2 <= 32 && N / (G * 2) >= 32) G *= 2;
    G = env_int("RB_V2_GROUPS", G);
    const int threads = nrt * G;
    int n_pow2 = 32;
    while (n_pow2 < N) n_pow2 <<= 1;
    const int ts_max = std::min(std::min(threads, 1024), std::max(G * 32, n_pow2));   // one staged entry per thread and tile
    int spg = 256;
    while (spg > 32 && G * spg > ts_max) spg >>= 1;
    s->v2_RB = RB;
    s->v2_R = R;
    s->v2_groups = G;
    s->v2_spg = spg;
    s->v2_TS = G * spg;
    s->v2_bpm = (rows + RB - 1) / RB;
    s->v2_total_blocks = B * s->v2_bpm;
    // source split: work items = row blocks x parts; the smallest split that keeps >= 95 % of the SMs busy in every round
    {
        const int NT = (N + s->v2_TS - 1) / s->v2_TS;
        int best = 1;
        double best_eff = 0.0;
        for (int sx = 1; sx <= 16 && sx <= NT; sx <<= 1) {
            const long items = (long)s->v2_total_blocks * sx;
            const long grid = std::min<long>(items, nSM);
            const long per = (items + grid - 1) / grid;
            const double eff = (double)items / ((double)per * nSM);
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best = sx;
            }
            if (eff >= 0.95) break;
        }
        // measured (B200): with the split the per-item overheads (un-overlapped first prefetch, reductions, fences, tickets) outweigh
        // the better balance -- N = 8192/16384/32768: 122/287/968 us against 82/250/891 us for the tiled kernel -- so it stays
        // off unless asked for; mid-size problems and multi-GPU shards use the tiled kernel
        if (!env_int("RB_V2_SPLIT_AUTO", 0)) best = 1;
        s->v2_split = std::max(1, std::min(env_int("RB_V2_SPLIT", best), NT));
        const long items = (long)s->v2_total_blocks * s->v2_split;
        const long grid = std::min<long>(items, nSM);
        s->v2_eff = (double)items / ((double)((items + grid - 1) / grid) * nSM);
    }
    // (R = 4 with 256 threads: two CTAs per SM, so that the staging / closing phases of one row block overlap the pair loop of another)
    s->v2l.grid = (int)std::min<long>((long)s->v2_total_blocks * s->v2_split, env_int("RB_V2_GRID", (R == 4 && threads <= 256) ? 2 * nSM : nSM));
    s->v2l.threads = threads;
    s->v2l.smem = (size_t)s->v2_TS * 32 * (s->use_local ? 2 : 1) + (size_t)threads * R * 16 + (size_t)threads * 8 +
                  (size_t)s->v2_TS * 8 + 16;   // + one padding entry behind g: the far loop loads one source ahead
    if (s->v2_rnorm_part) cudaFree(s->v2_rnorm_part);
    s->v2_rnorm_part = dmalloc<double>(s->v2_total_blocks);
    if (s->v2_partial) cudaFree(s->v2_partial);
    if (s->v2_xs_part) cudaFree(s->v2_xs_part);
    if (s->v2_blk_tickets) cudaFree(s->v2_blk_tickets);
    s->v2_partial = dmalloc<double2>((size_t)B * s->v2_split * std::max(rows, 1));
    s->v2_xs_part = dmalloc<double>((size_t)s->v2_total_blocks * s->v2_split);
    s->v2_blk_tickets = dmalloc<unsigned int>(s->v2_total_blocks);
    RB_CUDA(cudaMemset(s->v2_blk_tickets, 0, (size_t)s->v2_total_blocks * sizeof(unsigned int)));
    if (!s->v2_ticket) {
        s->v2_ticket = dmalloc<unsigned int>(1);
        RB_CUDA(cudaMemset(s->v2_ticket, 0, sizeof(unsigned int)));
    }
    if (threads > max_threads || threads % 32 || s->v2_TS > threads || RB % (32 * R))
        throw std::runtime_error("plan_sweep2: inconsistent schedule");
    if (env_int("RB_VERBOSE", 0))
        std::fprintf(stderr, "[roberts_b200] sweep2 plan: N=%d B=%d rows=[%d,%d) RB=%d R=%d groups=%d spg=%d TS=%d blocks=%d split=%d grid=%d threads=%d smem=%zu eff=%.3f\n",
                     N, B, row_begin, row_end, RB, R, G, spg, s->v2_TS, s->v2_total_blocks, s->v2_split, s->v2l.grid, threads,
                     s->v2l.smem, s->v2_eff);
}

// persistent kernel when its static schedule keeps (nearly) every SM busy or the problem is small; otherwise the tiled kernel,
// whose source chunking balances mid-size problems better
static void choose_sweep_kernel(rb_solver* s) {
    int nSM = 148;
    cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, s->device);
    const double eff = s->v2_eff;
    // measured on a B200 (solver sweep alone, us; tiled / persistent): N=256 18.5/12.3, 1024 18.5/18.5, 2048 21.8/25.2, 4096 32.8/37.3,
    // 8192 74/96, 16384 234/287, 32768 813/974, 65536 2875 (4 rows per thread)/3113; recorded RK4 steps per second, tiled / persistent:
    // N=1024 4074/4073, 2048 3678/3387, 4096 2546/2465, 8192 1126/902.
    // Ensembles (batch > 1) keep the persistent kernel whenever its schedule fills the SMs.
    bool v2 = !s->has_image && ((eff >= 0.95 && s->batch > 1) || (long)s->N * s->batch <= 1024);
    int force = env_int("RB_SWEEP_V2", -1);
    if (force >= 0) v2 = !s->has_image && force != 0;
    s->use_v2 = v2;
}

static void sweep(rb_solver* s, const SweepArgs& a, int mode) {
    if (s->use_v2) launch_sweep2(a, s->v2l, mode, s->stream);
    else launch_sweep(a, mode, s->stream);
    s->total_sweeps++;
}

static rb_solver* solver_create(int N, int batch, const rb_props* pin) {
    if (N < 2) throw std::runtime_error("rb_create: N must be >= 2");
    if (batch < 1) throw std::runtime_error("rb_create: batch must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("rb_create: no CUDA device (libroberts_b200 has no CPU path)");
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> up(new rb_solver, solver_free);
    rb_solver* s = up.get();
    RB_CUDA(cudaGetDevice(&s->device));
    s->N = N;
    s->batch = batch;
    s->BN = (size_t)N * batch;
    s->ncell = (N + kCell - 1) / kCell;
    if (pin) s->props = *pin; else rb_default_props(&s->props);
    rb_props& p = s->props;
    if (p.max_iterations <= 0) p.max_iterations = 200;
    if (p.tolerance <= 0) p.tolerance = 1e-13;

    switch (p.physics) {
        case RB_WATER:
            s->rhoM = p.rho;
            s->has_image = 0;
            s->rhs_phi_kind = (p.rho == 0.0) ? 1 : 0;
            break;
        case RB_HELIUM:
            s->rhoM = 0.0;   // createFiniteDepthMKernel: 1/2 and 1/(4 pi), L/createM.cuh:79,85
            s->has_image = p.infinite_depth ? 0 : 1;
            s->rhs_phi_kind = (p.use_expansions || p.kappa != 0.0) ? 0 : 2;
            break;
        case RB_HELIUM_INF:
            s->rhoM = p.depth;   // depth passed in the rho slot, L/HeliumBoundaryProblem.cuh:61
            s->has_image = 0;
            s->rhs_phi_kind = 2;
            break;
        default:
            throw std::runtime_error("rb_create: unknown physics");
    }
    s->cK = 0.25 * (1.0 - s->rhoM) / kPi;
    s->omega = 2.0 / (1.0 + s->rhoM);
    s->use_local = (s->ncell >= kMinCellsForLocal && !env_int("RB_NO_LOCAL", 0)) ? 1 : 0;
    // Richardson on (1/2) I + K converges for the water / infinite-depth operators; the finite-depth image term of the
    // reference (no Zp_k factor, L/createM.cuh:87-88) spreads the spectrum of M between 1/2 and N/(4 pi): that case is solved
    // by restarted GMRES on the same matrix-free operator, right-preconditioned with the flat-film symbol of the image term.
    s->matrix_free_solve = (p.solve_mode == RB_SOLVE_MATRIX_FREE);
    s->use_gmres = s->matrix_free_solve && s->has_image;
    choose_chunking(s);

    const size_t BN = s->BN;
    s->deriv = dmalloc<double2>(3 * BN);
    s->fwork = dmalloc<double2>(3 * BN);
    s->EG = dmalloc<double2>(BN);
    s->P0 = dmalloc<double2>(BN);
    s->Pm = dmalloc<double2>(BN);
    s->Pp = dmalloc<double2>(BN);
    if (s->has_image) s->EI = dmalloc<double2>(BN);
    s->V1diag = dmalloc<double2>(BN);
    s->V2 = dmalloc<double2>(BN);
    s->Mdiag = dmalloc<double>(BN);
    s->b = dmalloc<double>(BN);
    s->a = dmalloc<double>(BN);
    size_t pc = (size_t)s->ncell * batch;
    {
        auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
        size_t off = 0, ox[2], oxs[2], ok[4], oA[2];
        for (int i = 0; i < 2; ++i) { ox[i] = off; off += up256(BN * sizeof(double)); }
        for (int i = 0; i < 2; ++i) { oxs[i] = off; off += up256(pc * sizeof(double)); }
        s->comm.off_rn = off; off += up256(2 * kMaxRanks * sizeof(double));
        s->comm.off_flags = off; off += up256(kMaxRanks * sizeof(unsigned long long));
        for (int i = 0; i < 4; ++i) { ok[i] = off; off += up256(2 * BN * sizeof(double2)); }
        for (int i = 0; i < 2; ++i) { oA[i] = off; off += up256(BN * sizeof(double2)); }
        s->arena_bytes = off;
        s->arena = dmalloc<char>(off);
        RB_CUDA(cudaMemset(s->arena, 0, off));
        for (int i = 0; i < 2; ++i) {
            s->xbuf[i] = reinterpret_cast<double*>(s->arena + ox[i]);
            s->xsum_part[i] = reinterpret_cast<double*>(s->arena + oxs[i]);
        }
        for (int i = 0; i < 4; ++i) s->kbuf[i] = reinterpret_cast<double2*>(s->arena + ok[i]);
        for (int i = 0; i < 2; ++i) s->Abuf[i] = reinterpret_cast<double2*>(s->arena + oA[i]);
        s->epochs = dmalloc<unsigned long long>(4);
        RB_CUDA(cudaMemset(s->epochs, 0, 4 * sizeof(unsigned long long)));
        s->comm.nranks = 1;
        s->comm.rank = 0;
        s->comm.my_base = s->arena;
        s->comm.peer_base[0] = s->arena;
        s->comm.signal_epoch = s->epochs;
        s->comm.wait_epoch = s->epochs + 1;
        s->comm.error_flag = reinterpret_cast<int*>(s->epochs + 2);
        s->row_cell0 = 0;
        s->row_cells = s->ncell;
    }
    s->xsum_a = dmalloc<double>(pc);
    s->rnorm_part = dmalloc<double>(pc);
    s->bnorm_part = dmalloc<double>(pc);
    s->energies = dmalloc<double>(8);
    s->ac = dmalloc<double2>(BN);
    s->aprime = dmalloc<double2>(BN);
    s->vel_upper = dmalloc<double2>(BN);
    alloc_partials(s);
    s->cell_tickets = dmalloc<unsigned int>(pc);
    s->member_tickets = dmalloc<unsigned int>(batch);
    s->ctrl_all = dmalloc<SolveCtrl>(4);
    s->ctrl = s->ctrl_all;
    s->lu_info = dmalloc<int>(batch);   // one per member (dense validation path)
    RB_CUDA(cudaMemset(s->cell_tickets, 0, pc * sizeof(unsigned int)));
    RB_CUDA(cudaMemset(s->member_tickets, 0, batch * sizeof(unsigned int)));
    RB_CUDA(cudaMemset(s->ctrl_all, 0, 4 * sizeof(SolveCtrl)));
    RB_CUDA(cudaMemset(s->a, 0, BN * sizeof(double)));
    RB_CUDA(cudaMemset(s->energies, 0, 8 * sizeof(double)));
    RB_CUDA(cudaMallocHost(&s->h_ctrl, 4 * sizeof(SolveCtrl)));
    std::memset(s->h_ctrl, 0, 4 * sizeof(SolveCtrl));

    if (s->use_gmres) {
        s->gm_m = std::max(2, std::min(env_int("RB_GMRES_RESTART", 60), N));
        s->gm_ld = (BN + 31) / 32 * 32;
        s->gm_V = dmalloc<double>((size_t)(s->gm_m + 1) * s->gm_ld);
        s->gm_x = dmalloc<double>(BN);
        s->gm_t = dmalloc<double>(BN);
        s->gm_dev = dmalloc<double>(4 * (s->gm_m + 4));
        RB_CUDA(cudaMallocHost(&s->gm_host, 4 * (s->gm_m + 4) * sizeof(double)));
        s->gm_invP = dmalloc<double>(N);
        // flat-film symbol of M: 1/2 + (N/4pi) (e^{-2Hm} + e^{-2H(N-m)}) / (1 - e^{-2HN}),  H = depth
        std::vector<double> invP(N, 2.0);
        if (env_int("RB_HELIUM_PRECOND", 1)) {
            const double H = p.depth;
            for (int m = 0; m < N; ++m) {
                double sym = (std::exp(-2.0 * H * m) + std::exp(-2.0 * H * (N - m))) / (1.0 - std::exp(-2.0 * H * N));
                invP[m] = 1.0 / (0.5 + N / (4.0 * kPi) * sym);
            }
        } else {
            std::fill(invP.begin(), invP.end(), 1.0);
        }
        RB_CUDA(cudaMemcpy(s->gm_invP, invP.data(), N * sizeof(double), cudaMemcpyHostToDevice));
        s->gm_device = env_int("RB_DEVICE_GMRES", 1) != 0;
        s->gm_members = dmalloc<GmMember>(batch);
        RB_CUDA(cudaMemset(s->gm_members, 0, (size_t)batch * sizeof(GmMember)));
        static_assert(sizeof(SolveCtrl) >= sizeof(GmCtrl) && offsetof(SolveCtrl, done) == offsetof(GmCtrl, done),
                      "the sweeps read GmCtrl::done through a SolveCtrl pointer");
        s->gm_ctrl = reinterpret_cast<GmCtrl*>(dmalloc<SolveCtrl>(1));
        RB_CUDA(cudaMemset(s->gm_ctrl, 0, sizeof(SolveCtrl)));
    }

    int n[1] = {N};
    cufft_check(cufftPlanMany(&s->plan1, 1, n, nullptr, 1, N, nullptr, 1, N, CUFFT_Z2Z, batch), "cufftPlanMany(B)");
    cufft_check(cufftPlanMany(&s->plan2, 1, n, nullptr, 1, N, nullptr, 1, N, CUFFT_Z2Z, 2 * batch), "cufftPlanMany(2B)");
    cufft_check(cufftPlanMany(&s->plan3, 1, n, nullptr, 1, N, nullptr, 1, N, CUFFT_Z2Z, 3 * batch), "cufftPlanMany(3B)");
    cufft_check(cufftPlanMany(&s->plan_d2z, 1, n, nullptr, 1, N, nullptr, 1, N / 2 + 1, CUFFT_D2Z, batch), "cufftPlanMany(D2Z)");
    if (s->use_gmres)
        cufft_check(cufftPlanMany(&s->plan_z2d, 1, n, nullptr, 1, N / 2 + 1, nullptr, 1, N, CUFFT_Z2D, batch), "cufftPlanMany(Z2D)");
    s->plans = true;
    // the one-CTA radix-2 transform is shared-memory-bandwidth bound (~1.3k cycles per pass at N = 4096): it beats the library's
    // three launches only in the launch-bound regime (measured: faster at N <= 1024, slower at N = 4096)
    const int own_fft_max = env_int("RB_OWN_FFT_MAX", 1024);
    if ((N & (N - 1)) == 0 && N >= 4 && N <= 4096 && (long)N * batch <= 4096 && env_int("RB_OWN_FFT", 1)) {
        // up to own_fft_max the fused kernels replace the library everywhere; above it they are only used for the a' of the
        // surplus (normally skipped) rounds of a recorded step, because they can skip themselves and the library cannot
        s->own_fft = N <= own_fft_max;
        s->own_fft_skippable = true;
        while ((1 << s->logN) < N) s->logN++;
        // per-pass tables: exp(-i pi k / Ns), k < Ns, at offset Ns - 1 (see stage_twiddles in spectral.cu)
        std::vector<double2> tw(N);
        for (int Ns = 1; Ns < N; Ns <<= 1)
            for (int k = 0; k < Ns; ++k) {
                double ang = -kPi * (double)k / (double)Ns;
                tw[Ns - 1 + k] = make_double2(std::cos(ang), std::sin(ang));
            }
        tw[N - 1] = make_double2(0.0, 0.0);
        s->fft_tw = dmalloc<double2>(N);
        RB_CUDA(cudaMemcpy(s->fft_tw, tw.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
    }
    if (s->own_fft_skippable && env_int("RB_OVERLAP", 1)) {
        // every surplus a' of a recorded step can skip itself here, so a' buffers by parity are never overwritten after the solve ended
        s->overlap_ok = true;
        RB_CUDA(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        cufft_check(cufftPlanMany(&s->plan1_side, 1, n, nullptr, 1, N, nullptr, 1, N, CUFFT_Z2Z, batch), "cufftPlanMany(side)");
        cufft_check(cufftPlanMany(&s->plan_d2z_side, 1, n, nullptr, 1, N, nullptr, 1, N / 2 + 1, CUFFT_D2Z, batch), "cufftPlanMany(side D2Z)");
        s->side_plans = true;
        cufft_check(cufftSetStream(s->plan1_side, s->side_stream), "cufftSetStream");
        cufft_check(cufftSetStream(s->plan_d2z_side, s->side_stream), "cufftSetStream");
        s->aprime2[0] = dmalloc<double2>(BN);
        s->aprime2[1] = dmalloc<double2>(BN);
        s->half_side = dmalloc<double2>(BN);
    }
    s->combined_ok = env_int("RB_COMBINED", 1) != 0;
    s->hist_store_next = env_int("RB_HIST_NEXT", 1) != 0;
    plan_sweep2(s);
    choose_sweep_kernel(s);
    set_stream(s, nullptr);
    return up.release();
}

static void set_stream(rb_solver* s, cudaStream_t st) {
    // The legacy default stream cannot be captured into a CUDA graph.  A blocking stream created here is ordered against the
    // legacy stream in both directions (implicit synchronisation), so callers that work on stream 0 see the same semantics.
    if (st == nullptr || st == cudaStreamLegacy) {
        if (!s->own_stream) RB_CUDA(cudaStreamCreate(&s->own_stream));
        st = s->own_stream;
    }
    s->stream = st;
    cufft_check(cufftSetStream(s->plan1, st), "cufftSetStream");
    cufft_check(cufftSetStream(s->plan2, st), "cufftSetStream");
    cufft_check(cufftSetStream(s->plan3, st), "cufftSetStream");
    cufft_check(cufftSetStream(s->plan_d2z, st), "cufftSetStream");
    if (s->plan_z2d) cufft_check(cufftSetStream(s->plan_z2d, st), "cufftSetStream");
}

// ZPhiDerivative::exec into the solver's own buffers (Zp | Zpp | PhiPrime)
static void derivatives(rb_solver* s, const double2* Z, const double2* Phi, bool finish = true) {
    const size_t BN = s->BN;
    cudaStream_t st = s->stream;
    if (s->own_fft) {
        launch_fft_zphi(Z, Phi, s->Zp(), s->Zpp(), s->PhiPc(), s->N, s->logN, s->batch, s->fft_tw, s->props.rho, s->props.U, st);
        if (finish) launch_finish_zphi(s->Zp(), s->Zpp(), s->PhiPc(), s->N, s->batch, s->props.rho, s->props.U, st);
        return;
    }
    double2* zper = s->fwork;            // [0]
    double2* phiper = s->fwork + BN;     // [1]
    launch_sub_linear(Z, Phi, zper, phiper, s->N, s->batch, s->props.rho, s->props.U, st);
    cufft_check(cufftExecZ2Z(s->plan2, (cufftDoubleComplex*)zper, (cufftDoubleComplex*)zper, CUFFT_FORWARD), "fft fwd");
    // d1z -> Zp, d2z -> Zpp, d1phi -> PhiPc, then one inverse over the three
    launch_spectral_multiply_zphi(zper, phiper, s->Zp(), s->Zpp(), s->PhiPc(), s->N, s->batch, st);
    cufft_check(cufftExecZ2Z(s->plan3, (cufftDoubleComplex*)s->deriv, (cufftDoubleComplex*)s->deriv, CUFFT_INVERSE), "fft inv");
    // finish == false: the caller's geometry kernel applies the scaling and the linear parts (one launch less per RHS)
    if (finish) launch_finish_zphi(s->Zp(), s->Zpp(), s->PhiPc(), s->N, s->batch, s->props.rho, s->props.U, st);
}

// derivatives + geometry of one RHS (everything that depends on the surface only)
static void surface_stage(rb_solver* s, const double2* Z, const double2* Phi);

static Geometry make_geometry(rb_solver* s, const double2* Z) {
    Geometry g;
    g.Z = Z;
    g.Zp = s->Zp();
    g.Zpp = s->Zpp();
    g.EG = s->EG;
    g.P0 = s->P0;
    g.Pm = s->Pm;
    g.Pp = s->Pp;
    g.EI = s->EI;
    g.Mdiag = s->Mdiag;
    g.V1diag = s->V1diag;
    g.V2 = s->V2;
    g.b = s->b;
    return g;
}

static void surface_stage(rb_solver* s, const double2* Z, const double2* Phi) {
    derivatives(s, Z, Phi, false);
    Geometry g = make_geometry(s, Z);
    launch_geometry(g, s->PhiPc(), s->N, s->batch, s->ncell, s->props.physics, s->rhoM, s->props.depth, s->has_image,
                    s->use_local, 1, s->props.rho, s->props.U, s->stream);
}

// a' = (2 pi / N) D1(a) for the real vector a: D2Z, coefficient multiply with the scale folded in, inverse Z2Z
// (L/BaseBoundaryIntegrator.cuh:201-203 does real_to_complex + Z2Z + multiply + Z2Z + scale: five launches)
static void real_derivative(rb_solver* s, const double* x, double2* out, const SolveCtrl* skip_ctrl = nullptr) {
    if (s->own_fft) {
        launch_fft_real_derivative(x, out, s->N, s->logN, s->batch, s->fft_tw, 2.0 * kPi / s->N, skip_ctrl, s->stream);
        return;
    }
    double2* half = s->fwork + 2 * s->BN;
    cufft_check(cufftExecD2Z(s->plan_d2z, (cufftDoubleReal*)x, (cufftDoubleComplex*)half), "fft d2z");
    launch_spectral_multiply_real(half, out, s->N, s->batch, 2.0 * kPi / s->N, s->stream);
    cufft_check(cufftExecZ2Z(s->plan1, (cufftDoubleComplex*)out, (cufftDoubleComplex*)out, CUFFT_INVERSE), "fft inv");
}

// the same derivative on the side stream (own plans, own half-spectrum buffer): runs beside the sweep of the same round
static void real_derivative_side(rb_solver* s, const double* x, double2* out, const SolveCtrl* skip_ctrl, bool force_own) {
    cudaStream_t q = s->side_stream;
    if (s->own_fft || force_own) {
        launch_fft_real_derivative(x, out, s->N, s->logN, s->batch, s->fft_tw, 2.0 * kPi / s->N, skip_ctrl, q);
        return;
    }
    cufft_check(cufftExecD2Z(s->plan_d2z_side, (cufftDoubleReal*)x, (cufftDoubleComplex*)s->half_side), "fft d2z (side)");
    launch_spectral_multiply_real(s->half_side, out, s->N, s->batch, 2.0 * kPi / s->N, q);
    cufft_check(cufftExecZ2Z(s->plan1_side, (cufftDoubleComplex*)out, (cufftDoubleComplex*)out, CUFFT_INVERSE), "fft inv (side)");
}

static SweepArgs base_args(rb_solver* s, const double2* Z) {
    SweepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.N = s->N;
    a.batch = s->batch;
    a.ncell = s->ncell;
    a.rows_per_thread = s->v1_rows;
    a.tile = s->tile;
    a.tiles_per_chunk = s->tiles_per_chunk;
    a.nchunks = s->nchunks;
    a.row_cell0 = s->row_cell0;
    a.row_cells = s->row_cells;
    a.comm = s->comm;
    a.use_local = s->use_local;
    a.has_image = s->has_image;
    a.g = make_geometry(s, Z);
    a.partial = s->partial;
    a.partial_img = s->partial_img;
    a.chunk_group = s->chunk_group;
    a.ngroups = s->ngroups;
    a.gpartial = s->gpartial;
    a.gpartial_img = s->gpartial_img;
    a.group_tickets = s->group_tickets;
    a.cell_tickets = s->cell_tickets;
    a.member_tickets = s->member_tickets;
    a.ctrl = s->ctrl;
    a.cK = s->cK;
    a.omega = s->omega;
    a.rho = s->props.rho;
    a.depth = s->props.depth;
    a.tol2 = s->props.tolerance * s->props.tolerance;
    a.max_iters = s->props.max_iterations;
    a.kappa = s->props.kappa;
    a.expansion_order = s->props.expansion_order;
    a.bnorm_part = s->bnorm_part;
    a.rnorm_part = s->rnorm_part;
    a.v2_RB = s->v2_RB;
    a.v2_R = s->v2_R;
    a.v2_groups = s->v2_groups;
    a.v2_spg = s->v2_spg;
    a.v2_TS = s->v2_TS;
    a.v2_bpm = s->v2_bpm;
    a.v2_total_blocks = s->v2_total_blocks;
    a.v2_row_begin = s->row_cell0 * kCell;
    a.v2_row_end = std::min(s->N, (s->row_cell0 + s->row_cells) * kCell);
    a.v2_rnorm_part = s->v2_rnorm_part;
    a.v2_ticket = s->v2_ticket;
    a.v2_split = s->v2_split;
    a.v2_partial = s->v2_partial;
    a.v2_xs_part = s->v2_xs_part;
    a.v2_blk_tickets = s->v2_blk_tickets;
    return a;
}

static void launch_mv(rb_solver* s, const SweepArgs& base, int i, int skip) {
    SweepArgs a = base;
    a.x = s->xbuf[i & 1];
    a.x_out = s->xbuf[(i + 1) & 1];
    a.xsum_part = s->xsum_part[i & 1];
    a.xsum_part_out = s->xsum_part[(i + 1) & 1];
    a.out_buf = (i + 1) & 1;
    a.final_buf_on_done = (i + 1) & 1;
    a.skip_if_done = skip;
    sweep(s, a, kSweepMV);
    if (s->comm.nranks > 1)
        launch_comm_wait(s->comm, s->ctrl, skip ? 1 : 2, a.out_buf, a.final_buf_on_done, s->bnorm_part, s->ncell, a.tol2, a.max_iters,
                         s->stream);
}

static void read_ctrl(rb_solver* s) {
    RB_CUDA(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(SolveCtrl), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(cudaStreamSynchronize(s->stream));
    s->last_iters = s->h_ctrl->iters;
    s->last_converged = s->h_ctrl->converged;
    s->last_stagnated = s->h_ctrl->stagnated;
    s->last_rel = std::sqrt(std::max(0.0, s->h_ctrl->rel2));
}

// Book-keeping of how one solve ended.  The reference's direct LU cannot fail to converge (L/MatrixSolver.cuh:114-172); an iterative
// solve can, and must not hand an unconverged vortex-sheet strength to its caller silently: a failed solve throws (the C ABI call
// returns -1, rb_last_error says why) unless rb_set_strict(s, 0) was called; a stagnated one is accepted and counted.
static void note_solve_end(rb_solver* s, int converged, int stagnated, double rel, int iters, const char* what) {
    if (converged || stagnated) {
        if (rel == rel) s->worst_rel = std::max(s->worst_rel, rel);
        if (stagnated && !converged) s->stagnated_solves++;
        return;
    }
    s->failed_solves++;
    if (s->strict) {
        char msg[256];
        std::snprintf(msg, sizeof msg, "%s: the solve for the vortex-sheet strength did not converge (relative residual %.3e after %d "
                      "applications of M, tolerance %.1e%s)", what, rel, iters, s->props.tolerance,
                      s->comm.nranks > 1 ? "; on a row-sharded run also check rb_comm_error" : "");
        throw std::runtime_error(msg);
    }
}

static void account_solve(rb_solver* s) {
    s->sum_iters += s->last_iters;
    s->num_solves++;
}

// ---- restarted GMRES(m), right-preconditioned, classical Gram-Schmidt with re-orthogonalisation -------------------------
// y = M x for the rows of this rank (published to every rank when sharded); x: any device vector, result in xbuf[1]
// skip != nullptr: the sweep (and, row-sharded, the wait behind it) returns at once when skip->done is set (recorded GMRES cycle)
static void apply_M(rb_solver* s, const SweepArgs& base, const double* x, SolveCtrl* skip = nullptr) {
    cudaStream_t st = s->stream;
    launch_finish_solve(x, x, nullptr, nullptr, nullptr, s->xsum_part[0], HistoryRing(), s->N, s->batch, s->ncell, st);
    SweepArgs a = base;
    a.x = x;
    a.x_out = s->xbuf[1];
    a.xsum_part = s->xsum_part[0];
    a.xsum_part_out = s->xsum_part[1];
    a.apply_only = 1;
    a.skip_if_done = skip ? 1 : 0;
    if (skip) a.ctrl = skip;
    a.out_buf = 1;
    sweep(s, a, kSweepMV);
    if (s->comm.nranks > 1)
        launch_comm_wait(s->comm, skip ? skip : s->ctrl, skip ? 3 : 0, 0, 0, s->bnorm_part, s->ncell, 0.0, 0, st);
}

// out = P^{-1} v  (real FFT, divide by the flat-film symbol, inverse real FFT: three launches); out may alias v
static void apply_Pinv(rb_solver* s, const double* v, double* out) {
    cudaStream_t st = s->stream;
    double2* half = s->fwork;   // free between the derivative stage and the a' stage; (N/2 + 1) * batch complex values
    cufft_check(cufftExecD2Z(s->plan_d2z, (cufftDoubleReal*)const_cast<double*>(v), (cufftDoubleComplex*)half), "fft d2z");
    launch_precond_scale_half(half, s->gm_invP, s->N, s->batch, st);
    cufft_check(cufftExecZ2D(s->plan_z2d, (cufftDoubleComplex*)half, (cufftDoubleReal*)out), "fft z2d");
}

static void gmres_solve(rb_solver* s, const double2* Z) {
    cudaStream_t st = s->stream;
    const int n = (int)s->BN, m = s->gm_m;
    const size_t ld = s->gm_ld;
    const double tol = s->props.tolerance;
    double* V = s->gm_V;
    double* w = s->xbuf[1];
    double* dh = s->gm_dev;            // [0..m+1] pass 1, [m+2..2m+3] pass 2, then norm, then y
    double* hh = s->gm_host;
    const int stride = m + 2;
    SweepArgs base = base_args(s, Z);

    const double* warm = nullptr;
    if (s->props.guess_mode == RB_GUESS_WARM && !s->hist.base && s->have_prev_a) warm = s->a;
    launch_guess(s->b, warm, s->hist, s->gm_x, s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch, s->ncell, st);
    if (!warm && !s->hist.base) apply_Pinv(s, s->b, s->gm_x);   // cold start: x0 = P^{-1} b

    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    int applies = 0;
    double rel = 1e300, bnorm = 0.0, prev_cycle_rel = 1e300;
    bool converged = false, stagnated = false;
    const int max_applies = s->props.max_iterations;
    for (int restart = 0; restart < 50 && !converged && applies < max_applies; ++restart) {
        // true residual r = b - M x
        apply_M(s, base, s->gm_x);
        applies++;
        launch_axpby(V, s->b, -1.0, w, n, st);
        launch_multi_dot(V, ld, 0, V, dh, n, st);           // dh[0] = r.r
        launch_multi_dot(V, ld, 0, s->b, dh + 1, n, st);     // dh[1] = b.b
        RB_CUDA(cudaMemcpyAsync(hh, dh, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        const double beta = std::sqrt(hh[0]);
        bnorm = std::sqrt(hh[1]);
        rel = bnorm > 0 ? beta / bnorm : (beta == 0 ? 0.0 : 1e300);
        if (!(rel == rel)) break;                             // NaN: give up, reported as not converged
        if (rel <= tol) {
            converged = true;
            break;
        }
        // the true residual after a whole restart cycle no better than half the one before it, and already <= 1e-10: the iteration
        // sits on the round-off floor of this operator (eps x cond(M), cond ~ N / 2 pi for the thin film) -- same rule, same
        // status as the Richardson solver's (solve_decide): accepted, reported as stagnated, never as converged
        if (restart > 0 && rel <= 1e-10 && rel > 0.5 * prev_cycle_rel) {
            stagnated = true;
            break;
        }
        prev_cycle_rel = rel;
        launch_normalize(V, V, dh, n, st);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int k_used = 0;
        for (int k = 0; k < m && applies < max_applies; ++k) {
            apply_Pinv(s, V + (size_t)k * ld, s->gm_t);
            apply_M(s, base, s->gm_t);                        // w = M P^{-1} v_k
            applies++;
            launch_multi_dot(V, ld, k + 1, w, dh, n, st);
            launch_multi_axpy(w, V, ld, k + 1, dh, -1.0, n, st);
            launch_multi_dot(V, ld, k + 1, w, dh + stride, n, st);
            launch_multi_axpy(w, V, ld, k + 1, dh + stride, -1.0, n, st);
            launch_multi_dot(V, ld, 0, w, dh + 2 * stride, n, st);
            launch_normalize(V + (size_t)(k + 1) * ld, w, dh + 2 * stride, n, st);
            RB_CUDA(cudaMemcpyAsync(hh, dh, (2 * stride + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
            RB_CUDA(cudaStreamSynchronize(st));
            for (int j = 0; j <= k; ++j) H[(size_t)j * m + k] = hh[j] + hh[stride + j];
            H[(size_t)(k + 1) * m + k] = std::sqrt(std::max(0.0, hh[2 * stride]));
            for (int j = 0; j < k; ++j) {                    // previous rotations
                double a0 = H[(size_t)j * m + k], a1 = H[(size_t)(j + 1) * m + k];
                H[(size_t)j * m + k] = cs[j] * a0 + sn[j] * a1;
                H[(size_t)(j + 1) * m + k] = -sn[j] * a0 + cs[j] * a1;
            }
            double a0 = H[(size_t)k * m + k], a1 = H[(size_t)(k + 1) * m + k];
            double d = std::hypot(a0, a1);
            cs[k] = d > 0 ? a0 / d : 1.0;
            sn[k] = d > 0 ? a1 / d : 0.0;
            H[(size_t)k * m + k] = d;
            H[(size_t)(k + 1) * m + k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            k_used = k + 1;
            rel = std::fabs(g[k + 1]) / bnorm;
            if (!(rel == rel) || rel <= tol) break;
        }
        if (k_used == 0) break;
        for (int i = k_used - 1; i >= 0; --i) {               // back substitution
            double acc = g[i];
            for (int j = i + 1; j < k_used; ++j) acc -= H[(size_t)i * m + j] * y[j];
            y[i] = acc / H[(size_t)i * m + i];
        }
        std::memcpy(hh, y.data(), k_used * sizeof(double));
        double* dy = dh + 3 * stride;
        RB_CUDA(cudaMemcpyAsync(dy, hh, k_used * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_combine(s->gm_t, V, ld, k_used, dy, n, st);
        apply_Pinv(s, s->gm_t, s->gm_t);
        launch_axpby(s->gm_x, s->gm_x, 1.0, s->gm_t, n, st);
        RB_CUDA(cudaStreamSynchronize(st));                   // hh is reused by the next restart
        if (rel == rel && rel <= tol) converged = true;       // estimate; CGS2 keeps it within round-off of the true residual
    }
    s->last_iters = applies;
    s->last_converged = converged ? 1 : 0;
    s->last_stagnated = (stagnated && !converged) ? 1 : 0;
    s->last_rel = rel;
    account_solve(s);
    launch_finish_solve(s->gm_x, s->gm_x, nullptr, s->a, s->ac, s->xsum_a, s->hist, s->N, s->batch, s->ncell, st);
    s->have_prev_a = true;
    note_solve_end(s, s->last_converged, s->last_stagnated, rel, applies, "GMRES");
}

// M a = b.  On return a (real), ac (complex copy) and the per-cell sums of a are valid on the stream.
static void solve(rb_solver* s, const double2* Z) {
    cudaStream_t st = s->stream;
    if (!s->matrix_free_solve) {
        // validation path: materialise M exactly as the reference kernels do and factorise it
        const int N = s->N;
        if (!s->Mdense) s->Mdense = dmalloc<double>((size_t)N * N);
        for (int bm = 0; bm < s->batch; ++bm) {
            const size_t o = (size_t)bm * N;
            if (s->props.physics == RB_HELIUM)
                launch_create_finite_depth_M(s->Mdense, Z + o, s->Zp() + o, s->Zpp() + o, s->props.depth, N, 1,
                                             s->props.infinite_depth != 0, st);
            else
                launch_create_M(s->Mdense, Z + o, s->Zp() + o, s->Zpp() + o, s->rhoM, N, 1, st);
            RB_CUDA(cudaMemcpyAsync(s->xbuf[0] + o, s->b + o, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
            launch_lu_solve(s->Mdense, s->xbuf[0] + o, N, s->lu_info + bm, st);
        }
        launch_finish_solve(s->xbuf[0], s->xbuf[0], nullptr, s->a, s->ac, s->xsum_a, s->hist, s->N, s->batch, s->ncell, st);
        s->last_iters = 0;
        s->last_stagnated = 0;
        s->last_rel = 0;
        s->num_solves++;
        if (s->fixed_sweeps > 0) {   // being recorded into a graph: no host read here, the factorisation's info is left on the device
            s->last_converged = 1;
            return;
        }
        std::vector<int> infos(s->batch, 0);   // per member: first singular (or non-finite) pivot column + 1, else 0
        RB_CUDA(cudaMemcpyAsync(infos.data(), s->lu_info, s->batch * sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        int info = 0;
        for (int v : infos)
            if (v != 0 && info == 0) info = v;
        s->last_converged = info == 0 ? 1 : 0;
        if (info != 0) {
            s->failed_solves++;
            if (s->strict)
                throw std::runtime_error("dense LU: M is singular to working precision (zero or non-finite pivot in column " +
                                         std::to_string(info - 1) + ")");
        }
        return;
    }

    if (s->use_gmres) {
        gmres_solve(s, Z);
        return;
    }

    // matrix-free Richardson: x <- x + omega (b - M x), error contracts by rho(I - omega M) per sweep
    const double* warm = nullptr;
    if (s->props.guess_mode == RB_GUESS_WARM && !s->hist.base && s->have_prev_a) warm = s->a;
    launch_guess(s->b, warm, s->hist, s->xbuf[0], s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch, s->ncell,
                 st);
    SweepArgs base = base_args(s, Z);
    int launched = 0;
    if (s->fixed_sweeps > 0) {
        // capture mode: a fixed number of sweeps that skip themselves once the control block says done; the caller
        // inspects the control block after the whole step
        for (; launched < s->fixed_sweeps; ++launched) launch_mv(s, base, launched, 1);
    } else {
        int group = s->kpred > 0 ? s->kpred : 8;
        while (true) {
            for (int g = 0; g < group && launched < s->props.max_iterations; ++g, ++launched) launch_mv(s, base, launched, 1);
            read_ctrl(s);
            if (s->h_ctrl->done || launched >= s->props.max_iterations) break;
            group = 2;
        }
        s->kpred = s->last_iters;
        account_solve(s);
    }
    launch_finish_solve(s->xbuf[0], s->xbuf[1], s->ctrl, s->a, s->ac, s->xsum_a, s->hist, s->N, s->batch, s->ncell, st);
    s->have_prev_a = true;
    if (s->fixed_sweeps <= 0) note_solve_end(s, s->last_converged, s->last_stagnated, s->last_rel, s->last_iters, "Richardson iteration");
}

static void vorticities(rb_solver* s, const double2* state) {
    const double2* Z = state;
    const double2* Phi = state + s->BN;
    surface_stage(s, Z, Phi);
    solve(s, Z);
    s->cur_Z = Z;
    s->cur_Phi = Phi;
}

static void fft_derivative(rb_solver* s, const double2* in, double2* out, int second, double scaling) {
    double2* tmp = s->fwork + 2 * s->BN;
    cufft_check(cufftExecZ2Z(s->plan1, (cufftDoubleComplex*)in, (cufftDoubleComplex*)tmp, CUFFT_FORWARD), "fft fwd");
    launch_spectral_multiply(tmp, tmp, s->N, s->batch, second, s->stream);
    cufft_check(cufftExecZ2Z(s->plan1, (cufftDoubleComplex*)tmp, (cufftDoubleComplex*)out, CUFFT_INVERSE), "fft inv");
    if (scaling != 1.0) launch_scale(out, scaling, s->BN, s->stream);
}

// arena redirection of the RHS output when sharded: rows arrive from the peers inside the arena only
static double2* redirect_out(rb_solver* s, double2* out, double2** user_out) {
    *user_out = nullptr;
    if (s->comm.nranks > 1) {
        if (!s->rhs_phi_kind)
            throw std::runtime_error("row-sharded runs need a fused dPhi/dt (water with rho = 0, or helium without expansions / surface tension)");
        const char* p = reinterpret_cast<const char*>(out);
        if (p < s->arena || p >= s->arena + s->arena_bytes) {
            *user_out = out;
            out = s->kbuf[0];
        }
    }
    return out;
}

static void rhs_tail(rb_solver* s, const double2* state, double2* out, double2* user_out) {
    const size_t BN = s->BN;
    cudaStream_t st = s->stream;
    const double2* Z = state;
    if (user_out) {
        RB_CUDA(cudaMemcpyAsync(user_out, out, 2 * BN * sizeof(double2), cudaMemcpyDeviceToDevice, st));
        out = user_out;
    }
    if (!s->rhs_phi_kind) {
        const rb_props& p = s->props;
        if (p.physics == RB_WATER) {
            launch_rhs_phi_water(Z, out, s->vel_upper, out + BN, p.rho, (int)BN, st);   // L/WaterBoundaryProblem.cuh:37
        } else if (p.use_expansions) {
            launch_rhs_phi_helium_exp(Z, out, out + BN, p.depth, (int)BN, p.expansion_order, st);
        } else {
            launch_rhs_phi_helium_st(Z, s->Zp(), s->Zpp(), out, out + BN, p.depth, p.kappa, (int)BN, st);
        }
    }
    s->cur_vel = out;
    if (s->props.compute_energies)
        launch_energies(Z, s->Zp(), state + BN, out, s->energies, s->N, s->props.physics == RB_WATER ? 0 : 1, s->props.rho,
                        s->props.U, s->props.depth, s->props.kappa, st);
}

// RHS inside a recorded step: the sweep that verifies an iterate (r = b - M a) also produces its velocities from the same row
// sums, so a well-started solve costs two sweeps per RHS (one solver sweep + one combined sweep) instead of three.
// Sequence: guess x0 | sweep: x1, r0 | for each further recorded sweep i: a' of x_i, combined sweep on x_i: velocities(x_i), r_i,
// x_{i+1}; the first r_i within tolerance ends the solve with a = x_i (later sweeps skip themselves).
static void rhs_combined(rb_solver* s, const double2* state, double2* out) {
    const size_t BN = s->BN;
    cudaStream_t st = s->stream;
    const double2* Z = state;
    const double2* Phi = state + BN;
    s->cur_Z = Z;
    s->cur_Phi = Phi;
    double2* user_out = nullptr;
    out = redirect_out(s, out, &user_out);

    const double* warm = nullptr;
    if (s->props.guess_mode == RB_GUESS_WARM && !s->hist.base && s->have_prev_a) warm = s->a;
    // derivatives, then ONE kernel for the geometry of the surface and the start of the solve
    derivatives(s, Z, Phi, false);
    launch_geometry_guess(make_geometry(s, Z), s->PhiPc(), s->N, s->batch, s->ncell, s->rhoM, s->props.depth, s->has_image, s->use_local,
                          s->props.rho, s->props.U, warm, s->hist, s->xbuf[0], s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega,
                          s->cK, st);
    SweepArgs base = base_args(s, Z);
    const bool overlap = s->overlap_ok;
    // optimistic: the extrapolated guess is expected to verify as it stands, so sweep 0 is already a combined sweep and a
    // well-predicted RHS costs ONE O(N^2) sweep; otherwise sweep 0 is the cheaper solver sweep that cannot declare convergence
    // (nothing has produced velocities yet)
    const int first_combined = s->optimistic ? 0 : 1;
    if (!s->optimistic) {
        SweepArgs first = base;
        first.tol2 = -1.0;
        launch_mv(s, first, 0, 1);
    }
    for (int i = first_combined; i < s->fixed_sweeps; ++i) {
        const double* xi = s->xbuf[i & 1];
        const bool surplus = i >= first_combined + 1;   // almost always skipped -> use the kernel that can skip itself
        if (overlap) {
            // fork: a'(x_i) on the side stream while this round's sweep runs; finish_solve adds V2 a' once both have ended
            RB_CUDA(cudaEventRecord(s->ev_fork, st));
            RB_CUDA(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
            real_derivative_side(s, xi, s->aprime2[i & 1], s->ctrl, surplus);
        } else if (surplus && s->own_fft_skippable && !s->own_fft) {
            launch_fft_real_derivative(xi, s->aprime, s->N, s->logN, s->batch, s->fft_tw, 2.0 * kPi / s->N, s->ctrl, st);
        } else {
            real_derivative(s, xi, s->aprime, s->ctrl);   // own kernel: skips itself once the solve is finished
        }
        SweepArgs a = base;
        a.x = xi;
        a.x_out = s->xbuf[(i + 1) & 1];
        a.xsum_part = s->xsum_part[i & 1];
        a.xsum_part_out = s->xsum_part[(i + 1) & 1];
        a.out_buf = (i + 1) & 1;
        a.final_buf_on_done = i & 1;      // the verified iterate is the INPUT of this sweep
        a.combined = 1;
        a.skip_if_done = 1;
        a.aprime = s->aprime;
        a.vel_lower = out;
        a.vel_upper = s->vel_upper;
        a.rhs_phi_kind = s->rhs_phi_kind;
        a.dphi = s->rhs_phi_kind ? out + BN : nullptr;
        a.A_out = s->hist.Abase ? s->Abuf[i & 1] : nullptr;
        a.defer_aprime = overlap ? 1 : 0;
        sweep(s, a, kSweepVEL);
        if (s->comm.nranks > 1)
            launch_comm_wait(s->comm, s->ctrl, 1, a.out_buf, a.final_buf_on_done, s->bnorm_part, s->ncell, a.tol2, a.max_iters, st);
    }
    HistoryRing hist_out = s->hist;
    hist_out.store_next = s->hist_store_next ? 1 : 0;
    if (overlap) {   // join
        RB_CUDA(cudaEventRecord(s->ev_join, s->side_stream));
        RB_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
    }
    FinishPost post;
    const bool fuse_update = s->post_update.update != 0 && s->rhs_phi_kind != 0 && !user_out && !s->props.compute_energies;
    if (overlap || fuse_update) {
        post.vel = out;
        post.vel_upper = s->vel_upper;
        post.dphi = s->rhs_phi_kind ? out + BN : nullptr;
        post.V2 = s->V2;
        post.Z = Z;
        post.rhs_phi_kind = s->rhs_phi_kind;
        post.depth = s->props.depth;
        post.BN = BN;
        if (overlap) {
            post.aprime0 = s->aprime2[0];
            post.aprime1 = s->aprime2[1];
        }
        if (fuse_update) {
            post.update = s->post_update.update;
            post.c = s->post_update.c;
            post.y0 = s->post_update.y0;
            post.y_out = s->post_update.y_out;
            post.k1 = s->post_update.k1;
            post.k2 = s->post_update.k2;
            post.k3 = s->post_update.k3;
            s->post_update_done = true;
        }
    }
    launch_finish_solve(s->xbuf[0], s->xbuf[1], s->ctrl, s->a, nullptr, s->xsum_a, hist_out, s->N, s->batch, s->ncell, st,
                        s->hist.Abase ? s->Abuf[0] : nullptr, s->Abuf[1], post.vel ? &post : nullptr);
    s->have_prev_a = true;
    rhs_tail(s, state, out, user_out);
}

// RHS of the finite-depth helium operator inside a recorded step: one GMRES cycle of at most K = fixed_sweeps - 2 Arnoldi steps,
// driven entirely from the device (no host synchronisation), then the velocity sweep in combined mode, which also VERIFIES the
// solution with the true residual b - M a from the same row sums (the cycle itself stops on the Arnoldi estimate at half the
// tolerance).  A solve that does not verify leaves its control block "not done": the stepper rolls the step back and redoes it with
// the host-driven restarted GMRES of gmres_solve.  Sequence per RHS: r0 = b - M x0 | K x [P^-1 v_k, M (.), arnoldi] | x += P^-1 V y |
// a' | combined velocity sweep.
static void rhs_gmres_recorded(rb_solver* s, const double2* state, double2* out) {
    const size_t BN = s->BN;
    cudaStream_t st = s->stream;
    const double2* Z = state;
    const double2* Phi = state + BN;
    const int n = (int)BN;
    s->cur_Z = Z;
    s->cur_Phi = Phi;
    double2* user_out = nullptr;
    out = redirect_out(s, out, &user_out);
    surface_stage(s, Z, Phi);
    const double* warm = nullptr;
    if (s->props.guess_mode == RB_GUESS_WARM && !s->hist.base && s->have_prev_a) warm = s->a;
    launch_guess(s->b, warm, s->hist, s->gm_x, s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch, s->ncell, st);
    if (!warm && !s->hist.base) apply_Pinv(s, s->b, s->gm_x);   // cold start: x0 = P^{-1} b
    SweepArgs base = base_args(s, Z);
    double* w = s->xbuf[1];
    SolveCtrl* skip = reinterpret_cast<SolveCtrl*>(s->gm_ctrl);
    const double tol_in = 0.5 * s->props.tolerance;
    const int K = std::max(1, std::min(std::min(kGmMax, s->gm_m), s->fixed_sweeps - 2));
    apply_M(s, base, s->gm_x);                                                      // w = M x0
    launch_gm_start(s->b, w, s->gm_V, s->gm_members, s->gm_ctrl, s->ctrl, s->N, s->batch, tol_in, st);
    for (int k = 0; k < K; ++k) {
        apply_Pinv(s, s->gm_V + (size_t)k * s->gm_ld, s->gm_t);
        apply_M(s, base, s->gm_t, skip);                                            // w = M P^{-1} v_k
        launch_gm_arnoldi(s->gm_V, s->gm_ld, w, s->gm_members, s->gm_ctrl, s->ctrl, s->N, s->batch, k, K - 1, tol_in, st);
    }
    launch_gm_correction(s->gm_V, s->gm_ld, s->gm_t, s->gm_members, s->N, s->batch, st);
    apply_Pinv(s, s->gm_t, s->gm_t);
    launch_axpby(s->gm_x, s->gm_x, 1.0, s->gm_t, n, st);
    launch_finish_solve(s->gm_x, s->gm_x, nullptr, s->a, s->ac, s->xsum_a, s->hist, s->N, s->batch, s->ncell, st);
    s->have_prev_a = true;
    real_derivative(s, s->a, s->aprime);
    SweepArgs a = base;
    a.x = s->a;
    a.x_out = s->xbuf[1];                 // (the next Richardson iterate the combined mode also forms: not used here)
    a.xsum_part = s->xsum_a;
    a.xsum_part_out = s->xsum_part[1];
    a.out_buf = 1;
    a.final_buf_on_done = 0;
    a.combined = 1;
    a.skip_if_done = 0;
    a.max_iters = 1 << 30;                // not verified -> the block stays "not done" -> the stepper's fallback, never a silent accept
    a.aprime = s->aprime;
    a.vel_lower = out;
    a.vel_upper = s->vel_upper;
    a.rhs_phi_kind = s->rhs_phi_kind;
    a.dphi = s->rhs_phi_kind ? out + BN : nullptr;
    sweep(s, a, kSweepVEL);
    if (s->comm.nranks > 1)
        launch_comm_wait(s->comm, s->ctrl, 1, a.out_buf, a.final_buf_on_done, s->bnorm_part, s->ncell, a.tol2, a.max_iters, st);
    rhs_tail(s, state, out, user_out);
}

static void rhs(rb_solver* s, const double2* state, double2* out) {
    if (s->fixed_sweeps >= 2 && s->matrix_free_solve && !s->use_gmres && s->combined_ok) {
        rhs_combined(s, state, out);
        return;
    }
    if (s->fixed_sweeps >= 3 && s->matrix_free_solve && s->use_gmres && s->gm_device) {
        rhs_gmres_recorded(s, state, out);
        return;
    }
    const size_t BN = s->BN;
    cudaStream_t st = s->stream;
    vorticities(s, state);
    const double2* Z = state;
    real_derivative(s, s->a, s->aprime);   // L/BaseBoundaryIntegrator.cuh:201-203
    SweepArgs a = base_args(s, Z);
    double2* user_out = nullptr;
    out = redirect_out(s, out, &user_out);
    a.x = s->a;
    a.xsum_part = s->xsum_a;
    a.aprime = s->aprime;
    a.vel_lower = out;
    a.vel_upper = s->vel_upper;
    a.rhs_phi_kind = s->rhs_phi_kind;
    a.dphi = s->rhs_phi_kind ? out + BN : nullptr;
    sweep(s, a, kSweepVEL);
    s->vel_sweeps++;
    if (s->comm.nranks > 1) launch_comm_wait(s->comm, s->ctrl, 0, 0, 0, s->bnorm_part, s->ncell, 0.0, 0, st);
    rhs_tail(s, state, out, user_out);
}

// ------------------------------------------------------------------------------------------------
// the RK4 stepper
// ------------------------------------------------------------------------------------------------
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
    // "tight" recording: after a long run of steps that all needed the same number of sweeps, record exactly that many (no surplus,
    // self-skipping round per solve: at N <= 4096 such a round -- a skipped sweep, its a' transform, the fork / join around it -- is
    // ~8 % of a step); a step that then runs out of sweeps is rolled back and redone as always, and tight recording is banned for a while
    bool tight_ok = true;
    bool tight = false;
    int tight_hits = 0, tight_ban = 0;
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

static void stepper_free(rb_stepper* st) {
    if (!st) return;
    for (auto& g : st->graph_cache)
        if (g) cudaGraphExecDestroy(g);
    if (st->ev) cudaEventDestroy(st->ev);
    if (st->owns_y0 && st->y0) cudaFree(st->y0);
    if (st->ytmp) cudaFree(st->ytmp);
    if (st->ybackup) cudaFree(st->ybackup);
    for (auto p : st->hist)
        if (p) cudaFree(p);
    if (st->d_counter) cudaFree(st->d_counter);
    if (st->d_agg) cudaFree(st->d_agg);
    if (st->h_agg) cudaFreeHost(st->h_agg);
    if (st->ycheck) cudaFree(st->ycheck);
    if (st->log_states) cudaFree(st->log_states);
    delete st;
}

static void stepper_reset_history(rb_stepper* st) {
    st->h_counter = 0;
    RB_CUDA(cudaMemsetAsync(st->d_counter, 0, sizeof(int), st->s->stream));
}

// the kernel sequence of one RK4 step (L/AutonomousRungeKuttaStepper.cuh:124-307); identical whether it is launched
// directly or recorded into a graph.  fixed_sweeps > 0: no host synchronisation anywhere inside.
static void issue_step(rb_stepper* st, int fixed_sweeps) {
    rb_solver* s = st->s;
    const size_t n2 = 2 * s->BN;
    cudaStream_t cs = s->stream;
    const double h = st->dt;
    const bool warm = s->props.guess_mode == RB_GUESS_WARM && s->matrix_free_solve;
    s->fixed_sweeps = fixed_sweeps;
    int agg_iters = 0, agg_conv = 1, agg_stag = 0;
    double agg_rel = 0.0;
    auto stage = [&](int i, const double2* y) {
        s->ctrl = s->ctrl_all + i;
        if (warm) {
            s->hist.base = st->hist[i];
            s->hist.stride = s->BN;
            s->hist.ring = kHistRing;
            s->hist.order = st->order;
            s->hist.counter = st->d_counter;
            // row-sum history only where the combined sweep produces it (recorded steps of the Richardson path)
            const bool keepA = st->predict && fixed_sweeps >= 2 && !s->use_gmres && s->combined_ok && !s->has_image;
            s->hist.Abase = keepA ? reinterpret_cast<double2*>(st->hist[i] + (size_t)kHistRing * s->BN) : nullptr;
            s->hist.predict = keepA ? 1 : 0;
        }
        s->optimistic = fixed_sweeps > 0 && ((st->opt_mask >> i) & 1);
        rhs(s, y, st->k[i]);
        s->optimistic = false;
        if (fixed_sweeps <= 0) {   // synchronising path: each solve has just reported; keep the step's aggregate
            agg_iters = std::max(agg_iters, s->last_iters);
            agg_conv = agg_conv && s->last_converged;
            agg_stag = agg_stag || s->last_stagnated;
            agg_rel = std::max(agg_rel, s->last_rel);
        }
    };
    // the RK update after a stage is folded into the kernel that closes the stage's solve when the RHS allows it
    auto staged = [&](int i, const double2* y, int update, double c) {
        s->post_update = FinishPost();
        s->post_update.update = update;
        s->post_update.c = c;
        s->post_update.y0 = st->y0;
        s->post_update.y_out = update == 2 ? st->y0 : st->ytmp;
        s->post_update.k1 = st->k[0];
        s->post_update.k2 = st->k[1];
        s->post_update.k3 = st->k[2];
        s->post_update_done = false;
        stage(i, y);
        const bool done = s->post_update_done;
        s->post_update = FinishPost();
        s->post_update_done = false;
        return done;
    };
    auto restore = [&]() {
        s->hist = HistoryRing();
        s->ctrl = s->ctrl_all;
        s->fixed_sweeps = 0;
        s->optimistic = false;
        s->post_update = FinishPost();
        s->post_update_done = false;
    };
    try {
        if (!staged(0, st->y0, 1, h * 0.5)) launch_stage_update(st->ytmp, st->y0, st->k[0], h * 0.5, n2, cs);
        if (!staged(1, st->ytmp, 1, h * 0.5)) launch_stage_update(st->ytmp, st->y0, st->k[1], h * 0.5, n2, cs);
        if (!staged(2, st->ytmp, 1, h)) launch_stage_update(st->ytmp, st->y0, st->k[2], h, n2, cs);
        if (!staged(3, st->ytmp, 2, h / 6.0)) launch_final_update(st->y0, st->k[0], st->k[1], st->k[2], st->k[3], h, n2, cs);
        // recorded steps end with the kernel that advances the history counter and folds the stage solves' status into the chunk aggregate
        if (fixed_sweeps > 0) launch_step_end(warm ? st->d_counter : nullptr, s->ctrl_all, st->d_agg, st->opt_mask, cs);
        else if (warm) launch_advance_counter(st->d_counter, cs);
    } catch (...) {
        // a stage's solve failed (strict mode): y0 has not been touched yet (the final update is the last thing a step does)
        restore();
        throw;
    }
    restore();
    if (fixed_sweeps <= 0) {
        s->last_iters = agg_iters;
        s->last_converged = agg_conv;
        s->last_stagnated = agg_stag;
        s->last_rel = agg_rel;
    }
}

static void capture_graph(rb_stepper* st, int sweeps) {
    rb_solver* s = st->s;
    cudaGraphExec_t& slot = st->graph_cache[st->opt_mask & 15];
    if (slot) {
        cudaGraphExecDestroy(slot);
        slot = nullptr;
    }
    st->graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    const unsigned long long launches_before = rb::g_launch_count;
    RB_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    try {
        RB_CUDA(cudaMemcpyAsync(st->ybackup, st->y0, 2 * s->BN * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
        issue_step(st, sweeps);
        RB_CUDA(cudaMemcpyAsync(s->h_ctrl, s->ctrl_all, 4 * sizeof(SolveCtrl), cudaMemcpyDeviceToHost, s->stream));
    } catch (...) {
        cudaStreamEndCapture(s->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    RB_CUDA(cudaStreamEndCapture(s->stream, &graph));
    st->graph_kernels[st->opt_mask & 15] = (int)(rb::g_launch_count - launches_before);   // this library's kernels in one step
    rb::g_launch_count = launches_before;   // recorded, not launched
    RB_CUDA(cudaGraphInstantiate(&slot, graph, 0));
    RB_CUDA(cudaGraphDestroy(graph));
    st->graph_exec = slot;
    st->graph_mask = st->opt_mask;
    st->graph_sweeps = sweeps;
    st->graph_dt = st->dt;
    st->graph_y0 = st->y0;
    st->graph_hits_below = 0;
    st->graph_captures++;
}

static void after_step(rb_stepper* st) {
    st->t += st->dt;
    st->step_index++;
    rb_solver* s = st->s;
    if (st->log_every && st->log_states && (st->step_index % st->log_every) == 0 && st->log_count < st->log_capacity) {
        RB_CUDA(cudaMemcpyAsync(st->log_states + st->log_count * 2 * s->BN, st->y0, 2 * s->BN * sizeof(double2),
                                cudaMemcpyDeviceToDevice, s->stream));
        st->log_times.push_back(st->t);
        st->log_count++;
    }
}

static void invalidate_graphs(rb_stepper* st) {
    for (auto& g : st->graph_cache)
        if (g) {
            cudaGraphExecDestroy(g);
            g = nullptr;
        }
    st->graph_exec = nullptr;
}

static void stepper_step(rb_stepper* st) {
    rb_solver* s = st->s;
    const bool graphable = st->use_graph && s->matrix_free_solve && (!s->use_gmres || s->gm_device);   // host-driven GMRES cannot be recorded
    if (!graphable) {
        issue_step(st, 0);   // every stage's solve synchronises and is checked where it ends (note_solve_end)
        if (s->props.guess_mode == RB_GUESS_WARM && s->matrix_free_solve) st->h_counter++;
        after_step(st);
        return;
    }
    if (st->graph_dt != st->dt || st->graph_y0 != st->y0) invalidate_graphs(st);   // recorded constants changed
    st->graph_exec = st->graph_cache[st->opt_mask & 15];
    if (!st->graph_exec) {
        int sweeps = st->graph_sweeps > 0 ? st->graph_sweeps : std::min(s->props.max_iterations, 16);
        sweeps = std::max(sweeps, s->use_gmres ? 3 : 2);
        capture_graph(st, sweeps);
    }
    const int mask = st->opt_mask;
    RB_CUDA(cudaGraphLaunch(st->graph_exec, s->stream));
    rb::count_launch(st->graph_kernels[st->opt_mask & 15]);
    st->graph_launches++;
    RB_CUDA(cudaEventRecord(st->ev, s->stream));
    RB_CUDA(cudaEventSynchronize(st->ev));
    int worst = 0;
    bool all_done = true;
    for (int i = 0; i < 4; ++i) {
        const SolveCtrl& c = s->h_ctrl[i];
        all_done = all_done && c.done;
        // sweeps this solve occupied in the recorded sequence (an optimistic stage has no leading solver sweep)
        worst = std::max(worst, c.iters + (((mask >> i) & 1) ? 1 : 0));
    }
    if (!all_done) {
        // some solve ran out of recorded sweeps: roll the step back and redo it with the synchronising loop
        RB_CUDA(cudaMemcpyAsync(st->y0, st->ybackup, 2 * s->BN * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
        if (s->props.guess_mode == RB_GUESS_WARM) {
            // the counter was advanced by the failed graph: put it back (slots written by the failed step are rewritten)
            RB_CUDA(cudaMemcpyAsync(st->d_counter, &st->h_counter, sizeof(int), cudaMemcpyHostToDevice, s->stream));
            RB_CUDA(cudaStreamSynchronize(s->stream));
        }
        issue_step(st, 0);
        st->fallback_steps++;
        if (st->predict && s->props.guess_mode == RB_GUESS_WARM) {
            // the synchronising loop records no row sums: the rings are inconsistent for this step -> start the history afresh
            stepper_reset_history(st);
            st->h_counter = -1;   // incremented to 0 below, matching the device counter
        }
        worst = std::max(worst, s->kpred);
        int sweeps = std::min(s->props.max_iterations, worst + 4);
        if (st->tight) {   // a tightly recorded step ran out of sweeps: back to a surplus round, and no new attempt for a while
            st->tight = false;
            st->tight_ban = 512;
            st->tight_failures++;
            sweeps = std::min(s->props.max_iterations, std::max(worst, st->graph_sweeps) + 2);
        }
        st->tight_hits = 0;
        st->graph_sweeps = sweeps;
        st->opt_mask = st->opt_policy == 2 ? 15 : 0;
        invalidate_graphs(st);
    } else {
        const double tol2 = s->props.tolerance * s->props.tolerance;
        int next_mask = 0;
        for (int i = 0; i < 4; ++i) {
            const SolveCtrl& c = s->h_ctrl[i];
            s->sum_iters += c.iters;
            s->num_solves++;
            st->first_rel[i] = std::sqrt(std::max(0.0, c.first_rel2));
            if ((mask >> i) & 1) {
                st->opt_stage_solves++;
                if (c.iters == 1) st->one_sweep_solves++;
            }
            // adaptive policy: a failed optimistic stage costs 13 + 13 instead of 11 + 13 instructions per pair, a successful one 13
            // instead of 24, so a stage is optimistic whenever its last guess came within twice the tolerance
            if (c.first_rel2 <= 4.0 * tol2) next_mask |= 1 << i;
        }
        if (st->opt_policy == 0) next_mask = 0;
        if (st->opt_policy == 2) next_mask = 15;
        st->opt_mask = next_mask;
        // status of the step = status of its four stage solves together (a stage that ended on the iteration cap, a NaN or a peer
        // time-out has done = 1 and converged = stagnated = 0)
        int it_max = 0, all_conv = 1, any_stag = 0, failed = -1;
        double rel_max = 0.0;
        for (int i = 0; i < 4; ++i) {
            const SolveCtrl& c = s->h_ctrl[i];
            const double rel = std::sqrt(std::max(0.0, c.rel2));
            it_max = std::max(it_max, c.iters);
            all_conv = all_conv && c.converged;
            any_stag = any_stag || (c.stagnated && !c.converged);
            rel_max = (rel == rel) ? std::max(rel_max, rel) : 1e300;
            if (!c.converged && !c.stagnated && failed < 0) failed = i;
        }
        s->last_iters = it_max;
        s->last_converged = all_conv;
        s->last_stagnated = any_stag;
        s->last_rel = rel_max;
        if (failed >= 0 && s->strict) {
            // leave the state as it was before the step and tell the caller
            RB_CUDA(cudaMemcpyAsync(st->y0, st->ybackup, 2 * s->BN * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
            if (s->props.guess_mode == RB_GUESS_WARM)
                RB_CUDA(cudaMemcpyAsync(st->d_counter, &st->h_counter, sizeof(int), cudaMemcpyHostToDevice, s->stream));
            RB_CUDA(cudaStreamSynchronize(s->stream));
            const SolveCtrl& c = s->h_ctrl[failed];
            note_solve_end(s, 0, 0, std::sqrt(std::max(0.0, c.rel2)), c.iters, "RK4 step (state restored)");
        }
        for (int i = 0; i < 4; ++i) {
            const SolveCtrl& c = s->h_ctrl[i];   // (a failed stage in strict mode has thrown above)
            note_solve_end(s, c.converged, c.stagnated, std::sqrt(std::max(0.0, c.rel2)), c.iters, "RK4 stage");
        }
        // shrink the recorded sweep count when it has been clearly too large for a while (each skipped sweep costs a launch)
        if (st->tight_ban > 0) st->tight_ban--;
        if (st->graph_sweeps - worst >= 3) {
            st->tight_hits = 0;
            if (++st->graph_hits_below >= 8) {
                st->graph_sweeps = worst + 1;
                invalidate_graphs(st);
            }
        } else {
            st->graph_hits_below = 0;
            // ... and drop the last surplus round once the count has been the same for 24 steps in a row
            if (st->tight_ok && !st->tight && st->tight_ban == 0 && st->graph_sweeps - worst >= 1 && worst >= (s->use_gmres ? 3 : 2) &&
                s->props.guess_mode == RB_GUESS_WARM) {
                if (++st->tight_hits >= 24) {
                    st->graph_sweeps = worst;
                    st->tight = true;
                    st->tight_hits = 0;
                    invalidate_graphs(st);
                }
            } else if (!st->tight) {
                st->tight_hits = 0;
            }
        }
    }
    if (s->props.guess_mode == RB_GUESS_WARM) st->h_counter++;
    after_step(st);
}

// m recorded steps launched back to back, ONE host synchronisation at the end: in the launch-bound regime (N <= 8192: a step is a
// few hundred microseconds) the host round trip after every step (event wait, status check, next launch) is ~5-10 % of the step.
// The last kernel of each recorded step folds its four solves' status into a device aggregate; if any step of the chunk ran out of
// recorded sweeps or failed, the whole chunk is rolled back (state, history counter; the history ring is deep enough that the
// repeated steps never read a slot the failed attempt overwrote) and the caller redoes it step by step.  Returns false when rolled back.
static bool stepper_chunk(rb_stepper* st, int m) {
    rb_solver* s = st->s;
    cudaStream_t cs = s->stream;
    const bool warm = s->props.guess_mode == RB_GUESS_WARM;
    const int mask = st->opt_mask;
    RB_CUDA(cudaMemcpyAsync(st->ycheck, st->y0, 2 * s->BN * sizeof(double2), cudaMemcpyDeviceToDevice, cs));
    RB_CUDA(cudaMemsetAsync(st->d_agg, 0, sizeof(StepAgg), cs));
    for (int j = 0; j < m; ++j) RB_CUDA(cudaGraphLaunch(st->graph_exec, cs));
    RB_CUDA(cudaMemcpyAsync(st->h_agg, st->d_agg, sizeof(StepAgg), cudaMemcpyDeviceToHost, cs));
    RB_CUDA(cudaEventRecord(st->ev, cs));
    RB_CUDA(cudaEventSynchronize(st->ev));
    rb::count_launch(m * st->graph_kernels[mask & 15]);
    st->chunks_launched++;
    const StepAgg& a = *st->h_agg;
    if (a.steps != m || a.not_done || a.failed) {
        RB_CUDA(cudaMemcpyAsync(st->y0, st->ycheck, 2 * s->BN * sizeof(double2), cudaMemcpyDeviceToDevice, cs));
        if (warm) RB_CUDA(cudaMemcpyAsync(st->d_counter, &st->h_counter, sizeof(int), cudaMemcpyHostToDevice, cs));
        RB_CUDA(cudaStreamSynchronize(cs));
        st->chunks_rolled_back++;
        return false;
    }
    st->graph_launches += m;
    s->sum_iters += a.sum_iters;
    s->num_solves += 4LL * m;
    s->stagnated_solves += a.stagnated;
    const double wr = std::sqrt(std::max(0.0, a.worst_rel2));
    if (wr == wr) s->worst_rel = std::max(s->worst_rel, wr);
    const double tol2 = s->props.tolerance * s->props.tolerance;
    int next_mask = 0, it_max = 0, all_conv = 1, any_stag = 0;
    double rel_max = 0.0;
    for (int i = 0; i < 4; ++i) {
        st->first_rel[i] = std::sqrt(std::max(0.0, a.first_rel2[i]));
        if (a.first_rel2[i] <= 4.0 * tol2) next_mask |= 1 << i;
        it_max = std::max(it_max, a.iters_last[i]);
        all_conv = all_conv && a.conv_last[i];
        any_stag = any_stag || (a.stag_last[i] && !a.conv_last[i]);
        rel_max = std::max(rel_max, std::sqrt(std::max(0.0, a.rel2_last[i])));
        if ((mask >> i) & 1) {   // (per-step counts are not kept inside a chunk: the last step stands for all of them)
            st->opt_stage_solves += m;
            if (a.iters_last[i] == 1) st->one_sweep_solves += m;
        }
    }
    if (st->opt_policy == 0) next_mask = 0;
    if (st->opt_policy == 2) next_mask = 15;
    st->opt_mask = next_mask;
    s->last_iters = it_max;
    s->last_converged = all_conv;
    s->last_stagnated = any_stag;
    s->last_rel = rel_max;
    if (st->graph_sweeps - a.max_occupied >= 3) {
        st->graph_hits_below += m;
        if (st->graph_hits_below >= 8) {
            st->graph_sweeps = a.max_occupied + 1;
            invalidate_graphs(st);
        }
    } else {
        st->graph_hits_below = 0;
    }
    if (warm) st->h_counter += m;
    st->t += m * st->dt;      // (same rounding as m single additions is not required: the time is bookkeeping only)
    st->step_index += m;
    return true;
}

// n steps: asynchronous chunks once the stepper has settled (stage history filled, recorded sweep count tuned), single steps otherwise
static void stepper_run(rb_stepper* st, size_t n) {
    rb_solver* s = st->s;
    size_t i = 0;
    while (i < n) {
        const bool graphable = st->use_graph && s->matrix_free_solve && (!s->use_gmres || s->gm_device);
        const bool settled = graphable && st->chunk >= 2 && !st->log_every && st->graph_launches >= 8 && st->graph_dt == st->dt &&
                             st->graph_y0 == st->y0 && st->graph_cache[st->opt_mask & 15] != nullptr && st->graph_hits_below == 0;
        const int m = (int)std::min<size_t>(st->chunk, n - i);
        if (settled && m >= 2) {
            st->graph_exec = st->graph_cache[st->opt_mask & 15];
            if (stepper_chunk(st, m)) {
                i += m;
                continue;
            }
            for (int j = 0; j < m; ++j) stepper_step(st);   // rolled back: redo these steps with the per-step checks and fallbacks
            i += m;
            continue;
        }
        stepper_step(st);
        ++i;
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* rb_last_error(void) { return g_last_error.c_str(); }
int rb_version(void) { return 100; }
unsigned long long rb_launch_count(void) { return rb::g_launch_count; }

int rb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rb_set_device(int device) {
    RB_TRY
    RB_CUDA(cudaSetDevice(device));
    RB_CATCH
}

void rb_default_props(rb_props* p) {
    std::memset(p, 0, sizeof(*p));
    p->rho = 0.0;
    p->U = 0.0;
    p->kappa = 0.0;
    p->depth = 1.0;
    p->expansion_order = 1;
    p->physics = RB_WATER;
    p->solve_mode = RB_SOLVE_MATRIX_FREE;
    p->guess_mode = RB_GUESS_COLD;
    p->max_iterations = 200;
    p->compute_energies = 0;
    p->tolerance = 1e-13;
}

rb_solver* rb_create(int N, int batch, const rb_props* props) {
    try {
        return solver_create(N, batch, props);
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

int rb_destroy(rb_solver* s) {
    RB_TRY
    if (s) {
        cudaSetDevice(s->device);
        cudaDeviceSynchronize();
        solver_free(s);
    }
    RB_CATCH
}

int rb_set_stream(rb_solver* s, void* cuda_stream) {
    RB_TRY
    set_stream(s, (cudaStream_t)cuda_stream);
    RB_CATCH
}

void* rb_get_stream(rb_solver* s) { return s ? (void*)s->stream : nullptr; }

int rb_get_props(rb_solver* s, rb_props* out, int* N, int* batch) {
    RB_TRY
    if (!s) throw std::runtime_error("rb_get_props: null solver");
    if (out) *out = s->props;
    if (N) *N = s->N;
    if (batch) *batch = s->batch;
    RB_CATCH
}

int rb_rhs(rb_solver* s, const rb_complex* state_dev, rb_complex* rhs_dev) {
    RB_TRY
    rhs(s, (const double2*)state_dev, (double2*)rhs_dev);
    RB_CATCH
}

int rb_vorticities(rb_solver* s, const rb_complex* state_dev) {
    RB_TRY
    vorticities(s, (const double2*)state_dev);
    RB_CATCH
}

double* rb_dev_a(rb_solver* s) { return s->a; }
rb_complex* rb_dev_zp(rb_solver* s) { return (rb_complex*)s->Zp(); }
rb_complex* rb_dev_zpp(rb_solver* s) { return (rb_complex*)s->Zpp(); }
rb_complex* rb_dev_velocities_upper(rb_solver* s) { return (rb_complex*)s->vel_upper; }
double* rb_dev_phi_prime(rb_solver* s) { return s->b; }

int rb_synchronize(rb_solver* s) {
    RB_TRY
    RB_CUDA(cudaStreamSynchronize(s->stream));
    RB_CATCH
}

int rb_energies(rb_solver* s, double out_host[5]) {
    RB_TRY
    if (!s->cur_Z || !s->cur_vel) throw std::runtime_error("rb_energies: no RHS has been evaluated yet");
    if (!s->props.compute_energies)
        launch_energies(s->cur_Z, s->Zp(), s->cur_Phi, s->cur_vel, s->energies, s->N, s->props.physics == RB_WATER ? 0 : 1,
                        s->props.rho, s->props.U, s->props.depth, s->props.kappa, s->stream);
    RB_CUDA(cudaMemcpyAsync(out_host, s->energies, 5 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(cudaStreamSynchronize(s->stream));
    RB_CATCH
}

int rb_solve_stats(rb_solver* s, double out_host[6]) {
    RB_TRY
    out_host[0] = s->last_iters;
    out_host[1] = s->last_converged;
    out_host[2] = s->last_rel;
    out_host[3] = (double)s->sum_iters;
    out_host[4] = (double)s->num_solves;
    out_host[5] = (double)s->vel_sweeps;
    RB_CATCH
}

int rb_solve_status(rb_solver* s, double out_host[8]) {
    RB_TRY
    if (!s) throw std::runtime_error("rb_solve_status: null solver");
    out_host[0] = s->last_converged;
    out_host[1] = s->last_stagnated;
    out_host[2] = (double)s->stagnated_solves;
    out_host[3] = (double)s->failed_solves;
    out_host[4] = s->worst_rel;
    out_host[5] = s->strict ? 1.0 : 0.0;
    out_host[6] = out_host[7] = 0.0;
    RB_CATCH
}

int rb_set_strict(rb_solver* s, int strict) {
    RB_TRY
    if (!s) throw std::runtime_error("rb_set_strict: null solver");
    s->strict = strict != 0;
    RB_CATCH
}

int rb_zphi_derivative(rb_solver* s, const rb_complex* Z, const rb_complex* Phi, rb_complex* Zp, rb_complex* PhiPrime,
                       rb_complex* Zpp) {
    RB_TRY
    derivatives(s, (const double2*)Z, (const double2*)Phi);
    const size_t bytes = s->BN * sizeof(double2);
    if (Zp) RB_CUDA(cudaMemcpyAsync(Zp, s->Zp(), bytes, cudaMemcpyDeviceToDevice, s->stream));
    if (Zpp) RB_CUDA(cudaMemcpyAsync(Zpp, s->Zpp(), bytes, cudaMemcpyDeviceToDevice, s->stream));
    if (PhiPrime) RB_CUDA(cudaMemcpyAsync(PhiPrime, s->PhiPc(), bytes, cudaMemcpyDeviceToDevice, s->stream));
    RB_CATCH
}

int rb_fft_derivative(rb_solver* s, const rb_complex* in, rb_complex* out, int second, double scaling) {
    RB_TRY
    fft_derivative(s, (const double2*)in, (double2*)out, second, scaling);
    RB_CATCH
}

int rb_create_M(double* A, const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, double rho, int n, size_t batch,
                void* stream) {
    RB_TRY
    launch_create_M(A, (const double2*)Z, (const double2*)Zp, (const double2*)Zpp, rho, n, batch, (cudaStream_t)stream);
    RB_CATCH
}

int rb_create_finite_depth_M(double* A, const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, double h, int n,
                             size_t batch, int infinite_depth, void* stream) {
    RB_TRY
    launch_create_finite_depth_M(A, (const double2*)Z, (const double2*)Zp, (const double2*)Zpp, h, n, batch,
                                 infinite_depth != 0, (cudaStream_t)stream);
    RB_CATCH
}

int rb_velocity_matrices(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, int n, rb_complex* V1,
                         rb_complex* V2, int lower, size_t batch, void* stream) {
    RB_TRY
    launch_velocity_matrices((const double2*)Z, (const double2*)Zp, (const double2*)Zpp, n, (double2*)V1, (double2*)V2,
                             lower != 0, batch, false, 0.0, true, (cudaStream_t)stream);
    RB_CATCH
}

int rb_helium_velocity_matrices(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, double h, int n,
                                rb_complex* V1, rb_complex* V2, int lower, size_t batch, int infinite_depth, void* stream) {
    RB_TRY
    launch_velocity_matrices((const double2*)Z, (const double2*)Zp, (const double2*)Zpp, n, (double2*)V1, (double2*)V2,
                             lower != 0, batch, true, h, infinite_depth != 0, (cudaStream_t)stream);
    RB_CATCH
}

int rb_rhs_phi_water(const rb_complex* Z, const rb_complex* V1, const rb_complex* V2, rb_complex* result, double rho, int n,
                     void* stream) {
    RB_TRY
    launch_rhs_phi_water((const double2*)Z, (const double2*)V1, (const double2*)V2, (double2*)result, rho, n,
                         (cudaStream_t)stream);
    RB_CATCH
}

int rb_rhs_phi_helium(const rb_complex* Z, const rb_complex* V1, rb_complex* result, double h, int n, void* stream) {
    RB_TRY
    launch_rhs_phi_helium((const double2*)Z, (const double2*)V1, (double2*)result, h, n, (cudaStream_t)stream);
    RB_CATCH
}

int rb_rhs_phi_helium_surface_tension(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, const rb_complex* V1,
                                      rb_complex* result, double h, double kappa, int n, void* stream) {
    RB_TRY
    launch_rhs_phi_helium_st((const double2*)Z, (const double2*)Zp, (const double2*)Zpp, (const double2*)V1, (double2*)result,
                             h, kappa, n, (cudaStream_t)stream);
    RB_CATCH
}

int rb_rhs_phi_helium_expansion(const rb_complex* Z, const rb_complex* V1, rb_complex* result, double h, int n, int order,
                                void* stream) {
    RB_TRY
    launch_rhs_phi_helium_exp((const double2*)Z, (const double2*)V1, (double2*)result, h, n, order, (cudaStream_t)stream);
    RB_CATCH
}

int rb_cotangent_sum(rb_solver* s, const rb_complex* Z_dev, const double* x_dev, rb_complex* S_dev) {
    RB_TRY
    if (s->has_image) throw std::runtime_error("rb_cotangent_sum: not defined for the finite-depth image operator");
    const double2* Z = (const double2*)Z_dev;
    // geometry from the solver's current Zp/Zpp (diagonal terms are not used by the raw sum)
    Geometry g = make_geometry(s, Z);
    launch_geometry(g, nullptr, s->N, s->batch, s->ncell, s->props.physics, s->rhoM, s->props.depth, 0, s->use_local, 0, 0.0, 0.0,
                    s->stream);
    launch_finish_solve(x_dev, x_dev, nullptr, nullptr, nullptr, s->xsum_a, HistoryRing(), s->N, s->batch, s->ncell, s->stream);
    SweepArgs a = base_args(s, Z);
    a.x = x_dev;
    a.xsum_part = s->xsum_a;
    a.raw_out = (double2*)S_dev;
    sweep(s, a, kSweepRAW);
    RB_CATCH
}

// ---- stepper -----------------------------------------------------------------------------------
rb_stepper* rb_rk4_create(rb_solver* s, double tstep) {
    try {
        if (!s) throw std::runtime_error("rb_rk4_create: null solver");
        std::unique_ptr<rb_stepper, void (*)(rb_stepper*)> up(new rb_stepper, stepper_free);
        rb_stepper* st = up.get();
        st->s = s;
        st->dt = tstep;
        const size_t n2 = 2 * s->BN;
        for (int i = 0; i < 4; ++i) st->k[i] = s->kbuf[i];   // in the solver's arena: peers publish their rows there
        st->ytmp = dmalloc<double2>(n2);
        st->ybackup = dmalloc<double2>(n2);
        for (auto& p : st->hist) {
            p = dmalloc<double>((size_t)3 * kHistRing * s->BN);   // ring of solutions a | ring of their row sums A (complex)
            RB_CUDA(cudaMemset(p, 0, (size_t)3 * kHistRing * s->BN * sizeof(double)));
        }
        st->d_counter = dmalloc<int>(1);
        RB_CUDA(cudaMemset(st->d_counter, 0, sizeof(int)));
        RB_CUDA(cudaEventCreateWithFlags(&st->ev, cudaEventDisableTiming));
        st->d_agg = dmalloc<StepAgg>(1);
        RB_CUDA(cudaMemset(st->d_agg, 0, sizeof(StepAgg)));
        RB_CUDA(cudaMallocHost(&st->h_agg, sizeof(StepAgg)));
        st->ycheck = dmalloc<double2>(n2);
        // extrapolation order of the stage history: 4 points wins where the truncation error of the guess dominates; at large N
        // the round-off noise of the spectral derivatives (~N eps) dominates and the wider stencil amplifies it (measured at
        // N = 65536: 2.00 sweeps per solve with 3 points, 2.10 with 4)
        st->order = std::max(1, std::min(6, env_int("RB_GUESS_ORDER", s->N >= 32768 ? 3 : 4)));
        {
            const int pr = env_int("RB_GUESS_PREDICT", -1);
            st->predict = pr >= 0 ? (pr != 0) : (s->props.tolerance >= 4e-13);
        }
        st->use_graph = env_int("RB_NO_GRAPH", 0) == 0;
        st->tight_ok = env_int("RB_TIGHT_GRAPH", 1) != 0;
        // measured on a B200 (profiles/r02c_async_chunks.log): 4135 vs 4086 steps/s at N = 1024, 2451 vs 2514 at N = 4096, 430 vs 433 at
        // N = 16384 -- the per-step host round trip is already hidden behind the recorded step, so the chunks are OFF unless asked for
        st->chunk = std::max(0, std::min(kChunkMax, std::min(env_int("RB_ASYNC_STEPS", 0), kHistRing - st->order)));
        st->opt_policy = std::max(0, std::min(2, env_int("RB_OPTIMISTIC", 1)));
        st->opt_mask = st->opt_policy == 2 ? 15 : 0;
        return up.release();
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

int rb_rk4_destroy(rb_stepper* st) {
    RB_TRY
    if (st) {
        cudaDeviceSynchronize();
        stepper_free(st);
    }
    RB_CATCH
}

int rb_rk4_set_time_step(rb_stepper* st, double tstep) {
    RB_TRY
    st->dt = tstep;
    stepper_reset_history(st);
    RB_CATCH
}

int rb_rk4_initialize(rb_stepper* st, rb_complex* y0, int on_device) {
    RB_TRY
    const size_t n2 = 2 * st->s->BN;
    if (on_device) {
        if (st->owns_y0 && st->y0) cudaFree(st->y0);
        st->y0 = (double2*)y0;   // caller keeps ownership, L/AutonomousRungeKuttaStepper.cuh:312-318
        st->owns_y0 = false;
    } else {
        if (!st->owns_y0 || !st->y0) st->y0 = dmalloc<double2>(n2);
        st->owns_y0 = true;
        RB_CUDA(cudaMemcpyAsync(st->y0, y0, n2 * sizeof(double2), cudaMemcpyHostToDevice, st->s->stream));
        RB_CUDA(cudaStreamSynchronize(st->s->stream));
    }
    stepper_reset_history(st);
    st->t = 0.0;
    st->step_index = 0;
    st->log_count = 0;
    st->log_times.clear();
    RB_CATCH
}

int rb_rk4_step(rb_stepper* st) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_rk4_step: initialize() has not been called");
    stepper_step(st);
    RB_CATCH
}

int rb_rk4_run_steps(rb_stepper* st, size_t steps) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_rk4_run_steps: initialize() has not been called");
    stepper_run(st, steps);
    RB_CATCH
}

int rb_rk4_evolve(rb_stepper* st, double t0, double t1, size_t* steps_out) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_rk4_evolve: initialize() has not been called");
    st->t = t0;
    size_t steps = static_cast<size_t>((t1 - t0) / st->dt);   // truncation, L/AutonomousRungeKuttaStepper.cuh:421
    stepper_run(st, steps);
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    if (steps_out) *steps_out = steps;
    RB_CATCH
}

rb_complex* rb_rk4_dev_state(rb_stepper* st) { return (rb_complex*)st->y0; }

int rb_rk4_stats(rb_stepper* st, double out_host[4]) {
    out_host[0] = (double)st->graph_launches;
    out_host[1] = (double)st->graph_captures;
    out_host[2] = (double)st->fallback_steps;
    out_host[3] = (double)st->graph_sweeps;
    return 0;
}
int rb_rk4_chunk_stats(rb_stepper* st, double out_host[4]) {
    out_host[0] = (double)st->chunk;
    out_host[1] = (double)st->chunks_launched;
    out_host[2] = (double)st->chunks_rolled_back;
    out_host[3] = (double)st->tight_failures + (st->tight ? 0.5 : 0.0);   // tightly recorded steps that had to be redone (+ 0.5 while tight)
    return 0;
}
int rb_rk4_guess_stats(rb_stepper* st, double out_host[8]) {
    for (int i = 0; i < 4; ++i) out_host[i] = st->first_rel[i];
    out_host[4] = (double)st->opt_mask;
    out_host[5] = (double)st->opt_stage_solves;
    out_host[6] = (double)st->one_sweep_solves;
    out_host[7] = (double)st->opt_policy;
    return 0;
}
int rb_rk4_set_optimistic(rb_stepper* st, int policy) {
    if (policy < 0 || policy > 2) return -1;
    st->opt_policy = policy;
    st->opt_mask = policy == 2 ? 15 : 0;
    return 0;
}
int rb_rk4_set_guess(rb_stepper* st, int order, int predict) {
    RB_TRY
    if (order < 1 || order > 6) throw std::runtime_error("rb_rk4_set_guess: order must be in 1..6");
    st->order = order;
    st->chunk = std::max(0, std::min(st->chunk, kHistRing - order));
    st->predict = predict < 0 ? (st->s->props.tolerance >= 4e-13) : (predict != 0);
    stepper_reset_history(st);   // the rings of the two modes hold different iterates
    invalidate_graphs(st);
    RB_CATCH
}
double rb_rk4_current_time(rb_stepper* st) { return st->t; }

int rb_rk4_get_state(rb_stepper* st, rb_complex* y_host) {
    RB_TRY
    const size_t n2 = 2 * st->s->BN;
    RB_CUDA(cudaMemcpyAsync(y_host, st->y0, n2 * sizeof(double2), cudaMemcpyDeviceToHost, st->s->stream));
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    RB_CATCH
}

int rb_rk4_set_logging(rb_stepper* st, size_t every, size_t capacity) {
    RB_TRY
    if (st->log_states) {
        cudaFree(st->log_states);
        st->log_states = nullptr;
    }
    st->log_every = every;
    st->log_capacity = capacity;
    st->log_count = 0;
    st->log_times.clear();
    if (every && capacity) st->log_states = dmalloc<double2>(capacity * 2 * st->s->BN);
    RB_CATCH
}

int rb_rk4_copy_trajectory(rb_stepper* st, double** times_out, size_t* times_count, rb_complex** states_out,
                           size_t* states_count) {
    RB_TRY
    const size_t n2 = 2 * st->s->BN;
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    size_t cnt = st->log_count;
    if (times_out) {
        *times_out = (double*)std::malloc(std::max<size_t>(cnt, 1) * sizeof(double));
        std::memcpy(*times_out, st->log_times.data(), cnt * sizeof(double));
    }
    if (times_count) *times_count = cnt;
    if (states_out) {
        *states_out = (rb_complex*)std::malloc(std::max<size_t>(cnt * n2, 1) * sizeof(rb_complex));
        if (cnt) RB_CUDA(cudaMemcpy(*states_out, st->log_states, cnt * n2 * sizeof(double2), cudaMemcpyDeviceToHost));
    }
    if (states_count) *states_count = cnt;
    RB_CATCH
}

void rb_free(void* p) { std::free(p); }

int rb_rk4_stage_update(rb_complex* y_out, const rb_complex* y0, const rb_complex* k, double c, size_t n, void* stream) {
    RB_TRY
    launch_stage_update((double2*)y_out, (const double2*)y0, (const double2*)k, c, n, (cudaStream_t)stream);
    RB_CATCH
}

int rb_rk4_final_update(rb_complex* y0, const rb_complex* k1, const rb_complex* k2, const rb_complex* k3, const rb_complex* k4,
                        double h, size_t n, void* stream) {
    RB_TRY
    launch_final_update((double2*)y0, (const double2*)k1, (const double2*)k2, (const double2*)k3, (const double2*)k4, h, n,
                        (cudaStream_t)stream);
    RB_CATCH
}

// ---- multi-GPU: row cells of every O(N^2) sweep sharded over the ranks of one node ------------------------------------
int rb_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int rb_comm_row_range(int N, int rank, int nranks, int out_rows[2]) {
    // contiguous blocks of whole 256-row cells; host-only arithmetic (no device needed)
    if (N < 2 || nranks < 1 || rank < 0 || rank >= nranks) return -1;
    const int ncell = (N + kCell - 1) / kCell;
    const int per = (ncell + nranks - 1) / nranks;
    const int c0 = std::min(rank * per, ncell);
    const int c1 = std::min(c0 + per, ncell);
    out_rows[0] = std::min(c0 * kCell, N);
    out_rows[1] = std::min(c1 * kCell, N);
    return 0;
}

int rb_comm_export(rb_solver* s, char* handle_out) {
    RB_TRY
    cudaIpcMemHandle_t h;
    RB_CUDA(cudaIpcGetMemHandle(&h, s->arena));
    std::memcpy(handle_out, &h, sizeof(h));
    RB_CATCH
}

int rb_comm_init(rb_solver* s, int rank, int nranks, const char* handles) {
    RB_TRY
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) throw std::runtime_error("rb_comm_init: bad rank / nranks");
    if (s->batch != 1) throw std::runtime_error("rb_comm_init: row sharding is for batch == 1; ensembles are replicated per rank");
    if (!s->matrix_free_solve) throw std::runtime_error("rb_comm_init: row sharding needs the matrix-free solve");
    if (s->ncell < nranks) throw std::runtime_error("rb_comm_init: N too small to give every rank a 256-row cell");
    RB_CUDA(cudaStreamSynchronize(s->stream));
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            s->comm.peer_base[r] = s->arena;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        RB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_mapped[r] = p;
        s->comm.peer_base[r] = static_cast<char*>(p);
    }
    s->comm.nranks = nranks;
    s->comm.rank = rank;
    int rows[2];
    rb_comm_row_range(s->N, rank, nranks, rows);
    s->row_cell0 = rows[0] / kCell;
    s->row_cells = (rows[1] - rows[0] + kCell - 1) / kCell;
    if (s->row_cells < 1) throw std::runtime_error("rb_comm_init: this rank owns no rows");
    // the sweep's grid now covers the local rows only: re-balance the schedules and the partial workspace
    plan_sweep2(s);
    choose_sweep_kernel(s);
    choose_chunking(s);
    alloc_partials(s);
    RB_CATCH
}

// measurement aid: restrict the sweeps of a single-GPU solver to the row cells [cell0, cell0 + cells) a rank of a row-sharded run
// would own, without any peer (the other rows of the iterate simply stay as they are): the per-rank sweep of G ranks can be timed
// and tuned on one GPU with rb_bench_sweep.  cells <= 0 restores the whole surface.
int rb_debug_set_row_range(rb_solver* s, int cell0, int cells) {
    RB_TRY
    if (s->comm.nranks > 1) throw std::runtime_error("rb_debug_set_row_range: the solver is part of a row-sharded run");
    if (cells <= 0) {
        cell0 = 0;
        cells = s->ncell;
    }
    if (cell0 < 0 || cell0 + cells > s->ncell) throw std::runtime_error("rb_debug_set_row_range: range outside the surface");
    RB_CUDA(cudaStreamSynchronize(s->stream));
    s->row_cell0 = cell0;
    s->row_cells = cells;
    plan_sweep2(s);
    choose_sweep_kernel(s);
    choose_chunking(s);
    alloc_partials(s);
    RB_CATCH
}

int rb_sweep_plan(rb_solver* s, int out[8]) {
    RB_TRY
    out[0] = s->use_v2 ? 2 : 1;            // 1 tiled, 2 persistent
    out[1] = s->use_v2 ? s->v2_R : s->v1_rows;
    out[2] = s->tile;
    out[3] = s->tiles_per_chunk;
    out[4] = s->nchunks;
    out[5] = s->row_cells;
    out[6] = s->use_v2 ? s->v2l.grid : s->row_cells * s->nchunks * s->batch;   // CTAs per sweep
    out[7] = s->use_v2 ? s->v2l.threads : kCell / s->v1_rows;
    RB_CATCH
}

int rb_comm_error(rb_solver* s) {
    int e = 0;
    if (cudaMemcpy(&e, s->comm.error_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}

int rb_comm_destroy(rb_solver* s) {
    RB_TRY
    RB_CUDA(cudaStreamSynchronize(s->stream));
    for (int r = 0; r < kMaxRanks; ++r)
        if (s->peer_mapped[r]) {
            cudaIpcCloseMemHandle(s->peer_mapped[r]);
            s->peer_mapped[r] = nullptr;
        }
    s->comm.nranks = 1;
    s->comm.rank = 0;
    s->comm.peer_base[0] = s->arena;
    s->row_cell0 = 0;
    s->row_cells = s->ncell;
    plan_sweep2(s);
    choose_sweep_kernel(s);
    choose_chunking(s);
    alloc_partials(s);
    RB_CATCH
}

// ---- measurement -------------------------------------------------------------------------------
int rb_measure_fp64_peak(double* tflops_out, void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    launch_fp64_peak(sink, 256, blocks, st);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        RB_CUDA(cudaEventRecord(e0, st));
        launch_fp64_peak(sink, iters, blocks, st);
        RB_CUDA(cudaEventRecord(e1, st));
        RB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    double flops = (double)blocks * 256.0 * iters * 64.0 * 2.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

// DMMA (m8n8k4, 512 flop per warp instruction) beside DFMA (64 flop per warp instruction): out[2*i] = ms, out[2*i+1] = TFLOP/s of
// mix i in {8 mma, 32 fma, 8+32, 4+32, 2+32, 1+32} per loop iteration
int rb_measure_fp64_tensor_overlap(double out_host[12], void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 2048;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    const int mixes[6][2] = {{8, 0}, {0, 32}, {8, 32}, {4, 32}, {2, 32}, {1, 32}};
    for (int m = 0; m < 6; ++m) {
        launch_fp64_mix(sink, 64, blocks, mixes[m][0], mixes[m][1], st);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            RB_CUDA(cudaEventRecord(e0, st));
            launch_fp64_mix(sink, iters, blocks, mixes[m][0], mixes[m][1], st);
            RB_CUDA(cudaEventRecord(e1, st));
            RB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        const double warps = (double)blocks * 8.0;
        const double flops = warps * iters * (mixes[m][0] * 512.0 + mixes[m][1] * 64.0);
        out_host[2 * m] = best;
        out_host[2 * m + 1] = flops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

int rb_measure_fp64_rate_3operand(double* tflops_out, void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    launch_fp64_peak3(sink, 256, blocks, 1e-9, st);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        RB_CUDA(cudaEventRecord(e0, st));
        launch_fp64_peak3(sink, iters, blocks, 1e-9, st);
        RB_CUDA(cudaEventRecord(e1, st));
        RB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    *tflops_out = (double)blocks * 256.0 * iters * 64.0 * 2.0 / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

int rb_bench_sweep(rb_solver* s, const rb_complex* state_dev, int reps, float* ms_per_sweep_out, double* pairs_per_sweep_out) {
    RB_TRY
    cudaStream_t st = s->stream;
    const double2* Z = (const double2*)state_dev;
    surface_stage(s, Z, Z + s->BN);
    launch_guess(s->b, nullptr, HistoryRing(), s->xbuf[0], s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch,
                 s->ncell, st);
    SweepArgs base = base_args(s, Z);
    base.max_iters = 1 << 30;
    base.tol2 = 0.0;
    for (int i = 0; i < 3; ++i) launch_mv(s, base, i, 0);
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    RB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) launch_mv(s, base, i + 1, 0);
    RB_CUDA(cudaEventRecord(e1, st));
    RB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_sweep_out) *ms_per_sweep_out = ms / reps;
    if (pairs_per_sweep_out) *pairs_per_sweep_out = (double)s->N * s->N * s->batch * (s->has_image ? 2.0 : 1.0);
    RB_CATCH
}

// ---- legacy exports ----------------------------------------------------------------------------
static const double kAlphaHamaker = 3.5e-24;   // L/constants.cuh:11

struct Adim {
    double base_length, base_acceleration, base_time, base_energy, kappa, depth, rho;
};

// adimensionalizeProperties, L/Export.cu:1222-1246 (the stdout prints of the reference are dropped)
static Adim adimensionalize(double L, double rho, double kappa, double depth, double rhoHelium = 150.0) {
    Adim a;
    a.base_length = L / (2.0 * kPi);
    a.base_acceleration = 3 * kAlphaHamaker / std::pow(depth, 4);
    a.base_time = std::sqrt(a.base_length / a.base_acceleration);
    a.base_energy = 3.0 * rhoHelium * kAlphaHamaker * std::pow(a.base_length, 4) / std::pow(depth, 4);   // L/Export.cu:1228
    double surfaceTensionFactor = rhoHelium * a.base_length * a.base_length * a.base_length / (a.base_time * a.base_time);
    a.kappa = kappa / surfaceTensionFactor;
    a.depth = depth / a.base_length;
    a.rho = rho / rhoHelium;
    return a;
}

static rb_props helium_props(const Adim& ad, bool use_expansions, int expansion_order, bool infinite_depth) {
    rb_props p;
    rb_default_props(&p);
    p.physics = RB_HELIUM;   // every RHS export of the reference instantiates HeliumBoundaryProblem, L/Export.cu:207
    p.rho = ad.rho;
    p.kappa = ad.kappa;
    p.depth = ad.depth;
    p.use_expansions = use_expansions;
    p.expansion_order = expansion_order;
    p.infinite_depth = infinite_depth;
    return p;
}

static int rhs_from_vectors(const double* x, const double* y, const double* phi, double* vx, double* vy, double* rhsPhi,
                            double L, double rho, double kappa, double depth, size_t N, size_t batch) {
    RB_TRY
    Adim ad = adimensionalize(L, rho, kappa, depth);
    rb_props p = helium_props(ad, false, 1, false);
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, (int)batch, &p), solver_free);
    const size_t BN = N * batch;
    std::vector<double2> host(2 * BN);   // loadDataToDevice packing, L/SimulationRunner.cuh:180-242
    for (size_t i = 0; i < BN; ++i) {
        host[i] = make_double2(x[i], y[i]);
        host[BN + i] = make_double2(phi[i], 0.0);
    }
    double2* dstate = dmalloc<double2>(4 * BN);
    double2* drhs = dstate + 2 * BN;
    RB_CUDA(cudaMemcpy(dstate, host.data(), 2 * BN * sizeof(double2), cudaMemcpyHostToDevice));
    rhs(s.get(), dstate, drhs);
    RB_CUDA(cudaMemcpy(host.data(), drhs, 2 * BN * sizeof(double2), cudaMemcpyDeviceToHost));
    cudaFree(dstate);
    for (size_t i = 0; i < BN; ++i) {
        vx[i] = host[i].x;
        vy[i] = host[i].y;
        rhsPhi[i] = host[BN + i].x;
    }
    RB_CATCH
}

int calculateRHSFromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy, double* rhsPhi,
                            double L, double rho, double kappa, double depth, size_t N) {
    return rhs_from_vectors(x, y, phi, vx, vy, rhsPhi, L, rho, kappa, depth, N, 1);
}
int calculateRHS256FromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy, double* rhsPhi,
                               double L, double rho, double kappa, double depth) {
    return rhs_from_vectors(x, y, phi, vx, vy, rhsPhi, L, rho, kappa, depth, 256, 1);
}
int calculateRHS2048FromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy, double* rhsPhi,
                                double L, double rho, double kappa, double depth) {
    return rhs_from_vectors(x, y, phi, vx, vy, rhsPhi, L, rho, kappa, depth, 2048, 1);
}
int calculateRHS256FromVectorsBatched(const double* x, const double* y, const double* phi, double* vx, double* vy,
                                      double* rhsPhi, double L, double rho, double kappa, double depth, int batchSize) {
    if (batchSize < 1) {
        g_last_error = "calculateRHS256FromVectorsBatched: batchSize must be >= 1";
        return -1;
    }
    return rhs_from_vectors(x, y, phi, vx, vy, rhsPhi, L, rho, kappa, depth, 256, (size_t)batchSize);
}

int calculateVorticities256FromVectors(const rb_complex* Z, const rb_complex* phi, double* a, rb_complex* Zp, rb_complex* Zpp,
                                       double L, double rho, double kappa, double depth) {
    RB_TRY
    const size_t N = 256;
    Adim ad = adimensionalize(L, rho, kappa, depth);
    rb_props p = helium_props(ad, false, 1, false);
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    double2* dstate = dmalloc<double2>(2 * N);
    RB_CUDA(cudaMemcpy(dstate, Z, N * sizeof(double2), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(dstate + N, phi, N * sizeof(double2), cudaMemcpyHostToDevice));
    vorticities(s.get(), dstate);
    RB_CUDA(cudaMemcpy(a, s->a, N * sizeof(double), cudaMemcpyDeviceToHost));
    if (Zp) RB_CUDA(cudaMemcpy(Zp, s->Zp(), N * sizeof(double2), cudaMemcpyDeviceToHost));
    if (Zpp) RB_CUDA(cudaMemcpy(Zpp, s->Zpp(), N * sizeof(double2), cudaMemcpyDeviceToHost));
    cudaFree(dstate);
    RB_CATCH
}

int calculateDerivativeFFT256(const rb_complex* input, rb_complex* output) {
    RB_TRY
    const size_t N = 256;
    rb_props p;
    rb_default_props(&p);
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    double2* d = dmalloc<double2>(2 * N);
    RB_CUDA(cudaMemcpy(d, input, N * sizeof(double2), cudaMemcpyHostToDevice));
    fft_derivative(s.get(), d, d + N, 0, 1.0);   // L/Export.cu: FftDerivative<256,1>::exec(in, out)
    RB_CUDA(cudaMemcpy(output, d + N, N * sizeof(double2), cudaMemcpyDeviceToHost));
    cudaFree(d);
    RB_CATCH
}

static void integrate_host(const double* initialState, size_t N, size_t batch, const rb_props& p, double dt, size_t steps,
                           bool trajectory, std::vector<double>& states, std::vector<double>& times, double t0) {
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, (int)batch, &p), solver_free);
    std::unique_ptr<rb_stepper, void (*)(rb_stepper*)> st(rb_rk4_create(s.get(), dt), stepper_free);
    if (!st) throw std::runtime_error(g_last_error);
    const size_t BN = N * batch;
    std::vector<double2> host(2 * BN);
    for (size_t i = 0; i < BN; ++i) {
        host[i] = make_double2(initialState[i], initialState[BN + i]);
        host[BN + i] = make_double2(initialState[2 * BN + i], 0.0);
    }
    if (rb_rk4_initialize(st.get(), (rb_complex*)host.data(), 0) != 0) throw std::runtime_error(g_last_error);
    st->t = t0;
    auto unpack = [&](const double2* y, double* out) {
        for (size_t i = 0; i < BN; ++i) {
            out[i] = y[i].x;
            out[BN + i] = y[i].y;
            out[2 * BN + i] = y[BN + i].x;
        }
    };
    if (trajectory) {
        if (rb_rk4_set_logging(st.get(), 1, steps) != 0) throw std::runtime_error(g_last_error);
        for (size_t i = 0; i < steps; ++i) stepper_step(st.get());
        RB_CUDA(cudaStreamSynchronize(s->stream));
        std::vector<double2> all(st->log_count * 2 * BN);
        if (st->log_count)
            RB_CUDA(cudaMemcpy(all.data(), st->log_states, all.size() * sizeof(double2), cudaMemcpyDeviceToHost));
        states.resize(st->log_count * 3 * BN);
        for (size_t r = 0; r < st->log_count; ++r) unpack(all.data() + r * 2 * BN, states.data() + r * 3 * BN);
        times = st->log_times;
    } else {
        stepper_run(st.get(), steps);
        if (rb_rk4_get_state(st.get(), (rb_complex*)host.data()) != 0) throw std::runtime_error(g_last_error);
        states.resize(3 * BN);
        unpack(host.data(), states.data());
        times.clear();
    }
}

int integrateSimulationRK4(double* initialState, double** statesOut, size_t* statesCount, double** timesOut, size_t* timesCount,
                           SimProperties* simProperties, RK4SolverOptions* rkOptions, size_t N) {
    RB_TRY
    if (!initialState || !statesOut || !statesCount || !simProperties || !rkOptions)
        throw std::runtime_error("integrateSimulationRK4: null argument");
    Adim ad = adimensionalize(simProperties->L, simProperties->rho, simProperties->kappa, simProperties->depth);
    rb_props p = helium_props(ad, simProperties->use_expansions, simProperties->expansion_order, simProperties->infinite_depth);
    p.guess_mode = RB_GUESS_WARM;
    // adimensionalizeRK4SolverOptions, L/Export.cu:1213-1220
    const double dt = rkOptions->timeStep / ad.base_time, t0 = rkOptions->t0 / ad.base_time, t1 = rkOptions->t1 / ad.base_time;
    const size_t steps = static_cast<size_t>((t1 - t0) / dt);
    std::vector<double> states, times;
    integrate_host(initialState, N, 1, p, dt, steps, rkOptions->returnTrajectory, states, times, t0);
    double* so = (double*)std::malloc(std::max<size_t>(states.size(), 1) * sizeof(double));
    std::memcpy(so, states.data(), states.size() * sizeof(double));
    *statesOut = so;
    *statesCount = states.size() / (3 * N);
    if (timesOut) {
        double* to = (double*)std::malloc(std::max<size_t>(times.size(), 1) * sizeof(double));
        std::memcpy(to, times.data(), times.size() * sizeof(double));
        *timesOut = to;
    }
    if (timesCount) *timesCount = times.size();
    RB_CATCH
}

int integrateSimulationRK4_freeMemory(double* statesOut, double* timesOut) {
    std::free(statesOut);
    std::free(timesOut);
    return 0;
}

int rb_integrate_rk4_host(const double* initialState_host, double* finalState_host, size_t N, size_t batch,
                          const rb_props* props, double dt, size_t steps) {
    RB_TRY
    rb_props p;
    if (props) p = *props; else rb_default_props(&p);
    std::vector<double> states, times;
    integrate_host(initialState_host, N, batch, p, dt, steps, false, states, times, 0.0);
    std::memcpy(finalState_host, states.data(), states.size() * sizeof(double));
    RB_CATCH
}

}  // extern "C"

// services for implicit.cu (internal.cuh)
namespace rb {
int report_error(const std::exception& e) { return fail(e); }
rb_props helium_props_from_si(double L, double rho, double kappa, double depth, bool use_expansions, int expansion_order,
                              bool infinite_depth) {
    return helium_props(adimensionalize(L, rho, kappa, depth), use_expansions, expansion_order, infinite_depth);
}
}  // namespace rb


// ------------------------------------------------------------------------------------------------
// optomechanically driven film: the autonomous augmented system y = [Z | Phi | D] (drive_kernels.cu) and its classical RK4 stepper
// (AugmentedBoundaryIntegrator + AutonomousRungeKuttaStepper<std_complex, 3N>, A/kernel.cu:85-96, L/Export.cu:980-1209)
// ------------------------------------------------------------------------------------------------
static const double kHbar = 1.054571817e-34;   // L/constants.cuh:10

struct rb_aug_stepper {
    rb_solver* s = nullptr;
    rb_opto v;
    double dt = 1e-2;
    double t = 0.0;
    double2* y0 = nullptr;
    bool owns_y0 = false;
    double2* k[4] = {nullptr, nullptr, nullptr, nullptr};
    double2* ytmp = nullptr;
};

static void aug_rhs(rb_solver* s, const rb_opto& v, const double2* state, double2* out) {
    rhs(s, state, out);                                         // m_integrator->run, driven problem's base dPhi/dt
    launch_augmented_terms(state, out, v, s->BN, s->stream);    // drive + damping, then m_delayedIntensityIntegrator->run
}

static void aug_stepper_free(rb_aug_stepper* st) {
    if (!st) return;
    if (st->owns_y0 && st->y0) cudaFree(st->y0);
    for (auto& k : st->k)
        if (k) cudaFree(k);
    if (st->ytmp) cudaFree(st->ytmp);
    delete st;
}

// Y1 = Y0 + h/2 k1; Y2 = Y0 + h/2 k2; Y3 = Y0 + h k3; Y0 += h/6 (k1 + 2 k2 + 2 k3 + k4), L/AutonomousRungeKuttaStepper.cuh:124-307
static void aug_step(rb_aug_stepper* st) {
    rb_solver* s = st->s;
    const size_t n = 3 * s->BN;
    const double h = st->dt;
    aug_rhs(s, st->v, st->y0, st->k[0]);
    launch_stage_update(st->ytmp, st->y0, st->k[0], 0.5 * h, n, s->stream);
    aug_rhs(s, st->v, st->ytmp, st->k[1]);
    launch_stage_update(st->ytmp, st->y0, st->k[1], 0.5 * h, n, s->stream);
    aug_rhs(s, st->v, st->ytmp, st->k[2]);
    launch_stage_update(st->ytmp, st->y0, st->k[2], h, n, s->stream);
    aug_rhs(s, st->v, st->ytmp, st->k[3]);
    launch_final_update(st->y0, st->k[0], st->k[1], st->k[2], st->k[3], h, n, s->stream);
    st->t += h;
}

// adimensionalizeOptomechanicalVariables, L/Export.cu:1250-1275 (properties already nondimensional: rho = rho / rhoHelium)
static rb_opto adimensionalize_opto(const COptomechanicalVariables& c, double base_length, double base_time, double base_energy,
                                    double rho_adim) {
    rb_opto v;
    v.detuning = c.detuning * base_time;
    v.gamma = c.gamma * base_time;
    v.G = c.G * base_time * base_length;
    v.Tau = c.tau / base_time;
    v.max_intensity = c.max_intensity;
    v.initial_time = c.initial_time;
    v.location_x0_mode = c.location_x0_mode / base_length;
    v.sigma_optical_mode = c.sigma_optical_mode / base_length;
    const double hbar_adim = kHbar / base_energy / base_time;
    v.Beta = c.beta * (hbar_adim * v.G / (v.Tau) / (v.sigma_optical_mode * v.sigma_optical_mode * rho_adim));
    v.DampingStrength = c.damping_strength;
    v.drive_strength = rb_opto_drive_strength(&v, base_energy, base_time, rho_adim);
    return v;
}

static void aug_integrate_host(const double* initialState, size_t N, const rb_props& p, const rb_opto& v, double dt, size_t steps,
                               bool trajectory, double t0, std::vector<double>& states, std::vector<double>& times) {
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    std::unique_ptr<rb_aug_stepper, void (*)(rb_aug_stepper*)> st(rb_aug_rk4_create(s.get(), &v, dt), aug_stepper_free);
    if (!st) throw std::runtime_error(g_last_error);
    std::vector<double2> host(3 * N);
    for (size_t i = 0; i < N; ++i) {
        host[i] = make_double2(initialState[i], initialState[N + i]);
        host[N + i] = make_double2(initialState[2 * N + i], 0.0);
        host[2 * N + i] = make_double2(initialState[3 * N + i], 0.0);
    }
    if (rb_aug_rk4_initialize(st.get(), (rb_complex*)host.data(), 0) != 0) throw std::runtime_error(g_last_error);
    st->t = t0;
    auto unpack = [&](double* out) {
        if (rb_aug_rk4_get_state(st.get(), (rb_complex*)host.data()) != 0) throw std::runtime_error(g_last_error);
        for (size_t i = 0; i < N; ++i) {
            out[i] = host[i].x;
            out[N + i] = host[i].y;
            out[2 * N + i] = host[N + i].x;
            out[3 * N + i] = host[2 * N + i].x;
        }
    };
    states.clear();
    times.clear();
    for (size_t i = 0; i < steps; ++i) {
        aug_step(st.get());
        if (trajectory) {   // TrajectoryLogger::logTrajectory after every step, L/AutonomousRungeKuttaStepper.cuh:426-428
            states.resize(states.size() + 4 * N);
            unpack(states.data() + states.size() - 4 * N);
            times.push_back(st->t);
        }
    }
    if (!trajectory) {
        states.resize(4 * N);
        unpack(states.data());
    }
}

extern "C" {

void rb_default_opto(rb_opto* v) {
    std::memset(v, 0, sizeof(*v));
    v->gamma = 1.0;
    v->G = 1.0;
    v->Tau = 1.0;
    v->sigma_optical_mode = 1.0;
    v->DampingStrength = 0.01;
}

double rb_opto_drive_strength(const rb_opto* v, double base_energy, double base_time, double rho) {
    return kHbar / (base_energy * base_time * rho) * v->G / (v->sigma_optical_mode * v->sigma_optical_mode);
}

int rb_light_intensity(const rb_complex* Z_dev, double* intensity_dev, const rb_opto* v, size_t n, void* stream) {
    RB_TRY
    launch_light_intensity((const double2*)Z_dev, intensity_dev, *v, n, (cudaStream_t)stream);
    RB_CATCH
}

int rb_augmented_rhs(rb_solver* s, const rb_opto* v, const rb_complex* state_dev, rb_complex* rhs_dev) {
    RB_TRY
    if (!s || !v) throw std::runtime_error("rb_augmented_rhs: null argument");
    aug_rhs(s, *v, (const double2*)state_dev, (double2*)rhs_dev);
    RB_CATCH
}

rb_aug_stepper* rb_aug_rk4_create(rb_solver* s, const rb_opto* v, double tstep) {
    try {
        if (!s || !v) throw std::runtime_error("rb_aug_rk4_create: null argument");
        std::unique_ptr<rb_aug_stepper, void (*)(rb_aug_stepper*)> st(new rb_aug_stepper, aug_stepper_free);
        st->s = s;
        st->v = *v;
        st->dt = tstep;
        const size_t n = 3 * s->BN;
        for (auto& k : st->k) k = dmalloc<double2>(n);
        st->ytmp = dmalloc<double2>(n);
        return st.release();
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

int rb_aug_rk4_destroy(rb_aug_stepper* st) {
    RB_TRY
    if (st) {
        cudaDeviceSynchronize();
        aug_stepper_free(st);
    }
    RB_CATCH
}

int rb_aug_rk4_set_time_step(rb_aug_stepper* st, double tstep) {
    st->dt = tstep;
    return 0;
}

int rb_aug_rk4_initialize(rb_aug_stepper* st, rb_complex* y0, int on_device) {
    RB_TRY
    const size_t n = 3 * st->s->BN;
    if (on_device) {
        if (st->owns_y0 && st->y0) cudaFree(st->y0);
        st->y0 = (double2*)y0;   // caller keeps ownership, L/AutonomousRungeKuttaStepper.cuh:312-318
        st->owns_y0 = false;
    } else {
        if (!st->owns_y0 || !st->y0) st->y0 = dmalloc<double2>(n);
        st->owns_y0 = true;
        RB_CUDA(cudaMemcpyAsync(st->y0, y0, n * sizeof(double2), cudaMemcpyHostToDevice, st->s->stream));
        RB_CUDA(cudaStreamSynchronize(st->s->stream));
    }
    st->t = 0.0;
    RB_CATCH
}

int rb_aug_rk4_step(rb_aug_stepper* st) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_aug_rk4_step: initialize() has not been called");
    aug_step(st);
    RB_CATCH
}

int rb_aug_rk4_run_steps(rb_aug_stepper* st, size_t steps) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_aug_rk4_run_steps: initialize() has not been called");
    for (size_t i = 0; i < steps; ++i) aug_step(st);
    RB_CATCH
}

int rb_aug_rk4_evolve(rb_aug_stepper* st, double t0, double t1, size_t* steps_out) {
    RB_TRY
    if (!st->y0) throw std::runtime_error("rb_aug_rk4_evolve: initialize() has not been called");
    st->t = t0;
    const size_t steps = static_cast<size_t>((t1 - t0) / st->dt);   // truncation, L/AutonomousRungeKuttaStepper.cuh:421
    for (size_t i = 0; i < steps; ++i) aug_step(st);
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    if (steps_out) *steps_out = steps;
    RB_CATCH
}

rb_complex* rb_aug_rk4_dev_state(rb_aug_stepper* st) { return (rb_complex*)st->y0; }

int rb_aug_rk4_get_state(rb_aug_stepper* st, rb_complex* y_host) {
    RB_TRY
    const size_t n = 3 * st->s->BN;
    RB_CUDA(cudaMemcpyAsync(y_host, st->y0, n * sizeof(double2), cudaMemcpyDeviceToHost, st->s->stream));
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    RB_CATCH
}

double rb_aug_rk4_current_time(rb_aug_stepper* st) { return st->t; }

int calculateRhsAugmentedOptomechanical(double* state, double* rhs_out, SimProperties* simProperties,
                                        COptomechanicalVariables* optomechanicalVariables, size_t N) {
    RB_TRY
    if (!state || !rhs_out || !simProperties || !optomechanicalVariables)
        throw std::runtime_error("calculateRhsAugmentedOptomechanical: null argument");
    Adim ad = adimensionalize(simProperties->L, simProperties->rho, simProperties->kappa, simProperties->depth);
    rb_props p = helium_props(ad, simProperties->use_expansions, simProperties->expansion_order, simProperties->infinite_depth);
    rb_opto v = adimensionalize_opto(*optomechanicalVariables, ad.base_length, ad.base_time, ad.base_energy, ad.rho);
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    std::vector<double2> host(3 * N);
    for (size_t i = 0; i < N; ++i) {
        host[i] = make_double2(state[i], state[N + i]);
        host[N + i] = make_double2(state[2 * N + i], 0.0);
        host[2 * N + i] = make_double2(state[3 * N + i], 0.0);
    }
    double2* d = dmalloc<double2>(6 * N);
    RB_CUDA(cudaMemcpyAsync(d, host.data(), 3 * N * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
    aug_rhs(s.get(), v, d, d + 3 * N);
    RB_CUDA(cudaMemcpyAsync(host.data(), d + 3 * N, 3 * N * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d);
    for (size_t i = 0; i < N; ++i) {
        rhs_out[i] = host[i].x;
        rhs_out[N + i] = host[i].y;
        rhs_out[2 * N + i] = host[N + i].x;
        rhs_out[3 * N + i] = host[2 * N + i].x;
    }
    RB_CATCH
}

int integrateAugmentedOptomechanicalSimulationRK4(double* initialState, double** statesOut, size_t* statesCount, double** timesOut,
                                                  size_t* timesCount, SimProperties* simProperties, RK4SolverOptions* rkOptions,
                                                  COptomechanicalVariables* optomechanicalVariables, size_t N) {
    RB_TRY
    if (!initialState || !statesOut || !statesCount || !simProperties || !rkOptions || !optomechanicalVariables)
        throw std::runtime_error("integrateAugmentedOptomechanicalSimulationRK4: null argument");
    Adim ad = adimensionalize(simProperties->L, simProperties->rho, simProperties->kappa, simProperties->depth);
    rb_props p = helium_props(ad, simProperties->use_expansions, simProperties->expansion_order, simProperties->infinite_depth);
    p.guess_mode = RB_GUESS_WARM;
    rb_opto v = adimensionalize_opto(*optomechanicalVariables, ad.base_length, ad.base_time, ad.base_energy, ad.rho);
    const double dt = rkOptions->timeStep / ad.base_time, t0 = rkOptions->t0 / ad.base_time, t1 = rkOptions->t1 / ad.base_time;
    const size_t steps = static_cast<size_t>((t1 - t0) / dt);
    std::vector<double> states, times;
    aug_integrate_host(initialState, N, p, v, dt, steps, rkOptions->returnTrajectory, t0, states, times);
    double* so = (double*)std::malloc(std::max<size_t>(states.size(), 1) * sizeof(double));
    std::memcpy(so, states.data(), states.size() * sizeof(double));
    *statesOut = so;
    *statesCount = states.size() / (4 * N);
    if (timesOut) {
        double* to = (double*)std::malloc(std::max<size_t>(times.size(), 1) * sizeof(double));
        std::memcpy(to, times.data(), times.size() * sizeof(double));
        *timesOut = to;
    }
    if (timesCount) *timesCount = times.size();
    RB_CATCH
}

int integrateAugmentedOptomechanicalSimulationRK4_freeMemory(double* statesOut, double* timesOut) {
    std::free(statesOut);
    std::free(timesOut);
    return 0;
}

int rb_integrate_aug_rk4_host(const double* initialState_host, double* finalState_host, size_t N, const rb_props* props,
                              const rb_opto* v, double dt, size_t steps) {
    RB_TRY
    if (!v) throw std::runtime_error("rb_integrate_aug_rk4_host: null optomechanical variables");
    rb_props p;
    if (props) p = *props; else rb_default_props(&p);
    std::vector<double> states, times;
    aug_integrate_host(initialState_host, N, p, *v, dt, steps, false, 0.0, states, times);
    std::memcpy(finalState_host, states.data(), states.size() * sizeof(double));
    RB_CATCH
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// the same drive in its explicitly time-dependent form: TimedBoundaryIntegrator<N,B> over HeliumWithOptomechanicalDrivingProblem<N>
// and RungeKuttaStepper<std_complex, 2N>(TimedProblem&) (L/RK4_Time_Dependent.cuh, L/TimedBoundaryIntegrator.cuh,
// L/HeliumWithDrivingBoundaryProblem.cuh, L/DelayedIntensityTerm.cuh; assembled as L/Export.cu:797-826, A/kernel.cu:281-366).
// State [Z | Phi]; the delayed intensity and its reference time belong to the stepper.  The reference time is kept on the host and
// handed to the kernel as an argument (see timed_drive_kernel); the trajectory is appended on the device, without a host sync.
// ------------------------------------------------------------------------------------------------
struct rb_timed_stepper {
    rb_solver* s = nullptr;
    rb_opto v;
    double dt = 1e-2;
    double t = 0.0;                 // RungeKuttaStepperBase::currentTime
    double prev_time = 0.0;         // DelayedIntensityTerm::prev_time
    double* delayed = nullptr;      // DelayedIntensityTerm::delayed_intensity, BN doubles
    double2* y0 = nullptr;
    bool owns_y0 = false;
    double2* k[4] = {nullptr, nullptr, nullptr, nullptr};
    double2* ytmp = nullptr;
    bool trajectory = true;         // RK4Options::returnTrajectory (default true, L/RK4Options.h)
    std::vector<double> times;      // devTimes
    double2* log = nullptr;         // devYs: log_count states of 2 BN complex
    size_t log_count = 0, log_cap = 0;
};

static void timed_stepper_free(rb_timed_stepper* st) {
    if (!st) return;
    if (st->owns_y0 && st->y0) cudaFree(st->y0);
    for (auto& k : st->k)
        if (k) cudaFree(k);
    if (st->ytmp) cudaFree(st->ytmp);
    if (st->delayed) cudaFree(st->delayed);
    if (st->log) cudaFree(st->log);
    delete st;
}

// TimedBoundaryIntegrator::run with currentTime = time, saveProgress = save: the boundary-integral RHS, then the drive terms
// (calculateRhsPhi override, L/TimedBoundaryIntegrator.cuh:21-26)
static void timed_rhs(rb_timed_stepper* st, double time, bool save, const double2* state, double2* out) {
    rb_solver* s = st->s;
    rhs(s, state, out);
    launch_timed_drive(out + s->BN, state, out, st->delayed, st->v, time, st->prev_time, save ? 1 : 0, s->BN, s->stream);
    if (save) st->prev_time = time;   // save_value, L/DelayedIntensityTerm.cuh:29-33
}

// runStep, L/RK4_Time_Dependent.cuh:145-283: stages at t, t + h/2, t + h/2, t + h; setSaveProgress(false) after the first stage is
// never undone within the step, so only the first stage advances the delayed intensity
static void timed_step(rb_timed_stepper* st) {
    rb_solver* s = st->s;
    const size_t n = 2 * s->BN;
    const double h = st->dt, half = st->dt * 0.5;
    timed_rhs(st, st->t, true, st->y0, st->k[0]);
    launch_stage_update(st->ytmp, st->y0, st->k[0], half, n, s->stream);
    timed_rhs(st, st->t + half, false, st->ytmp, st->k[1]);
    launch_stage_update(st->ytmp, st->y0, st->k[1], half, n, s->stream);
    timed_rhs(st, st->t + half, false, st->ytmp, st->k[2]);
    launch_stage_update(st->ytmp, st->y0, st->k[2], h, n, s->stream);
    timed_rhs(st, st->t + h, false, st->ytmp, st->k[3]);
    launch_final_update(st->y0, st->k[0], st->k[1], st->k[2], st->k[3], h, n, s->stream);
}

static void timed_log_reserve(rb_timed_stepper* st, size_t extra) {
    const size_t n = 2 * st->s->BN;
    if (st->log_count + extra <= st->log_cap) return;
    const size_t cap = std::max(st->log_count + extra, 2 * st->log_cap);
    double2* grown = dmalloc<double2>(cap * n);
    if (st->log_count)
        RB_CUDA(cudaMemcpyAsync(grown, st->log, st->log_count * n * sizeof(double2), cudaMemcpyDeviceToDevice, st->s->stream));
    if (st->log) {
        RB_CUDA(cudaStreamSynchronize(st->s->stream));
        cudaFree(st->log);
    }
    st->log = grown;
    st->log_cap = cap;
}

// runEvolution, L/RK4_Time_Dependent.cuh:307-328
static size_t timed_evolve(rb_timed_stepper* st, double t0, double t1) {
    const size_t n = 2 * st->s->BN;
    st->t = t0;
    const size_t steps = static_cast<size_t>((t1 - t0) / st->dt);
    st->prev_time = t0;   // timedProblem.setStartingTime -> DelayedIntensityTerm::setInitialTime
    if (st->trajectory) timed_log_reserve(st, steps);
    for (size_t i = 0; i < steps; ++i) {
        timed_step(st);
        if (st->trajectory) {   // the time at the START of the step with the state after it, :318-322
            st->times.push_back(st->t);
            RB_CUDA(cudaMemcpyAsync(st->log + st->log_count * n, st->y0, n * sizeof(double2), cudaMemcpyDeviceToDevice,
                                    st->s->stream));
            ++st->log_count;
        }
        st->t += st->dt;
    }
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    return steps;
}

extern "C" {

rb_timed_stepper* rb_timed_rk4_create(rb_solver* s, const rb_opto* v, double tstep) {
    try {
        if (!s || !v) throw std::runtime_error("rb_timed_rk4_create: null argument");
        std::unique_ptr<rb_timed_stepper, void (*)(rb_timed_stepper*)> st(new rb_timed_stepper, timed_stepper_free);
        st->s = s;
        st->v = *v;
        st->dt = tstep;
        st->prev_time = v->initial_time;   // DelayedIntensityTerm ctor, L/DelayedIntensityTerm.cuh:43-50
        const size_t n = 2 * s->BN;
        for (auto& k : st->k) k = dmalloc<double2>(n);
        st->ytmp = dmalloc<double2>(n);
        st->delayed = dmalloc<double>(s->BN);
        RB_CUDA(cudaMemsetAsync(st->delayed, 0, s->BN * sizeof(double), s->stream));
        return st.release();
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

int rb_timed_rk4_destroy(rb_timed_stepper* st) {
    RB_TRY
    if (st) {
        cudaDeviceSynchronize();
        timed_stepper_free(st);
    }
    RB_CATCH
}

int rb_timed_rk4_set_time_step(rb_timed_stepper* st, double tstep) {
    RB_TRY
    if (!st) throw std::runtime_error("rb_timed_rk4_set_time_step: null stepper");
    st->dt = tstep;
    RB_CATCH
}

int rb_timed_rk4_initialize(rb_timed_stepper* st, rb_complex* y0, int on_device) {
    RB_TRY
    if (!st || !y0) throw std::runtime_error("rb_timed_rk4_initialize: null argument");
    const size_t n = 2 * st->s->BN;
    if (on_device) {
        if (st->owns_y0 && st->y0) cudaFree(st->y0);
        st->y0 = (double2*)y0;   // caller keeps ownership, L/RK4_Time_Dependent.cuh:292-298
        st->owns_y0 = false;
    } else {
        if (!st->owns_y0 || !st->y0) st->y0 = dmalloc<double2>(n);
        st->owns_y0 = true;
        RB_CUDA(cudaMemcpyAsync(st->y0, y0, n * sizeof(double2), cudaMemcpyHostToDevice, st->s->stream));
        RB_CUDA(cudaStreamSynchronize(st->s->stream));
    }
    RB_CATCH
}

int rb_timed_rk4_set_starting_time(rb_timed_stepper* st, double time) {
    RB_TRY
    if (!st) throw std::runtime_error("rb_timed_rk4_set_starting_time: null stepper");
    st->t = time;
    st->prev_time = time;
    RB_CATCH
}

int rb_timed_rhs(rb_timed_stepper* st, double time, int save_progress, const rb_complex* state_dev, rb_complex* rhs_dev) {
    RB_TRY
    if (!st || !state_dev || !rhs_dev) throw std::runtime_error("rb_timed_rhs: null argument");
    timed_rhs(st, time, save_progress != 0, (const double2*)state_dev, (double2*)rhs_dev);
    RB_CATCH
}

int rb_timed_rk4_step(rb_timed_stepper* st, int advance_time) {
    RB_TRY
    if (!st || !st->y0) throw std::runtime_error("rb_timed_rk4_step: initialize() has not been called");
    timed_step(st);
    if (advance_time) st->t += st->dt;
    RB_CATCH
}

int rb_timed_rk4_evolve(rb_timed_stepper* st, double t0, double t1, size_t* steps_out) {
    RB_TRY
    if (!st || !st->y0) throw std::runtime_error("rb_timed_rk4_evolve: initialize() has not been called");
    const size_t steps = timed_evolve(st, t0, t1);
    if (steps_out) *steps_out = steps;
    RB_CATCH
}

int rb_timed_rk4_set_logging(rb_timed_stepper* st, int return_trajectory) {
    RB_TRY
    if (!st) throw std::runtime_error("rb_timed_rk4_set_logging: null stepper");
    st->trajectory = return_trajectory != 0;
    RB_CATCH
}

int rb_timed_rk4_copy_trajectory(rb_timed_stepper* st, double** times_out, size_t* times_count, rb_complex** states_out,
                                 size_t* states_count) {
    RB_TRY
    if (!st || !states_out || !states_count) throw std::runtime_error("rb_timed_rk4_copy_trajectory: null argument");
    const size_t n = 2 * st->s->BN;
    const bool traj = st->trajectory;
    if (!traj && !st->y0) throw std::runtime_error("rb_timed_rk4_copy_trajectory: initialize() has not been called");
    // copyTimesToHost, L/RK4_Time_Dependent.cuh:80-103
    if (times_out) {
        *times_out = nullptr;
        if (traj) {
            double* t = (double*)std::malloc(std::max<size_t>(st->times.size(), 1) * sizeof(double));
            if (!t) throw std::runtime_error("rb_timed_rk4_copy_trajectory: out of host memory");
            std::memcpy(t, st->times.data(), st->times.size() * sizeof(double));
            *times_out = t;
        }
    }
    if (times_count) *times_count = traj ? st->times.size() : 0;
    // copyStatesToHost, :105-131: the trajectory, or the latest state alone
    const size_t count = traj ? st->log_count : 1;
    double2* h = (double2*)std::malloc(std::max<size_t>(count, 1) * n * sizeof(double2));
    if (!h) throw std::runtime_error("rb_timed_rk4_copy_trajectory: out of host memory");
    if (count)
        RB_CUDA(cudaMemcpyAsync(h, traj ? st->log : st->y0, count * n * sizeof(double2), cudaMemcpyDeviceToHost, st->s->stream));
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    *states_out = (rb_complex*)h;
    *states_count = count;
    RB_CATCH
}

rb_complex* rb_timed_rk4_dev_state(rb_timed_stepper* st) { return st ? (rb_complex*)st->y0 : nullptr; }

double* rb_timed_rk4_dev_delayed_intensity(rb_timed_stepper* st) { return st ? st->delayed : nullptr; }

int rb_timed_rk4_get_state(rb_timed_stepper* st, rb_complex* y_host) {
    RB_TRY
    if (!st || !st->y0 || !y_host) throw std::runtime_error("rb_timed_rk4_get_state: initialize() has not been called");
    const size_t n = 2 * st->s->BN;
    RB_CUDA(cudaMemcpyAsync(y_host, st->y0, n * sizeof(double2), cudaMemcpyDeviceToHost, st->s->stream));
    RB_CUDA(cudaStreamSynchronize(st->s->stream));
    RB_CATCH
}

double rb_timed_rk4_current_time(rb_timed_stepper* st) { return st ? st->t : 0.0; }

// L/Export.cu:779-975
int integrateOptomechanicalSimulationRK4(double* initialState, double** statesOut, size_t* statesCount, double** timesOut,
                                         size_t* timesCount, SimProperties* simProperties, RK4SolverOptions* rkOptions,
                                         COptomechanicalVariables* optomechanicalVariables, size_t N) {
    RB_TRY
    if (!initialState || !statesOut || !statesCount || !simProperties || !rkOptions || !optomechanicalVariables)
        throw std::runtime_error("integrateOptomechanicalSimulationRK4: null argument");
    Adim ad = adimensionalize(simProperties->L, simProperties->rho, simProperties->kappa, simProperties->depth);
    rb_props p = helium_props(ad, simProperties->use_expansions, simProperties->expansion_order, simProperties->infinite_depth);
    p.guess_mode = RB_GUESS_WARM;
    rb_opto v = adimensionalize_opto(*optomechanicalVariables, ad.base_length, ad.base_time, ad.base_energy, ad.rho);
    const double dt = rkOptions->timeStep / ad.base_time, t0 = rkOptions->t0 / ad.base_time, t1 = rkOptions->t1 / ad.base_time;
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    std::unique_ptr<rb_timed_stepper, void (*)(rb_timed_stepper*)> st(rb_timed_rk4_create(s.get(), &v, dt), timed_stepper_free);
    if (!st) throw std::runtime_error(g_last_error);
    st->trajectory = rkOptions->returnTrajectory;
    std::vector<double2> host(2 * N);
    for (size_t i = 0; i < N; ++i) {
        host[i] = make_double2(initialState[i], initialState[N + i]);
        host[N + i] = make_double2(initialState[2 * N + i], 0.0);
    }
    if (rb_timed_rk4_initialize(st.get(), (rb_complex*)host.data(), 0) != 0) throw std::runtime_error(g_last_error);
    timed_evolve(st.get(), t0, t1);
    double* times = nullptr;
    rb_complex* states = nullptr;
    size_t tcount = 0, scount = 0;
    if (rb_timed_rk4_copy_trajectory(st.get(), &times, &tcount, &states, &scount) != 0) throw std::runtime_error(g_last_error);
    const double2* hs = (const double2*)states;
    double* so = (double*)std::malloc(std::max<size_t>(3 * scount * N, 1) * sizeof(double));
    for (size_t j = 0; j < scount; ++j)
        for (size_t i = 0; i < N; ++i) {
            so[j * 3 * N + i] = hs[j * 2 * N + i].x;
            so[j * 3 * N + N + i] = hs[j * 2 * N + i].y;
            so[j * 3 * N + 2 * N + i] = hs[j * 2 * N + N + i].x;
        }
    std::free(states);
    *statesOut = so;
    *statesCount = scount;
    if (timesOut) *timesOut = times; else std::free(times);
    if (timesCount) *timesCount = tcount;
    RB_CATCH
}

int integrateOptomechanicalSimulationRK4_freeMemory(double* statesOut, double* timesOut) {
    std::free(statesOut);
    std::free(timesOut);
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// adaptive Runge-Kutta-Fehlberg 4(5) stepper (L/RK45.cuh): the reference's RK45Base<T,N> / RK45_std_complex<N> over either the
// boundary-integral RHS of a solver or a caller-supplied AutonomousProblem::run
// ------------------------------------------------------------------------------------------------
struct rb_rk45 {
    rb_solver* s = nullptr;          // RHS = rhs(s, .) on the solver's stream; nullptr: generic problem
    rb_rhs_fn fn = nullptr;
    void* user = nullptr;
    cudaStream_t stream = nullptr;   // generic problems only
    bool own_stream = false;
    size_t n = 0;                    // complex components of the state
    double2* raw = nullptr;          // k1..k6 | y | ytmp   (RK45WorkspaceGpu, L/RK45.cuh:36-66)
    double2* k[6] = {};
    double2 *y = nullptr, *ytmp = nullptr;
    double* partial = nullptr;
    unsigned int* ticket = nullptr;
    double* sumsq = nullptr;
    double* h_sumsq = nullptr;       // pinned
    double atol = 1e-6, rtol = 1e-3, h_min = 1e-16, h_max = 1e10;
    double h = 1e-2, t = 0.0;
    bool accepted_prev = true;
    double scaled_error = 0.0;
    size_t max_rejected = 500;
    long long n_accepted = 0, n_rejected = 0, n_rhs = 0;
};

static cudaStream_t rk45_stream(rb_rk45* r) { return r->s ? r->s->stream : r->stream; }

static void rk45_free(rb_rk45* r) {
    if (!r) return;
    if (r->raw) cudaFree(r->raw);
    if (r->partial) cudaFree(r->partial);
    if (r->ticket) cudaFree(r->ticket);
    if (r->sumsq) cudaFree(r->sumsq);
    if (r->h_sumsq) cudaFreeHost(r->h_sumsq);
    if (r->own_stream && r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

static void rk45_apply_options(rb_rk45* r, const rb_rk45_options* o) {
    if (!o) return;
    r->atol = o->atol;
    r->rtol = o->rtol;
    r->h_min = o->h_min;
    r->h_max = o->h_max;
    r->h = o->initial_timestep;
}

static rb_rk45* rk45_make(rb_solver* s, size_t n, rb_rhs_fn fn, void* user, const rb_rk45_options* opt, cudaStream_t stream) {
    std::unique_ptr<rb_rk45, void (*)(rb_rk45*)> up(new rb_rk45, rk45_free);
    rb_rk45* r = up.get();
    r->s = s;
    r->fn = fn;
    r->user = user;
    r->n = n;
    if (!s) {
        if (stream == nullptr || stream == cudaStreamLegacy) {
            RB_CUDA(cudaStreamCreate(&r->stream));   // blocking stream: ordered against the legacy stream in both directions
            r->own_stream = true;
        } else {
            r->stream = stream;
        }
    }
    r->raw = dmalloc<double2>(8 * n);
    RB_CUDA(cudaMemset(r->raw, 0, 8 * n * sizeof(double2)));
    for (int i = 0; i < 6; ++i) r->k[i] = r->raw + (size_t)i * n;
    r->y = r->raw + 6 * n;
    r->ytmp = r->raw + 7 * n;
    r->partial = dmalloc<double>(rb::rk45_error_blocks(n));
    r->ticket = dmalloc<unsigned int>(1);
    RB_CUDA(cudaMemset(r->ticket, 0, sizeof(unsigned int)));
    r->sumsq = dmalloc<double>(1);
    RB_CUDA(cudaMallocHost(&r->h_sumsq, sizeof(double)));
    rk45_apply_options(r, opt);
    return up.release();
}

static void rk45_rhs(rb_rk45* r, const double2* y, double2* k) {
    if (r->s) rhs(r->s, y, k);
    else r->fn(r->user, (const rb_complex*)y, (rb_complex*)k, (void*)r->stream);
    r->n_rhs++;
}

// L/RK45.cuh:306-330
static double rk45_new_timestep(const rb_rk45* r, double old_h, double error, bool accepting) {
    const double safety = 0.9, minfac = 0.2, maxfac = 5.0, expo = 1.0 / 5.0;
    if (error == 0.0) return old_h * (accepting ? maxfac : 1.0);
    double fac = safety * std::pow(error, -expo);
    if (!(fac == fac)) fac = minfac;   // NaN error estimate: shrink as far as allowed
    fac = accepting ? std::min(std::max(fac, minfac), maxfac) : std::min(std::max(fac, minfac), 1.0);
    return std::min(std::max(old_h * fac, r->h_min), r->h_max);
}

// one attempt (L/RK45.cuh:258-304); returns true when the step was accepted
static bool rk45_step(rb_rk45* r) {
    // Fehlberg tableau (L/RK45_Kernels.cuh:18-32)
    static const double A[5][5] = {{1.0 / 4.0, 0, 0, 0, 0},
                                   {3.0 / 32.0, 9.0 / 32.0, 0, 0, 0},
                                   {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0, 0},
                                   {439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0, 0},
                                   {-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0}};
    cudaStream_t st = rk45_stream(r);
    const double h = r->h;
    if (r->accepted_prev) rk45_rhs(r, r->y, r->k[0]);   // a rejected attempt keeps k1
    const double2* ks[5] = {r->k[0], r->k[1], r->k[2], r->k[3], r->k[4]};
    for (int stage = 0; stage < 5; ++stage) {
        double c[5];
        for (int j = 0; j < 5; ++j) c[j] = A[stage][j] * h;
        rb::launch_rk45_stage(r->y, ks, r->ytmp, c, stage + 1, r->n, st);
        rk45_rhs(r, r->ytmp, r->k[stage + 1]);
    }
    rb::launch_rk45_error_y5(r->y, r->k[0], r->k[2], r->k[3], r->k[4], r->k[5], r->ytmp, h, r->atol, r->rtol, r->partial, r->ticket,
                             r->sumsq, r->n, st);
    RB_CUDA(cudaMemcpyAsync(r->h_sumsq, r->sumsq, sizeof(double), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    r->scaled_error = std::sqrt((1.0 / (double)r->n) * *r->h_sumsq);
    const bool accepted = r->scaled_error <= 1.0;
    r->accepted_prev = accepted;
    const double h_new = rk45_new_timestep(r, h, r->scaled_error, accepted);
    if (accepted) {
        RB_CUDA(cudaMemcpyAsync(r->y, r->ytmp, r->n * sizeof(double2), cudaMemcpyDeviceToDevice, st));   // calculateWeightedY :370-374
        r->t += h;
        r->n_accepted++;
    } else {
        r->n_rejected++;
    }
    r->h = h_new;
    return accepted;
}

extern "C" {

rb_rk45* rb_rk45_create(rb_solver* s, const rb_rk45_options* opt) {
    try {
        if (!s) throw std::runtime_error("rb_rk45_create: null solver");
        return rk45_make(s, 2 * s->BN, nullptr, nullptr, opt, nullptr);
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

rb_rk45* rb_rk45_create_generic(size_t n, rb_rhs_fn f, void* user, const rb_rk45_options* opt, void* stream) {
    try {
        if (!f || n == 0) throw std::runtime_error("rb_rk45_create_generic: need a right-hand side and n > 0");
        return rk45_make(nullptr, n, f, user, opt, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

int rb_rk45_destroy(rb_rk45* r) {
    rk45_free(r);
    return 0;
}

int rb_rk45_set_options(rb_rk45* r, const rb_rk45_options* opt) {
    RB_TRY
    if (!opt) throw std::runtime_error("rb_rk45_set_options: null options");
    rk45_apply_options(r, opt);
    RB_CATCH
}

int rb_rk45_set_tolerance(rb_rk45* r, double atol, double rtol) {
    r->atol = atol;
    r->rtol = rtol;
    return 0;
}

int rb_rk45_set_max_rejected(rb_rk45* r, size_t max_rejected) {
    r->max_rejected = max_rejected;
    return 0;
}

int rb_rk45_initialize(rb_rk45* r, const rb_complex* y0, int on_device) {
    RB_TRY
    cudaStream_t st = rk45_stream(r);
    RB_CUDA(cudaMemcpyAsync(r->y, y0, r->n * sizeof(double2), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaStreamSynchronize(st));
    r->accepted_prev = true;
    RB_CATCH
}

int rb_rk45_step(rb_rk45* r, int* accepted) {
    RB_TRY
    const bool a = rk45_step(r);
    if (accepted) *accepted = a ? 1 : 0;
    RB_CATCH
}

// L/RK45.cuh:194-247; *result: 0 = ReachedEndTime, 1 = StiffnessDetected (L/OdeSolver.h:2-5)
int rb_rk45_evolve(rb_rk45* r, double t0, double t1, int* result) {
    RB_TRY
    r->t = t0;
    int res = 0;
    for (;;) {
        if (r->t >= t1) break;
        const double max_step = t1 - r->t;   // never overshoot the end time
        if (max_step < r->h) r->h = max_step;
        size_t rejected = 0;
        bool stiff = false;
        for (;;) {
            if (rk45_step(r)) break;
            if (++rejected > r->max_rejected) {
                stiff = true;
                break;
            }
        }
        if (stiff) {
            res = 1;
            break;
        }
    }
    RB_CUDA(cudaStreamSynchronize(rk45_stream(r)));
    if (result) *result = res;
    RB_CATCH
}

rb_complex* rb_rk45_dev_state(rb_rk45* r) { return (rb_complex*)r->y; }

int rb_rk45_get_state(rb_rk45* r, rb_complex* y_host) {
    RB_TRY
    cudaStream_t st = rk45_stream(r);
    RB_CUDA(cudaMemcpyAsync(y_host, r->y, r->n * sizeof(double2), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    RB_CATCH
}

double rb_rk45_current_time(rb_rk45* r) { return r->t; }
double rb_rk45_current_timestep(rb_rk45* r) { return r->h; }

int rb_rk45_stats(rb_rk45* r, double out_host[4]) {
    out_host[0] = (double)r->n_accepted;
    out_host[1] = (double)r->n_rejected;
    out_host[2] = (double)r->n_rhs;
    out_host[3] = r->scaled_error;
    return 0;
}

}  // extern "C"
