// solver.cu -- host orchestration of one RHS evaluation (the assembler object, its solves, its C ABI in include/roberts_b200.h).
// The stepper, the multi-GPU layer, the probes, the legacy exports, the driven film and RKF45 live in their own translation units
// (host.cuh lists them).
//
// Mirrors (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   BaseBoundaryIntegralCalculator<N,B>::runTimeStep / calculateVorticities   L/BaseBoundaryIntegrator.cuh:138-306
//   ZPhiDerivative<N,B>::exec, FftDerivative<N,B>::exec                        L/Derivatives.cuh:190-257, 311-384
// N and the batch size are runtime values.  No CPU fallback: everything below needs a CUDA device.
#include "host.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
thread_local std::string g_last_error;
namespace rb {
unsigned long long g_launch_count = 0;
int report_error(const std::exception& e) { return fail(e); }   // service for implicit.cu (internal.cuh)
}

int fail(const std::exception& e) {
    g_last_error = e.what();
    std::fprintf(stderr, "Error: %s\n", e.what());   // L/Export.cu: std::cerr << "Error: " << e.what()
    return -1;
}

// ------------------------------------------------------------------------------------------------
// the RHS assembler (struct rb_solver: host.cuh)
// ------------------------------------------------------------------------------------------------
void solver_free(rb_solver* s) {
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
                    s->gm_dev, s->gm_invP, s->gm_members, s->gm_ctrl, s->gm_part, s->v2_rnorm_part, s->v2_ticket, s->fft_tw, s->v2_partial, s->v2_xs_part,
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
void choose_chunking(rb_solver* s) {
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
void alloc_partials(rb_solver* s) {
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
void plan_sweep2(rb_solver* s) {
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
    // launch bounds of sweep2_kernel<., R>.  R = 4 (ensembles), measured (profiles/r02h_ensemble.log): whole-member row blocks of 512 rows
    // want one CTA of 512 threads per SM (1024 x N=512: 357 steps/s against 264 with two CTAs of 256), smaller members two CTAs of
    // 256 threads per SM (2048 x N=256: 582 against 447)
    const int max_threads = R == 4 ? (env_int("RB_V2_R4_THREADS", RB >= 512 ? 512 : 256) <= 256 ? 256 : 512) : (R == 2 ? 896 : 1024);
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
void choose_sweep_kernel(rb_solver* s) {
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
    // warp-per-row-group kernel (pair_kernels3.cu): one member, no image sum, at most 16 row groups per SM.  R rows per warp by the
    // number of rows of this rank, so that the row groups fill the warps of one CTA per SM.
    s->use_v3 = false;
    s->v3l.batched = false;
    if (!s->has_image && s->batch == 1 && s->N <= 8192) {
        const int rows = std::min(s->N, (s->row_cell0 + s->row_cells) * kCell) - s->row_cell0 * kCell;
        int R = rows >= 3072 ? 4 : (rows >= 1536 ? 2 : 1);
        R = env_int("RB_V3_R", R);
        R = R >= 4 ? 4 : (R >= 2 ? 2 : 1);
        const int NG = (rows + R - 1) / R;
        // one CTA per SM with ~200 KB of shared memory: when the a' transform of a recorded round runs beside the sweep on the side
        // stream (overlap_ok), one SM is left to it -- otherwise whichever of the two starts second waits for the other's whole CTA
        const int G = std::max(1, std::min(nSM - (s->overlap_ok ? 1 : 0), NG));
        const int Wg = (NG + G - 1) / G;                        // row groups per CTA
        // S warps per row group (they split the cells between them), as long as the CTA stays at 8 warps -- the variant compiled
        // without a register cap.  Measured (profiles/r02s_sweep3.log, sweep alone, us): N = 4096: R = 4 / S = 1 (7 warps) 26.7,
        // S = 2 (14 warps, 128 registers, spills) 32.8, R = 2 (14 warps) 28.7; N = 2048: R = 2 / S = 1 14.4, R = 4 / S = 2 14.4,
        // R = 2 / S = 2 16.6; N = 1024: R = 1 12.3, R = 2 / S = 2 11.7
        int S = 1;
        while (S * 2 <= s->ncell && Wg * S * 2 <= 8 && S < 4) S *= 2;
        S = env_int("RB_V3_S", S);
        S = std::max(1, std::min(S, std::min(s->ncell, 8)));
        const int W = Wg * S;
        if (W <= 16) {
            s->v3l.R = R;
            s->v3l.S = S;
            s->v3l.grid = G;
            s->v3l.threads = 32 * W;
            s->v3l.TS = std::min(4096, ((s->N + kCell - 1) / kCell) * kCell);
            s->v3l.smem = sweep3_smem(s->v3l.TS);
            // measured on a B200 (profiles/r02s_sweep3.log)
            const int lo = env_int("RB_V3_MIN_N", 2);
            s->use_v3 = s->N >= lo;
            const int f3 = env_int("RB_SWEEP_V3", -1);
            if (f3 >= 0) s->use_v3 = f3 != 0;
            if (s->use_v3 && (int)std::max(s->v2_total_blocks, 1) < G) {   // residual partials: one per CTA
                if (s->v2_rnorm_part) cudaFree(s->v2_rnorm_part);
                s->v2_rnorm_part = dmalloc<double>(std::max(G, s->v2_total_blocks));
            }
            if (s->use_v3) s->use_v2 = false;
        }
    }
    // ensembles of small members (no cell-local coordinates: N <= 768): the same mapping, one member at a time per CTA
    if (!s->has_image && s->batch > 1 && !s->use_local && s->comm.nranks <= 1 && s->N >= 32) {
        int R = s->N >= 128 ? 4 : 1;
        R = env_int("RB_V3_R", R);
        R = R >= 4 ? 4 : (R >= 2 ? 2 : 1);
        const int NG = (s->N + R - 1) / R;
        const int W = std::max(1, std::min(16, NG));
        const int NP = ((s->N + kCell - 1) / kCell) * kCell;
        if (NP <= 4 * 32 * W) {
            // measured (profiles/r02v_ensemble_sweep3b.log, steps/s, this kernel / the persistent or tiled one): 1024 x N=512 394 / 356,
            // 512 x N=768 359 / 197 (every solve in 2 sweeps instead of 3), 128 x N=512 2071 / 1843; one-cell members and small
            // batches lose (2048 x N=256 513 / 597, 37 x N=300 2784 / 3368)
            const bool on = env_int("RB_SWEEP_V3B", (s->N > kCell && s->batch >= 64) ? 1 : 0) != 0;
            if (on) {
                s->v3l.R = R;
                s->v3l.S = 1;
                s->v3l.batched = true;
                s->v3l.grid = std::min(nSM, s->batch);
                s->v3l.threads = 32 * W;
                s->v3l.TS = NP;
                s->v3l.smem = sweep3b_smem(NP);
                s->use_v3 = true;
                s->use_v2 = false;
            }
        }
    }
    if (env_int("RB_VERBOSE", 0) && s->use_v3)
        std::fprintf(stderr, "[roberts_b200] sweep3 plan: N=%d R=%d S=%d grid=%d threads=%d TS=%d smem=%zu\n", s->N, s->v3l.R, s->v3l.S,
                     s->v3l.grid, s->v3l.threads, s->v3l.TS, s->v3l.smem);
}

static void sweep(rb_solver* s, const SweepArgs& a, int mode) {
    if (s->use_v3) launch_sweep3(a, s->v3l, mode, s->stream);
    else if (s->use_v2) launch_sweep2(a, s->v2l, mode, s->stream);
    else launch_sweep(a, mode, s->stream);
    s->total_sweeps++;
}

rb_solver* solver_create(int N, int batch, const rb_props* pin) {
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
        s->gm_m = std::max(2, std::min(env_int("RB_GMRES_RESTART", kGmMax), std::min(N, kGmMax)));   // Krylov vectors per cycle
        s->gm_ld = (BN + 31) / 32 * 32;
        s->gm_V = dmalloc<double>((size_t)(s->gm_m + 1) * s->gm_ld);
        s->gm_x = dmalloc<double>(BN);
        s->gm_t = dmalloc<double>(BN);
        RB_CUDA(cudaMallocHost(&s->gm_host, std::max(sizeof(GmCtrl), sizeof(SolveCtrl))));
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
        // slice partials of the multi-CTA Arnoldi kernel: [batch][C][2 (kGmMax + 1) + 2]
        s->gm_part = dmalloc<double>((size_t)batch * gm_arnoldi_slices(N, batch) * (2 * (kGmMax + 1) + 2));
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
    // Own one-CTA transforms fused with the coefficient multiply (spectral.cu): radix-2 Stockham below N = 256, the radix-8
    // register-resident transform for 256 <= N <= 8192.  (The radix-2 one is shared-memory-bandwidth bound, ~1.3k cycles per pass at
    // N = 4096, and lost against the library's launches above N = 1024; the radix-8 one needs 4 passes there.)
    const int own_fft_max = env_int("RB_OWN_FFT_MAX", 8192);
    if ((N & (N - 1)) == 0 && N >= 4 && N <= 8192 && (long)N * batch <= 8192 && env_int("RB_OWN_FFT", 1)) {
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

void surface_stage(rb_solver* s, const double2* Z, const double2* Phi) {
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

SweepArgs base_args(rb_solver* s, const double2* Z) {
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

void launch_mv(rb_solver* s, const SweepArgs& base, int i, int skip) {
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
void note_solve_end(rb_solver* s, int converged, int stagnated, double rel, int iters, const char* what) {
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

// ---- restarted GMRES, right-preconditioned, classical Gram-Schmidt with re-orthogonalisation (device-driven cycles) ------------
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

// The restarted solver outside recorded steps (standalone rb_rhs / rb_vorticities, the batched RHS behind the finite-difference
// Jacobian, the stepper's fallback): cycles of at most K Arnoldi steps run on the device exactly as inside a recorded step
// (krylov_kernels.cu); the host looks ONCE per cycle (round 1: once per Arnoldi step, with the Givens rotations on the host).  Every
// cycle starts from the TRUE residual b - M x of every member, each member stops on its own residual against its own ||b||, and
// the solve ends when the worst member has converged -- or stagnated (the true residual after a whole cycle no better than half
// the one before it and already <= 1e-10: the iteration sits on the round-off floor of this operator, eps x cond(M) with
// cond ~ N / 2 pi for the thin film; same status as the Richardson solver's) -- or the cap on applications of M is reached.
static void gmres_solve(rb_solver* s, const double2* Z) {
    cudaStream_t st = s->stream;
    const int n = (int)s->BN;
    const double tol = s->props.tolerance;
    double* w = s->xbuf[1];
    SweepArgs base = base_args(s, Z);
    SolveCtrl* skip = reinterpret_cast<SolveCtrl*>(s->gm_ctrl);

    const double* warm = nullptr;
    if (s->props.guess_mode == RB_GUESS_WARM && !s->hist.base && s->have_prev_a) warm = s->a;
    launch_guess(s->b, warm, s->hist, s->gm_x, s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch, s->ncell, st);
    if (!warm && !s->hist.base) apply_Pinv(s, s->b, s->gm_x);   // cold start: x0 = P^{-1} b
    RB_CUDA(cudaMemsetAsync(&s->gm_ctrl->k_total, 0, sizeof(int), st));   // (the other fields are (re)set by gm_start_kernel)

    GmCtrl* hc = reinterpret_cast<GmCtrl*>(s->gm_host);          // pinned
    const int K = std::max(1, std::min(kGmMax, s->gm_m));
    const int max_applies = s->props.max_iterations;
    int applies = 0;
    double rel = 1e300, prev_cycle_rel = 1e300;
    bool converged = false, stagnated = false;
    for (int cycle = 0; cycle < 64 && applies < max_applies; ++cycle) {
        apply_M(s, base, s->gm_x);                                                  // w = M x
        launch_gm_start(s->b, w, s->gm_V, s->gm_members, s->gm_ctrl, s->ctrl, s->N, s->batch, tol, st);
        RB_CUDA(cudaMemcpyAsync(hc, s->gm_ctrl, sizeof(GmCtrl), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        applies = 1 + cycle + hc->k_total;                                          // one residual per cycle + the Arnoldi steps so far
        rel = hc->worst_rel;
        if (!(rel == rel) || rel >= 1e300) break;                                   // NaN: give up, reported as failed
        if (hc->done) {                                                             // every member's true residual <= tolerance
            converged = true;
            break;
        }
        if (cycle > 0 && rel <= 1e-10 && rel > 0.5 * prev_cycle_rel) {
            stagnated = true;
            break;
        }
        prev_cycle_rel = rel;
        const int Kc = std::min(K, max_applies - applies);
        if (Kc < 1) break;
        for (int k = 0; k < Kc; ++k) {
            apply_Pinv(s, s->gm_V + (size_t)k * s->gm_ld, s->gm_t);
            apply_M(s, base, s->gm_t, skip);                                        // w = M P^{-1} v_k (skips itself once the cycle has ended)
            launch_gm_arnoldi(s->gm_V, s->gm_ld, w, s->gm_members, s->gm_ctrl, s->ctrl, s->gm_part, s->N, s->batch, k, Kc - 1, tol, st);
        }
        launch_gm_correction(s->gm_V, s->gm_ld, s->gm_t, s->gm_members, s->N, s->batch, st);
        apply_Pinv(s, s->gm_t, s->gm_t);
        launch_axpby(s->gm_x, s->gm_x, 1.0, s->gm_t, n, st);
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

void vorticities(rb_solver* s, const double2* state) {
    const double2* Z = state;
    const double2* Phi = state + s->BN;
    surface_stage(s, Z, Phi);
    solve(s, Z);
    s->cur_Z = Z;
    s->cur_Phi = Phi;
}

void fft_derivative(rb_solver* s, const double2* in, double2* out, int second, double scaling) {
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
        launch_gm_arnoldi(s->gm_V, s->gm_ld, w, s->gm_members, s->gm_ctrl, s->ctrl, s->gm_part, s->N, s->batch, k, K - 1, tol_in, st);
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

void rhs(rb_solver* s, const double2* state, double2* out) {
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
// C ABI of the assembler
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

}  // extern "C"
