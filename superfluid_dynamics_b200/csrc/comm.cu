// comm.cu -- row sharding of every O(N^2) sweep over the GPUs of one node: arenas mapped into every process by CUDA IPC, rows published
// by peer stores from the sweep epilogues (pair_kernels.cu), epoch flags instead of collectives; plus the sweep-plan queries.
// No reference counterpart (L/utilities.cuh:20 is single-GPU).
#include "host.cuh"

extern "C" {

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
    out[0] = s->use_v3 ? 3 : (s->use_v2 ? 2 : 1);            // 1 tiled, 2 persistent, 3 warp per row group
    out[1] = s->use_v3 ? s->v3l.R : (s->use_v2 ? s->v2_R : s->v1_rows);
    out[2] = s->tile;
    out[3] = s->tiles_per_chunk;
    out[4] = s->nchunks;
    out[5] = s->row_cells;
    out[6] = s->use_v3 ? s->v3l.grid : (s->use_v2 ? s->v2l.grid : s->row_cells * s->nchunks * s->batch);   // CTAs per sweep
    out[7] = s->use_v3 ? s->v3l.threads : (s->use_v2 ? s->v2l.threads : kCell / s->v1_rows);
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

}  // extern "C"
