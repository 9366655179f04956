// lu_kernels.cu -- blocked right-looking LU with partial pivoting for ONE right-hand side, column-major, in place: the dense solve
// of the path wherever a matrix is really materialised (the 6N x 6N Newton matrix of the Gauss-Legendre integrator, the N x N
// system M a = Re(Phi') in the dense validation mode).  The trailing update A22 -= L21 U12 (K = 32) is the one true dense
// contraction on this path and runs on the FP64 tensor path (mma.sync m8n8k4.f64, "DMMA"); everything else is latency work.
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/): MatrixSolver<N,1>::solve = cusolverDnDgetrf + cusolverDnDgetrs,
// L/MatrixSolver.cuh:114-125.
//
// Per panel of 32 columns, four launches:
//   lu_panel_kernel      one CTA: pivot search (warp shuffles), row swap inside the panel, scale, rank-1 update of the panel
//   lu_swap_trsm_kernel  one thread per column right of the panel and for b: the panel's row interchanges (the left part, L, is
//                        never read again because b is eliminated on the fly, exactly as the unblocked kernels of dense_kernels.cu
//                        treat it), then U12 = L11^{-1} A12 with L11 in shared memory
//   lu_gemm_kernel       A22 -= L21 U12: 64 x 64 tiles, 8 warps, 2 x 4 DMMA tiles of 8 x 8 per warp, 8 k-steps of 4
//   lu_gemv_kernel       b2  -= L21 b1
// then lu_backsolve_blocked_kernel: the back substitution with U, one CTA, 32 columns per pair of barriers.
// Measured on a B200 (profiles/r02c_implicit_report_before_cluster_panel.json, CUDA events, ms; unblocked / blocked with the one-CTA
// panel / cuSOLVER getrf+getrs): n = 384: 4.2 / 3.3 / 0.98; 1536: 25.4 / 17.3 / 4.4; 4096: 166 / 67 / 18.0; 6144: - / 165 / 32.3 --
// the blocked factorisation is the default (dense_kernels.cu: launch_lu_solve).  ncu launch list at n = 2048 (profiles/
// r02p_lu_launches_n2048_summary.txt): panel 73 % (215 us per 32 columns), swap + TRSM 12 %, trailing DMMA GEMM 7 %, back
// substitution 4.5 %, GEMV 3 %: the panel's per-column chain of L2 round trips and barriers is what separates this factorisation
// from cuSOLVER's (3.7x at n = 4096), not the tensor-core update.
// The kernels and the launch sequence below also compile under g++ with tests/cpp/cuda_emu.h standing in for the CUDA headers
// (RB_EMULATE): the CPU test tier runs them thread for thread against LAPACK (tests/test_kernel_emulation.py).
#ifdef RB_EMULATE
#include "cuda_emu.h"
#else
#include <cooperative_groups.h>

#include <cstring>

#include "internal.cuh"
#endif
#include "launch.cuh"

namespace rb {

namespace {

constexpr int kNB = 32;              // panel width == K of the trailing update
constexpr int kPanelThreads = 1024;

__global__ void __launch_bounds__(kPanelThreads) lu_panel_kernel(double* A, int n, int k0, int kb, int* __restrict__ piv,
                                                                  int* __restrict__ info) {
    // Three block-wide barriers per column: (1) after the per-warp pivot candidates are in shared memory -- every warp then reduces
    // the 32 candidates itself, so the winner needs no broadcast (the candidate arrays are double-buffered against the next column);
    // (2) after the row interchange, which also leaves the pivot row in shared memory; (3) after the rank-1 update of the panel.
    __shared__ double s_val[2][32];
    __shared__ int s_idx[2][32];
    __shared__ double s_row[kNB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < kb; ++j) {
        const int col = k0 + j, buf = j & 1;
        double* cptr = A + (size_t)col * n;
        // pivot: largest |A[i, col]|, i >= col; ties to the lowest row (the choice of the unblocked kernel)
        double best = -1.0;
        int bi = col;
        for (int i = col + tid; i < n; i += kPanelThreads) {
            const double v = fabs(cptr[i]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[buf][warp] = best; s_idx[buf][warp] = bi; }
        __syncthreads();
        best = s_val[buf][lane];
        bi = s_idx[buf][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        const int p = __shfl_sync(0xffffffffu, bi, 0);
        const bool usable = best > 0.0;
        if (tid == 0) {
            piv[col] = p;
            if (!usable && *info == 0) *info = col + 1;   // zero (or NaN) column: singular, as getrf's info
        }
        if (tid < kb) {   // row interchange inside the panel; the thread that moves column k0 + tid also knows the new pivot-row entry
            double* q = A + (size_t)(k0 + tid) * n;
            const double top = q[col], low = q[p];
            if (p != col) {
                q[col] = low;
                q[p] = top;
            }
            s_row[tid] = low;
        }
        __syncthreads();
        if (usable) {
            const double pv = s_row[j];
            for (int i = col + 1 + tid; i < n; i += kPanelThreads) {
                const double l = cptr[i] / pv;
                cptr[i] = l;
                for (int c = j + 1; c < kb; ++c) A[(size_t)(k0 + c) * n + i] -= l * s_row[c];
            }
        }
        __syncthreads();
    }
}

#ifndef RB_EMULATE
// ---- the panel on a thread-block cluster (sm_100a): the n x 32 panel lives in the DISTRIBUTED shared memory of 8 (or 16) CTAs ----
// The one-CTA panel above streams the panel through one SM's path to L2 for every column (n = 4096: ~0.5 ms per panel, 64 of the
// 67 ms of a factorisation).  Here CTA c of the cluster keeps rows [c * rpc, (c + 1) * rpc) of the panel in its shared memory for the
// whole panel; per column the CTAs exchange ONE pivot candidate each through distributed shared memory and meet at ONE cluster
// barrier:
//   * rows are never moved: the row chosen as the pivot of column j is RETIRED where it lies (no CTA writes it again), everybody
//     reads its 32 entries from the owner's shared memory, and the interchanges LAPACK would have made are replayed on index maps
//     (pos / loc below) -- the panel is written back to global memory in its final row order, and piv[] holds position-based
//     pivots exactly as the one-CTA kernel leaves them (the swap + TRSM kernel applies them to the columns right of the panel);
//   * the candidates are double-buffered by column parity, so a fast CTA can publish column j + 1 while a slow one still reads j;
//   * the rank-1 update of column j also finds the pivot candidate of column j + 1 (one pass over the rows per column).
// Pivot choice: largest |a|, ties to the lowest CURRENT POSITION, as getrf / the other two kernels do.
constexpr int kClusterPanelThreads = 256;

struct PivotCand {
    double val;
    int row;       // physical row (global index) of the candidate
    int pos;       // its current position (global row index after the interchanges so far)
};

template <int CS>
__global__ void __launch_bounds__(kClusterPanelThreads) lu_panel_cluster_kernel(double* __restrict__ A, int n, int k0, int kb, int rpc,
                                                                                 int* __restrict__ piv, int* __restrict__ info) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: panel P[kNB][rpc] | slot[rpc] (int: -1 active, j retired as pivot of column j) | cand[2] | s_row[kNB] | wcand[8]
    double* P = reinterpret_cast<double*>(smem_raw);
    int* slot = reinterpret_cast<int*>(P + (size_t)kNB * rpc);
    PivotCand* cand = reinterpret_cast<PivotCand*>(slot + rpc);           // [2]
    double* s_row = reinterpret_cast<double*>(cand + 2);                   // [kNB]
    PivotCand* wcand = reinterpret_cast<PivotCand*>(s_row + kNB);         // [warps]
    __shared__ int pos_top[kNB];      // physical row at position k0 + i
    __shared__ int loc_top[kNB];      // position of physical row k0 + i
    __shared__ int s_p[2];            // winner of the column: physical row, its position

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int row0 = k0 + rank * rpc;                        // first physical row of this CTA
    const int nloc = max(0, min(rpc, n - row0));             // rows held here
    for (int idx = tid; idx < kb * rpc; idx += kClusterPanelThreads) {
        const int c = idx / rpc, r = idx - c * rpc;
        P[idx] = r < nloc ? A[(size_t)(k0 + c) * n + row0 + r] : 0.0;
    }
    for (int r = tid; r < rpc; r += kClusterPanelThreads) slot[r] = -1;
    if (tid < kNB) {
        pos_top[tid] = k0 + tid;
        loc_top[tid] = k0 + tid;
    }
    __syncthreads();
    // candidate of column 0
    PivotCand mine;
    mine.val = -1.0; mine.row = -1; mine.pos = 0x7fffffff;
    for (int r = tid; r < nloc; r += kClusterPanelThreads) {
        const double v = fabs(P[r]);
        if (v > mine.val) { mine.val = v; mine.row = row0 + r; mine.pos = row0 + r; }   // (positions = rows before any interchange;
    }                                                                                    //  the strided scan visits them in order)

    for (int j = 0; j < kb; ++j) {
        const int buf = j & 1;
        // ---- this CTA's candidate: warp shuffles, then warp 0 over the warps ----
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, mine.val, o);
            const int orow = __shfl_down_sync(0xffffffffu, mine.row, o);
            const int opos = __shfl_down_sync(0xffffffffu, mine.pos, o);
            if (ov > mine.val || (ov == mine.val && opos < mine.pos)) { mine.val = ov; mine.row = orow; mine.pos = opos; }
        }
        if (lane == 0) wcand[warp] = mine;
        __syncthreads();
        if (warp == 0) {
            PivotCand c = lane < kClusterPanelThreads / 32 ? wcand[lane] : PivotCand{-1.0, -1, 0x7fffffff};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, c.val, o);
                const int orow = __shfl_down_sync(0xffffffffu, c.row, o);
                const int opos = __shfl_down_sync(0xffffffffu, c.pos, o);
                if (ov > c.val || (ov == c.val && opos < c.pos)) { c.val = ov; c.row = orow; c.pos = opos; }
            }
            if (lane == 0) cand[buf] = c;
        }
        cluster.sync();   // every CTA's candidate of column j is published (release / acquire over the cluster)
        // ---- the winner, identically on every CTA ----
        if (warp == 0) {
            PivotCand c{-1.0, -1, 0x7fffffff};
            if (lane < CS) c = *cluster.map_shared_rank(&cand[buf], lane);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, c.val, o);
                const int orow = __shfl_down_sync(0xffffffffu, c.row, o);
                const int opos = __shfl_down_sync(0xffffffffu, c.pos, o);
                if (ov > c.val || (ov == c.val && opos < c.pos)) { c.val = ov; c.row = orow; c.pos = opos; }
            }
            c.val = __shfl_sync(0xffffffffu, c.val, 0);
            c.row = __shfl_sync(0xffffffffu, c.row, 0);
            c.pos = __shfl_sync(0xffffffffu, c.pos, 0);
            const bool usable = c.val > 0.0;
            int q = c.row, qpos = c.pos;
            if (!usable) {   // zero (or NaN) column: singular as getrf reports it; "pivot" = the row at the diagonal position, no elimination
                q = pos_top[j];
                qpos = k0 + j;
            }
            // the pivot row's panel entries, from its owner's shared memory
            const int owner = (q - k0) / rpc, ql = (q - k0) - owner * rpc;
            const double* Pown = cluster.map_shared_rank(P, owner);
            if (lane < kb) s_row[lane] = Pown[(size_t)lane * rpc + ql];
            if (lane == 0) {
                s_p[0] = q;
                s_p[1] = usable ? 1 : 0;
                // replay of the interchange (positions k0 + j <-> qpos) on the index maps
                const int u = pos_top[j];                       // physical row now at the diagonal position (a top-origin row)
                if (qpos - k0 < kb) pos_top[qpos - k0] = u;
                loc_top[u - k0] = qpos;
                if (q - k0 < kb) loc_top[q - k0] = k0 + j;
                pos_top[j] = q;
                if (rank == 0) {
                    piv[k0 + j] = qpos;
                    if (!usable && *info == 0) *info = k0 + j + 1;
                }
            }
        }
        __syncthreads();
        const int q = s_p[0];
        const bool usable = s_p[1] != 0;
        if (tid == 0 && q >= row0 && q < row0 + nloc) slot[q - row0] = j;   // retired: nobody writes this row again
        __syncthreads();
        // ---- rank-1 update of the active rows; candidate of column j + 1 on the way ----
        mine.val = -1.0; mine.row = -1; mine.pos = 0x7fffffff;
        if (usable) {
            const double pv = s_row[j];
            double* Pj = P + (size_t)j * rpc;
            for (int r = tid; r < nloc; r += kClusterPanelThreads) {
                if (slot[r] >= 0) continue;
                const double l = Pj[r] / pv;
                Pj[r] = l;
                for (int c = j + 1; c < kb; ++c) P[(size_t)c * rpc + r] = fma(-l, s_row[c], P[(size_t)c * rpc + r]);
                if (j + 1 < kb) {
                    const double v = fabs(P[(size_t)(j + 1) * rpc + r]);
                    const int phys = row0 + r;
                    const int ps = (phys - k0 < kb) ? loc_top[phys - k0] : phys;
                    if (v > mine.val || (v == mine.val && ps < mine.pos)) { mine.val = v; mine.row = phys; mine.pos = ps; }
                }
            }
        } else if (j + 1 < kb) {
            for (int r = tid; r < nloc; r += kClusterPanelThreads) {
                if (slot[r] >= 0) continue;
                const double v = fabs(P[(size_t)(j + 1) * rpc + r]);
                const int phys = row0 + r;
                const int ps = (phys - k0 < kb) ? loc_top[phys - k0] : phys;
                if (v > mine.val || (v == mine.val && ps < mine.pos)) { mine.val = v; mine.row = phys; mine.pos = ps; }
            }
        }
        // (no barrier here: the next column's reduction starts with __syncthreads after the warp candidates, and the candidate
        //  buffers alternate; s_row / s_p are rewritten only after the next cluster barrier, which every thread of every CTA passes
        //  after finishing this loop)
    }
    cluster.sync();   // nobody reads a peer's panel any more
    // ---- write the panel back in its final row order ----
    for (int idx = tid; idx < kb * rpc; idx += kClusterPanelThreads) {
        const int c = idx / rpc, r = idx - c * rpc;
        if (r >= nloc) continue;
        const int phys = row0 + r;
        const int sj = slot[r];
        const int where = sj >= 0 ? k0 + sj : ((phys - k0 < kb) ? loc_top[phys - k0] : phys);
        A[(size_t)(k0 + c) * n + where] = P[idx];
    }
}

// ---- the panel on P co-resident CTAs that meet once per column at a barrier in global memory ---------------------------------------
// Same algorithm as the cluster kernel above (rows never move, the pivot row is retired where it lies, LAPACK's interchanges are
// replayed on index maps, the panel is written back in its final row order), without clusters: CTA c keeps rows
// [c * kCoopRows, (c + 1) * kCoopRows) of the panel in its OWN shared memory (64 KB), and what the CTAs exchange per column goes
// through L2: every CTA publishes its pivot candidate TOGETHER WITH that row's 32 panel entries (272 bytes), arrives at a monotonic
// counter, spins until all P have arrived, and reads the P candidates -- the winner's record already holds the pivot row, so one
// round trip per column is all there is.  Records are double-buffered by column parity.  P <= 128 CTAs of 256 threads are always
// co-resident on a B200.  Measured (profiles/r02q_lu_panels.txt): ~4.4 us per column against ~12 us for the one-CTA kernel at n = 4096
// (whose every access to the panel is a round trip to L2); whole factorisation 37 ms against 70 at n = 4096, 287 against 618 at
// n = 12288, 3.8 against 8.3 at n = 512, equal at n = 700 ... 1300; the same pivots and the same solution in every case.
constexpr int kCoopRows = 256;             // panel rows per CTA
constexpr int kCoopThreads = 256;
constexpr int kCoopMaxCtas = 128;          // panels of up to 32768 rows; 128 CTAs of 256 threads and 64 KB are co-resident on 148 SMs

struct CoopCand {
    double val;
    int row, pos;
    double entries[kNB];
};

__global__ void __launch_bounds__(kCoopThreads) lu_panel_coop_kernel(double* __restrict__ A, int n, int k0, int kb, int* __restrict__ piv,
                                                                     int* __restrict__ info, CoopCand* __restrict__ cands /* [2][P] */,
                                                                     unsigned int* bar, unsigned int bar_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* P = reinterpret_cast<double*>(smem_raw);                 // [kNB][kCoopRows]
    __shared__ int slot[kCoopRows];                                   // -1 active, j: retired as pivot of column j
    __shared__ PivotCand wcand[kCoopThreads / 32];
    __shared__ double s_row[kNB];
    __shared__ int pos_top[kNB], loc_top[kNB], s_p[2];
    constexpr int rpc = kCoopRows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncta = gridDim.x, rank = blockIdx.x;
    const int row0 = k0 + rank * rpc;
    const int nloc = max(0, min(rpc, n - row0));
    for (int idx = tid; idx < kb * rpc; idx += kCoopThreads) {
        const int c = idx / rpc, r = idx - c * rpc;
        P[idx] = r < nloc ? A[(size_t)(k0 + c) * n + row0 + r] : 0.0;
    }
    for (int r = tid; r < rpc; r += kCoopThreads) slot[r] = -1;
    if (tid < kNB) {
        pos_top[tid] = k0 + tid;
        loc_top[tid] = k0 + tid;
    }
    __syncthreads();
    PivotCand mine;
    mine.val = -1.0; mine.row = -1; mine.pos = 0x7fffffff;
    for (int r = tid; r < nloc; r += kCoopThreads) {
        const double v = fabs(P[r]);
        if (v > mine.val) { mine.val = v; mine.row = row0 + r; mine.pos = row0 + r; }
    }
    for (int j = 0; j < kb; ++j) {
        const int buf = j & 1;
        // ---- this CTA's candidate ----
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, mine.val, o);
            const int orow = __shfl_down_sync(0xffffffffu, mine.row, o);
            const int opos = __shfl_down_sync(0xffffffffu, mine.pos, o);
            if (ov > mine.val || (ov == mine.val && opos < mine.pos)) { mine.val = ov; mine.row = orow; mine.pos = opos; }
        }
        if (lane == 0) wcand[warp] = mine;
        __syncthreads();
        if (warp == 0) {
            PivotCand c = lane < kCoopThreads / 32 ? wcand[lane] : PivotCand{-1.0, -1, 0x7fffffff};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, c.val, o);
                const int orow = __shfl_down_sync(0xffffffffu, c.row, o);
                const int opos = __shfl_down_sync(0xffffffffu, c.pos, o);
                if (ov > c.val || (ov == c.val && opos < c.pos)) { c.val = ov; c.row = orow; c.pos = opos; }
            }
            c.val = __shfl_sync(0xffffffffu, c.val, 0);
            c.row = __shfl_sync(0xffffffffu, c.row, 0);
            c.pos = __shfl_sync(0xffffffffu, c.pos, 0);
            // publish: (value, row, position) and the candidate row's panel entries
            CoopCand* out = cands + (size_t)buf * ncta + rank;
            if (lane == 0) { out->val = c.val; out->row = c.row; out->pos = c.pos; }
            if (lane < kb) out->entries[lane] = c.row >= 0 ? P[(size_t)lane * rpc + (c.row - row0)] : 0.0;
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicAdd(bar, 1u);
                const unsigned int target = bar_base + (unsigned)(j + 1) * (unsigned)ncta;
                while (*reinterpret_cast<volatile unsigned int*>(bar) < target) {   // (a load, not an atomic: the pollers must not queue
                }                                                                    //  in front of the arrivals at the L2 atomic unit)
                __threadfence();
            }
            __syncwarp();
            // ---- the winner, identically on every CTA ----
            PivotCand w{-1.0, -1, 0x7fffffff};
            int wcta = lane;
            for (int c2 = lane; c2 < ncta; c2 += 32) {
                const CoopCand* cc = cands + (size_t)buf * ncta + c2;
                const double v = __ldcg(&cc->val);
                const int vpos = __ldcg(&cc->pos);
                if (v > w.val || (v == w.val && vpos < w.pos)) { w.val = v; w.row = __ldcg(&cc->row); w.pos = vpos; wcta = c2; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, w.val, o);
                const int orow = __shfl_down_sync(0xffffffffu, w.row, o);
                const int opos = __shfl_down_sync(0xffffffffu, w.pos, o);
                const int octa = __shfl_down_sync(0xffffffffu, wcta, o);
                if (ov > w.val || (ov == w.val && opos < w.pos)) { w.val = ov; w.row = orow; w.pos = opos; wcta = octa; }
            }
            w.val = __shfl_sync(0xffffffffu, w.val, 0);
            w.row = __shfl_sync(0xffffffffu, w.row, 0);
            w.pos = __shfl_sync(0xffffffffu, w.pos, 0);
            wcta = __shfl_sync(0xffffffffu, wcta, 0);
            const bool usable = w.val > 0.0;
            int q = w.row, qpos = w.pos;
            if (usable) {
                if (lane < kb) s_row[lane] = __ldcg(&(cands + (size_t)buf * ncta + wcta)->entries[lane]);
            } else {
                // zero (or NaN) column: singular as getrf reports it; "pivot" = the row at the diagonal position, no elimination
                q = pos_top[j];
                qpos = k0 + j;
                if (lane < kb) s_row[lane] = 0.0;
            }
            if (lane == 0) {
                s_p[0] = q;
                s_p[1] = usable ? 1 : 0;
                const int u = pos_top[j];
                if (qpos - k0 < kb) pos_top[qpos - k0] = u;
                loc_top[u - k0] = qpos;
                if (q - k0 < kb) loc_top[q - k0] = k0 + j;
                pos_top[j] = q;
                if (rank == 0) {
                    piv[k0 + j] = qpos;
                    if (!usable && *info == 0) *info = k0 + j + 1;
                }
            }
        }
        __syncthreads();
        const int q = s_p[0];
        const bool usable = s_p[1] != 0;
        if (tid == 0 && q >= row0 && q < row0 + nloc) slot[q - row0] = j;
        __syncthreads();
        mine.val = -1.0; mine.row = -1; mine.pos = 0x7fffffff;
        if (usable) {
            const double pv = s_row[j];
            double* Pj = P + (size_t)j * rpc;
            for (int r = tid; r < nloc; r += kCoopThreads) {
                if (slot[r] >= 0) continue;
                const double l = Pj[r] / pv;
                Pj[r] = l;
                for (int c = j + 1; c < kb; ++c) P[(size_t)c * rpc + r] = fma(-l, s_row[c], P[(size_t)c * rpc + r]);
                if (j + 1 < kb) {
                    const double v = fabs(P[(size_t)(j + 1) * rpc + r]);
                    const int phys = row0 + r;
                    const int ps = (phys - k0 < kb) ? loc_top[phys - k0] : phys;
                    if (v > mine.val || (v == mine.val && ps < mine.pos)) { mine.val = v; mine.row = phys; mine.pos = ps; }
                }
            }
        } else if (j + 1 < kb) {
            for (int r = tid; r < nloc; r += kCoopThreads) {
                if (slot[r] >= 0) continue;
                const double v = fabs(P[(size_t)(j + 1) * rpc + r]);
                const int phys = row0 + r;
                const int ps = (phys - k0 < kb) ? loc_top[phys - k0] : phys;
                if (v > mine.val || (v == mine.val && ps < mine.pos)) { mine.val = v; mine.row = phys; mine.pos = ps; }
            }
        }
    }
    __syncthreads();
    // ---- write the panel back in its final row order (every CTA has loaded its rows long ago: it passed the first barrier) ----
    for (int idx = tid; idx < kb * rpc; idx += kCoopThreads) {
        const int c = idx / rpc, r = idx - c * rpc;
        if (r >= nloc) continue;
        const int phys = row0 + r;
        const int sj = slot[r];
        const int where = sj >= 0 ? k0 + sj : ((phys - k0 < kb) ? loc_top[phys - k0] : phys);
        A[(size_t)(k0 + c) * n + where] = P[idx];
    }
}
#endif   // RB_EMULATE

// column c of the augmented matrix [A | b]: c == n addresses b
__device__ __forceinline__ double* aug_column(double* A, double* b, int n, int c) { return c < n ? A + (size_t)c * n : b; }

// one thread per column right of the panel (b included): the panel's row interchanges, then U12 = L11^{-1} A12 with L11 (unit lower
// triangular) in shared memory and the 32 entries of the column in registers
// (columns c_begin <= c <= c_last of the augmented matrix: the look-ahead schedule does the next panel's columns first)
__global__ void __launch_bounds__(128) lu_swap_trsm_kernel(double* __restrict__ A, double* __restrict__ b, int n, int k0, int kb,
                                                           const int* __restrict__ piv, int c_begin, int c_last) {
    __shared__ double L[kNB][kNB + 1];
    __shared__ int s_piv[kNB];
    for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) {
        const int i = idx % kb, j = idx / kb;
        L[i][j] = A[(size_t)(k0 + j) * n + k0 + i];
    }
    if ((int)threadIdx.x < kb) s_piv[threadIdx.x] = piv[k0 + threadIdx.x];
    __syncthreads();
    const int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (c > c_last) return;
    double* colp = aug_column(A, b, n, c);
    for (int j = 0; j < kb; ++j) {
        const int r = k0 + j, p = s_piv[j];
        if (p != r) {
            const double t = colp[r];
            colp[r] = colp[p];
            colp[p] = t;
        }
    }
    colp += k0;
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = i < kb ? colp[i] : 0.0;
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
#pragma unroll
        for (int i = j + 1; i < kNB; ++i)
            if (i < kb) x[i] -= L[i][j] * x[j];   // unit lower triangular
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i)
        if (i < kb) colp[i] = x[i];
}

#ifndef RB_EMULATE
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

constexpr int kTile = 64;
constexpr int kGemmThreads = 256;

// C[r0.., c0..] -= L21[r0.., 0..kb) * U12[0..kb, c0..), the trailing block starting at row/column t0 = k0 + kb of the n x n matrix;
// blockIdx.y + ct0 is the column tile
__global__ void __launch_bounds__(kGemmThreads) lu_gemm_kernel(double* __restrict__ A, int n, int k0, int kb, int ct0) {
    __shared__ double As[kTile][kNB + 1];   // -L21 tile, As[row][k]
    __shared__ double Bs[kNB][kTile + 1];   //  U12 tile, Bs[k][col]
    const int t0 = k0 + kb;
    const int r0 = t0 + blockIdx.x * kTile, c0 = t0 + (blockIdx.y + ct0) * kTile;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < kTile * kNB; idx += kGemmThreads) {
        const int i = idx % kTile, j = idx / kTile;   // consecutive threads: consecutive rows of one column of L21
        const int r = r0 + i;
        As[i][j] = (r < n && j < kb) ? -A[(size_t)(k0 + j) * n + r] : 0.0;
    }
    for (int idx = tid; idx < kNB * kTile; idx += kGemmThreads) {
        const int j = idx % kNB, c = idx / kNB;       // consecutive threads: consecutive rows of one column of U12
        const int cc = c0 + c;
        Bs[j][c] = (cc < n && j < kb) ? A[(size_t)cc * n + k0 + j] : 0.0;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;            // fragment coordinates of mma.m8n8k4.f64
    const int rb = (warp & 3) * 16, cb = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int r = r0 + rb + rt * 8 + g, c = c0 + cb + ct * 8 + 2 * t;
            acc[rt][ct][0] = (r < n && c < n) ? A[(size_t)c * n + r] : 0.0;
            acc[rt][ct][1] = (r < n && c + 1 < n) ? A[(size_t)(c + 1) * n + r] : 0.0;
        }
#pragma unroll
    for (int kk = 0; kk < kNB / 4; ++kk) {
        double a[2], bf[4];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) a[rt] = As[rb + rt * 8 + g][kk * 4 + t];
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) bf[ct] = Bs[kk * 4 + t][cb + ct * 8 + g];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) dmma_m8n8k4(acc[rt][ct][0], acc[rt][ct][1], a[rt], bf[ct]);
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int r = r0 + rb + rt * 8 + g, c = c0 + cb + ct * 8 + 2 * t;
            if (r < n && c < n) A[(size_t)c * n + r] = acc[rt][ct][0];
            if (r < n && c + 1 < n) A[(size_t)(c + 1) * n + r] = acc[rt][ct][1];
        }
}

__global__ void lu_gemv_kernel(const double* __restrict__ A, double* __restrict__ b, int n, int k0, int kb) {
    __shared__ double bs[kNB];
    if ((int)threadIdx.x < kb) bs[threadIdx.x] = b[k0 + threadIdx.x];
    __syncthreads();
    const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = b[i];
    for (int j = 0; j < kb; ++j) acc -= A[(size_t)(k0 + j) * n + i] * bs[j];
    b[i] = acc;
}

// U x = b in place, 32 columns at a time from the bottom: warp 0 solves the 32 x 32 triangle in shared memory, then every thread
// subtracts that block's contribution from the rows above it (coalesced down the columns).  Two block-wide barriers per 32 columns
// instead of two per column.
constexpr int kBackThreads = 1024;
__global__ void __launch_bounds__(kBackThreads) lu_backsolve_blocked_kernel(const double* __restrict__ A, double* __restrict__ b, int n) {
    __shared__ double Us[kNB][kNB + 1];
    __shared__ double xs[kNB];
    const int tid = threadIdx.x;
    for (int hi = n; hi > 0; hi -= kNB) {
        const int k0 = hi > kNB ? hi - kNB : 0, w = hi - k0;
        for (int idx = tid; idx < w * w; idx += kBackThreads) {
            const int i = idx % w, j = idx / w;
            Us[i][j] = A[(size_t)(k0 + j) * n + k0 + i];
        }
        if (tid < w) xs[tid] = b[k0 + tid];
        __syncthreads();
        if (tid < 32) {
            for (int j = w - 1; j >= 0; --j) {
                if (tid == j) xs[j] /= Us[j][j];
                __syncwarp();
                if (tid < j) xs[tid] -= Us[tid][j] * xs[j];
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += kBackThreads) {
            double acc = b[i];
            for (int j = 0; j < w; ++j) acc -= A[(size_t)(k0 + j) * n + i] * xs[j];
            b[i] = acc;
        }
        if (tid < w) b[k0 + tid] = xs[tid];
        __syncthreads();
    }
}

}  // namespace

#ifndef RB_EMULATE
namespace {

size_t cluster_panel_smem(int rpc) {
    return (size_t)kNB * rpc * sizeof(double) + (size_t)rpc * sizeof(int) + 2 * sizeof(PivotCand) + kNB * sizeof(double) +
           (kClusterPanelThreads / 32) * sizeof(PivotCand) + 64;
}

// 0: not usable (too many rows for the cluster's shared memory, or the launch was refused), else the cluster size used
template <int CS>
bool try_cluster_panel(double* A, int n, int k0, int kb, int* piv, int* info, cudaStream_t st) {
    static int state = 0;   // 0 untried, 1 configured, -1 refused by the device
    if (state < 0) return false;
    const int m = n - k0;
    const int rpc = std::max(32, ((m + CS - 1) / CS + 31) / 32 * 32);
    const size_t smem = cluster_panel_smem(rpc);
    if (smem > 227 * 1024) return false;
    auto kern = lu_panel_cluster_kernel<CS>;
    if (state == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess && CS > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) {
            cudaGetLastError();
            state = -1;
            return false;
        }
        state = 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(kClusterPanelThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, A, n, k0, kb, rpc, piv, info);
    if (e != cudaSuccess) {
        cudaGetLastError();
        state = -1;   // e.g. a device partition that cannot co-schedule this cluster size: use the one-CTA panel from now on
        return false;
    }
    return true;
}

// Measured on a B200 (profiles/r02e_implicit_report.json, r02p_lu_profile_*.log): correct (tests/test_zz_gpu_implicit.py runs it at
// n = 257 ... 7000, clusters of 8 and 16), but NOT faster than the one-CTA panel -- n = 1536: 26 ms against 17, n = 4096: 70 against 67,
// n = 6144: 145 against 165, and erratic from call to call (72 ... 600 ms at n = 4096; a launch of one 8-CTA cluster with > 100 KB
// of dynamic shared memory per CTA between small ordinary launches appears to cost far more than the kernel itself).  It is therefore
// opt-in (RB_LU_PANEL=cluster, or the older RB_LU_CLUSTER_PANEL=1); the default panel is the cooperative kernel above.
bool cluster_panels_enabled() {
    const char* v = std::getenv("RB_LU_CLUSTER_PANEL");
    return v && std::atoi(v) != 0;
}

}  // namespace
#endif

#ifndef RB_EMULATE
namespace {
bool env_flag(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) != 0 : dflt != 0;
}

// which panel kernel: RB_LU_PANEL = coop (default: P CTAs meeting through L2), one (one CTA), cluster (distributed shared memory)
int panel_choice() {
    const char* v = std::getenv("RB_LU_PANEL");
    if (v && !std::strcmp(v, "one")) return 0;
    if (v && !std::strcmp(v, "cluster")) return 2;
    if (cluster_panels_enabled()) return 2;
    return 1;
}
}  // namespace
#endif

// the factorisation with b eliminated on the fly: on return A holds U in its upper triangle and b holds L^{-1} P b
void lu_factor_blocked(double* A, double* b, int n, int* info, cudaStream_t st) {
    int* piv = nullptr;   // pivot rows, stream-ordered allocation: no synchronisation, no state shared between streams
#ifndef RB_EMULATE
    {   // keep the pool's memory across synchronisations (the default threshold of 0 hands it back to the OS at every one, and the
        // next factorisation's first allocation then maps fresh memory while kernels are already queued)
        static thread_local bool pool_configured = false;
        if (!pool_configured) {
            int dev = 0;
            cudaMemPool_t pool = nullptr;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = 64ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            pool_configured = true;
        }
    }
#endif
    RB_CUDA(cudaMallocAsync(&piv, (size_t)std::max(n, 1) * sizeof(int), st));
    RB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    int launches = 0;
#ifndef RB_EMULATE
    const int choice = panel_choice();
    CoopCand* cands = nullptr;
    unsigned int* bar = nullptr;
    unsigned int bar_total = 0;
    if (choice == 1) {
        static bool configured = false;
        if (!configured) {
            RB_CUDA(cudaFuncSetAttribute(lu_panel_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNB * kCoopRows * (int)sizeof(double)));
            configured = true;
        }
        RB_CUDA(cudaMallocAsync(&cands, (size_t)2 * kCoopMaxCtas * sizeof(CoopCand), st));
        RB_CUDA(cudaMallocAsync(&bar, sizeof(unsigned int), st));
        RB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
    }
#endif
    // one panel factorisation, on stream q
    auto panel = [&](int k0, int kb, cudaStream_t q) {
        bool done = false;
#ifndef RB_EMULATE
        const int m_panel = n - k0;
        const int ncta = (m_panel + kCoopRows - 1) / kCoopRows;
        if (choice == 1 && m_panel > kCoopRows && ncta <= kCoopMaxCtas) {
            // P co-resident CTAs, one barrier through L2 per column (a panel of <= 256 rows is one CTA's work anyway)
            lu_panel_coop_kernel<<<ncta, kCoopThreads, kNB * kCoopRows * sizeof(double), q>>>(A, n, k0, kb, piv, info, cands, bar, bar_total);
            bar_total += (unsigned)ncta * (unsigned)kb;
            done = true;
        } else if (choice == 2 && m_panel >= 256) {
            // the panel in the distributed shared memory of a cluster of 8 CTAs (16 when the panel is too tall for 8)
            done = try_cluster_panel<8>(A, n, k0, kb, piv, info, q);
            if (!done) done = try_cluster_panel<16>(A, n, k0, kb, piv, info, q);
        }
#endif
        if (!done) RB_LAUNCH(lu_panel_kernel, 1, kPanelThreads, q, A, n, k0, kb, piv, info);
        launches++;
    };
    auto swap_trsm = [&](int k0, int kb, int c_begin, int c_last, cudaStream_t q) {
        if (c_last < c_begin) return;
        RB_LAUNCH(lu_swap_trsm_kernel, (c_last - c_begin + 1 + 127) / 128, 128, q, A, b, n, k0, kb, (const int*)piv, c_begin, c_last);
        launches++;
    };
    bool lookahead = false;
#ifndef RB_EMULATE
    // Look-ahead of one panel: the row interchanges, the triangular solve and the trailing update of panel k are done FIRST for
    // the 64 columns that hold the next two panels; then panel k + 1 is factorised on a second stream while this stream does the
    // same for the rest of the matrix (disjoint columns; the columns left of a panel are never touched again, b is eliminated
    // on the fly).  The critical path per panel is the panel itself plus two small launches instead of panel + swap/TRSM + GEMM.
    static thread_local cudaStream_t side = nullptr;
    static thread_local cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    lookahead = env_flag("RB_LU_LOOKAHEAD", n >= 4096 ? 1 : 0);   // measured: 13.0 -> 17.3 ms at n = 2048, 23.3 -> 27.6 at 3072, 44 -> 39 at 4096, 108 -> 79 at 7000, 272 -> 231 at 12288
    if (lookahead && !side) {
        RB_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
    }
    if (lookahead) {
        RB_CUDA(cudaEventRecord(ev_main, st));              // allocations and memsets above
        RB_CUDA(cudaStreamWaitEvent(side, ev_main, 0));
        panel(0, std::min(kNB, n), side);
        RB_CUDA(cudaEventRecord(ev_side, side));
        for (int k0 = 0; k0 < n; k0 += kNB) {
            const int kb = std::min(kNB, n - k0);
            const int t0 = k0 + kb, m = n - t0;
            RB_CUDA(cudaStreamWaitEvent(st, ev_side, 0));   // panel k is factorised
            const int a_last = std::min(t0 + kTile, n) - 1;  // first column tile: the next two panels' columns
            const int tiles = (m + kTile - 1) / kTile;
            if (m > 0) {
                swap_trsm(k0, kb, t0, a_last, st);
                RB_LAUNCH(lu_gemm_kernel, dim3(tiles, 1), kGemmThreads, st, A, n, k0, kb, 0);
                launches++;
                RB_CUDA(cudaEventRecord(ev_main, st));
                RB_CUDA(cudaStreamWaitEvent(side, ev_main, 0));
                panel(t0, std::min(kNB, n - t0), side);
                RB_CUDA(cudaEventRecord(ev_side, side));
            }
            swap_trsm(k0, kb, std::max(a_last + 1, t0), n, st);   // the other columns and b
            if (m > 0) {
                if (tiles > 1) {
                    RB_LAUNCH(lu_gemm_kernel, dim3(tiles, tiles - 1), kGemmThreads, st, A, n, k0, kb, 1);
                    launches++;
                }
                RB_LAUNCH(lu_gemv_kernel, (m + 255) / 256, 256, st, (const double*)A, b, n, k0, kb);
                launches++;
            }
        }
    }
#endif
    for (int k0 = 0; !lookahead && k0 < n; k0 += kNB) {
        const int kb = std::min(kNB, n - k0);
        panel(k0, kb, st);
        swap_trsm(k0, kb, k0 + kb, n, st);
        const int m = n - k0 - kb;
        if (m > 0) {
            const int tiles = (m + kTile - 1) / kTile;
            RB_LAUNCH(lu_gemm_kernel, dim3(tiles, tiles), kGemmThreads, st, A, n, k0, kb, 0);
            RB_LAUNCH(lu_gemv_kernel, (m + 255) / 256, 256, st, (const double*)A, b, n, k0, kb);
            launches += 2;
        }
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(piv, st);
#ifndef RB_EMULATE
    if (cands) cudaFreeAsync(cands, st);
    if (bar) cudaFreeAsync(bar, st);
#endif
    RB_CUDA(e);
    count_launch(launches);
}

void launch_lu_solve_blocked(double* A, double* b, int n, int* info, cudaStream_t st) {
    lu_factor_blocked(A, b, n, info, st);
    RB_LAUNCH(lu_backsolve_blocked_kernel, 1, kBackThreads, st, (const double*)A, b, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
