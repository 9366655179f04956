// pair_kernels3.cu -- warp-per-row-group variant of the O(N^2) cotangent-sum sweep for small surfaces (one member, N <= 8192).
//
// Same mathematics as pair_kernels.cu / pair_kernels2.cu (far form: 10 FP64-pipe instructions per pair through the global
// exponentials; near cells: cell-local expm1 coordinates, 13 instructions).  What differs is who adds up a row:
//   tiled kernel (pair_kernels.cu):    a row's sources are cut into chunks handled by different CTAs; partial sums go through global
//                                      memory and two levels of tickets (~ 12 dependent L2 round trips after the last pair);
//   persistent kernel (pair_kernels2): lanes = rows, warps = source groups, combined through shared memory; a row block is 32 R rows,
//                                      so N = 4096 has only 128 (R = 1) or 32 (R = 4) blocks for 148 SMs;
//   here:                              a WARP owns R consecutive rows and walks over ALL sources, lane l taking sources l, l + 32, ...
//                                      of every 256-point cell from shared memory (structure of arrays: conflict-free 16-byte loads,
//                                      each load feeds R pairs); the row sums are finished by one xor-butterfly over the lanes.  No
//                                      partial sums in global memory, no cross-warp combine, one ticket per CTA at the very end.
// Grid: one CTA per SM; CTA c owns the row groups [c NG / G, (c + 1) NG / G), one per warp (W = ceil(NG / G) <= 16 warps).  The
// sources are staged once per tile of <= 4096 points (N <= 4096: the whole surface, 160 KB); the cell-local coordinates only for the
// <= 4 cells that are near the CTA's rows.  At N = 4096 the N^2 pair loop is ~10 us of FP64-pipe time per sweep; the tiled kernel
// needed 35 us per sweep there because of its serial tail.
#include "internal.cuh"
#include "pair_common.cuh"

namespace rb {

namespace {

constexpr int kV3MaxWarps = 16;
constexpr int kV3NearCells = 4;                      // cells around the CTA's rows staged in cell-local coordinates
constexpr int kV3Near = kV3NearCells * kCell;
constexpr int kV3Pad = 32;                           // the pair loops load one source (stride 32) ahead
constexpr int kV3Iter = kCell / 32;                  // sources per lane and cell

// far form, lane-strided sources: U_k = sum_j F_j / |d|^2, V_k = sum_j g_j / |d|^2 (see accumulate2_far in pair_kernels2.cu)
template <int R, bool DIAG>
__device__ __forceinline__ void accumulate3_far(const double2* __restrict__ se, const double2* __restrict__ sf, const double* __restrict__ sg,
                                                const double2 (&ek)[R], const int (&sd)[R], double2 (&U)[R], double (&V)[R]) {
    double2 e = se[0], f = sf[0];
    double gj = sg[0];
#pragma unroll
    for (int s = 0; s < kV3Iter; ++s) {
        const double2 en = se[32 * (s + 1)], fn = sf[32 * (s + 1)];   // (past the last cell: padding, loaded, never used)
        const double gn = sg[32 * (s + 1)];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double dr = ek[r].x - e.x;
            const double di = ek[r].y - e.y;
            const double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp2(n2);
            if (DIAG) inv = (s == sd[r]) ? 0.0 : inv;
            U[r].x = fma(f.x, inv, U[r].x);
            U[r].y = fma(f.y, inv, U[r].y);
            V[r] = fma(gj, inv, V[r]);
        }
        e = en;
        f = fn;
        gj = gn;
    }
}

// near form in cell-local coordinates (see accumulate2 in pair_kernels2.cu)
template <int R, bool DIAG>
__device__ __forceinline__ void accumulate3_near(const double2* __restrict__ ne, const double2* __restrict__ nf, const double2 (&ek)[R],
                                                 const int (&sd)[R], double2 (&acc)[R]) {
    double2 e = ne[0], f = nf[0];
#pragma unroll
    for (int s = 0; s < kV3Iter; ++s) {
        const double2 en = ne[32 * (s + 1)], fn = nf[32 * (s + 1)];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double dr = ek[r].x - e.x;
            const double di = ek[r].y - e.y;
            const double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp2(n2);
            if (DIAG) inv = (s == sd[r]) ? 0.0 : inv;
            const double tr = fma(f.y, di, f.x * dr);
            const double ti = fma(f.y, dr, -(f.x * di));
            acc[r].x = fma(tr, inv, acc[r].x);
            acc[r].y = fma(ti, inv, acc[r].y);
        }
        e = en;
        f = fn;
    }
}

// deterministic sum over the block: warp butterflies, then the warp sums in warp order (identical in every CTA of a launch)
__device__ __forceinline__ double block_sum3(double v, double* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, W = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sred[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < W; ++i) r += sred[i];
    return r;
}

// what the epilogue of one row reads besides the row sum (loaded before the pair loops where registers allow)
struct RowOps {
    double xk = 0.0, md = 0.0, b = 0.0, y = 0.0;
    double2 zp = {1.0, 0.0}, v1d = {0.0, 0.0}, v2 = {0.0, 0.0}, ap = {0.0, 0.0};
};

// o: index of the row in the (batched) arrays
template <int MODE>
__device__ __forceinline__ void load_row_ops(const SweepArgs& a, size_t o, RowOps& q) {
    q.xk = a.x[o];
    if (MODE != kSweepRAW) {
        q.zp = a.g.Zp[o];
        q.md = a.g.Mdiag[o];
        if (!a.apply_only) q.b = a.g.b[o];
    }
    if (MODE == kSweepVEL) {
        q.v1d = a.g.V1diag[o];
        q.v2 = a.g.V2[o];
        if (!a.defer_aprime) q.ap = a.aprime[o];   // deferred: finish_solve adds V2 a'
        if (a.dphi && !a.defer_aprime) q.y = a.g.Z[o].y;
    }
}

// the epilogue of row o from its finished sum `mine` (MV: Re(Zp T) in mine.x; otherwise T); returns the squared residual of the row
template <int MODE>
__device__ __forceinline__ double finish_row(const SweepArgs& a, size_t o, const RowOps& q, double sumx, double2 mine) {
    const double inv4pi = 0.25 / kPi;
    const double xk = q.xk;
    const double Ar = (sumx - xk) + 2.0 * mine.x;
    const double Ai = 2.0 * mine.y;
    double sr = 0.0;
    if (MODE == kSweepMV) {
        const double Mx = fma(q.md, xk, a.cK * fma(q.zp.x, sumx - xk, 2.0 * mine.x));
        if (a.apply_only) {
            mirror_store2(a.comm, a.x_out + o, Mx);
        } else {
            const double res = q.b - Mx;
            mirror_store2(a.comm, a.x_out + o, fma(a.omega, res, xk));
            sr = res * res;
        }
    } else if (MODE == kSweepVEL) {
        const double2 zp = q.zp;
        const double wr = inv4pi * Ar + q.v1d.x * xk + (q.v2.x * q.ap.x - q.v2.y * q.ap.y);
        const double wi = inv4pi * Ai + q.v1d.y * xk + (q.v2.x * q.ap.y + q.v2.y * q.ap.x);
        mirror_store2(a.comm, a.vel_lower + o, make_double2(wr, -wi));
        const double inv = 1.0 / (zp.x * zp.x + zp.y * zp.y);
        const double azx = xk * zp.x * inv, azy = -xk * zp.y * inv;     // a_k / Zp_k
        a.vel_upper[o] = make_double2(wr - azx, -(wi - azy));
        if (a.dphi && !a.defer_aprime) {
            double d;
            if (a.rhs_phi_kind == 1) {
                d = -q.y + 0.5 * (wr * wr + wi * wi);
            } else {
                const double vdw = a.depth / 3.0;
                d = vdw * pow(1.0 + q.y / a.depth, -3.0) - vdw + (0.5 * wr * wr + 0.5 * wi * wi);
            }
            mirror_store2(a.comm, a.dphi + o, make_double2(d, 0.0));
        }
        if (a.combined) {   // verify the iterate with the same row sum: r = b - M a; next iterate in case it is needed
            if (a.A_out) mirror_store2(a.comm, a.A_out + o, make_double2(Ar, Ai));
            const double res = q.b - fma(q.md, xk, a.cK * (zp.x * Ar - zp.y * Ai));
            mirror_store2(a.comm, a.x_out + o, fma(a.omega, res, xk));
            sr = res * res;
        }
    } else {
        a.raw_out[o] = make_double2(-Ai, Ar);
    }
    return sr;
}

}  // namespace

template <int MODE, int R, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) sweep3_kernel(const SweepArgs a, const int TS, const int S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* se = reinterpret_cast<double2*>(smem_raw);        // [TS + pad]   E_j
    double2* sf = se + TS + kV3Pad;                            // [TS + pad]   F_j = x_j E_j
    double2* ne = sf + TS + kV3Pad;                            // [near + pad] P_j (cell-local)
    double2* nf = ne + kV3Near + kV3Pad;                       // [near + pad] x_j (1 + P_j)
    double* sg = reinterpret_cast<double*>(nf + kV3Near + kV3Pad);   // [TS + pad] g_j = x_j |E_j|^2
    double* sred = sg + TS + kV3Pad;                           // [32]
    double2* comb = reinterpret_cast<double2*>(sred + 32);     // [warps][R] row sums of the warps that share a row group
    __shared__ unsigned int s_ticket;
    constexpr bool REALPATH = (MODE == kSweepMV);

    // "is this solve finished already?": the flag is loaded now and looked at after the prologue's loads have been issued (before
    // the first barrier), so that it does not cost a round trip of its own
    int solve_done = 0;
    if ((MODE == kSweepMV || MODE == kSweepVEL) && a.skip_if_done) solve_done = *reinterpret_cast<volatile int*>(&a.ctrl->done);
    const int t = threadIdx.x, T = blockDim.x, lane = t & 31, w = t >> 5;
    const int N = a.N;
    const int rows_total = a.v2_row_end - a.v2_row_begin;
    const int NG = (rows_total + R - 1) / R;                                        // row groups of this rank
    const int i0 = (int)(((long long)blockIdx.x * NG) / gridDim.x);
    const int i1 = (int)(((long long)(blockIdx.x + 1) * NG) / gridDim.x);
    // S warps share a row group: warp (group, h) takes the cells c = h (mod S) of every tile
    const int item = i0 + w / S;
    const int h = w - (w / S) * S;
    const bool active = item < i1;
    const double* __restrict__ x = a.x;
    const double2* __restrict__ EG = a.g.EG;
    const double2* __restrict__ P0 = a.g.P0;

    // rows of this CTA and the window of cells staged in cell-local coordinates
    const int cta_row0 = a.v2_row_begin + i0 * R;
    const int cta_rend = min(a.v2_row_begin + i1 * R, a.v2_row_end);
    const int cB0 = min(cta_row0, N - 1) / kCell, cB1 = (max(cta_rend, cta_row0 + 1) - 1) / kCell;
    const int nc0 = (cB0 - 1 + a.ncell) % a.ncell;
    const int nnear = a.use_local ? min(a.ncell, min(kV3NearCells, cB1 - cB0 + 3)) : 0;

    int krow[R];
    bool valid[R];
    double2 acc[R], ek[R], ekG[R], U[R];
    double V[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        krow[r] = a.v2_row_begin + item * R + r;
        valid[r] = active && krow[r] < a.v2_row_end;
        acc[r] = make_double2(0.0, 0.0);
        U[r] = make_double2(0.0, 0.0);
        V[r] = 0.0;
        ek[r] = make_double2(3.0e150, 0.0);
        ekG[r] = valid[r] ? EG[krow[r]] : make_double2(3.0e150, 0.0);
    }
    const int cellK = min(active ? krow[0] : cta_row0, N - 1) / kCell;
    int cur_variant = -1;
    // operands of the epilogue, loaded now (they do not depend on the row sums): no L2 round trip after the last pair.
    // Lane r < R finishes row r of the warp's group.
    int myk = krow[0];
    bool myok = valid[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if (lane == r) {
            myk = krow[r];
            myok = valid[r];
        }
    myok = myok && lane < R && h == 0;
    // (under a register cap -- more than 8 warps -- the operands are loaded after the loops instead: see load_operands below)
    constexpr bool PREFETCH = MAXT <= 256;
    double2 zpr[R];
    RowOps ops;
    auto load_operands = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) zpr[r] = (REALPATH && valid[r]) ? a.g.Zp[krow[r]] : make_double2(0.0, 0.0);
        if (myok) load_row_ops<MODE>(a, (size_t)myk, ops);
    };
    if (PREFETCH) load_operands();

    // ---- the near window, once: cell-local coordinates of the sources in the cells around this CTA's rows ----
    // (all staging loops load a batch of entries per thread into registers before the first store: one L2 round trip per batch
    // instead of one per entry; the loads of the near window fly together with the first batch of the first tile)
    // (under a register cap: smaller batches, and the near window is staged by the plain loop further down)
    constexpr int kV3Batch = MAXT <= 256 ? 10 : 6;        // staged entries per thread in flight
    constexpr int kV3NearBatch = MAXT <= 256 ? 5 : 1;     // near-window entries per thread held across the first tile's staging
    double nxv[kV3NearBatch];
    double2 npv[kV3NearBatch];
#pragma unroll
    for (int u = 0; u < kV3NearBatch; ++u) {
        const int sl = t + u * T;
        const int c = (nc0 + sl / kCell) % a.ncell;
        const int j = c * kCell + (sl & (kCell - 1));
        nxv[u] = 0.0;
        npv[u] = make_double2(1.0e150, 0.0);               // padding contributes exactly 0
        if (sl < nnear * kCell && j < N) {
            nxv[u] = x[j];
            npv[u] = P0[j];
        }
    }
    bool near_pending = true;
    if (solve_done) return;      // (uniform over the launch: nothing has been stored yet)

    double xs = 0.0;
    for (int j0 = 0; j0 < N; j0 += TS) {
        if (j0 > 0) __syncthreads();            // everyone is done with the previous tile
        for (int s0 = t; s0 < TS; s0 += kV3Batch * T) {
            double xv[kV3Batch];
            double2 ev[kV3Batch];
#pragma unroll
            for (int u = 0; u < kV3Batch; ++u) {
                const int s = s0 + u * T;
                const int j = j0 + s;
                xv[u] = 0.0;
                ev[u] = make_double2(1.0e150, 0.0);
                if (s < TS && j < N) {
                    xv[u] = x[j];
                    ev[u] = EG[j];
                }
            }
#pragma unroll
            for (int u = 0; u < kV3Batch; ++u) {
                const int s = s0 + u * T;
                if (s < TS) {
                    xs += xv[u];
                    se[s] = ev[u];
                    sf[s] = make_double2(xv[u] * ev[u].x, xv[u] * ev[u].y);                // (padding: x = 0, E = 1e150 -> exactly 0)
                    sg[s] = xv[u] * (ev[u].x * ev[u].x + ev[u].y * ev[u].y);
                }
            }
        }
        if (near_pending) {   // (the near window's loads have been in flight since before the first tile)
            near_pending = false;
#pragma unroll
            for (int u = 0; u < kV3NearBatch; ++u) {
                const int sl = t + u * T;
                if (sl < nnear * kCell) {
                    ne[sl] = npv[u];
                    nf[sl] = make_double2(nxv[u] * (1.0 + npv[u].x), nxv[u] * npv[u].y);   // (padding: x = 0)
                }
            }
            for (int sl = t + kV3NearBatch * T; sl < nnear * kCell; sl += T) {   // (what the held batch does not cover)
                const int c = (nc0 + sl / kCell) % a.ncell;
                const int j = c * kCell + (sl & (kCell - 1));
                const double xj = j < N ? x[j] : 0.0;
                const double2 pj = j < N ? P0[j] : make_double2(1.0e150, 0.0);
                ne[sl] = pj;
                nf[sl] = make_double2(xj * (1.0 + pj.x), xj * pj.y);
            }
        }
        __syncthreads();
        if (!active) continue;
        const int ncells_tile = min(TS, N - j0 + kCell - 1) / kCell;
        for (int c = h; c < ncells_tile; c += S) {
            const int jj = j0 + c * kCell;
            if (jj >= N) break;
            const int cellJ = jj / kCell;
            int dist = cellJ - cellK;
            if (dist < 0) dist += a.ncell;
            int sd[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = krow[r] - jj - lane;           // source index of the target inside this cell, minus the lane
                sd[r] = (d >= 0 && (d & 31) == 0) ? (d >> 5) : -1;
            }
            const int so = c * kCell + lane;
            if (!a.use_local) {   // few cells: global exponentials and the 10-instruction form everywhere
                if (cellJ == cellK) accumulate3_far<R, true>(se + so, sf + so, sg + so, ekG, sd, U, V);
                else accumulate3_far<R, false>(se + so, sf + so, sg + so, ekG, sd, U, V);
                continue;
            }
            const bool near = dist == 0 || dist == 1 || dist == a.ncell - 1;
            if (!near) {
                accumulate3_far<R, false>(se + so, sf + so, sg + so, ekG, sd, U, V);
                continue;
            }
            const int variant = dist == 0 ? 1 : (dist == 1 ? 2 : 3);
            if (variant != cur_variant) {
                const double2* tk = variant == 1 ? P0 : (variant == 2 ? a.g.Pp : a.g.Pm);
#pragma unroll
                for (int r = 0; r < R; ++r) ek[r] = valid[r] ? tk[krow[r]] : make_double2(3.0e150, 0.0);
                cur_variant = variant;
            }
            int dn = cellJ - nc0;
            if (dn < 0) dn += a.ncell;
            const int no = dn * kCell + lane;
            if (cellJ == cellK) accumulate3_near<R, true>(ne + no, nf + no, ek, sd, acc);
            else accumulate3_near<R, false>(ne + no, nf + no, ek, sd, acc);
        }
    }

    // ---- sum_j x_j (identical in every CTA: same staging pattern, same tree) and the row sums over the lanes ----
    if (!PREFETCH) load_operands();
    const double sumx = block_sum3(xs, sred);
#pragma unroll
    for (int r = 0; r < R; ++r) {   // far part of this lane's sources: T += conj(E_k) U - V
        acc[r].x += fma(ekG[r].x, U[r].x, ekG[r].y * U[r].y) - V[r];
        acc[r].y += fma(ekG[r].x, U[r].y, -(ekG[r].y * U[r].x));
        if (REALPATH) acc[r] = make_double2(zpr[r].x * acc[r].x - zpr[r].y * acc[r].y, 0.0);   // solver sweeps only need Re(Zp T)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, o);
            if (!REALPATH) acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, o);
        }
    }

    // ---- epilogue: lane r of the group's first warp finishes row r ----
    double sr = 0.0;
    {
        double2 mine = acc[0];
#pragma unroll
        for (int r = 1; r < R; ++r)
            if (lane == r) mine = acc[r];
        if (S > 1) {   // the other warps of the group hand their share over through shared memory, added in warp order
            if (h > 0 && lane < R) comb[w * R + lane] = mine;
            __syncthreads();
            if (h == 0 && lane < R)
                for (int q = 1; q < S; ++q) {
                    const double2 v = comb[(w + q) * R + lane];
                    mine.x += v.x;
                    mine.y += v.y;
                }
        }
        if (myok) sr = finish_row<MODE>(a, (size_t)myk, ops, sumx, mine);
    }

    // ---- the last CTA of the launch closes the sweep ----
    const bool solver_sweep = (MODE == kSweepMV && !a.apply_only) || (MODE == kSweepVEL && a.combined);
    const bool need_close = solver_sweep || a.comm.nranks > 1;
    if (!need_close) return;
    if (solver_sweep) {
        sr = block_sum3(sr, sred);
        if (t == 0) a.v2_rnorm_part[blockIdx.x] = sr;
    }
    if (a.comm.nranks > 1) __threadfence_system(); else __threadfence();
    __syncthreads();
    if (t == 0) s_ticket = atomicAdd(a.v2_ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (t == 0) *a.v2_ticket = 0u;
    if (!solver_sweep) {
        if (t == 0) {
            __threadfence_system();
            comm_signal2(a.comm);
        }
        return;
    }
    double rn = 0.0, bn = 0.0;
    int iters_before = 0;
    double prev_rel2 = 0.0;
    if (t == 0) {                 // (the control block's history, in the same round trip as the partials)
        iters_before = *reinterpret_cast<volatile int*>(&a.ctrl->iters);
        prev_rel2 = *reinterpret_cast<volatile double*>(&a.ctrl->prev_rel2);
    }
    for (int i = t; i < (int)gridDim.x; i += T) rn += __ldcg(a.v2_rnorm_part + i);
    if (a.comm.nranks <= 1)
        for (int i = t; i < a.ncell; i += T) bn += __ldcg(a.bnorm_part + i);       // (same round trip as the residual partials)
    rn = block_sum3(rn, sred);
    if (a.comm.nranks > 1) {
        if (t == 0) {
            double* slot = reinterpret_cast<double*>(a.comm.my_base + a.comm.off_rn) + a.out_buf * kMaxRanks + a.comm.rank;
            mirror_store2(a.comm, slot, rn);
            __threadfence_system();
            comm_signal2(a.comm);
        }
        return;
    }
    bn = block_sum3(bn, sred);
    const double worst = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
    if (t == 0) solve_decide(a.ctrl, worst, a.tol2, a.max_iters, a.final_buf_on_done, iters_before, prev_rel2);
}

// ------------------------------------------------------------------------------------------------
// Ensembles of small surfaces (batch > 1, N <= 768: no cell-local coordinates, the 10-instruction form on every cell): the same
// mapping, one MEMBER at a time per CTA.  CTA c steps through the members c, c + G, ...; a member's N sources are staged once
// (the next member's loads fly during this member's pairs: two shared buffers), the 16 warps walk over the member's row groups
// (R rows each), every row is finished by the lane butterfly and the epilogue where it was computed; one residual partial per
// member, the last CTA of the launch takes the decision over the members.  Against sweep2_kernel (lanes = rows, warps = source
// groups): no cross-group combine through shared memory, no idle phase between staging, pairs and epilogue of a row block.
// ------------------------------------------------------------------------------------------------
constexpr int kV3bStage = 4;   // staged entries per thread and member (NP <= 4 T)

__device__ __forceinline__ double block_max3(double v, double* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, W = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sred[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < W; ++i) r = fmax(r, sred[i]);
    return r;
}

template <int MODE, int R>
__global__ void __launch_bounds__(kV3MaxWarps * 32, 1) sweep3b_kernel(const SweepArgs a, const int NP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int LD = NP + kV3Pad;
    double2* se = reinterpret_cast<double2*>(smem_raw);        // [2][LD]  E_j
    double2* sf = se + 2 * LD;                                 // [2][LD]  F_j = x_j E_j
    double* sg = reinterpret_cast<double*>(sf + 2 * LD);       // [2][LD]  g_j = x_j |E_j|^2
    double* sred = sg + 2 * LD;                                // [32]
    __shared__ unsigned int s_ticket;
    constexpr bool REALPATH = (MODE == kSweepMV);

    int solve_done = 0;
    if ((MODE == kSweepMV || MODE == kSweepVEL) && a.skip_if_done) solve_done = *reinterpret_cast<volatile int*>(&a.ctrl->done);
    const int t = threadIdx.x, T = blockDim.x, lane = t & 31, w = t >> 5, W = T >> 5;
    const int N = a.N;
    const int NG = (N + R - 1) / R;                            // row groups per member
    const bool solver_sweep = (MODE == kSweepMV && !a.apply_only) || (MODE == kSweepVEL && a.combined);

    double pxv[kV3bStage];
    double2 pev[kV3bStage];
    auto prefetch = [&](int m) {
        const size_t boff = (size_t)m * N;
#pragma unroll
        for (int u = 0; u < kV3bStage; ++u) {
            const int s = t + u * T;
            pxv[u] = 0.0;
            pev[u] = make_double2(1.0e150, 0.0);               // padding contributes exactly 0
            if (m < a.batch && s < N) {
                pxv[u] = a.x[boff + s];
                pev[u] = a.g.EG[boff + s];
            }
        }
    };
    prefetch(blockIdx.x);
    if (solve_done) return;      // (uniform over the launch: nothing has been stored yet)

    int it = 0;
    for (int m = blockIdx.x; m < a.batch; m += gridDim.x, ++it) {
        const int buf = (it & 1) * LD;
        double xs = 0.0;
#pragma unroll
        for (int u = 0; u < kV3bStage; ++u) {
            const int s = t + u * T;
            if (s < NP) {
                xs += pxv[u];
                se[buf + s] = pev[u];
                sf[buf + s] = make_double2(pxv[u] * pev[u].x, pxv[u] * pev[u].y);
                sg[buf + s] = pxv[u] * (pev[u].x * pev[u].x + pev[u].y * pev[u].y);
            }
        }
        const double sumx = block_sum3(xs, sred);              // (its barriers publish the staged member)
        prefetch(m + gridDim.x);                               // the next member's loads fly during this member's pairs
        const size_t boff = (size_t)m * N;
        double sr = 0.0;
        for (int g = w; g < NG; g += W) {
            int krow[R];
            bool valid[R];
            double2 ekG[R], zpr[R], U[R];
            double V[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                krow[r] = g * R + r;
                valid[r] = krow[r] < N;
                ekG[r] = valid[r] ? se[buf + krow[r]] : make_double2(3.0e150, 0.0);   // (the targets are the staged sources themselves)
                zpr[r] = (REALPATH && valid[r]) ? a.g.Zp[boff + krow[r]] : make_double2(0.0, 0.0);
                U[r] = make_double2(0.0, 0.0);
                V[r] = 0.0;
            }
            int myk = krow[0];
            bool myok = valid[0];
#pragma unroll
            for (int r = 1; r < R; ++r)
                if (lane == r) {
                    myk = krow[r];
                    myok = valid[r];
                }
            myok = myok && lane < R;
            RowOps ops;
            if (myok) load_row_ops<MODE>(a, boff + myk, ops);
            const int cellK = krow[0] / kCell;
            for (int c = 0; c < a.ncell; ++c) {
                const int jj = c * kCell;
                int sd[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int d = krow[r] - jj - lane;
                    sd[r] = (d >= 0 && (d & 31) == 0) ? (d >> 5) : -1;
                }
                const int so = buf + jj + lane;
                if (c == cellK) accumulate3_far<R, true>(se + so, sf + so, sg + so, ekG, sd, U, V);
                else accumulate3_far<R, false>(se + so, sf + so, sg + so, ekG, sd, U, V);
            }
            double2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {   // T = conj(E_k) U - V
                acc[r].x = fma(ekG[r].x, U[r].x, ekG[r].y * U[r].y) - V[r];
                acc[r].y = fma(ekG[r].x, U[r].y, -(ekG[r].y * U[r].x));
                if (REALPATH) acc[r] = make_double2(zpr[r].x * acc[r].x - zpr[r].y * acc[r].y, 0.0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, o);
                    if (!REALPATH) acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, o);
                }
            }
            double2 mine = acc[0];
#pragma unroll
            for (int r = 1; r < R; ++r)
                if (lane == r) mine = acc[r];
            if (myok) sr += finish_row<MODE>(a, boff + myk, ops, sumx, mine);
        }
        if (solver_sweep) {
            sr = block_sum3(sr, sred);
            if (t == 0) a.v2_rnorm_part[m] = sr;
        }
    }

    // ---- the last CTA of the launch closes the sweep: max over the members of ||r||^2 / ||b||^2 ----
    if (!solver_sweep) return;
    __threadfence();
    __syncthreads();
    if (t == 0) s_ticket = atomicAdd(a.v2_ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (t == 0) *a.v2_ticket = 0u;
    int iters_before = 0;
    double prev_rel2 = 0.0;
    if (t == 0) {
        iters_before = *reinterpret_cast<volatile int*>(&a.ctrl->iters);
        prev_rel2 = *reinterpret_cast<volatile double*>(&a.ctrl->prev_rel2);
    }
    double wm = 0.0;
    for (int m = t; m < a.batch; m += T) {
        const double rn = __ldcg(a.v2_rnorm_part + m);
        double bn = 0.0;
        for (int c = 0; c < a.ncell; ++c) bn += __ldcg(a.bnorm_part + (size_t)m * a.ncell + c);
        double rel2 = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
        if (!(rel2 == rel2)) rel2 = 1e300;
        wm = fmax(wm, rel2);
    }
    const double worst = block_max3(wm, sred);
    if (t == 0) solve_decide(a.ctrl, worst, a.tol2, a.max_iters, a.final_buf_on_done, iters_before, prev_rel2);
}

template <int MODE, int R>
static void launch_one3b(const SweepArgs& a, const Sweep3Launch& l, cudaStream_t st) {
    static size_t configured = 0;
    if (l.smem > configured) {
        RB_CUDA(cudaFuncSetAttribute(sweep3b_kernel<MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
        configured = l.smem;
    }
    sweep3b_kernel<MODE, R><<<l.grid, l.threads, l.smem, st>>>(a, l.TS);
}

template <int R>
static void launch_mode3b(const SweepArgs& a, const Sweep3Launch& l, int mode, cudaStream_t st) {
    if (mode == kSweepMV) launch_one3b<kSweepMV, R>(a, l, st);
    else if (mode == kSweepVEL) launch_one3b<kSweepVEL, R>(a, l, st);
    else launch_one3b<kSweepRAW, R>(a, l, st);
}

size_t sweep3b_smem(int NP) { return (size_t)2 * (NP + kV3Pad) * (16 + 16 + 8) + 32 * sizeof(double); }

template <int MODE, int R, int MAXT>
static void launch_one3(const SweepArgs& a, const Sweep3Launch& l, cudaStream_t st) {
    static size_t configured = 0;
    if (l.smem > configured) {
        RB_CUDA(cudaFuncSetAttribute(sweep3_kernel<MODE, R, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
        configured = l.smem;
    }
    sweep3_kernel<MODE, R, MAXT><<<l.grid, l.threads, l.smem, st>>>(a, l.TS, l.S);
}

// (register budget by the launch bound: 8 warps 255, 14 warps 144, 16 warps 128; 4 rows per warp want ~180)
template <int R>
static void launch_mode3(const SweepArgs& a, const Sweep3Launch& l, int mode, cudaStream_t st) {
    if (l.threads <= 256) {
        if (mode == kSweepMV) launch_one3<kSweepMV, R, 256>(a, l, st);
        else if (mode == kSweepVEL) launch_one3<kSweepVEL, R, 256>(a, l, st);
        else launch_one3<kSweepRAW, R, 256>(a, l, st);
    } else if (l.threads <= 448) {
        if (mode == kSweepMV) launch_one3<kSweepMV, R, 448>(a, l, st);
        else if (mode == kSweepVEL) launch_one3<kSweepVEL, R, 448>(a, l, st);
        else launch_one3<kSweepRAW, R, 448>(a, l, st);
    } else {
        if (mode == kSweepMV) launch_one3<kSweepMV, R, kV3MaxWarps * 32>(a, l, st);
        else if (mode == kSweepVEL) launch_one3<kSweepVEL, R, kV3MaxWarps * 32>(a, l, st);
        else launch_one3<kSweepRAW, R, kV3MaxWarps * 32>(a, l, st);
    }
}

size_t sweep3_smem(int TS) {
    return (size_t)(TS + kV3Pad) * (16 + 16 + 8) + (size_t)(kV3Near + kV3Pad) * 32 + 32 * sizeof(double) + (size_t)kV3MaxWarps * 4 * 16;
}

void launch_sweep3(const SweepArgs& a, const Sweep3Launch& l, int mode, cudaStream_t st) {
    if (l.batched) {
        if (a.has_image || a.use_local || a.comm.nranks > 1) throw std::runtime_error("sweep3 (ensembles): no image sum, N <= 768, one rank");
        if (l.R == 4) launch_mode3b<4>(a, l, mode, st);
        else if (l.R == 2) launch_mode3b<2>(a, l, mode, st);
        else launch_mode3b<1>(a, l, mode, st);
        RB_CUDA(cudaGetLastError());
        count_launch();
        return;
    }
    if (a.has_image || a.batch != 1) throw std::runtime_error("sweep3: one member, no image sum");
    if (l.R == 4) launch_mode3<4>(a, l, mode, st);
    else if (l.R == 2) launch_mode3<2>(a, l, mode, st);
    else launch_mode3<1>(a, l, mode, st);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
