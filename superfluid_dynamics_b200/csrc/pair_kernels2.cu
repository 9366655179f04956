// pair_kernels2.cu -- persistent one-wave variant of the O(N^2) cotangent-sum sweep (same mathematics as pair_kernels.cu).
//
// Schedule: the rows of this rank are cut into row blocks of RB rows (RB a multiple of 32*R, so a warp never straddles a
// 256-point cell); CTA c owns the contiguous range [c*TB/G, (c+1)*TB/G) of the TB row blocks (G = grid size ~ number of SMs), and
// for each of its blocks walks over ALL sources of that batch member.  Threads = (RB/R) row-threads x `groups` source groups:
// every staged tile of TS = groups*spg sources is split between the groups, whose accumulators are combined through shared
// memory at the end of the block.  Consequences: no partial sums in global memory, no per-row-cell tickets, sum_j x_j is obtained
// for free while staging, the result is deterministic, and the work per SM is balanced to 1/(blocks per CTA).
// The next tile is prefetched into registers while the current one is being evaluated.
#include "internal.cuh"
#include "pair_common.cuh"

namespace rb {

namespace {

template <bool DIAG, int R>
__device__ __forceinline__ void accumulate2(const Src2* __restrict__ sh, int len, const double2 (&ek)[R], const int (&sd)[R],
                                            double2 (&acc)[R]) {
#pragma unroll 4
    for (int s = 0; s < len; ++s) {
        const double2 e = *reinterpret_cast<const double2*>(&sh[s].p);
        const double2 f = *reinterpret_cast<const double2*>(&sh[s].fr);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double dr = ek[r].x - e.x;
            double di = ek[r].y - e.y;
            double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp2(n2);
            if (DIAG) inv = (s == sd[r]) ? 0.0 : inv;
            double tr = fma(f.y, di, f.x * dr);
            double ti = fma(f.y, dr, -(f.x * di));
            acc[r].x = fma(tr, inv, acc[r].x);
            acc[r].y = fma(ti, inv, acc[r].y);
        }
    }
}

// Far tiles (no cell-local coordinates, no j == k), every mode: with conj(d) = conj(E_k) - conj(E_j),
//   T_k = sum_j F_j conj(d)/|d|^2 = conj(E_k) U_k - V_k ,   U_k = sum_j F_j / |d|^2 (complex),  V_k = sum_j g_j / |d|^2 (real),
//   g_j = x_j |E_j|^2, so the pair costs 2 DADD + DMUL + DFMA + (MUFU + 3 DFMA) + 3 DFMA = 10 FP64-pipe instructions for the 20
//   algorithmic flops, and the target-side product is applied once per row after the loop.  The two sums cancel by at most
//   |E|/|E_k - E_j| <= 1/(cell width) ~ 40 (far tiles only), i.e. ~1e-15 relative in the row sum.
template <int R>
__device__ __forceinline__ void accumulate2_far(const Src2* __restrict__ sh, const double* __restrict__ sg, int len,
                                                const double2 (&ek)[R], double2 (&U)[R], double (&V)[R]) {
    // software pipeline: the shared-memory loads of source s + 1 are issued before the arithmetic of source s (ncu: with the
    // loads placed next to their first use, short-scoreboard stalls on LDS were the top stall reason of this loop)
    double2 e = *reinterpret_cast<const double2*>(&sh[0].p);
    double2 f = *reinterpret_cast<const double2*>(&sh[0].fr);
    double gj = sg[0];
#pragma unroll 4
    for (int s = 0; s < len; ++s) {
        // (entry `len` is the next piece's first entry or the padding behind the staged tile: loaded, never used)
        const double2 en = *reinterpret_cast<const double2*>(&sh[s + 1].p);
        const double2 fn = *reinterpret_cast<const double2*>(&sh[s + 1].fr);
        const double gn = sg[s + 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double dr = ek[r].x - e.x;
            double di = ek[r].y - e.y;
            double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp2(n2);
            U[r].x = fma(f.x, inv, U[r].x);
            U[r].y = fma(f.y, inv, U[r].y);
            V[r] = fma(gj, inv, V[r]);
        }
        e = en;
        f = fn;
        gj = gn;
    }
}

// The same 10-instruction form for a piece that contains the targets themselves (j == k masked out by a select on the integer
// pipe): used when the surface has too few cells for cell-local coordinates (N <= 768, every tile is "near").  There the global
// exponentials are accurate to eps N / 2 pi <= 1.2e-14 per entry anyway, and the cancellation of conj(E_k) U - V is bounded by the
// same N / 2 pi <= 122.
template <int R>
__device__ __forceinline__ void accumulate2_far_diag(const Src2* __restrict__ sh, const double* __restrict__ sg, int len,
                                                     const double2 (&ek)[R], const int (&sd)[R], double2 (&U)[R], double (&V)[R]) {
    double2 e = *reinterpret_cast<const double2*>(&sh[0].p);
    double2 f = *reinterpret_cast<const double2*>(&sh[0].fr);
    double gj = sg[0];
#pragma unroll 4
    for (int s = 0; s < len; ++s) {
        const double2 en = *reinterpret_cast<const double2*>(&sh[s + 1].p);
        const double2 fn = *reinterpret_cast<const double2*>(&sh[s + 1].fr);
        const double gn = sg[s + 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double dr = ek[r].x - e.x;
            double di = ek[r].y - e.y;
            double n2 = fma(di, di, dr * dr);
            double inv = fast_rcp2(n2);
            inv = (s == sd[r]) ? 0.0 : inv;
            U[r].x = fma(f.x, inv, U[r].x);
            U[r].y = fma(f.y, inv, U[r].y);
            V[r] = fma(gj, inv, V[r]);
        }
        e = en;
        f = fn;
        gj = gn;
    }
}

__device__ __forceinline__ bool cells_near(int cT0, int cT1, int cB0, int cB1, int ncell) {
    for (int cT = cT0; cT <= cT1; ++cT)
        for (int cB = cB0; cB <= cB1; ++cB) {
            int d = cT - cB;
            if (d < 0) d += ncell;
            if (d <= 1 || d == ncell - 1) return true;
        }
    return false;
}

constexpr int kMaxStage = 1;   // staged entries per thread and tile (TS <= threads)

}  // namespace

// MAXT: launch bound (R = 2: 896 threads = 4 source groups at 72 registers; measured at N = 65536: 3 groups / 672 threads / 80
// registers 3.56 ms against 3.11 ms -- resident warps matter more than registers here)
template <int MODE, int R, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? 2 : 1) sweep2_kernel(const SweepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int TS = a.v2_TS;
    Src2* sh_far = reinterpret_cast<Src2*>(smem_raw);
    Src2* sh_near = sh_far + TS;
    double2* red = reinterpret_cast<double2*>(sh_far + (a.use_local ? 2 : 1) * TS);   // [T * R]
    double* sred = reinterpret_cast<double*>(red + (size_t)T * R);                    // [T]
    double* sh_g = sred + T;                                                          // [TS]  g_j = x_j |E_j|^2
    constexpr bool REALPATH = (MODE == kSweepMV);
    __shared__ unsigned int s_ticket;

    if ((MODE == kSweepMV || MODE == kSweepVEL) && a.skip_if_done) {
        if (*reinterpret_cast<volatile int*>(&a.ctrl->done)) return;
    }
    int P2 = 1;
    while (P2 < T) P2 <<= 1;
    const int t = threadIdx.x;
    const int N = a.N;
    const int nrt = a.v2_RB / R;
    const int rt = t % nrt;
    const int g = t / nrt;
    const int SPG = a.v2_spg;
    const int sublen = SPG < kCell ? SPG : kCell;
    const double inv4pi = 0.25 / kPi;
    const int TB = a.v2_total_blocks;
    const int SX = a.v2_split;                       // source parts per row block (1: no split)
    const int TI = TB * SX;                          // work items
    const int NT = (N + TS - 1) / TS;                // staged tiles per member
    const int i0 = (int)(((long long)blockIdx.x * TI) / gridDim.x);
    const int i1 = (int)(((long long)(blockIdx.x + 1) * TI) / gridDim.x);
    __shared__ unsigned int s_part_ticket;

    for (int item = i0; item < i1; ++item) {
        const int blk = item / SX;
        const int part = item - blk * SX;
        const int tile_begin = (int)(((long long)part * NT) / SX);
        const int tile_end = (int)(((long long)(part + 1) * NT) / SX);
        const int bm = blk / a.v2_bpm;
        const int ib = blk - bm * a.v2_bpm;
        const int row0 = a.v2_row_begin + ib * a.v2_RB;
        const int rend = min(row0 + a.v2_RB, a.v2_row_end);
        const size_t boff = (size_t)bm * N;
        const double* __restrict__ x = a.x + boff;
        const double2* __restrict__ EG = a.g.EG + boff;
        const double2* __restrict__ P0 = a.g.P0 + boff;
        int krow[R];
        bool valid[R];
        double2 acc[R], ek[R], ekG[R], U[R];
        double V[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            krow[r] = row0 + rt * R + r;
            valid[r] = krow[r] < rend;
            acc[r] = make_double2(0.0, 0.0);
            U[r] = make_double2(0.0, 0.0);
            V[r] = 0.0;
            ek[r] = make_double2(3.0e150, 0.0);
            ekG[r] = valid[r] ? EG[krow[r]] : make_double2(3.0e150, 0.0);
        }
        const int cellK = min(krow[0], N - 1) / kCell;
        const int cB0 = row0 / kCell, cB1 = (rend - 1) / kCell;
        int cur_variant = -1;
        double xs = 0.0;

        // ---- software pipeline: registers hold the staged entries of the next tile ------------------------------------------
        double2 pe[kMaxStage], pp[kMaxStage];
        double px[kMaxStage];
        bool pnear = false;
        auto prefetch = [&](int j0) {
            pnear = false;
            if (a.use_local) {
                int cT0 = j0 / kCell, cT1 = (min(j0 + TS, N) - 1) / kCell;
                pnear = cells_near(cT0, cT1, cB0, cB1, a.ncell);
            }
#pragma unroll
            for (int u = 0; u < kMaxStage; ++u) {
                int s = t + u * T;
                int j = j0 + s;
                if (s < TS && j < N) {
                    px[u] = x[j];
                    pe[u] = EG[j];
                    if (pnear) pp[u] = P0[j];
                }
            }
        };
        auto commit = [&](int j0) {
#pragma unroll
            for (int u = 0; u < kMaxStage; ++u) {
                int s = t + u * T;
                int j = j0 + s;
                if (s < TS) {
                    Src2 e;
                    if (j < N) {
                        double xj = px[u];
                        xs += xj;
                        e.p = pe[u].x; e.q = pe[u].y; e.fr = xj * pe[u].x; e.fi = xj * pe[u].y;
                        sh_far[s] = e;
                        sh_g[s] = xj * (pe[u].x * pe[u].x + pe[u].y * pe[u].y);
                        if (pnear) {
                            Src2 n;
                            n.p = pp[u].x; n.q = pp[u].y; n.fr = xj * (1.0 + pp[u].x); n.fi = xj * pp[u].y;
                            sh_near[s] = n;
                        }
                    } else {
                        e.p = 1.0e150; e.q = 0.0; e.fr = 0.0; e.fi = 0.0;   // contributes exactly 0
                        sh_far[s] = e;
                        sh_g[s] = 0.0;
                        if (pnear) sh_near[s] = e;
                    }
                }
            }
        };

        prefetch(tile_begin * TS);
        for (int tl = tile_begin; tl < tile_end; ++tl) {
            const int j0 = tl * TS;
            __syncthreads();            // everyone is done with the previous tile
            const bool tile_near = pnear;
            commit(j0);
            __syncthreads();
            if (tl + 1 < tile_end) prefetch(j0 + TS);   // loads fly while this tile is evaluated
            // ---- this group's share of the tile, in pieces that stay inside one 256-point cell --------------------------------
            for (int sub0 = g * SPG; sub0 < (g + 1) * SPG; sub0 += sublen) {
                const int jj = j0 + sub0;
                if (jj >= N) break;
                const int cellJ = jj / kCell;
                int dist = cellJ - cellK;
                if (dist < 0) dist += a.ncell;
                const bool near = a.use_local && tile_near && (dist == 0 || dist == 1 || dist == a.ncell - 1);
                int sd[R];
#pragma unroll
                for (int r = 0; r < R; ++r) sd[r] = krow[r] - jj;
                if (!a.use_local) {   // few cells: global exponentials everywhere, 10-instruction form everywhere
                    if (cellJ == cellK) accumulate2_far_diag<R>(sh_far + sub0, sh_g + sub0, sublen, ekG, sd, U, V);
                    else accumulate2_far<R>(sh_far + sub0, sh_g + sub0, sublen, ekG, U, V);
                    continue;
                }
                const int variant = near ? (dist == 0 ? 1 : (dist == 1 ? 2 : 3)) : 0;
                if (variant != cur_variant) {
                    const double2* tk = variant == 0 ? EG : (variant == 1 ? P0 : (variant == 2 ? a.g.Pp + boff : a.g.Pm + boff));
#pragma unroll
                    for (int r = 0; r < R; ++r) ek[r] = valid[r] ? tk[krow[r]] : make_double2(3.0e150, 0.0);
                    cur_variant = variant;
                }
                const Src2* src = (near ? sh_near : sh_far) + sub0;
                if (cellJ == cellK) accumulate2<true, R>(src, sublen, ek, sd, acc);
                else if (a.use_local && !near) accumulate2_far<R>(src, sh_g + sub0, sublen, ekG, U, V);
                else                accumulate2<false, R>(src, sublen, ek, sd, acc);
            }
        }

        // ---- sum_j x_j (identical in every CTA: same staging pattern, same tree) and the cross-group combine ------------------
        double sumx = block_sum_any(xs, sred, T, P2);
#pragma unroll
        for (int r = 0; r < R; ++r) {   // far part of this group's sources: T += conj(E_k) U - V
            if (!valid[r]) continue;
            acc[r].x += fma(ekG[r].x, U[r].x, ekG[r].y * U[r].y) - V[r];
            acc[r].y += fma(ekG[r].x, U[r].y, -(ekG[r].y * U[r].x));
        }
        if (REALPATH) {
#pragma unroll
            for (int r = 0; r < R; ++r) {   // solver sweeps only need Re(Zp T)
                const double2 zp = valid[r] ? a.g.Zp[boff + krow[r]] : make_double2(0.0, 0.0);
                acc[r] = make_double2(zp.x * acc[r].x - zp.y * acc[r].y, 0.0);
            }
        }
        if (a.v2_groups > 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) red[(size_t)g * a.v2_RB + rt * R + r] = acc[r];
            __syncthreads();
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double2 s = acc[r];
                    for (int gg = 1; gg < a.v2_groups; ++gg) {
                        double2 v = red[(size_t)gg * a.v2_RB + rt * R + r];
                        s.x += v.x; s.y += v.y;
                    }
                    acc[r] = s;
                }
            }
        }

        // ---- source split: publish this part's row sums; the last part of the row block adds them up in part order -----------------
        if (SX > 1) {
            const size_t rows_total = (size_t)(a.v2_row_end - a.v2_row_begin);
            double2* pbase = a.v2_partial + ((size_t)bm * SX) * rows_total;
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (valid[r]) pbase[(size_t)part * rows_total + (krow[r] - a.v2_row_begin)] = acc[r];
            }
            if (t == 0) a.v2_xs_part[item] = sumx;
            __threadfence();
            __syncthreads();
            if (t == 0) s_part_ticket = atomicAdd(a.v2_blk_tickets + blk, 1u);
            __syncthreads();
            if (s_part_ticket != (unsigned)(SX - 1)) continue;   // another CTA finishes this row block
            __threadfence();
            if (t == 0) a.v2_blk_tickets[blk] = 0u;
            sumx = 0.0;
            for (int p = 0; p < SX; ++p) sumx += __ldcg(a.v2_xs_part + (size_t)blk * SX + p);
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double2 sacc = make_double2(0.0, 0.0);
                    if (valid[r]) {
                        for (int p = 0; p < SX; ++p) {
                            double2 v = __ldcg(pbase + (size_t)p * rows_total + (krow[r] - a.v2_row_begin));
                            sacc.x += v.x; sacc.y += v.y;
                        }
                    }
                    acc[r] = sacc;
                }
            }
        }

        // ---- epilogue (group 0 holds the row sums) --------------------------------------------------------------------------------
        double sr = 0.0;
        if (g == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (!valid[r]) continue;
                const size_t o = boff + krow[r];
                const double xk = x[krow[r]];
                const double Ar = (sumx - xk) + 2.0 * acc[r].x;
                const double Ai = 2.0 * acc[r].y;
                if (MODE == kSweepMV) {
                    double2 zp = a.g.Zp[o];
                    double Mx = fma(a.g.Mdiag[o], xk, a.cK * fma(zp.x, sumx - xk, 2.0 * acc[r].x));   // acc.x = Re(Zp T)
                    if (a.apply_only) {
                        mirror_store2(a.comm, a.x_out + o, Mx);
                    } else {
                        double res = a.g.b[o] - Mx;
                        mirror_store2(a.comm, a.x_out + o, fma(a.omega, res, xk));
                        sr += res * res;
                    }
                } else if (MODE == kSweepVEL) {
                    double2 zp = a.g.Zp[o];
                    double2 v1d = a.g.V1diag[o];
                    double2 v2 = a.g.V2[o];
                    const double2 ap = a.defer_aprime ? make_double2(0.0, 0.0) : a.aprime[o];   // deferred: finish_solve adds V2 a'
                    double wr = inv4pi * Ar + v1d.x * xk + (v2.x * ap.x - v2.y * ap.y);
                    double wi = inv4pi * Ai + v1d.y * xk + (v2.x * ap.y + v2.y * ap.x);
                    mirror_store2(a.comm, a.vel_lower + o, make_double2(wr, -wi));
                    double inv = 1.0 / (zp.x * zp.x + zp.y * zp.y);
                    double azx = xk * zp.x * inv, azy = -xk * zp.y * inv;     // a_k / Zp_k
                    a.vel_upper[o] = make_double2(wr - azx, -(wi - azy));
                    if (a.dphi && !a.defer_aprime) {
                        double y = a.g.Z[o].y;
                        double d;
                        if (a.rhs_phi_kind == 1) {
                            d = -y + 0.5 * (wr * wr + wi * wi);
                        } else {
                            double vdw = a.depth / 3.0;
                            d = vdw * pow(1.0 + y / a.depth, -3.0) - vdw + (0.5 * wr * wr + 0.5 * wi * wi);
                        }
                        mirror_store2(a.comm, a.dphi + o, make_double2(d, 0.0));
                    }
                    if (a.combined) {   // verify the iterate with the same row sum: r = b - M a; next iterate in case it is needed
                        if (a.A_out) mirror_store2(a.comm, a.A_out + o, make_double2(Ar, Ai));
                        double res = a.g.b[o] - fma(a.g.Mdiag[o], xk, a.cK * (zp.x * Ar - zp.y * Ai));
                        mirror_store2(a.comm, a.x_out + o, fma(a.omega, res, xk));
                        sr += res * res;
                    }
                } else {
                    a.raw_out[o] = make_double2(-Ai, Ar);
                }
            }
        }
        if ((MODE == kSweepMV && !a.apply_only) || (MODE == kSweepVEL && a.combined)) {
            sr = block_sum_any(sr, sred, T, P2);
            if (t == 0) a.v2_rnorm_part[blk] = sr;
        }
    }

    // ---- end of this CTA's schedule: the last CTA of the launch closes the sweep ------------------------------------------------
    const bool solver_sweep = (MODE == kSweepMV && !a.apply_only) || (MODE == kSweepVEL && a.combined);
    const bool need_close = solver_sweep || a.comm.nranks > 1;
    if (!need_close) return;
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
    // residual: max over batch members of ||r||^2 / ||b||^2
    double worst;
    if (a.batch == 1) {
        double rn = 0.0, bn = 0.0;
        for (int i = t; i < TB; i += T) rn += __ldcg(a.v2_rnorm_part + i);
        rn = block_sum_any(rn, sred, T, P2);
        if (a.comm.nranks > 1) {
            if (t == 0) {
                double* slot = reinterpret_cast<double*>(a.comm.my_base + a.comm.off_rn) + a.out_buf * kMaxRanks + a.comm.rank;
                mirror_store2(a.comm, slot, rn);
                __threadfence_system();
                comm_signal2(a.comm);
            }
            return;
        }
        for (int i = t; i < a.ncell; i += T) bn += __ldcg(a.bnorm_part + i);
        bn = block_sum_any(bn, sred, T, P2);
        worst = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
    } else {
        double w = 0.0;
        for (int m = t; m < a.batch; m += T) {
            double rn = 0.0, bn = 0.0;
            for (int i = 0; i < a.v2_bpm; ++i) rn += __ldcg(a.v2_rnorm_part + (size_t)m * a.v2_bpm + i);
            for (int c = 0; c < a.ncell; ++c) bn += __ldcg(a.bnorm_part + (size_t)m * a.ncell + c);
            double rel2 = bn > 0.0 ? rn / bn : (rn == 0.0 ? 0.0 : 1e300);
            if (!(rel2 == rel2)) rel2 = 1e300;
            w = fmax(w, rel2);
        }
        worst = block_max_any(w, sred, T, P2);
    }
    if (t == 0) {
        solve_decide(a.ctrl, worst, a.tol2, a.max_iters, a.final_buf_on_done);
        __threadfence();
    }
}

template <int MODE, int R, int MAXT>
static void launch_one(const SweepArgs& a, const Sweep2Launch& l, cudaStream_t st) {
    static size_t configured = 0;
    if (l.smem > configured) {
        RB_CUDA(cudaFuncSetAttribute(sweep2_kernel<MODE, R, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
        configured = l.smem;
    }
    sweep2_kernel<MODE, R, MAXT><<<l.grid, l.threads, l.smem, st>>>(a);
}

void launch_sweep2(const SweepArgs& a, const Sweep2Launch& l, int mode, cudaStream_t st) {
    if (a.has_image) throw std::runtime_error("sweep2: the image (finite-depth) sum uses the tiled kernel");
    if (a.v2_R == 4 && l.threads <= 256) {   // ensembles: 4 rows per thread, two CTAs of 256 threads per SM (one stages / closes
        if (mode == kSweepMV) launch_one<kSweepMV, 4, 256>(a, l, st);          // its row block while the other evaluates pairs)
        else if (mode == kSweepVEL) launch_one<kSweepVEL, 4, 256>(a, l, st);
        else launch_one<kSweepRAW, 4, 256>(a, l, st);
    } else if (a.v2_R == 4) {
        if (mode == kSweepMV) launch_one<kSweepMV, 4, 512>(a, l, st);
        else if (mode == kSweepVEL) launch_one<kSweepVEL, 4, 512>(a, l, st);
        else launch_one<kSweepRAW, 4, 512>(a, l, st);
    } else if (a.v2_R == 2) {
        if (mode == kSweepMV) launch_one<kSweepMV, 2, 896>(a, l, st);
        else if (mode == kSweepVEL) launch_one<kSweepVEL, 2, 896>(a, l, st);
        else launch_one<kSweepRAW, 2, 896>(a, l, st);
    } else {
        if (mode == kSweepMV) launch_one<kSweepMV, 1, 1024>(a, l, st);
        else if (mode == kSweepVEL) launch_one<kSweepVEL, 1, 1024>(a, l, st);
        else launch_one<kSweepRAW, 1, 1024>(a, l, st);
    }
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
