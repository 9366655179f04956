// stepper.cu -- the classical RK4 stepper over the boundary-integral RHS: one recorded CUDA graph per step with rollback, adaptive
// number of recorded sweeps, optional asynchronous chunks; and its C ABI (rb_rk4_*).
//
// Mirrors (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   AutonomousRungeKuttaStepperBase::runStep / initialize / runEvolution       L/AutonomousRungeKuttaStepper.cuh:124-437
#include "host.cuh"

void stepper_free(rb_stepper* st) {
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

void stepper_step(rb_stepper* st) {
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
            // ... and drop the last surplus round once the count has been the same for 8 steps in a row
            if (st->tight_ok && !st->tight && st->tight_ban == 0 && st->graph_sweeps - worst >= 1 && worst >= (s->use_gmres ? 3 : 2) &&
                s->props.guess_mode == RB_GUESS_WARM) {
                st->tight_max = st->tight_hits == 0 ? worst : std::max(st->tight_max, worst);   // the most any of these steps needed
                if (++st->tight_hits >= 8) {
                    if (st->tight_max < st->graph_sweeps) {
                        st->graph_sweeps = st->tight_max;
                        st->tight = true;
                        invalidate_graphs(st);
                    }
                    st->tight_hits = 0;
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
void stepper_run(rb_stepper* st, size_t n) {
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

extern "C" {

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

}  // extern "C"
