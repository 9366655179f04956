// drive.cu -- the optomechanically driven film: the autonomous augmented system [Z | Phi | D] and the
// explicitly time-dependent form of the same drive, their RK4 steppers and legacy exports.
#include "host.cuh"

// ------------------------------------------------------------------------------------------------
// optomechanically driven film: the autonomous augmented system y = [Z | Phi | D] (drive_kernels.cu) and its classical RK4 stepper
// (AugmentedBoundaryIntegrator + AutonomousRungeKuttaStepper<std_complex, 3N>, A/kernel.cu:85-96, L/Export.cu:980-1209)
// ------------------------------------------------------------------------------------------------

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
    device_ptr<double2> d_owner = dmalloc_scoped<double2>(6 * N);
    double2* d = d_owner.get();
    RB_CUDA(cudaMemcpyAsync(d, host.data(), 3 * N * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
    aug_rhs(s.get(), v, d, d + 3 * N);
    RB_CUDA(cudaMemcpyAsync(host.data(), d + 3 * N, 3 * N * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
    RB_CUDA(cudaStreamSynchronize(s->stream));
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
