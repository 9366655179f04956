// exports.cu -- the legacy export names of the reference (L/Export.cuh:27-79) with identical argument lists: SI in, nondimensionalised
// inside (adimensionalizeProperties, L/Export.cu:1213-1246), host vectors packed as loadDataToDevice does (L/SimulationRunner.cuh:180-242).
#include "host.cuh"

// ---- nondimensionalisation ---------------------------------------------------------------------


// adimensionalizeProperties, L/Export.cu:1222-1246 (the stdout prints of the reference are dropped)
Adim adimensionalize(double L, double rho, double kappa, double depth, double rhoHelium) {
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

rb_props helium_props(const Adim& ad, bool use_expansions, int expansion_order, bool infinite_depth) {
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

extern "C" {

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
    device_ptr<double2> dstate_owner = dmalloc_scoped<double2>(4 * BN);
    double2* dstate = dstate_owner.get();
    double2* drhs = dstate + 2 * BN;
    RB_CUDA(cudaMemcpy(dstate, host.data(), 2 * BN * sizeof(double2), cudaMemcpyHostToDevice));
    rhs(s.get(), dstate, drhs);
    RB_CUDA(cudaMemcpy(host.data(), drhs, 2 * BN * sizeof(double2), cudaMemcpyDeviceToHost));
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
    device_ptr<double2> dstate_owner = dmalloc_scoped<double2>(2 * N);
    double2* dstate = dstate_owner.get();
    RB_CUDA(cudaMemcpy(dstate, Z, N * sizeof(double2), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(dstate + N, phi, N * sizeof(double2), cudaMemcpyHostToDevice));
    vorticities(s.get(), dstate);
    RB_CUDA(cudaMemcpy(a, s->a, N * sizeof(double), cudaMemcpyDeviceToHost));
    if (Zp) RB_CUDA(cudaMemcpy(Zp, s->Zp(), N * sizeof(double2), cudaMemcpyDeviceToHost));
    if (Zpp) RB_CUDA(cudaMemcpy(Zpp, s->Zpp(), N * sizeof(double2), cudaMemcpyDeviceToHost));
    RB_CATCH
}

int calculateDerivativeFFT256(const rb_complex* input, rb_complex* output) {
    RB_TRY
    const size_t N = 256;
    rb_props p;
    rb_default_props(&p);
    std::unique_ptr<rb_solver, void (*)(rb_solver*)> s(solver_create((int)N, 1, &p), solver_free);
    device_ptr<double2> d_owner = dmalloc_scoped<double2>(2 * N);
    double2* d = d_owner.get();
    RB_CUDA(cudaMemcpy(d, input, N * sizeof(double2), cudaMemcpyHostToDevice));
    fft_derivative(s.get(), d, d + N, 0, 1.0);   // L/Export.cu: FftDerivative<256,1>::exec(in, out)
    RB_CUDA(cudaMemcpy(output, d + N, N * sizeof(double2), cudaMemcpyDeviceToHost));
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

// service for implicit.cu (internal.cuh)
namespace rb {
rb_props helium_props_from_si(double L, double rho, double kappa, double depth, bool use_expansions, int expansion_order,
                              bool infinite_depth) {
    return helium_props(adimensionalize(L, rho, kappa, depth), use_expansions, expansion_order, infinite_depth);
}
}  // namespace rb
