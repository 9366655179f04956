// rk45.cu -- adaptive Runge-Kutta-Fehlberg 4(5) stepper (L/RK45.cuh, L/RK45_Kernels.cuh) over the
// boundary-integral RHS or a caller-supplied AutonomousProblem::run callback.
#include "host.cuh"

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
