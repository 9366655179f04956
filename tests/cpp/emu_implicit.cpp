// emu_implicit.cpp -- TEST INFRASTRUCTURE: the host logic of csrc/implicit.cu (finite-difference Jacobian, Gauss-Legendre Newton
// iteration, trajectory logging, legacy exports) compiled by g++ against tests/cpp/cuda_emu.h, with a MOCK of the RHS assembler behind
// it: rb_rhs forwards to a callback the Python test installs (the oracle's batched RHS), the dense solve is a textbook host LU.
// What runs unchanged is everything implicit.cu itself does: which kernels it launches with which grids and buffers, in which
// order, what it decides from the residual norms, what it logs and returns.  Used by tests/test_kernel_emulation.py only.
#include "cuda_emu.h"

#include "../../include/roberts_b200.h"

namespace rb {
int report_error(const std::exception& e);
rb_props helium_props_from_si(double L, double rho, double kappa, double depth, bool use_expansions, int expansion_order, bool infinite_depth);
void launch_lu_solve(double* A, double* b, int n, int* info, cudaStream_t st);
void launch_lu_solve_blocked(double* A, double* b, int n, int* info, cudaStream_t st);
void launch_lu_solve_unblocked(double* A, double* b, int n, int* info, cudaStream_t st);
}  // namespace rb

#include "../../superfluid_dynamics_b200/csrc/implicit.cu"

// ---- the mock behind the C ABI implicit.cu is written against ---------------------------------------------------------------------
struct rb_solver {
    int N, batch;
    rb_props props;
    long long rhs_calls;
};

typedef void (*emu_rhs_fn)(const rb_props* props, int N, int batch, const double* state, double* out);
typedef void (*emu_adim_fn)(double L, double rho, double kappa, double depth, double* out3);   // -> rho, kappa, depth (nondimensional)
static emu_rhs_fn g_rhs = nullptr;
static emu_adim_fn g_adim = nullptr;
static std::string g_error;
static int g_force_unconverged = 0;

namespace rb {
int report_error(const std::exception& e) {
    g_error = e.what();
    return -1;
}
rb_props helium_props_from_si(double L, double rho, double kappa, double depth, bool use_expansions, int expansion_order, bool infinite_depth) {
    rb_props p;
    rb_default_props(&p);
    double out[3] = {0, 0, 0};
    g_adim(L, rho, kappa, depth, out);
    p.physics = RB_HELIUM;
    p.rho = out[0];
    p.kappa = out[1];
    p.depth = out[2];
    p.use_expansions = use_expansions;
    p.expansion_order = expansion_order;
    p.infinite_depth = infinite_depth;
    return p;
}
// column-major in-place LU with partial pivoting, one right-hand side (the library's own factorisations are tested elsewhere)
void launch_lu_solve(double* A, double* b, int n, int* info, cudaStream_t) {
    *info = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A[(size_t)k * n + i]) > std::fabs(A[(size_t)k * n + p])) p = i;
        if (A[(size_t)k * n + p] == 0.0) {
            if (!*info) *info = k + 1;
            continue;
        }
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(A[(size_t)j * n + k], A[(size_t)j * n + p]);
            std::swap(b[k], b[p]);
        }
        for (int i = k + 1; i < n; ++i) {
            const double l = A[(size_t)k * n + i] /= A[(size_t)k * n + k];
            for (int j = k + 1; j < n; ++j) A[(size_t)j * n + i] -= l * A[(size_t)j * n + k];
            b[i] -= l * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        b[k] /= A[(size_t)k * n + k];
        for (int i = 0; i < k; ++i) b[i] -= A[(size_t)k * n + i] * b[k];
    }
}
void launch_lu_solve_blocked(double* A, double* b, int n, int* info, cudaStream_t st) { launch_lu_solve(A, b, n, info, st); }
void launch_lu_solve_unblocked(double* A, double* b, int n, int* info, cudaStream_t st) { launch_lu_solve(A, b, n, info, st); }
}  // namespace rb

extern "C" {

const char* rb_last_error(void) { return g_error.c_str(); }

void rb_default_props(rb_props* p) {
    std::memset(p, 0, sizeof(*p));
    p->physics = RB_WATER;
    p->max_iterations = 200;
    p->tolerance = 1e-13;
}

rb_solver* rb_create(int N, int batch, const rb_props* props) {
    rb_solver* s = new rb_solver;
    s->N = N;
    s->batch = batch;
    s->props = *props;
    s->rhs_calls = 0;
    return s;
}

int rb_destroy(rb_solver* s) {
    delete s;
    return 0;
}

int rb_set_stream(rb_solver*, void*) { return 0; }
void* rb_get_stream(rb_solver*) { return nullptr; }

int rb_get_props(rb_solver* s, rb_props* out, int* N, int* batch) {
    if (out) *out = s->props;
    if (N) *N = s->N;
    if (batch) *batch = s->batch;
    return 0;
}

int rb_rhs(rb_solver* s, const rb_complex* state, rb_complex* rhs) {
    g_rhs(&s->props, s->N, s->batch, (const double*)state, (double*)rhs);
    ++s->rhs_calls;
    return 0;
}

int rb_solve_stats(rb_solver* s, double out[6]) {
    out[0] = 1;
    out[1] = g_force_unconverged ? 0 : 1;
    out[2] = 0;
    out[3] = (double)s->rhs_calls;
    out[4] = (double)s->rhs_calls;
    out[5] = 0;
    return 0;
}

int rb_solve_status(rb_solver*, double out[8]) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    out[0] = g_force_unconverged ? 0 : 1;
    out[5] = 1;
    return 0;
}

int rb_set_strict(rb_solver*, int) { return 0; }

void rb_free(void* p) { std::free(p); }

// test hooks
void emu_set_callbacks(emu_rhs_fn rhs, emu_adim_fn adim) {
    g_rhs = rhs;
    g_adim = adim;
}
void emu_force_unconverged(int on) { g_force_unconverged = on; }
long long emu_rhs_calls(rb_solver* s) { return s->rhs_calls; }

}  // extern "C"
