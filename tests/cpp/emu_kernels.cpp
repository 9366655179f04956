// emu_kernels.cpp -- TEST INFRASTRUCTURE: the library's small kernels compiled by g++ against tests/cpp/cuda_emu.h and exported
// with a C interface for tests/test_kernel_emulation.py (CPU tier, no GPU).  The kernel sources are the ones the library ships
// (#included below, unchanged); grids and blocks are the ones implicit.cu / lu_kernels.cu launch them with.
#include "cuda_emu.h"

#include "../../superfluid_dynamics_b200/csrc/implicit_kernels.cuh"
#include "../../superfluid_dynamics_b200/csrc/lu_kernels.cu"

namespace {
inline unsigned blocks_for(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }
}  // namespace

extern "C" {

void emu_real_to_complex_state(const double* y, double* s, int N) {
    emu::launch_seq(real_to_complex_state_kernel, dim3(blocks_for(N)), dim3(256), y, (double2*)s, N);
}

void emu_complex_to_real_rhs(const double* r, double* out, int N) {
    emu::launch_seq(complex_to_real_rhs_kernel, dim3(blocks_for(N)), dim3(256), (const double2*)r, out, N);
}

// both signs in one launch, as jacobian_calculate does; neg may be null (then one sign, as rb_perturbed_states does)
void emu_perturbed_states(const double* state, double* pos, double* neg, double eps, int N) {
    emu::launch_seq(perturbed_states_kernel, dim3(blocks_for(N), 3 * N, neg ? 2 : 1), dim3(256), (const double2*)state, (double2*)pos,
                    (double2*)neg, eps, N);
}

void emu_jacobian_from_perturbed(const double* pos, const double* neg, double* C, int N, double eps) {
    emu::launch_seq(jacobian_from_perturbed_kernel, dim3(blocks_for((size_t)6 * N * N)), dim3(256), (const double2*)pos,
                    (const double2*)neg, C, N, eps);
}

// both stage states as real vectors (ystage = y1 | y2) and as the batch-2 complex state cstate = [Z_1 | Z_2 | Phi_1 | Phi_2]
void emu_gl2_stage_states_batched(const double* y, double h, const double* k, double* ystage, double* cstate, int N) {
    const size_t n = (size_t)3 * N;
    emu::launch_seq(gl2_stage_states_batched_kernel, dim3(blocks_for(N)), dim3(256), y, h, k, k + n, ystage, ystage + n, (double2*)cstate, N);
}

void emu_gl2_batched_rhs_to_real(const double* crhs, double* fy, int N) {
    emu::launch_seq(gl2_batched_rhs_to_real_kernel, dim3(blocks_for(N)), dim3(256), (const double2*)crhs, fy, N);
}

void emu_gl2_residual(const double* fy, const double* k, double* R, size_t n2, double* sums) {
    emu::launch_coop(gl2_residual_kernel, dim3(1), dim3(kNormThreads), fy, k, R, n2, sums);
}

void emu_gl2_trial(const double* k, double alpha, const double* dK, double* kt, size_t n2) {
    emu::launch_seq(gl2_trial_kernel, dim3(blocks_for(n2)), dim3(256), k, alpha, dK, kt, n2);
}

void emu_gl2_negate(const double* R, double* out, size_t n2) {
    emu::launch_seq(gl2_negate_kernel, dim3(blocks_for(n2)), dim3(256), R, out, n2);
}

void emu_gl2_newton_matrix(const double* J1, const double* J2, double h, double* M, size_t n) {
    emu::launch_seq(gl2_newton_matrix_kernel, dim3(blocks_for(n, 32), blocks_for(n, 8)), dim3(32, 8), J1, J2, h, M, n);
}

void emu_gl2_next_state(const double* y, double h, const double* k, double* out, size_t n) {
    emu::launch_seq(gl2_next_state_kernel, dim3(blocks_for(n)), dim3(256), y, h, k, k + n, out, n);
}

// lu_kernels.cu: the blocked factorisation with b eliminated on the fly (panel, swap, trsm, gemm on the emulated mma, gemv);
// returns getrf's info.  The back substitution with U is left to the caller.
int emu_lu_factor_blocked(double* A, double* b, int n) {
    int info = 0;
    rb::lu_factor_blocked(A, b, n, &info, nullptr);
    return info;
}

// the whole blocked solve: factorisation + blocked back substitution; b returns x
int emu_lu_solve_blocked(double* A, double* b, int n) {
    int info = 0;
    rb::launch_lu_solve_blocked(A, b, n, &info, nullptr);
    return info;
}

}  // extern "C"
