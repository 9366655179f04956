// emu_compat.cpp -- TEST INFRASTRUCTURE: the functions and kernels include/cusuperhelium_compat.cuh carries under the reference's
// names for the reference's own tests (T/ComplexFunctionsTests.cuh, T/MatrixMTests.cuh:284-390), compiled by g++ with the toolkit's
// headers and tests/cpp/cuda_emu.h supplying the execution model, exported with a C interface for tests/test_kernel_emulation.py.
#define RB_EMU_CUDA_HEADERS 1
#include "cuda_emu.h"

#include "cusuperhelium_compat.cuh"

namespace {
__global__ void k_sin(const cuDoubleComplex* z, cuDoubleComplex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sin(z[i], out[i]);
}
__global__ void k_cos(const cuDoubleComplex* z, cuDoubleComplex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cos(z[i], out[i]);
}
__global__ void k_green(const std_complex* zk, const std_complex* zj, std_complex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cotangent_green_function(zk[i], zj[i]);
}
__global__ void k_cot(const std_complex* z, std_complex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cot(z[i]);
}
__global__ void k_inv_sub(const std_complex* zk, const std_complex* zj, std_complex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = PrecisionMath::fastPreciseInvSub(zk[i], zj[i]);
}
__global__ void k_two_diff(const std_complex* zk, const std_complex* zj, std_complex* hi, std_complex* lo, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        PrecisionMath::dd_complex d = PrecisionMath::c_twoDiff(zk[i], zj[i]);
        hi[i] = std_complex(d.real.hi, d.imag.hi);
        lo[i] = std_complex(d.real.lo, d.imag.lo);
    }
}
inline unsigned blocks_for(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }
}  // namespace

extern "C" {

void emuc_sin(const double* z, double* out, int n) { emu::launch_seq(k_sin, dim3(blocks_for(n)), dim3(256), (const cuDoubleComplex*)z, (cuDoubleComplex*)out, n); }
void emuc_cos(const double* z, double* out, int n) { emu::launch_seq(k_cos, dim3(blocks_for(n)), dim3(256), (const cuDoubleComplex*)z, (cuDoubleComplex*)out, n); }
// the kernel overload of cotangent_complex, launched as T/ComplexFunctionsTests.cuh:125 does
void emuc_cotangent_complex(const double* z, double* out, int n) {
    void (*kernel)(const cuDoubleComplex*, cuDoubleComplex*, int) = cotangent_complex;
    emu::launch_seq(kernel, dim3(blocks_for(n)), dim3(256), (const cuDoubleComplex*)z, (cuDoubleComplex*)out, n);
}
void emuc_cot(const double* z, double* out, int n) { emu::launch_seq(k_cot, dim3(blocks_for(n)), dim3(256), (const std_complex*)z, (std_complex*)out, n); }
void emuc_green(const double* zk, const double* zj, double* out, int n) {
    emu::launch_seq(k_green, dim3(blocks_for(n)), dim3(256), (const std_complex*)zk, (const std_complex*)zj, (std_complex*)out, n);
}
void emuc_inv_sub(const double* zk, const double* zj, double* out, int n) {
    emu::launch_seq(k_inv_sub, dim3(blocks_for(n)), dim3(256), (const std_complex*)zk, (const std_complex*)zj, (std_complex*)out, n);
}
void emuc_two_diff(const double* zk, const double* zj, double* hi, double* lo, int n) {
    emu::launch_seq(k_two_diff, dim3(blocks_for(n)), dim3(256), (const std_complex*)zk, (const std_complex*)zj, (std_complex*)hi,
                    (std_complex*)lo, n);
}
// the reference's launch geometries: createInitialState<<<N, 1>>>, createInitialBatchedZ<<<(ceil(2N/256), 3N), 256>>>
void emuc_create_initial_state(const double* state, double* cstate, size_t N) {
    emu::launch_seq(createInitialState, dim3((unsigned)N), dim3(1), state, (std_complex*)cstate, N);
}
void emuc_create_initial_batched_z(const double* cstate, double* batched, double eps, size_t N) {
    emu::launch_seq(createInitialBatchedZ, dim3((unsigned)((2 * N + 255) / 256), (unsigned)(3 * N), 1), dim3(256, 1, 1),
                    (const std_complex*)cstate, (std_complex*)batched, eps, N);
}
void emuc_jacobian_from_perturbed(const double* pos, const double* neg, double* C, size_t N, double eps) {
    emu::launch_seq(createJacobianMatrixFromPerturbedRhs, dim3(blocks_for(6 * N * N)), dim3(256), (const std_complex*)pos,
                    (const std_complex*)neg, C, N, eps);
}

}  // extern "C"
