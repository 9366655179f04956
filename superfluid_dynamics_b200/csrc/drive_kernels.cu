// drive_kernels.cu -- optomechanical drive of the helium film: the element-wise add-ons that turn the boundary-integral RHS into
// the reference's autonomous "augmented" system  y = [Z | Phi | D]  (D = delayed light intensity felt by the film).
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/), three launches there, one here:
//   add_optical_field_drive_terms_no_time_depence   L/createM.cuh:138-149   (HeliumDrivenAutonomousProblem::CalculateRhsPhi,
//                                                                            L/HeliumDrivenAutonomousProblem.cuh:20-25)
//   calculate_intensity_delayed_rhs                 L/createM.cuh:161-169   (DelayedIntensityIntegrator::run,
//   add_delayed_intensity_phi_rhs                   L/createM.cuh:151-159    L/DelayedIntensityIntegrator.cuh:21-33)
//   LightIntensity::compute_intensity / compute_x_profile                   L/LightIntensity.cuh:17-29
//   add_optical_field_drive_terms + DelayedIntensityTermDevice              L/createM.cuh:119-136, L/DelayedIntensityTerm.cuh:9-34
//                                                                            (HeliumWithOptomechanicalDrivingProblem, time-dependent)
// HBM-bound and tiny: per point 3 complex reads (Z, w, D), one read-modify-write (dPhi/dt) and one write (dD/dt) = 96 B.
#include "../../include/roberts_b200.h"
#include "internal.cuh"

namespace rb {

// Lorentzian cavity response times the Gaussian transverse profile of the optical mode
__device__ __forceinline__ double light_intensity(double height, double x, const rb_opto& v) {
    const double delta_f = v.detuning - v.G * height;
    const double half = v.gamma / 2;
    const double dx = x - v.location_x0_mode;
    const double profile = exp(-(dx * dx) / (2 * (v.sigma_optical_mode * v.sigma_optical_mode)));
    return 0.25 * (v.gamma * v.gamma) * v.max_intensity / (delta_f * delta_f + half * half) * profile;
}

__global__ void light_intensity_kernel(const double2* __restrict__ Z, double* __restrict__ out, rb_opto v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 z = Z[i];
    out[i] = light_intensity(z.y, z.x, v);
}

void launch_light_intensity(const double2* Z, double* out, const rb_opto& v, size_t n, cudaStream_t st) {
    light_intensity_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Z, out, v, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// state = [Z | Phi | D], rhs = [w | dPhi/dt | dD/dt], each block BN complex; w and dPhi/dt hold the boundary-integral RHS on entry.
//   dPhi/dt += DampingStrength * Im w + drive_strength * I(Z) + D        (same order of additions as the reference's launches)
//   dD/dt    = Beta * I(Z) - D / Tau
__global__ void augmented_terms_kernel(const double2* __restrict__ state, double2* __restrict__ rhs, rb_opto v, size_t BN) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BN) return;
    const double2 z = state[i];
    const double2 d = state[2 * BN + i];
    const double2 w = rhs[i];
    double2 r = rhs[BN + i];
    const double I = light_intensity(z.y, z.x, v);
    r.x += v.DampingStrength * w.y;
    r.x += v.drive_strength * I;
    r.x += d.x;
    r.y += d.y;
    rhs[BN + i] = r;
    const double inv_tau = 1.0 / v.Tau;
    rhs[2 * BN + i] = make_double2(v.Beta * I - inv_tau * d.x, -(inv_tau * d.y));
}

void launch_augmented_terms(const double2* state, double2* rhs, const rb_opto& v, size_t BN, cudaStream_t st) {
    augmented_terms_kernel<<<(unsigned)((BN + 255) / 256), 256, 0, st>>>(state, rhs, v, BN);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// The explicitly time-dependent drive: add_optical_field_drive_terms (L/createM.cuh:119-136) with the exponential integrator of
// DelayedIntensityTermDevice (L/DelayedIntensityTerm.cuh:16-33).  `delayed` holds the delayed intensity saved at time prev_time:
//   D(t) = I                                              if t == prev_time   (the reference's first-call convention)
//        = a D_saved + Beta Tau (1 - a) I,  a = exp(-(t - prev_time)/Tau)     otherwise
//   dPhi/dt += DampingStrength * Im w + Beta * D(t) + drive_strength * I
// and, when `save` is set (first RK stage only), D_saved <- D(t).  The reference's kernel reads *prev_time from device memory
// while other threads of the same launch overwrite it (save_value): here prev_time is a launch argument kept by the host, i.e.
// every thread sees the value from before the launch -- the intended semantics, without the race.
__global__ void timed_drive_kernel(double2* __restrict__ rhs_phi, const double2* __restrict__ Z, const double2* __restrict__ w,
                                   double* __restrict__ delayed, rb_opto v, double time, double prev_time, int save, size_t BN) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BN) return;
    const double2 z = Z[i];
    const double I = light_intensity(z.y, z.x, v);
    double d;
    if (time == prev_time) {
        d = I;
    } else {
        const double a = exp(-(time - prev_time) / v.Tau);
        d = a * delayed[i] + v.Beta * v.Tau * (1 - a) * I;
    }
    if (save) delayed[i] = d;
    double2 r = rhs_phi[i];
    r.x += v.DampingStrength * w[i].y;
    r.x += v.Beta * d;
    r.x += v.drive_strength * I;
    rhs_phi[i] = r;
}

void launch_timed_drive(double2* rhs_phi, const double2* Z, const double2* w, double* delayed, const rb_opto& v, double time,
                        double prev_time, int save, size_t BN, cudaStream_t st) {
    timed_drive_kernel<<<(unsigned)((BN + 255) / 256), 256, 0, st>>>(rhs_phi, Z, w, delayed, v, time, prev_time, save, BN);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
