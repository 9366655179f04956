// roberts_b200_device.cuh -- per-entry device formulas of the materialised operators, shared by libroberts_b200's own
// kernels (csrc/dense_kernels.cu) and by the kernels the compatibility header exposes under the reference's names.
// Reference: createMKernel / createFiniteDepthMKernel L/createM.cuh:43-92, createVelocityMatrices /
// createHeliumVelocityMatrices L/WaterVelocities.cuh:38-107 (L/ = CuSuperHelium/CuSuperHelium/).
#pragma once
#include <cuda_runtime.h>

namespace rb_dev {

constexpr double kPi = 3.14159265358979323846;

// cot((a + i b)/2), accurate for every argument (no cosh - cos cancellation):
//   cot(u + i v) = (sin u cos u - i sinh v cosh v) / (sin^2 u + sinh^2 v)
__device__ __forceinline__ double2 cot_half(double a, double b) {
    double u = 0.5 * a, v = 0.5 * b;
    if (fabs(v) > 300.0) return make_double2(0.0, v > 0 ? -1.0 : 1.0);
    double su, cu;
    sincos(u, &su, &cu);
    double sh = sinh(v), ch = cosh(v);
    double inv = 1.0 / (su * su + sh * sh);
    return make_double2(su * cu * inv, -sh * ch * inv);
}

__device__ __forceinline__ double2 cdivd(double2 a, double2 b) {
    double inv = 1.0 / (b.x * b.x + b.y * b.y);
    return make_double2((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}

// entry (row k, col j) of M; FINITE selects createFiniteDepthMKernel (coefficients 1/2 and 1/(4 pi), image term without Zp_k)
template <bool FINITE>
__device__ __forceinline__ double M_entry(int k, int j, const double2* Z, const double2* Zp, const double2* Zpp, double rho,
                                          double h, bool infinite_depth) {
    const double2 zk = Z[k], zpk = Zp[k];
    const double coef = FINITE ? 0.25 / kPi : 0.25 * (1 - rho) / kPi;
    double v;
    if (k == j) {
        double2 q = cdivd(Zpp[k], zpk);
        v = (FINITE ? 0.5 : 0.5 * (1 + rho)) + coef * q.y;
        if (FINITE && !infinite_depth) v -= 0.25 / kPi * cot_half(0.0, 2.0 * (zk.y + h)).y;   // cot(i (Y + h))
    } else {
        const double2 zj = Z[j];
        double2 c = cot_half(zk.x - zj.x, zk.y - zj.y);
        v = coef * (zpk.x * c.y + zpk.y * c.x);                                               // Im(Zp_k * cot)
        if (FINITE && !infinite_depth) {
            // 0.5 (Z_k - conj Z_j) + i h  =  ((x_k - x_j) + i (y_k + y_j + 2h)) / 2
            v -= 0.25 / kPi * cot_half(zk.x - zj.x, zk.y + zj.y + 2.0 * h).y;
        }
    }
    return v;
}

// entry (row k, col j) of V1; writes V2[k] on the diagonal
__device__ __forceinline__ double2 V1_entry(int k, int j, const double2* Z, const double2* Zp, const double2* Zpp, double2* V2,
                                            bool lower, bool helium, double h, bool infinite_depth) {
    const double2 zk = Z[k];
    const double q4 = 1.0 / (4.0 * kPi);
    double2 v;
    if (k == j) {
        const double2 zpk = Zp[k];
        double2 q2 = cdivd(cdivd(Zpp[k], zpk), zpk);               // Zpp / Zp^2
        v = make_double2(q4 * q2.y, -q4 * q2.x);                    // multiply_by_i(-q4 * q2)
        if (helium && !infinite_depth) {
            double2 c = cot_half(0.0, 2.0 * (zk.y + h));
            v.x += -q4 * c.y;                                       // multiply_by_i(q4 * c)
            v.y += q4 * c.x;
        }
        double2 hz = cdivd(make_double2(0.5, 0.0), zpk);
        if (lower) { v.x += hz.x; v.y += hz.y; } else { v.x -= hz.x; v.y -= hz.y; }
        double2 iz = cdivd(make_double2(1.0 / (2.0 * kPi), 0.0), zpk);
        V2[k] = make_double2(-iz.y, iz.x);
    } else {
        const double2 zj = Z[j];
        double2 c = cot_half(zk.x - zj.x, zk.y - zj.y);
        v = make_double2(q4 * c.y, -q4 * c.x);                      // multiply_by_i(-q4 * c)
        if (helium && !infinite_depth) {
            double2 ci = cot_half(zk.x - zj.x, zk.y + zj.y + 2.0 * h);
            v.x += -q4 * ci.y;
            v.y += q4 * ci.x;
        }
    }
    return v;
}

}  // namespace rb_dev
