// compat_test.cu -- the reference's own kernel tests, written against the reference's class / kernel names as provided by
// include/cusuperhelium_compat.cuh, linked with libroberts_b200.so.  Fixtures and tolerances follow
// CuSuperHelium.Tests/MatrixMTests.cuh (Kernels.TwoByTwoMMatrix :160-223, Kernels.MMatrixKernel :225-282, Kernels.Velocities
// :396-468, Kernels.ZPhiDerivatives :477-607, Kernels.RhsPhi :748-819) and the App's stepper usage (CuSuperHelium.App/kernel.cu:85-96).
// Prints one line per test and exits non-zero on any failure.  Needs a GPU.
#include <cmath>
#include <complex>
#include <cstdio>
#include <string>
#include <vector>

#include "cusuperhelium_compat.cuh"

using cd = std::complex<double>;
static int failures = 0;
#define EXPECT_NEAR(a, b, tol, what)                                                            \
    do {                                                                                        \
        if (!(std::fabs((a) - (b)) <= (tol))) {                                                 \
            if (failures < 20) std::printf("  FAIL %s: %.17g vs %.17g\n", what, (double)(a), (double)(b)); \
            ++failures;                                                                         \
        }                                                                                       \
    } while (0)

static double X(double j, double h, double w, double t) { return j - h * std::sin(j - w * t); }
static double Y(double j, double h, double w, double t) { return h * std::cos(j - w * t); }
static double Xp(double j, double h, double w, double t) { return 1 - h * std::cos(j - w * t); }
static double Yp(double j, double h, double w, double t) { return -h * std::sin(j - w * t); }
static double Xpp(double j, double h, double w, double t) { return h * std::sin(j - w * t); }
static double Ypp(double j, double h, double w, double t) { return -h * std::cos(j - w * t); }
static double PhiF(double j, double h, double w, double t, double rho) { return h * (1 + rho) * w * std::sin(j - w * t); }

static void prepareZPhi(std::vector<cd>& Z, std::vector<cd>& Phi, std::vector<cd>& Zp, std::vector<cd>& Zpp, std::vector<cd>& PhiP,
                        double h, double w, double t, double rho, int N) {
    for (int i = 0; i < N; i++) {
        double j = 2 * PI_d * i / (double)N, s = 2.0 * PI_d / N;
        Z[i] = cd(X(j, h, w, t), Y(j, h, w, t));
        Phi[i] = PhiF(j, h, w, t, rho);
        Zp[i] = cd(Xp(j, h, w, t) * s, Yp(j, h, w, t) * s);
        Zpp[i] = cd(Xpp(j, h, w, t) * s * s, Ypp(j, h, w, t) * s * s);
        PhiP[i] = h * (1.0 + rho) * w * std::cos(j - w * t) * s;
    }
}

template <typename T>
static T* toDevice(const std::vector<T>& v) {
    T* d;
    cudaMalloc(&d, v.size() * sizeof(T));
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

static void test_matrices(int N, double t) {
    const double h = 0.5, w = 10, rho = 0;
    std::vector<cd> Z(N), Phi(N), Zp(N), Zpp(N), PhiP(N);
    prepareZPhi(Z, Phi, Zp, Zpp, PhiP, h, w, t, rho, N);
    auto* dZ = (std_complex*)toDevice(Z);
    auto* dZp = (std_complex*)toDevice(Zp);
    auto* dZpp = (std_complex*)toDevice(Zpp);
    double* dM;
    cudaMalloc(&dM, N * N * sizeof(double));
    dim3 th(16, 16), bl((N + 15) / 16, (N + 15) / 16);
    createMKernel<<<bl, th>>>(dM, dZ, dZp, dZpp, 0.0, N, 1);
    cudaDeviceSynchronize();
    std::vector<double> M(N * N);
    cudaMemcpy(M.data(), dM, N * N * sizeof(double), cudaMemcpyDeviceToHost);
    if (N == 2) {   // closed form, MatrixMTests.cuh:69-87
        double th2 = std::sinh(2.0 * h) / (std::cosh(2.0 * h) + 1.0);
        EXPECT_NEAR(M[0 + 0 * N], 0.5 - 0.25 * h / (1.0 - h), 1e-14, "M00");
        EXPECT_NEAR(M[1 + 1 * N], 0.5 + 0.25 * h / (1.0 + h), 1e-14, "M11");
        EXPECT_NEAR(M[1 + 0 * N], (h + 1) / 4.0 * th2, 1e-14, "M10");
        EXPECT_NEAR(M[0 + 1 * N], (h - 1) / 4.0 * th2, 1e-14, "M01");
    }
    for (int i = 0; i < N; i++)       // host loop, MatrixMTests.cuh:89-115
        for (int j = 0; j < N; j++) {
            double e = (i == j) ? 0.5 * (1 + rho) + 0.25 * (1 - rho) / PI_d * std::imag(Zpp[j] / Zp[j])
                                : 0.25 * (1.0 - rho) / PI_d * std::imag(Zp[i] / std::tan(0.5 * (Z[i] - Z[j])));
            EXPECT_NEAR(M[i + j * N], e, 1e-14, "M host loop");
        }
    std_complex *dV1, *dV2;
    cudaMalloc(&dV1, N * N * sizeof(std_complex));
    cudaMalloc(&dV2, N * sizeof(std_complex));
    createVelocityMatrices<<<bl, th>>>(dZ, dZp, dZpp, N, dV1, dV2, true);
    cudaDeviceSynchronize();
    std::vector<cd> V1(N * N), V2(N);
    cudaMemcpy(V1.data(), dV1, N * N * sizeof(cd), cudaMemcpyDeviceToHost);
    cudaMemcpy(V2.data(), dV2, N * sizeof(cd), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++)       // MatrixMTests.cuh:117-151
        for (int j = 0; j < N; j++) {
            cd e = (i == j) ? cd(0, -0.25 / PI_d) * Zpp[j] / (Zp[j] * Zp[j]) + 0.5 / Zp[j]
                            : cd(0, -0.25 / PI_d) / std::tan(0.5 * (Z[i] - Z[j]));
            EXPECT_NEAR(V1[i + j * N].real(), e.real(), 1e-12, "V1 re");
            EXPECT_NEAR(V1[i + j * N].imag(), e.imag(), 1e-12, "V1 im");
            if (i == j) {
                cd e2 = cd(0, 0.5) / (PI_d * Zp[j]);
                EXPECT_NEAR(V2[i].real(), e2.real(), 1e-12, "V2 re");
                EXPECT_NEAR(V2[i].imag(), e2.imag(), 1e-12, "V2 im");
            }
        }
    cudaFree(dZ); cudaFree(dZp); cudaFree(dZpp); cudaFree(dM); cudaFree(dV1); cudaFree(dV2);
    std::printf("Kernels.M/Velocities N=%d: failures so far %d\n", N, failures);
}

static void test_zphi_derivatives() {
    constexpr int N = 1024;
    const double h = 0.5, w = 10, t = 0;
    ProblemProperties properties;
    properties.rho = 0.0;
    std::vector<cd> Z(N), Phi(N), Zp(N), Zpp(N), PhiP(N);
    prepareZPhi(Z, Phi, Zp, Zpp, PhiP, h, w, t, properties.rho, N);
    auto* dZ = (std_complex*)toDevice(Z);
    auto* dPhi = (std_complex*)toDevice(Phi);
    std_complex *dZp, *dZpp, *dPhiP;
    cudaMalloc(&dZp, N * sizeof(std_complex));
    cudaMalloc(&dZpp, N * sizeof(std_complex));
    cudaMalloc(&dPhiP, N * sizeof(std_complex));
    ZPhiDerivative<N, 1> zPhiDerivative(properties);
    zPhiDerivative.exec(dZ, dPhi, dZp, dPhiP, dZpp);
    cudaDeviceSynchronize();
    std::vector<cd> cZp(N), cZpp(N), cPhiP(N);
    cudaMemcpy(cZp.data(), dZp, N * sizeof(cd), cudaMemcpyDeviceToHost);
    cudaMemcpy(cZpp.data(), dZpp, N * sizeof(cd), cudaMemcpyDeviceToHost);
    cudaMemcpy(cPhiP.data(), dPhiP, N * sizeof(cd), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++) {
        EXPECT_NEAR(cZp[i].real(), Zp[i].real(), 1e-14, "Xprime");
        EXPECT_NEAR(cZp[i].imag(), Zp[i].imag(), 1e-14, "Yprime");
        EXPECT_NEAR(cZpp[i].real(), Zpp[i].real(), 1e-14, "Xpp");
        EXPECT_NEAR(cZpp[i].imag(), Zpp[i].imag(), 1e-14, "Ypp");
        EXPECT_NEAR(cPhiP[i].real(), PhiP[i].real(), 1e-14, "PhiPrime");
    }
    std::printf("Kernels.ZPhiDerivatives: failures so far %d\n", failures);
}

static void test_rhs_phi() {
    const int N = 32;
    const double w = 10.0, h = 0.5, t = 0.1;
    std::vector<cd> Z(N), VL(N), VU(N, 0.0);
    std::vector<double> e(N);
    for (int i = 0; i < N; i++) {
        double j = 2 * PI_d * i / (double)N;
        double y = Y(i, h, w, t);
        Z[i] = cd(X(j, h, w, t), y);
        VL[i] = cd(2 * PI_d / N * Xp(j, h, w, t), 2 * PI_d / N * Yp(j, h, w, t));
        e[i] = -y + 0.5 * std::norm(VL[i]);
    }
    auto* dZ = (std_complex*)toDevice(Z);
    auto* dVL = (std_complex*)toDevice(VL);
    auto* dVU = (std_complex*)toDevice(VU);
    std_complex* dR;
    cudaMalloc(&dR, N * sizeof(std_complex));
    compute_rhs_phi_expression<<<1, 256>>>(dZ, dVL, dVU, dR, 0.0, N);
    cudaDeviceSynchronize();
    std::vector<cd> r(N);
    cudaMemcpy(r.data(), dR, N * sizeof(cd), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++) {
        EXPECT_NEAR(r[i].real(), e[i], 1e-14, "RhsPhi");
        EXPECT_NEAR(r[i].imag(), 0.0, 1e-14, "RhsPhi imag");
    }
    std::printf("Kernels.RhsPhi: failures so far %d\n", failures);
}

static void test_stepper() {
    // App pattern (CuSuperHelium.App/kernel.cu:85-96) on a deep-water wave: after t the profile is eps cos(x - t) + O(eps^2)
    constexpr int N = 256;
    const double eps = 1e-4, dt = 1e-2;
    ProblemProperties properties;
    properties.rho = 0.0;
    WaterBoundaryProblem<N, 1> problem(properties);
    BaseBoundaryIntegralCalculator<N, 1> integrator(properties, problem);
    auto logger = std::make_shared<TrajectoryLogger<std_complex, 2 * N>>(10, 16);
    AutonomousRungeKuttaStepper<std_complex, 2 * N> stepper(integrator, dt, logger);
    RK4Options opts;
    opts.initial_timestep = dt;
    stepper.setOptions(opts);
    std::vector<std_complex> y0(2 * N);
    for (int i = 0; i < N; i++) {
        double a = 2 * PI_d * i / N;
        y0[i] = std_complex(a, eps * std::cos(a));
        y0[N + i] = std_complex(eps * std::sin(a), 0.0);
    }
    stepper.initialize(y0.data(), false);
    stepper.runEvolution(0.0, 1.0);
    std::vector<std_complex> y(2 * N);
    stepper.getState(y.data());
    const double t = 1.0;
    for (int i = 0; i < N; i++) EXPECT_NEAR(y[i].imag(), eps * std::cos(y[i].real() - t), 5 * eps * eps, "dispersion");
    double* times; size_t nt;
    logger->copyTimesToHost(&times, &nt);
    EXPECT_NEAR((double)nt, 10.0, 0.0, "logged states");
    if (nt == 10) EXPECT_NEAR(times[9], 1.0, 1e-12, "last logged time");
    rb_free(times);
    double ekin = integrator.kineticEnergy.getEnergy(), epot = integrator.potentialEnergy.getEnergy();
    EXPECT_NEAR(ekin + epot, 0.5 * eps * eps, 0.02 * eps * eps, "energy of a linear wave = eps^2/2 per unit (g = 1, period 2 pi)/(2 pi)");
    std::printf("Stepper (App pattern): failures so far %d, E = %.6e\n", failures, ekin + epot);
}

// TEST(ODE_Solvers, RK45), T/ODESolverTests.cuh:248-421: dz/dt = i z through RK45_std_complex with the reference's own problem class
// shape (AutonomousProblem<std_complex, N>::run launching a kernel on the stepper's stream); checked against the closed form
__global__ void flip_x_y(const std_complex* in, std_complex* out, int n) {   // T/ODESolverTests.cuh:25-33: dx/dt = -y, dy/dt = x
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = std_complex(-in[i].imag(), in[i].real());
}
template <int N>
class OscillatoryProblemStdComplex : public AutonomousProblem<std_complex, N> {
    cudaStream_t stream = nullptr;
public:
    void run(std_complex* state, std_complex* rhs) override { flip_x_y<<<(N + 255) / 256, 256, 0, stream>>>(state, rhs, N); }
    void setStream(cudaStream_t s) override { stream = s; }
};

static void test_rk45() {
    constexpr int N = 256;
    std::vector<std_complex> z0(N), z(N);
    for (int i = 0; i < N; ++i) z0[i] = std_complex(2 * M_PI * i / N, std::sin(2 * M_PI * i * 0.01));
    std_complex* dev = nullptr;
    cudaMalloc(&dev, N * sizeof(std_complex));
    cudaMemcpy(dev, z0.data(), N * sizeof(std_complex), cudaMemcpyHostToDevice);
    OscillatoryProblemStdComplex<N> problem;
    RK45_std_complex<N> stepper(problem, 1e-3);
    stepper.initialize(dev, true);
    stepper.setTolerance(1e-8, 1e-8);
    OdeSolverResult res = stepper.runEvolution(0, 10.0);
    EXPECT_NEAR(res == OdeSolverResult::ReachedEndTime ? 0.0 : 1.0, 0.0, 0.5, "RK45 reaches the end time");
    EXPECT_NEAR(stepper.getCurrentTime(), 10.0, 1e-12, "RK45 end time");
    cudaMemcpy(z.data(), stepper.getY(), N * sizeof(std_complex), cudaMemcpyDeviceToHost);
    const double c = std::cos(10.0), s = std::sin(10.0);
    double worst = 0;
    for (int i = 0; i < N; ++i) {
        worst = std::max(worst, std::abs(z[i].real() - (z0[i].real() * c - z0[i].imag() * s)));
        worst = std::max(worst, std::abs(z[i].imag() - (z0[i].real() * s + z0[i].imag() * c)));
    }
    EXPECT_NEAR(worst, 0.0, 1e-5, "RK45 rotation by 10 rad (the reference compares with 1e-2)");
    cudaFree(dev);
    std::printf("RK45 (reference test pattern): failures so far %d, max error %.3e\n", failures, worst);
}

// App pattern of the optomechanically driven film (CuSuperHelium.App/kernel.cu:60-96, augmentedSystem()): a flat film at rest has
// w = 0 and a vanishing van-der-Waals term, so the augmented RHS is the drive alone:
//   dPhi/dt = drive_strength * I(x) + D,   dD/dt = Beta * I(x) - D / Tau,   I = Lorentzian(detuning - G y) * Gaussian(x - x0)
static void test_augmented() {
    constexpr int N = 128;
    ProblemProperties properties;
    properties.depth = 0.0942478;
    properties.rho = 1;
    OptomechanicalVariables opto;
    opto.Beta = 2e-33;
    opto.detuning = 0.5;
    opto.G = 3.0;
    opto.gamma = 2.0;
    opto.location_x0_mode = PI_d;
    opto.sigma_optical_mode = 0.8;
    opto.max_intensity = 1e32;
    opto.Tau = 0.7;
    opto.DampingStrength = 0.01;
    std::vector<std_complex> y0(3 * N), rhs(3 * N);
    std::vector<double> inten(N);
    for (int i = 0; i < N; i++) {
        const double x = 2.0 * PI_d * i / N;
        y0[i] = std_complex(x, 0.0);
        y0[N + i] = 0.0;
        const double df = opto.detuning - opto.G * 0.0;
        inten[i] = 0.25 * opto.gamma * opto.gamma * opto.max_intensity / (df * df + 0.25 * opto.gamma * opto.gamma) *
                   std::exp(-(x - opto.location_x0_mode) * (x - opto.location_x0_mode) / (2 * opto.sigma_optical_mode * opto.sigma_optical_mode));
        y0[2 * N + i] = std_complex(0.3 * opto.Beta * inten[i], 0.0);
    }
    HeliumDrivenAutonomousProblem<N, 1> heliumProblem(properties, opto);
    auto calculator = std::make_unique<BaseBoundaryIntegralCalculator<N, 1>>(properties, heliumProblem);
    AugmentedBoundaryIntegrator<N, 1> integrator(std::move(calculator), std::make_unique<DelayedIntensityIntegrator<N, 1>>(opto), properties);
    std_complex *dState = nullptr, *dRhs = nullptr;
    cudaMalloc(&dState, 3 * N * sizeof(std_complex));
    cudaMalloc(&dRhs, 3 * N * sizeof(std_complex));
    cudaMemcpy(dState, y0.data(), 3 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
    integrator.run(dState, dRhs);
    cudaDeviceSynchronize();
    cudaMemcpy(rhs.data(), dRhs, 3 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
    const double strength = 1.054571817e-34 / (properties.base_energy * properties.base_time * properties.rho) * opto.G /
                            (opto.sigma_optical_mode * opto.sigma_optical_mode);
    for (int i = 0; i < N; i++) {
        EXPECT_NEAR(std::hypot(rhs[i].real(), rhs[i].imag()), 0.0, 1e-12, "flat film at rest does not move");
        EXPECT_NEAR(rhs[N + i].real(), strength * inten[i] + y0[2 * N + i].real(), 1e-12, "driven dPhi/dt");
        EXPECT_NEAR(rhs[2 * N + i].real(), opto.Beta * inten[i] - y0[2 * N + i].real() / opto.Tau, 1e-12, "dD/dt");
    }
    AutonomousRungeKuttaStepper<std_complex, 3 * N> stepper(integrator, 1e-3);
    stepper.initialize(dState, true);
    for (int i = 0; i < 10; i++) stepper.runStep(i);
    std::vector<std_complex> y(3 * N);
    stepper.getState(y.data());
    double dmax = 0, moved = 0;
    for (int i = 0; i < N; i++) {
        // D relaxes towards Beta Tau I with rate 1/Tau (the film has barely moved after 10 steps of 1e-3)
        const double target = opto.Beta * opto.Tau * inten[i];
        const double expect = target + (y0[2 * N + i].real() - target) * std::exp(-0.01 / opto.Tau);
        dmax = std::max(dmax, std::abs(y[2 * N + i].real() - expect));
        moved = std::max(moved, std::abs(y[N + i].real()));
    }
    EXPECT_NEAR(dmax, 0.0, 1e-6, "delayed intensity relaxes exponentially");
    EXPECT_NEAR(moved > 1e-5 ? 1.0 : 0.0, 1.0, 0.5, "the drive has acted on the potential");
    cudaFree(dState);
    cudaFree(dRhs);
    std::printf("Augmented optomechanical system (App pattern): failures so far %d, |D - analytic| = %.3e, max |Phi| = %.3e\n", failures, dmax, moved);
}

// The explicitly time-dependent drive, assembled as L/Export.cu:797-826 / A/kernel.cu:281-366: a flat film at rest feels
//   dPhi/dt = Beta * D(t) + drive_strength * I(x),   D(t0) = I,   D(t) = a D_saved + Beta Tau (1 - a) I,  a = exp(-(t - t_saved)/Tau)
static void test_timed_drive() {
    constexpr int N = 128;
    ProblemProperties properties;
    properties.depth = 0.0942478;
    properties.rho = 1;
    OptomechanicalVariables opto;
    opto.Beta = 2e-33;
    opto.detuning = 0.5;
    opto.G = 3.0;
    opto.gamma = 2.0;
    opto.location_x0_mode = PI_d;
    opto.sigma_optical_mode = 0.8;
    opto.max_intensity = 1e32;
    opto.Tau = 0.7;
    opto.DampingStrength = 0.01;
    std::vector<std_complex> y0(2 * N), rhs(2 * N);
    std::vector<double> inten(N);
    double imax = 0;
    for (int i = 0; i < N; i++) {
        const double x = 2.0 * PI_d * i / N;
        y0[i] = std_complex(x, 0.0);
        y0[N + i] = 0.0;
        inten[i] = 0.25 * opto.gamma * opto.gamma * opto.max_intensity / (opto.detuning * opto.detuning + 0.25 * opto.gamma * opto.gamma) *
                   std::exp(-(x - opto.location_x0_mode) * (x - opto.location_x0_mode) / (2 * opto.sigma_optical_mode * opto.sigma_optical_mode));
        imax = std::max(imax, inten[i]);
    }
    HeliumWithOptomechanicalDrivingProblem<N> heliumProblem(properties, opto);
    TimedBoundaryIntegrator<N, 1> integrator(properties, heliumProblem);
    std_complex *dState = nullptr, *dRhs = nullptr;
    cudaMalloc(&dState, 2 * N * sizeof(std_complex));
    cudaMalloc(&dRhs, 2 * N * sizeof(std_complex));
    cudaMemcpy(dState, y0.data(), 2 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
    const double strength = 1.054571817e-34 / (properties.base_energy * properties.base_time * properties.rho) * opto.G /
                            (opto.sigma_optical_mode * opto.sigma_optical_mode);
    const double t0 = 0.5, dt = 1e-3;
    // first call at the starting time: D = I, saved; a later call without saving decays it and leaves the saved value alone
    integrator.setStartingTime(t0);
    integrator.setSaveProgress(true);
    integrator.run(dState, dRhs);
    cudaDeviceSynchronize();
    cudaMemcpy(rhs.data(), dRhs, 2 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++) {
        EXPECT_NEAR(std::hypot(rhs[i].real(), rhs[i].imag()), 0.0, 1e-12, "flat film at rest does not move (timed)");
        EXPECT_NEAR(rhs[N + i].real(), (opto.Beta + strength) * inten[i], 1e-12, "timed dPhi/dt at the starting time");
    }
    integrator.setCurrentTime(t0 + 0.1);
    integrator.setSaveProgress(false);
    integrator.run(dState, dRhs);
    cudaDeviceSynchronize();
    cudaMemcpy(rhs.data(), dRhs, 2 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
    std::vector<double> dsaved(N);
    cudaMemcpy(dsaved.data(), integrator.delayedIntensity(), N * sizeof(double), cudaMemcpyDeviceToHost);
    const double a1 = std::exp(-0.1 / opto.Tau);
    for (int i = 0; i < N; i++) {
        const double d = a1 * inten[i] + opto.Beta * opto.Tau * (1 - a1) * inten[i];
        EXPECT_NEAR(rhs[N + i].real(), opto.Beta * d + strength * inten[i], 1e-12, "timed dPhi/dt with a decayed delayed intensity");
        EXPECT_NEAR(dsaved[i] / imax, inten[i] / imax, 1e-14, "saveProgress = false leaves the saved delayed intensity alone");
    }
    // evolution with the trajectory: 10 steps from t0, times logged at the START of each step
    RungeKuttaStepper<std_complex, 2 * N> stepper(integrator);
    RK4Options options;
    options.initial_timestep = dt;
    options.returnTrajectory = true;
    stepper.setOptions(options);
    stepper.initialize(dState, true);
    stepper.runEvolution(t0, t0 + 10.5 * dt);
    double* times = nullptr;
    size_t nt = 0, ns = 0;
    std_complex* states = nullptr;
    stepper.copyTimesToHost(&times, &nt);
    stepper.copyStatesToHost(&states, &ns);
    EXPECT_NEAR((double)nt, 10.0, 0.0, "ten times logged");
    EXPECT_NEAR((double)ns, 10.0, 0.0, "ten states logged");
    for (size_t i = 0; i < nt && i < 10; i++) EXPECT_NEAR(times[i], t0 + i * dt, 1e-12, "time at the start of the step");
    std::vector<std_complex> y(2 * N);
    stepper.getState(y.data());
    double moved = 0, dmax = 0, last = 0;
    cudaMemcpy(dsaved.data(), integrator.delayedIntensity(), N * sizeof(double), cudaMemcpyDeviceToHost);
    const double a9 = std::exp(-9 * dt / opto.Tau);
    for (int i = 0; i < N; i++) {
        moved = std::max(moved, std::abs(y[N + i].real()));
        if (ns == 10) last = std::max(last, std::fabs(states[9 * 2 * N + N + i].real() - y[N + i].real()));
        // the first step saves D = I at t0, the nine later ones decay it towards Beta Tau I (the film has barely moved)
        const double expect = a9 * inten[i] + opto.Beta * opto.Tau * (1 - a9) * inten[i];
        dmax = std::max(dmax, std::abs(dsaved[i] - expect) / imax);
    }
    EXPECT_NEAR(dmax, 0.0, 1e-4, "delayed intensity follows the exponential integrator");
    EXPECT_NEAR(last, 0.0, 0.0, "last logged state is the current state");
    EXPECT_NEAR(moved > 1e-4 ? 1.0 : 0.0, 1.0, 0.5, "the timed drive has acted on the potential");
    EXPECT_NEAR(stepper.getCurrentTime(), t0 + 10 * dt, 1e-12, "current time after the evolution");
    rb_free(times);
    rb_free(states);
    cudaFree(dState);
    cudaFree(dRhs);
    std::printf("Time-dependent optomechanical drive (Export pattern): failures so far %d, |D - analytic|/max I = %.3e, max |Phi| = %.3e\n",
                failures, dmax, moved);
}

// The implicit side, assembled as L/Export.cu:680-700: RealBoundaryItegralCalculator + JacobianCalculator + GaussLegendre2.
// Run with `compat_test --implicit` (kept apart from the default run until its first pass on hardware).
static void test_implicit() {
    constexpr int N = 16;
    ProblemProperties properties;
    properties.depth = 0.3;
    properties.rho = 1;
    HeliumBoundaryProblem<N, 1> heliumProblem(properties);
    BaseBoundaryIntegralCalculator<N, 1> calculator(properties, heliumProblem);
    RealBoundaryItegralCalculator<N> realCalculator(calculator);
    HeliumBoundaryProblem<N, 3 * N> heliumJacProblem(properties);
    JacobianCalculator<N> jacobianCalculator(properties, heliumJacProblem);

    // flat film at rest: d(dPhi_i/dt)/dy_j = -delta_ij (van-der-Waals stiffness 3 vdw / h with vdw = h / 3, L/createM.cuh:109-117)
    std::vector<double> flat(3 * N, 0.0), jac(9 * N * N);
    for (int i = 0; i < N; i++) flat[i] = 2.0 * PI_d * i / N;
    double* dFlat = toDevice(flat);
    double* dJac = nullptr;
    cudaMalloc(&dJac, 9 * N * N * sizeof(double));
    jacobianCalculator.setEpsilon(1e-6);
    jacobianCalculator.calculateJacobian(dFlat, dJac);
    cudaDeviceSynchronize();
    cudaMemcpy(jac.data(), dJac, 9 * N * N * sizeof(double), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)   // column N + j (y_j), row 2N + i (dPhi_i/dt), column-major
            EXPECT_NEAR(jac[(size_t)(N + j) * 3 * N + 2 * N + i], i == j ? -1.0 : 0.0, 1e-6, "d(dPhi/dt)/dy of the flat film");

    // a small standing wave: 10 implicit steps of 0.05 against 500 explicit RK4 steps of 1e-3
    std::vector<double> y0(3 * N);
    std::vector<std_complex> c0(2 * N);
    for (int i = 0; i < N; i++) {
        const double a = 2.0 * PI_d * i / N;
        y0[i] = a - 0.3 * 0.05 * 0.3 * std::sin(a);
        y0[N + i] = 0.05 * 0.3 * std::cos(a);
        y0[2 * N + i] = 0.2 * 0.05 * 0.3 * std::sin(a);
        c0[i] = std_complex(y0[i], y0[N + i]);
        c0[N + i] = std_complex(y0[2 * N + i], 0.0);
    }
    GaussLegendre2Options options;
    options.stepSize = 0.05;
    GaussLegendre2<N> integrator(realCalculator, jacobianCalculator, options);
    integrator.initialize(y0.data(), false);
    integrator.runEvolution(0.0, 0.5);
    double *times = nullptr, *states = nullptr;
    size_t nt = 0, ns = 0;
    integrator.copyTimesToHost(&times, &nt);
    integrator.copyStatesToHost(&states, &ns);
    EXPECT_NEAR((double)nt, (double)ns, 0.0, "as many times as states");
    EXPECT_NEAR(nt >= 11 ? 1.0 : 0.0, 1.0, 0.5, "initial state + at least ten steps logged");
    if (nt) {
        EXPECT_NEAR(times[0], 0.0, 0.0, "trajectory starts at t0");
        EXPECT_NEAR(times[nt - 1], 0.5, 1e-12, "trajectory ends at t1");
    }
    AutonomousRungeKuttaStepper<std_complex, 2 * N> stepper(calculator, 1e-3);
    stepper.initialize(c0.data(), false);
    for (int i = 0; i < 500; i++) stepper.runStep(i);
    std::vector<std_complex> ye(2 * N);
    stepper.getState(ye.data());
    double diff = 0, moved = 0;
    if (ns)
        for (int i = 0; i < N; i++) {
            const double* yl = states + (ns - 1) * 3 * N;
            diff = std::max(diff, std::fabs(yl[i] - ye[i].real()));
            diff = std::max(diff, std::fabs(yl[N + i] - ye[i].imag()));
            diff = std::max(diff, std::fabs(yl[2 * N + i] - ye[N + i].real()));
            moved = std::max(moved, std::fabs(yl[N + i] - y0[N + i]));
        }
    EXPECT_NEAR(diff, 0.0, 1e-8, "Gauss-Legendre-2 agrees with explicit RK4");
    EXPECT_NEAR(moved > 1e-4 ? 1.0 : 0.0, 1.0, 0.5, "the film has moved");
    rb_free(times);
    rb_free(states);
    cudaFree(dFlat);
    cudaFree(dJac);
    std::printf("Implicit Gauss-Legendre-2 + FD Jacobian (Export pattern): failures so far %d, |GL2 - RK4| = %.3e, moved %.3e\n", failures,
                diff, moved);
}

// ---- the complex elementary functions and double-double helpers the reference's tests call by name
// (CuSuperHelium.Tests/ComplexFunctionsTests.cuh) and the perturbed-state kernels of Kernels.BatchedFormationForJacobian
// (MatrixMTests.cuh:284-390).  Run with `compat_test --functions`. ----------------------------------------------------------------
__global__ void complexSinKernel(cuDoubleComplex* zs, cuDoubleComplex* out, int N) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) sin(zs[idx], out[idx]);
}
__global__ void complexCosKernel(cuDoubleComplex* zs, cuDoubleComplex* out, int N) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) cos(zs[idx], out[idx]);
}
__global__ void complexCotKernel(std_complex* zsk, std_complex* zsj, std_complex* out, int N) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) out[idx] = cotangent_green_function(zsk[idx], zsj[idx]);
}
__global__ void test_precision_inversion(std_complex* zsk, std_complex* zsj, std_complex* out, int N) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) out[idx] = PrecisionMath::fastPreciseInvSub(zsk[idx], zsj[idx]);
}
__global__ void test_precision_substraction(std_complex* zsk, std_complex* zsj, std_complex* out, std_complex* low, int N) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N) {
        PrecisionMath::dd_complex d = PrecisionMath::c_twoDiff(zsk[idx], zsj[idx]);
        out[idx] = std_complex(d.real.hi, d.imag.hi);
        low[idx] = std_complex(d.real.lo, d.imag.lo);
    }
}

static void test_functions() {
    constexpr int N = 256;
    // sin, cos, cot along the diagonal of the first period (imaginary parts up to 2 pi: cosh ~ 268)
    std::vector<cuDoubleComplex> zs(N), r1(N), r2(N), r3(N);
    for (int i = 0; i < N; ++i) {
        const double v = 2 * PI_d / (N + 1) * (i + 1);
        zs[i] = make_cuDoubleComplex(v, v);
    }
    cuDoubleComplex *dz = toDevice(zs), *dout = nullptr;
    cudaMalloc(&dout, N * sizeof(cuDoubleComplex));
    complexSinKernel<<<1, N>>>(dz, dout, N);
    cudaMemcpy(r1.data(), dout, N * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost);
    complexCosKernel<<<1, N>>>(dz, dout, N);
    cudaMemcpy(r2.data(), dout, N * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost);
    cotangent_complex<<<1, N>>>(dz, dout, N);
    cudaMemcpy(r3.data(), dout, N * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; ++i) {
        const cd z(zs[i].x, zs[i].y), s = std::sin(z), c = std::cos(z), ct = c / s;
        EXPECT_NEAR(std::abs(cd(r1[i].x, r1[i].y) - s) / std::abs(s), 0.0, 4e-15, "complex sin");
        EXPECT_NEAR(std::abs(cd(r2[i].x, r2[i].y) - c) / std::abs(c), 0.0, 4e-15, "complex cos");
        EXPECT_NEAR(std::abs(cd(r3[i].x, r3[i].y) - ct) / std::abs(ct), 0.0, 1e-14, "complex cot");
    }
    // known answers (mpmath, 40 digits): cot(1e-3 + 2e-3 i), cot(0.7 - 1.3 i), cot(3 + 40 i)
    {
        std::vector<cuDoubleComplex> q = {make_cuDoubleComplex(1e-3, 2e-3), make_cuDoubleComplex(0.7, -1.3), make_cuDoubleComplex(3.0, 40.0)};
        const cd expect[3] = {cd(199.99966666691110686, -400.00066666662221382), cd(0.14933231644907048137, 1.0145011371447459055),
                              cd(-1.0086068994196989300e-35, -1.0)};
        cuDoubleComplex* dq = toDevice(q);
        cotangent_complex<<<1, 32>>>(dq, dq, 3);
        cudaMemcpy(q.data(), dq, 3 * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 3; ++i) {
            EXPECT_NEAR(q[i].x, expect[i].real(), 2e-15 * std::abs(expect[i]), "cot known answer (real)");
            EXPECT_NEAR(q[i].y, expect[i].imag(), 2e-15 * std::abs(expect[i]), "cot known answer (imag)");
        }
        cudaFree(dq);
    }
    // ComplexFunctionsTests.ComplexCotKernel: Zk = pi (1 + i), Zj = Zk + 0.1 (1 + i); the reference's 50-digit value, its 1e-13
    // ComplexFunctionsTests.PrecisionInv: Zk = i (1 + i), Zj = Zk - pi (1 + i) -> 1 / (pi (1 + i)), its 1e-13
    // ComplexFunctionsTests.SubstractionPrecise: hi = -pi, lo = 0 (the subtraction is exact there)
    {
        constexpr int M = 16;
        std::vector<std_complex> zk(M), zj(M), zk2(M), zj2(M), zj3(M), out(M), low(M);
        for (int i = 0; i < M; ++i) {
            zk[i] = std_complex(PI_d, PI_d);
            zj[i] = std_complex(PI_d + 0.1, PI_d + 0.1);
            zk2[i] = std_complex(i, i);
            zj2[i] = std_complex(i - PI_d, i - PI_d);
            zj3[i] = std_complex(i + PI_d, i + PI_d);
        }
        std_complex *dk = toDevice(zk), *dj = toDevice(zj), *dk2 = toDevice(zk2), *dj2 = toDevice(zj2), *dj3 = toDevice(zj3);
        std_complex *dres = nullptr, *dlow = nullptr;
        cudaMalloc(&dres, M * sizeof(std_complex));
        cudaMalloc(&dlow, M * sizeof(std_complex));
        complexCotKernel<<<1, 256>>>(dk, dj, dres, M);
        cudaMemcpy(out.data(), dres, M * sizeof(std_complex), cudaMemcpyDeviceToHost);
        for (int i = 0; i < M; ++i) {
            EXPECT_NEAR(out[i].real(), -9.9833388915330681153509188355277398344111330657474, 1e-13, "cotangent_green_function (real)");
            EXPECT_NEAR(out[i].imag(), 10.016672219575397493791065794281138271298250914604, 1e-13, "cotangent_green_function (imag)");
        }
        test_precision_inversion<<<1, 256>>>(dk2, dj2, dres, M);
        cudaMemcpy(out.data(), dres, M * sizeof(std_complex), cudaMemcpyDeviceToHost);
        for (int i = 0; i < M; ++i) {
            EXPECT_NEAR(out[i].real(), 0.15915494309189533576888376337251436203445964574046, 1e-13, "fastPreciseInvSub (real)");
            EXPECT_NEAR(out[i].imag(), -0.15915494309189533576888376337251436203445964574046, 1e-13, "fastPreciseInvSub (imag)");
        }
        test_precision_substraction<<<1, 256>>>(dk2, dj3, dres, dlow, M);
        cudaMemcpy(out.data(), dres, M * sizeof(std_complex), cudaMemcpyDeviceToHost);
        cudaMemcpy(low.data(), dlow, M * sizeof(std_complex), cudaMemcpyDeviceToHost);
        for (int i = 0; i < M; ++i) {
            const double exact = (double)i - ((double)i + PI_d);   // exact in double: the operands are within a factor 2..8 of each other
            EXPECT_NEAR(out[i].real(), exact, 0.0, "c_twoDiff hi (real)");
            EXPECT_NEAR(out[i].imag(), exact, 0.0, "c_twoDiff hi (imag)");
            EXPECT_NEAR(out[i].real(), -PI_d, 4e-15, "c_twoDiff hi is -pi to the rounding of i + pi");
            EXPECT_NEAR(low[i].real(), 0.0, 0.0, "c_twoDiff lo (real)");
            EXPECT_NEAR(low[i].imag(), 0.0, 0.0, "c_twoDiff lo (imag)");
        }
        // and where the subtraction is NOT exact the low part carries what the high part lost
        double hi, lo;
        PrecisionMath::twoDiff(1.0, 1e-20, hi, lo);
        EXPECT_NEAR(hi, 1.0, 0.0, "twoDiff hi");
        EXPECT_NEAR(lo, -1e-20, 0.0, "twoDiff lo");
        for (auto p : {dk, dj, dk2, dj2, dj3, dres, dlow}) cudaFree(p);
    }
    // Kernels.BatchedFormationForJacobian: the reference's launch geometry, checked entry by entry (the reference test only prints)
    {
        constexpr int NB = 64;
        std::vector<double> state(3 * NB);
        for (int i = 0; i < NB; ++i) {
            const double j = 2 * PI_d * i / NB;
            state[i] = X(j, 0.5, 10, 0.1);
            state[NB + i] = Y(j, 0.5, 10, 0.1);
            state[2 * NB + i] = PhiF(j, 0.5, 10, 0.1, 0.0);
        }
        const double eps = 1e-6;
        double* dState = toDevice(state);
        std_complex *dInit = nullptr, *dB = nullptr;
        cudaMalloc(&dInit, 2 * NB * sizeof(std_complex));
        cudaMalloc(&dB, 6 * NB * NB * sizeof(std_complex));
        createInitialState<<<NB, 1>>>(dState, dInit, NB);
        createInitialBatchedZ<<<dim3((2 * NB + 255) / 256, 3 * NB, 1), dim3(256, 1, 1)>>>(dInit, dB, eps, NB);
        std::vector<std_complex> zb(6 * NB * NB);
        cudaMemcpy(zb.data(), dB, zb.size() * sizeof(std_complex), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int b = 0; b < 3 * NB; ++b)
            for (int i = 0; i < NB; ++i) {
                const int c = b / NB, j = b % NB;
                cd z(state[i], state[NB + i]), p(state[2 * NB + i], 0.0);
                if (i == j) {
                    if (c == 0) z += cd(eps, 0);
                    if (c == 1) z += cd(0, eps);
                    if (c == 2) p += cd(eps, 0);
                }
                const std_complex gz = zb[(size_t)b * NB + i], gp = zb[(size_t)3 * NB * NB + (size_t)b * NB + i];
                if (gz.real() != z.real() || gz.imag() != z.imag() || gp.real() != p.real() || gp.imag() != p.imag()) ++bad;
            }
        EXPECT_NEAR((double)bad, 0.0, 0.0, "createInitialBatchedZ entries");
        cudaFree(dState);
        cudaFree(dInit);
        cudaFree(dB);
    }
    cudaFree(dz);
    cudaFree(dout);
    std::printf("Complex functions, double-double helpers, perturbed-state kernels: failures so far %d\n", failures);
}

int main(int argc, char** argv) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        std::printf("no CUDA device\n");
        return 2;
    }
    if (argc > 1 && std::string(argv[1]) == "--functions") {
        test_functions();
        std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL PASSED", failures);
        return failures ? 1 : 0;
    }
    if (argc > 1 && std::string(argv[1]) == "--implicit") {
        test_implicit();
        std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL PASSED", failures);
        return failures ? 1 : 0;
    }
    test_matrices(2, 0.0);
    test_matrices(4, 0.1);
    test_matrices(8, 0.1);
    test_zphi_derivatives();
    test_rhs_phi();
    test_stepper();
    test_rk45();
    test_augmented();
    test_timed_drive();
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL PASSED", failures);
    return failures ? 1 : 0;
}
