// cusuperhelium_compat.cuh -- the reference's C++ names for the Roberts BIE RK4 path, as thin shims over the C ABI of
// libroberts_b200.so (include/roberts_b200.h).  Include this instead of the reference's BoundaryProblem.cuh /
// BaseBoundaryIntegrator.cuh / AutonomousRungeKuttaStepper.cuh / Derivatives.cuh / createM.cuh / WaterVelocities.cuh and link
// with -lroberts_b200.  Template parameters N / batchSize are kept for source compatibility; underneath N is a runtime value.
//
// Reference declarations mirrored (L/ = CuSuperHelium/CuSuperHelium/):
//   std_complex                         L/constants.cuh:12
//   ProblemProperties                   L/ProblemProperties.hpp:5-32
//   AutonomousProblem<T,N>              L/AutonomousProblem.h:9-28
//   BoundaryProblem<N,B> + Water / Helium / HeliumInfiniteDepth   L/BoundaryProblem.cuh:24-62, L/WaterBoundaryProblem.cuh,
//                                       L/HeliumBoundaryProblem.cuh
//   BaseBoundaryIntegralCalculator<N,B> L/BaseBoundaryIntegrator.cuh:10-85
//   ZPhiDerivative<N,B>, FftDerivative<N,B>   L/Derivatives.cuh:51-108
//   AutonomousRungeKuttaStepper<T,N>, RK4Options, OdeSolverResult  L/AutonomousRungeKuttaStepper.cuh:24-89, L/RK4Options.h, L/OdeSolver.h
//   TrajectoryLogger<T,N>               L/TrajectoryLogger.cuh:7-73
//   RK45_Options, RK45_std_complex<N>, RK45StepResult   L/RK45.cuh:21-27, 95-180, 333-400
//   createMKernel, createFiniteDepthMKernel, createVelocityMatrices, createHeliumVelocityMatrices,
//   compute_rhs_phi_expression, compute_rhs_helium_phi_expression   (the __global__ kernels the reference's tests launch)
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <cuda/std/complex>

#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "roberts_b200.h"
#include "roberts_b200_device.cuh"

typedef cuda::std::complex<double> std_complex;
constexpr double PI_d = 3.14159265358979323846;

struct ProblemProperties {
    double L = 1.0;
    double rho = 1.0;
    double U = 0.0;
    double kappa = 0.0;
    double depth = 1.0;
    double initial_amplitude = 1.0;
    double y_min = 0, y_max = 0;
    bool use_expansions = false;
    int expansion_order = 1;
    bool infinite_depth = false;
    double base_length = 1.0, base_time = 1.0, base_energy = 1.0, base_acceleration = 1.0;
};

struct RK4Options {
    double initial_timestep = 1e-2;
    bool returnTrajectory = true;
};

enum class OdeSolverResult { ReachedEndTime, StiffnessDetected };

inline void rb_compat_check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + rb_last_error());
}

template <typename T, int N>
class AutonomousProblem {
public:
    virtual ~AutonomousProblem() {}
    virtual void run(T* initialState, T* rhs) = 0;
    virtual void setStream(cudaStream_t stream) = 0;
};

// ---- physics plugins: here they only select which kernels the solver uses -----------------------------------------------
template <int N, size_t batchSize>
class BoundaryProblem {
public:
    virtual ~BoundaryProblem() {}
    virtual int physics() const = 0;
};
template <int N, size_t batchSize>
class WaterBoundaryProblem : public BoundaryProblem<N, batchSize> {
public:
    explicit WaterBoundaryProblem(ProblemProperties&) {}
    int physics() const override { return RB_WATER; }
};
template <int N, size_t batchSize>
class HeliumBoundaryProblem : public BoundaryProblem<N, batchSize> {
public:
    explicit HeliumBoundaryProblem(ProblemProperties&) {}
    int physics() const override { return RB_HELIUM; }
};
template <int N, size_t batchSize>
class HeliumInfiniteDepthBoundaryProblem : public BoundaryProblem<N, batchSize> {
public:
    explicit HeliumInfiniteDepthBoundaryProblem(ProblemProperties&) {}
    int physics() const override { return RB_HELIUM_INF; }
};

// ---- optomechanically driven film (L/OptomechanicalVariables.h:3-28, L/HeliumDrivenAutonomousProblem.cuh:10-26) ------------------
struct OptomechanicalVariables {
    double detuning = 0.0;
    double gamma = 1.0;
    double G = 1.0;
    double Tau = 1.0;
    double max_intensity = 0.0;
    double initial_time = 0.0;
    double location_x0_mode = 0.0;
    double sigma_optical_mode = 1.0;
    double Beta = 0.0;
    double DampingStrength = 0.01;
};
inline rb_opto rb_compat_opto(const OptomechanicalVariables& v, const ProblemProperties& p) {
    rb_opto o;
    o.detuning = v.detuning; o.gamma = v.gamma; o.G = v.G; o.Tau = v.Tau; o.max_intensity = v.max_intensity;
    o.initial_time = v.initial_time; o.location_x0_mode = v.location_x0_mode; o.sigma_optical_mode = v.sigma_optical_mode;
    o.Beta = v.Beta; o.DampingStrength = v.DampingStrength;
    o.drive_strength = rb_opto_drive_strength(&o, p.base_energy, p.base_time, p.rho);   // L/LightIntensity.cuh:30-33
    return o;
}
template <int N, size_t batchSize>
class HeliumDrivenAutonomousProblem : public HeliumBoundaryProblem<N, batchSize> {
public:
    ProblemProperties& properties;
    OptomechanicalVariables& variables;
    HeliumDrivenAutonomousProblem(ProblemProperties& p, OptomechanicalVariables& v)
        : HeliumBoundaryProblem<N, batchSize>(p), properties(p), variables(v) {}
};

// energy handles with the reference's getEnergy() (L/Energies.cuh:229-235)
class RbEnergyView {
    rb_solver* s_;
    int which_;
public:
    RbEnergyView(rb_solver* s = nullptr, int which = 0) : s_(s), which_(which) {}
    double getEnergy() const {
        double e[5];
        rb_compat_check(rb_energies(s_, e), "rb_energies");
        return e[which_];
    }
};

template <int N, size_t batchSize>
class BaseBoundaryIntegralCalculator : public AutonomousProblem<std_complex, 2 * N * (int)batchSize> {
    rb_solver* s_ = nullptr;
public:
    RbEnergyView kineticEnergy, potentialEnergy, surfaceEnergy, volumeFlux;
    std_complex* devVelocitiesUpper = nullptr;
    double* devPhiPrime = nullptr;

    BaseBoundaryIntegralCalculator(ProblemProperties& p, BoundaryProblem<N, batchSize>& problem, int guess_mode = RB_GUESS_WARM) {
        rb_props rp;
        rb_default_props(&rp);
        rp.rho = p.rho; rp.U = p.U; rp.kappa = p.kappa; rp.depth = p.depth;
        rp.use_expansions = p.use_expansions; rp.expansion_order = p.expansion_order; rp.infinite_depth = p.infinite_depth;
        rp.physics = problem.physics();
        rp.guess_mode = guess_mode;
        s_ = rb_create(N, (int)batchSize, &rp);
        if (!s_) throw std::runtime_error(std::string("rb_create: ") + rb_last_error());
        kineticEnergy = RbEnergyView(s_, 0); potentialEnergy = RbEnergyView(s_, 1); surfaceEnergy = RbEnergyView(s_, 2);
        volumeFlux = RbEnergyView(s_, 3);
        devVelocitiesUpper = reinterpret_cast<std_complex*>(rb_dev_velocities_upper(s_));
        devPhiPrime = rb_dev_phi_prime(s_);
    }
    ~BaseBoundaryIntegralCalculator() override { rb_destroy(s_); }
    BaseBoundaryIntegralCalculator(const BaseBoundaryIntegralCalculator&) = delete;
    BaseBoundaryIntegralCalculator& operator=(const BaseBoundaryIntegralCalculator&) = delete;

    void runTimeStep(const std_complex* initialState, std_complex* rhs) {
        rb_compat_check(rb_rhs(s_, reinterpret_cast<const rb_complex*>(initialState), reinterpret_cast<rb_complex*>(rhs)), "rb_rhs");
    }
    void run(std_complex* initialState, std_complex* rhs) override { runTimeStep(initialState, rhs); }
    void calculateVorticities(const std_complex* initialState) {
        rb_compat_check(rb_vorticities(s_, reinterpret_cast<const rb_complex*>(initialState)), "rb_vorticities");
    }
    double* getDevA() { return rb_dev_a(s_); }
    std_complex* getDevZp() { return reinterpret_cast<std_complex*>(rb_dev_zp(s_)); }
    std_complex* getDevZpp() { return reinterpret_cast<std_complex*>(rb_dev_zpp(s_)); }
    void setStream(cudaStream_t stream) override { rb_compat_check(rb_set_stream(s_, stream), "rb_set_stream"); }
    rb_solver* handle() { return s_; }
};

template <int N, size_t batchSize>
class ZPhiDerivative {
    rb_solver* s_ = nullptr;
public:
    explicit ZPhiDerivative(ProblemProperties& p) {
        rb_props rp;
        rb_default_props(&rp);
        rp.rho = p.rho; rp.U = p.U;
        s_ = rb_create(N, (int)batchSize, &rp);
        if (!s_) throw std::runtime_error(std::string("rb_create: ") + rb_last_error());
    }
    ~ZPhiDerivative() { rb_destroy(s_); }
    void exec(const std_complex* Z, const std_complex* Phi, std_complex* ZPrime, std_complex* PhiPrime, std_complex* Zpp) {
        rb_compat_check(rb_zphi_derivative(s_, (const rb_complex*)Z, (const rb_complex*)Phi, (rb_complex*)ZPrime,
                                           (rb_complex*)PhiPrime, (rb_complex*)Zpp), "rb_zphi_derivative");
        rb_synchronize(s_);
    }
};

template <int N, int batchSize>
class FftDerivative {
    rb_solver* s_ = nullptr;
public:
    FftDerivative() {}
    ~FftDerivative() { if (s_) rb_destroy(s_); }
    cudaError_t initialize(bool = false) {
        rb_props rp;
        rb_default_props(&rp);
        s_ = rb_create(N, batchSize, &rp);
        return s_ ? cudaSuccess : cudaErrorUnknown;
    }
    void exec(const std_complex* in, std_complex* out, const bool doubleDev = false, double scaling = 1.0, bool = false) {
        if (!s_) throw std::runtime_error("The FFT class wasn't initialized!");
        rb_compat_check(rb_fft_derivative(s_, (const rb_complex*)in, (rb_complex*)out, doubleDev, scaling), "rb_fft_derivative");
        rb_synchronize(s_);
    }
};

// ---- trajectory logger + stepper ------------------------------------------------------------------------------------------
// DelayedIntensityIntegrator<N,B> (L/DelayedIntensityIntegrator.cuh:9-39) and AugmentedBoundaryIntegrator<N,B>
// (L/AugmentedBoundaryIntegrator.cuh:10-40): y = [Z | Phi | D] -> [w | dPhi/dt | dD/dt] through rb_augmented_rhs
template <int N, size_t batchSize>
class DelayedIntensityIntegrator {
public:
    OptomechanicalVariables& variables;
    explicit DelayedIntensityIntegrator(OptomechanicalVariables& v) : variables(v) {}
};
template <int N, size_t batchSize>
class AugmentedBoundaryIntegrator : public AutonomousProblem<std_complex, 3 * N * (int)batchSize> {
    std::unique_ptr<BaseBoundaryIntegralCalculator<N, batchSize>> integrator_;
    std::unique_ptr<DelayedIntensityIntegrator<N, batchSize>> delayed_;
    ProblemProperties props_;
public:
    // the properties are those the calculator was built with (the reference's driven problem keeps a reference to them)
    AugmentedBoundaryIntegrator(std::unique_ptr<BaseBoundaryIntegralCalculator<N, batchSize>> integrator,
                                std::unique_ptr<DelayedIntensityIntegrator<N, batchSize>> delayedIntegrator,
                                const ProblemProperties& properties = ProblemProperties())
        : integrator_(std::move(integrator)), delayed_(std::move(delayedIntegrator)), props_(properties) {}
    rb_opto opto() const { return rb_compat_opto(delayed_->variables, props_); }
    void run(std_complex* initialState, std_complex* rhs) override {
        rb_opto o = opto();
        rb_compat_check(rb_augmented_rhs(integrator_->handle(), &o, reinterpret_cast<const rb_complex*>(initialState),
                                         reinterpret_cast<rb_complex*>(rhs)), "rb_augmented_rhs");
    }
    void setStream(cudaStream_t stream) override { integrator_->setStream(stream); }
    rb_solver* handle() { return integrator_->handle(); }
};

template <typename T, int N>
class TrajectoryLogger {
public:
    size_t every, capacity;
    rb_stepper* st = nullptr;
    explicit TrajectoryLogger(size_t log_every = 1, size_t max_states = 1024) : every(log_every), capacity(max_states) {}
    void copyTimesToHost(double** times, size_t* count) {
        rb_complex* states = nullptr; size_t ns = 0;
        rb_compat_check(rb_rk4_copy_trajectory(st, times, count, &states, &ns), "rb_rk4_copy_trajectory");
        rb_free(states);
    }
    void copyStatesToHost(T** states, size_t* count) {
        double* times = nullptr; size_t nt = 0;
        rb_compat_check(rb_rk4_copy_trajectory(st, &times, &nt, reinterpret_cast<rb_complex**>(states), count),
                        "rb_rk4_copy_trajectory");
        rb_free(times);
    }
};

template <typename T, int N>
class AutonomousRungeKuttaStepper {
    rb_stepper* st_ = nullptr;
    rb_aug_stepper* aug_ = nullptr;   // set instead of st_ when the problem is an AugmentedBoundaryIntegrator (3 N B state)
    std::shared_ptr<TrajectoryLogger<T, N>> logger_;
public:
    template <int NP, size_t B>
    AutonomousRungeKuttaStepper(AugmentedBoundaryIntegrator<NP, B>& problem, double tstep = 1e-2) {
        static_assert(3 * NP * (int)B == N, "state size must be 3 * N * batchSize");
        rb_opto o = problem.opto();
        aug_ = rb_aug_rk4_create(problem.handle(), &o, tstep);
        if (!aug_) throw std::runtime_error(std::string("rb_aug_rk4_create: ") + rb_last_error());
    }
    // the problem must be a BaseBoundaryIntegralCalculator<N/2/B, B>; its solver handle is what the stepper binds to
    template <int NP, size_t B>
    AutonomousRungeKuttaStepper(BaseBoundaryIntegralCalculator<NP, B>& problem, double tstep = 1e-2,
                                std::shared_ptr<TrajectoryLogger<T, N>> logger = nullptr)
        : logger_(logger) {
        static_assert(2 * NP * (int)B == N, "state size must be 2 * N * batchSize");
        st_ = rb_rk4_create(problem.handle(), tstep);
        if (!st_) throw std::runtime_error(std::string("rb_rk4_create: ") + rb_last_error());
        if (logger_) {
            logger_->st = st_;
            rb_compat_check(rb_rk4_set_logging(st_, logger_->every, logger_->capacity), "rb_rk4_set_logging");
        }
    }
    ~AutonomousRungeKuttaStepper() {
        if (st_) rb_rk4_destroy(st_);
        if (aug_) rb_aug_rk4_destroy(aug_);
    }
    void setTimeStep(double tstep) {
        if (aug_) rb_compat_check(rb_aug_rk4_set_time_step(aug_, tstep), "rb_aug_rk4_set_time_step");
        else rb_compat_check(rb_rk4_set_time_step(st_, tstep), "rb_rk4_set_time_step");
    }
    void setOptions(const RK4Options& o) { setTimeStep(o.initial_timestep); }
    void initialize(T* devY0, bool onDevice = false) {
        if (aug_) {
            rb_compat_check(rb_aug_rk4_initialize(aug_, reinterpret_cast<rb_complex*>(devY0), onDevice), "rb_aug_rk4_initialize");
            return;
        }
        rb_compat_check(rb_rk4_initialize(st_, reinterpret_cast<rb_complex*>(devY0), onDevice), "rb_rk4_initialize");
        if (logger_) rb_compat_check(rb_rk4_set_logging(st_, logger_->every, logger_->capacity), "rb_rk4_set_logging");
    }
    void runStep(int = 0) {
        if (aug_) rb_compat_check(rb_aug_rk4_step(aug_), "rb_aug_rk4_step");
        else rb_compat_check(rb_rk4_step(st_), "rb_rk4_step");
    }
    OdeSolverResult runEvolution(double startTime, double endTime) {
        size_t n = 0;
        if (aug_) rb_compat_check(rb_aug_rk4_evolve(aug_, startTime, endTime, &n), "rb_aug_rk4_evolve");
        else rb_compat_check(rb_rk4_evolve(st_, startTime, endTime, &n), "rb_rk4_evolve");
        return OdeSolverResult::ReachedEndTime;
    }
    void getState(T* host) {
        if (aug_) rb_compat_check(rb_aug_rk4_get_state(aug_, reinterpret_cast<rb_complex*>(host)), "rb_aug_rk4_get_state");
        else rb_compat_check(rb_rk4_get_state(st_, reinterpret_cast<rb_complex*>(host)), "rb_rk4_get_state");
    }
};

// ---- the explicitly time-dependent drive: TimedProblem (L/AutonomousProblem.h:30-51), TimedBoundaryProblem
//      (L/TimedBoundaryProblem.cuh:6-19), HeliumWithOptomechanicalDrivingProblem<N> (L/HeliumWithDrivingBoundaryProblem.cuh:11-67),
//      TimedBoundaryIntegrator<N,B> (L/TimedBoundaryIntegrator.cuh:7-49) and RungeKuttaStepper<std_complex, 2N>
//      (L/RK4_Time_Dependent.cuh:18-460), used as in A/kernel.cu:281-366 and L/Export.cu:797-826 ----------------------------------
template <typename T, int N>
class TimedProblem {
protected:
    bool saveProgress = true;
    double currentTime = 0.0;
public:
    virtual ~TimedProblem() {}
    virtual void run(T* initialState, T* rhs) = 0;
    void setCurrentTime(double time) { currentTime = time; }
    virtual void setStartingTime(double time) { currentTime = time; }
    void setSaveProgress(bool save) { saveProgress = save; }
    virtual void setStream(cudaStream_t stream) = 0;
};
template <int N, size_t batchSize>
class TimedBoundaryProblem : public HeliumBoundaryProblem<N, batchSize> {
public:
    ProblemProperties& properties;
    OptomechanicalVariables variables;
    TimedBoundaryProblem(ProblemProperties& p, OptomechanicalVariables v)
        : HeliumBoundaryProblem<N, batchSize>(p), properties(p), variables(v) {}
};
template <int N>
class HeliumWithOptomechanicalDrivingProblem : public TimedBoundaryProblem<N, 1> {
public:
    HeliumWithOptomechanicalDrivingProblem(ProblemProperties& p, OptomechanicalVariables v) : TimedBoundaryProblem<N, 1>(p, v) {}
};
// The delayed-intensity term (the reference keeps it in the boundary problem, L/DelayedIntensityTerm.cuh) lives in the library's
// timed stepper object; the integrator owns that object and the RungeKuttaStepper below drives the same one.
template <int N, size_t batchSize>
class TimedBoundaryIntegrator : public BaseBoundaryIntegralCalculator<N, batchSize>, public TimedProblem<std_complex, 2 * N * (int)batchSize> {
    rb_timed_stepper* timed_ = nullptr;
public:
    TimedBoundaryIntegrator(ProblemProperties& p, TimedBoundaryProblem<N, batchSize>& problem)
        : BaseBoundaryIntegralCalculator<N, batchSize>(p, problem) {
        rb_opto o = rb_compat_opto(problem.variables, p);
        timed_ = rb_timed_rk4_create(this->handle(), &o, 1e-2);
        if (!timed_) throw std::runtime_error(std::string("rb_timed_rk4_create: ") + rb_last_error());
    }
    ~TimedBoundaryIntegrator() override { rb_timed_rk4_destroy(timed_); }
    void run(std_complex* initialState, std_complex* rhs) override {
        rb_compat_check(rb_timed_rhs(timed_, this->currentTime, this->saveProgress, reinterpret_cast<const rb_complex*>(initialState),
                                     reinterpret_cast<rb_complex*>(rhs)), "rb_timed_rhs");
    }
    void setStream(cudaStream_t stream) override { BaseBoundaryIntegralCalculator<N, batchSize>::setStream(stream); }
    void setStartingTime(double time) override {
        this->currentTime = time;
        rb_compat_check(rb_timed_rk4_set_starting_time(timed_, time), "rb_timed_rk4_set_starting_time");
    }
    double* delayedIntensity() { return rb_timed_rk4_dev_delayed_intensity(timed_); }
    rb_timed_stepper* timedHandle() { return timed_; }
};

template <typename T, int N>
class RungeKuttaStepper {
    rb_timed_stepper* st_ = nullptr;   // owned by the TimedBoundaryIntegrator
    RK4Options options;
public:
    template <int NP, size_t B>
    RungeKuttaStepper(TimedBoundaryIntegrator<NP, B>& timedProblem, double tstep = 1e-2) : st_(timedProblem.timedHandle()) {
        static_assert(2 * NP * (int)B == N, "state size must be 2 * N * batchSize");
        setTimeStep(tstep);
    }
    void setTimeStep(double tstep) { rb_compat_check(rb_timed_rk4_set_time_step(st_, tstep), "rb_timed_rk4_set_time_step"); }
    void setOptions(const RK4Options& o) {
        setTimeStep(o.initial_timestep);
        options = o;
        rb_compat_check(rb_timed_rk4_set_logging(st_, o.returnTrajectory), "rb_timed_rk4_set_logging");
    }
    void initialize(T* devY0, bool onDevice = false) {
        rb_compat_check(rb_timed_rk4_initialize(st_, reinterpret_cast<rb_complex*>(devY0), onDevice), "rb_timed_rk4_initialize");
    }
    void setCurrentStream(cudaStream_t) {}   // L/RK4_Time_Dependent.cuh:41-44: the stepper works on the stream of its integrator
    void runStep(int = 0) { rb_compat_check(rb_timed_rk4_step(st_, 0), "rb_timed_rk4_step"); }
    OdeSolverResult runEvolution(double startTime, double endTime) {
        size_t n = 0;
        rb_compat_check(rb_timed_rk4_evolve(st_, startTime, endTime, &n), "rb_timed_rk4_evolve");
        return OdeSolverResult::ReachedEndTime;
    }
    int copyTimesToHost(double** hostTimes, size_t* countHost) {
        rb_complex* states = nullptr; size_t ns = 0;
        if (rb_timed_rk4_copy_trajectory(st_, hostTimes, countHost, &states, &ns) != 0) return -1;
        rb_free(states);
        return 0;
    }
    int copyStatesToHost(T** hostStates, size_t* countHost) {
        double* times = nullptr; size_t nt = 0;
        if (rb_timed_rk4_copy_trajectory(st_, &times, &nt, reinterpret_cast<rb_complex**>(hostStates), countHost) != 0) return -1;
        rb_free(times);
        return 0;
    }
    void getState(T* host) { rb_compat_check(rb_timed_rk4_get_state(st_, reinterpret_cast<rb_complex*>(host)), "rb_timed_rk4_get_state"); }
    double getCurrentTime() { return rb_timed_rk4_current_time(st_); }
};

// ---- implicit side: RealBoundaryItegralCalculator<N> (L/RealBoundaryIntegralCalculator.cuh:37-89), JacobianCalculator<N>
//      (L/JacobianCalculator.cuh:168-284), GaussLegendre2Options / GaussLegendre2<N> (L/GaussLegendre.cuh:70-612), assembled as
//      L/Export.cu:680-700 ------------------------------------------------------------------------------------------------
template <size_t N>
class RealBoundaryItegralCalculator final : public AutonomousProblem<double, 3 * (int)N> {
    BaseBoundaryIntegralCalculator<(int)N, 1>& calculator_;
public:
    explicit RealBoundaryItegralCalculator(BaseBoundaryIntegralCalculator<(int)N, 1>& boundaryIntegralCalculator)
        : calculator_(boundaryIntegralCalculator) {}
    void run(double* initialState, double* rhs) override {
        rb_compat_check(rb_real_rhs(calculator_.handle(), initialState, rhs), "rb_real_rhs");
    }
    void setStream(cudaStream_t stream) override { calculator_.setStream(stream); }
    rb_solver* handle() { return calculator_.handle(); }
};

// The reference hands the calculator a BaseBoundaryIntegralCalculator<N, 3N> it built itself; here the batch-3N assembler is
// built inside from the same two arguments that one would have been built from.
template <size_t N>
class JacobianCalculator final {
    rb_jacobian* j_ = nullptr;
public:
    JacobianCalculator(ProblemProperties& p, BoundaryProblem<(int)N, 3 * N>& problem) {
        rb_props rp;
        rb_default_props(&rp);
        rp.rho = p.rho; rp.U = p.U; rp.kappa = p.kappa; rp.depth = p.depth;
        rp.use_expansions = p.use_expansions; rp.expansion_order = p.expansion_order; rp.infinite_depth = p.infinite_depth;
        rp.physics = problem.physics();
        j_ = rb_jacobian_create((int)N, &rp);
        if (!j_) throw std::runtime_error(std::string("rb_jacobian_create: ") + rb_last_error());
    }
    ~JacobianCalculator() { rb_jacobian_destroy(j_); }
    JacobianCalculator(const JacobianCalculator&) = delete;
    JacobianCalculator& operator=(const JacobianCalculator&) = delete;
    void setEpsilon(double eps) { rb_compat_check(rb_jacobian_set_epsilon(j_, eps), "rb_jacobian_set_epsilon"); }
    void setStream(cudaStream_t stream) { rb_compat_check(rb_jacobian_set_stream(j_, stream), "rb_jacobian_set_stream"); }
    void calculateJacobian(const double* devState, double* devJacobian) {
        rb_compat_check(rb_jacobian_calculate(j_, devState, devJacobian), "rb_jacobian_calculate");
    }
    rb_jacobian* handle() { return j_; }
};

struct GaussLegendre2Options {
    double stepSize = 0.01;
    double newtonTolerance = 1e-10;
    size_t maxNewtonIterations = 20;
    bool allowSimplifiedFallback = false;
    bool returnTrajectory = true;
    double armijo_c = 1e-4;
    double backtrack = 0.5;
    double minAlpha = 1e-6;
    size_t maxStepsHalves = 6;
};

template <size_t N>
class GaussLegendre2 {
    rb_gl2* g_ = nullptr;
public:
    GaussLegendre2(RealBoundaryItegralCalculator<N>& problem, JacobianCalculator<N>& jacobianCalculator,
                   GaussLegendre2Options options = GaussLegendre2Options()) {
        rb_gl2_options o;
        o.stepSize = options.stepSize; o.newtonTolerance = options.newtonTolerance; o.maxNewtonIterations = options.maxNewtonIterations;
        o.allowSimplifiedFallback = options.allowSimplifiedFallback; o.returnTrajectory = options.returnTrajectory;
        o.armijo_c = options.armijo_c; o.backtrack = options.backtrack; o.minAlpha = options.minAlpha;
        o.maxStepsHalves = options.maxStepsHalves;
        g_ = rb_gl2_create(problem.handle(), jacobianCalculator.handle(), &o);
        if (!g_) throw std::runtime_error(std::string("rb_gl2_create: ") + rb_last_error());
    }
    ~GaussLegendre2() { rb_gl2_destroy(g_); }
    GaussLegendre2(const GaussLegendre2&) = delete;
    GaussLegendre2& operator=(const GaussLegendre2&) = delete;
    void setStream(cudaStream_t) {}   // the integrator works on the stream of its RHS assembler
    void initialize(double* initialState, bool onDevice = false) {
        rb_compat_check(rb_gl2_initialize(g_, initialState, onDevice), "rb_gl2_initialize");
    }
    OdeSolverResult runEvolution(double startTime, double endTime) {
        rb_compat_check(rb_gl2_evolve(g_, startTime, endTime), "rb_gl2_evolve");   // throws where the reference throws
        return OdeSolverResult::ReachedEndTime;
    }
    int copyTimesToHost(double** hostTimes, size_t* countHost) {
        double* states = nullptr; size_t ns = 0;
        if (rb_gl2_copy_trajectory(g_, hostTimes, countHost, &states, &ns) != 0) return -1;
        rb_free(states);
        return 0;
    }
    int copyStatesToHost(double** hostStates, size_t* countHost) {
        double* times = nullptr; size_t nt = 0;
        if (rb_gl2_copy_trajectory(g_, &times, &nt, hostStates, countHost) != 0) return -1;
        rb_free(times);
        return 0;
    }
    rb_gl2* handle() { return g_; }
};

// ---- adaptive RKF45 (L/RK45.cuh): RK45_Options and RK45_std_complex<N> over any AutonomousProblem<std_complex, N> ---------------
struct RK45_Options {   // L/RK45.cuh:21-27
    double atol = 1e-6;
    double rtol = 1e-3;
    double h_min = 1e-16;
    double h_max = 1e10;
    double initial_timestep = 1e-2;
};
enum class RK45StepResult { StepAccepted, StepRejected };

template <size_t N>
class RK45_std_complex {
    rb_rk45* r_ = nullptr;
    AutonomousProblem<std_complex, (int)N>* problem_ = nullptr;
    static void trampoline(void* user, const rb_complex* state, rb_complex* rhs, void* stream) {
        auto* p = static_cast<AutonomousProblem<std_complex, (int)N>*>(user);
        p->setStream(static_cast<cudaStream_t>(stream));
        p->run(reinterpret_cast<std_complex*>(const_cast<rb_complex*>(state)), reinterpret_cast<std_complex*>(rhs));
    }
    static rb_rk45_options make(double tstep, double h_max, double h_min) {
        rb_rk45_options o{1e-6, 1e-3, h_min, h_max, tstep};
        return o;
    }
public:
    // the reference's ctor also takes a DataLogger and value loggers (L/RK45.cuh:106); logging is outside this path
    explicit RK45_std_complex(AutonomousProblem<std_complex, (int)N>& problem, double tstep = 1e-2, double h_max = 1e10,
                              double h_min = 1e-16, cudaStream_t stream = nullptr)
        : problem_(&problem) {
        rb_rk45_options o = make(tstep, h_max, h_min);
        r_ = rb_rk45_create_generic(N, &RK45_std_complex::trampoline, problem_, &o, stream);
        if (!r_) throw std::runtime_error(std::string("rb_rk45_create_generic: ") + rb_last_error());
    }
    // the boundary-integral RHS runs inside the library (no callback)
    template <int NP, size_t B>
    explicit RK45_std_complex(BaseBoundaryIntegralCalculator<NP, B>& problem, double tstep = 1e-2, double h_max = 1e10,
                              double h_min = 1e-16) {
        static_assert(2 * NP * B == N, "state size must be 2 * N * batchSize");
        rb_rk45_options o = make(tstep, h_max, h_min);
        r_ = rb_rk45_create(problem.handle(), &o);
        if (!r_) throw std::runtime_error(std::string("rb_rk45_create: ") + rb_last_error());
    }
    ~RK45_std_complex() { rb_rk45_destroy(r_); }
    RK45_std_complex(const RK45_std_complex&) = delete;
    RK45_std_complex& operator=(const RK45_std_complex&) = delete;
    void setTolerance(double atol, double rtol) { rb_compat_check(rb_rk45_set_tolerance(r_, atol, rtol), "rb_rk45_set_tolerance"); }
    void setOptions(const RK45_Options& o) {
        rb_rk45_options c{o.atol, o.rtol, o.h_min, o.h_max, o.initial_timestep};
        rb_compat_check(rb_rk45_set_options(r_, &c), "rb_rk45_set_options");
    }
    void setMaxRejectedSteps(size_t m) { rb_compat_check(rb_rk45_set_max_rejected(r_, m), "rb_rk45_set_max_rejected"); }
    void initialize(std_complex* initialState, bool onDevice = false) {
        rb_compat_check(rb_rk45_initialize(r_, reinterpret_cast<rb_complex*>(initialState), onDevice), "rb_rk45_initialize");
    }
    RK45StepResult runStep(int = 0) {
        int acc = 0;
        rb_compat_check(rb_rk45_step(r_, &acc), "rb_rk45_step");
        return acc ? RK45StepResult::StepAccepted : RK45StepResult::StepRejected;
    }
    OdeSolverResult runEvolution(double startTime, double endTime) {
        int res = 0;
        rb_compat_check(rb_rk45_evolve(r_, startTime, endTime, &res), "rb_rk45_evolve");
        return res == 0 ? OdeSolverResult::ReachedEndTime : OdeSolverResult::StiffnessDetected;
    }
    std_complex* getY() { return reinterpret_cast<std_complex*>(rb_rk45_dev_state(r_)); }
    double getCurrentTime() const { return rb_rk45_current_time(r_); }
};

// ---- the kernels the reference's tests launch with <<< >>> (same names, same signatures) ----------------------------------
__global__ void createMKernel(double* A, const std_complex* const Z, const std_complex* const Zp, const std_complex* const Zpp,
                              double rho, int n, size_t batchSize) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    size_t b = blockIdx.z;
    if (b >= batchSize || k >= n || j >= n) return;
    A[(size_t)k + (size_t)j * n + b * (size_t)n * n] = rb_dev::M_entry<false>(
        k, j, reinterpret_cast<const double2*>(Z) + b * n, reinterpret_cast<const double2*>(Zp) + b * n,
        reinterpret_cast<const double2*>(Zpp) + b * n, rho, 0.0, true);
}

__global__ void createFiniteDepthMKernel(double* A, const std_complex* const Z, const std_complex* const Zp,
                                         const std_complex* const Zpp, double h, int n, size_t batchSize,
                                         bool infinite_depth = false) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    size_t b = blockIdx.z;
    if (b >= batchSize || k >= n || j >= n) return;
    A[(size_t)k + (size_t)j * n + b * (size_t)n * n] = rb_dev::M_entry<true>(
        k, j, reinterpret_cast<const double2*>(Z) + b * n, reinterpret_cast<const double2*>(Zp) + b * n,
        reinterpret_cast<const double2*>(Zpp) + b * n, 0.0, h, infinite_depth);
}

__global__ void createVelocityMatrices(const std_complex* Z, const std_complex* Zp, const std_complex* Zpp, int N,
                                       std_complex* out1, std_complex* out2, bool lower = true, size_t batchSize = 1) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    size_t b = blockIdx.z;
    if (b >= batchSize || k >= N || j >= N) return;
    double2 v = rb_dev::V1_entry(k, j, reinterpret_cast<const double2*>(Z) + b * N, reinterpret_cast<const double2*>(Zp) + b * N,
                                 reinterpret_cast<const double2*>(Zpp) + b * N, reinterpret_cast<double2*>(out2) + b * N, lower,
                                 false, 0.0, true);
    reinterpret_cast<double2*>(out1)[(size_t)k + (size_t)j * N + b * (size_t)N * N] = v;
}

__global__ void createHeliumVelocityMatrices(const std_complex* const Z, const std_complex* const Zp,
                                             const std_complex* const Zpp, double h, int N, std_complex* const out1,
                                             std_complex* const out2, bool lower, size_t batchSize,
                                             bool infinite_depth = false) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    size_t b = blockIdx.z;
    if (b >= batchSize || k >= N || j >= N) return;
    double2 v = rb_dev::V1_entry(k, j, reinterpret_cast<const double2*>(Z) + b * N, reinterpret_cast<const double2*>(Zp) + b * N,
                                 reinterpret_cast<const double2*>(Zpp) + b * N, reinterpret_cast<double2*>(out2) + b * N, lower,
                                 true, h, infinite_depth);
    reinterpret_cast<double2*>(out1)[(size_t)k + (size_t)j * N + b * (size_t)N * N] = v;
}

__global__ void compute_rhs_phi_expression(const std_complex* Z, const std_complex* V1, const std_complex* V2,
                                           std_complex* result, double rho, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        double Z_imag = Z[i].imag();
        double V1_abs2 = V1[i].real() * V1[i].real() + V1[i].imag() * V1[i].imag();
        double V2_abs2 = V2[i].real() * V2[i].real() + V2[i].imag() * V2[i].imag();
        double V1_dot_V2 = V1[1].real() * V2[i].real() + V1[i].imag() * V2[i].imag();   // index 1: as in L/createM.cuh:104
        result[i] = -(1 + rho) * Z_imag + 0.5 * V1_abs2 + 0.5 * rho * V2_abs2 - rho * V1_dot_V2;
    }
}

__global__ void compute_rhs_helium_phi_expression(const std_complex* Z, const std_complex* V1, std_complex* result, double h,
                                                  int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        double vdw = h / 3.0;
        result[i] = vdw * pow(1.0 + Z[i].imag() / h, -3.0) - vdw + 0.5 * V1[i].real() * V1[i].real() +
                    0.5 * V1[i].imag() * V1[i].imag();
    }
}

// ---- the perturbed-state kernels of the Jacobian (T/MatrixMTests.cuh:284-390 launches them as
//      createInitialState<<<N, 1>>> and createInitialBatchedZ<<<(ceil(2N/256), 3N), 256>>>; L/JacobianCalculator.cuh:11-166) ----------
__global__ void createInitialState(const double* initialState, std_complex* complexState, size_t N) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    complexState[i] = std_complex(initialState[i], initialState[i + N]);
    complexState[i + N] = std_complex(initialState[i + 2 * N], 0.0);
}

// one thread per entry of the 2N-entry state and per batch member b = blockIdx.y = c N + j: a copy of the state with coordinate
// c (0 x, 1 y, 2 phi) of point j moved by eps, written as [Z of member 0 .. Z of member 3N-1 | Phi of member 0 ..]
__global__ void createInitialBatchedZ(const std_complex* __restrict__ initialState, std_complex* __restrict__ ZBatched, double eps,
                                      size_t N) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= 2 * N) return;
    const size_t b = blockIdx.y, c = b / N, j = b % N;
    const bool position = tid < N;
    const size_t i = position ? tid : tid - N;
    std_complex v = initialState[tid];
    if (i == j) {
        if (position && c == 0) v += std_complex(eps, 0.0);
        else if (position && c == 1) v += std_complex(0.0, eps);
        else if (!position && c == 2) v += std_complex(eps, 0.0);
    }
    ZBatched[(position ? 0 : 3 * N * N) + b * N + i] = v;
}

// pos / neg: batched RHS (6 N^2 complex) at +eps / -eps; C: 3N x 3N column-major, C[c * 3N + r] = d f_r / d y_c
__global__ void createJacobianMatrixFromPerturbedRhs(const std_complex* __restrict__ pos, const std_complex* __restrict__ neg,
                                                     double* __restrict__ C, size_t N, double eps) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * N * N) return;
    const std_complex d = (pos[i] - neg[i]) / (2.0 * eps);
    const size_t k = i / N, p = i % N;
    if (k < 3 * N) {
        C[k * 3 * N + p] = d.real();
        C[k * 3 * N + p + N] = d.imag();
    } else {
        C[(k - 3 * N) * 3 * N + p + 2 * N] = d.real();
    }
}

// ---- complex elementary functions the reference's tests call by name (T/ComplexFunctionsTests.cuh:8-158; L/utilities.cuh:59-63,
//      270-286, 312-369).  The cotangent is evaluated as (sin u cos u - i sinh v cosh v) / (sin^2 u + sinh^2 v): every term of the
//      denominator is non-negative, so there is no cancellation anywhere in the plane. ------------------------------------------------
__device__ inline void sin(cuDoubleComplex z, cuDoubleComplex& zout) {
    double s, c;
    sincos(z.x, &s, &c);
    zout.x = s * cosh(z.y);
    zout.y = c * sinh(z.y);
}
__device__ inline void cos(cuDoubleComplex z, cuDoubleComplex& out) {
    double s, c;
    sincos(z.x, &s, &c);
    out.x = c * cosh(z.y);
    out.y = -s * sinh(z.y);
}
__device__ inline cuDoubleComplex cotangent_complex(cuDoubleComplex a) {
    const double2 r = rb_dev::cot_half(2.0 * a.x, 2.0 * a.y);
    return make_cuDoubleComplex(r.x, r.y);
}
__global__ void cotangent_complex(const cuDoubleComplex* a, cuDoubleComplex* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cotangent_complex(a[i]);
}
__device__ __forceinline__ std_complex cot(std_complex z) {
    const double2 r = rb_dev::cot_half(2.0 * z.real(), 2.0 * z.imag());
    return std_complex(r.x, r.y);
}
// cot((Zk - Zj) / 2), the kernel of every operator on this path
__device__ inline std_complex cotangent_green_function(std_complex Zk, std_complex Zj) {
    const double2 r = rb_dev::cot_half(Zk.real() - Zj.real(), Zk.imag() - Zj.imag());
    return std_complex(r.x, r.y);
}

// ---- double-double helpers pinned by T/ComplexFunctionsTests.cuh:247-405 (L/PrecisionMath.cuh) -----------------------------------------
namespace PrecisionMath {
struct doubledouble {
    double hi, lo;
};
struct dd_complex {
    doubledouble real;
    doubledouble imag;
};
// a - b = hi + lo exactly (Knuth's branch-free two-sum applied to a and -b)
__host__ __device__ __forceinline__ void twoDiff(double a, double b, double& hi, double& lo) {
    hi = a - b;
    const double bb = a - hi;          // the part of b that was actually subtracted
    lo = (a - (hi + bb)) + (bb - b);
}
// a * b = hi + lo exactly
__device__ __forceinline__ void twoProd(double a, double b, double& hi, double& lo) {
    hi = a * b;
    lo = fma(a, b, -hi);
}
__device__ __forceinline__ dd_complex c_twoDiff(std_complex z1, std_complex z2) {
    dd_complex d;
    twoDiff(z1.real(), z2.real(), d.real.hi, d.real.lo);
    twoDiff(z1.imag(), z2.imag(), d.imag.hi, d.imag.lo);
    return d;
}
// 1 / (Z1 - Z2) with the difference and its squared modulus carried in double-double: conj(d) / |d|^2
__device__ inline std_complex fastPreciseInvSub(std_complex Z1, std_complex Z2) {
    const dd_complex d = c_twoDiff(Z1, Z2);
    double rh, rl, ih, il;
    twoProd(d.real.hi, d.real.hi, rh, rl);
    rl = fma(2.0 * d.real.hi, d.real.lo, rl);
    twoProd(d.imag.hi, d.imag.hi, ih, il);
    il = fma(2.0 * d.imag.hi, d.imag.lo, il);
    double sh, sl;
    twoDiff(rh, -ih, sh, sl);          // rh + ih = sh + sl
    sl += rl + il;
    const double den = sh + sl, den_lo = sl - (den - sh);
    double inv = 1.0 / den;
    inv = fma(inv, fma(-den, inv, 1.0) - den_lo * inv, inv);   // one Newton step against den + den_lo
    return std_complex((d.real.hi + d.real.lo) * inv, -(d.imag.hi + d.imag.lo) * inv);
}
}  // namespace PrecisionMath

