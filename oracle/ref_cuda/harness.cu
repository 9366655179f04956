// TEST INFRASTRUCTURE (oracle/): C entry points over the UNMODIFIED reference classes, compiled from the sources where they lie
// under /root/reference (see oracle/build_ref.py) into oracle/_ref/libcusuperhelium_ref.so.  Nothing in the product links or
// loads this; only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may.
//
// What is the reference's and what is ours: every class and kernel used below (BaseBoundaryIntegralCalculator,
// WaterBoundaryProblem, HeliumBoundaryProblem, HeliumInfiniteDepthBoundaryProblem, AutonomousRungeKuttaStepper, the energy
// classes) comes from the reference's headers; this file only instantiates them for a few N, moves host arrays in and out and
// reads a clock -- the way L/Export.cu:194-265 (calculateRHSNFromVectors) and T/ODESolverTests.cuh:95-110 do.  L/Export.cu
// itself is not compiled: it drags in OpenCV, HighFive/HDF5 (L/VideoMaking.h, L/SimulationRunner.cuh), absent from this image.
//
// Two accommodations for a conforming compiler (the reference is an MSVC project):
//  * -std=c++20 (std::same_as in L/cuDoubleComplexOperators.cuh) and the matplotlibcpp.h stand-in in shims/;
//  * L/BaseBoundaryIntegrator.cuh:89 names its base class as AutonomousProblem<cufftDoubleComplex, 2*N> while the class derives
//    from AutonomousProblem<std_complex, 2*N*batchSize>; MSVC accepts that, nvcc/EDG does not.  The header is included with
//    `cufftDoubleComplex` defined as `std_complex` for the span of that one #include (same size and layout; all its other uses in
//    that header are sizeof()), after cufft.h has already been seen.  batchSize = 1 only, where 2*N == 2*N*batchSize.
#include "BoundaryProblem.cuh"
#include "WaterVelocities.cuh"
#include "MatrixSolver.cuh"
#include "AutonomousProblem.h"
#define cufftDoubleComplex std_complex
#include "BaseBoundaryIntegrator.cuh"
#undef cufftDoubleComplex
#include "WaterBoundaryProblem.cuh"
#include "HeliumBoundaryProblem.cuh"
#include "VectorUtilities.cuh"   // appendToVector, used (not included) by L/TrajectoryLogger.cuh:71
#include "AutonomousRungeKuttaStepper.cuh"
#include "HeliumDrivenAutonomousProblem.cuh"      // the optomechanically driven film (what A/kernel.cu:60-96 runs)
#include "AugmentedBoundaryIntegrator.cuh"
#include "HeliumWithDrivingBoundaryProblem.cuh"   // the explicitly time-dependent drive (A/kernel.cu:281-366, L/Export.cu:779-975)
#include "RK4_Time_Dependent.cuh"

#include <chrono>
#include <cstring>

extern "C" {
struct ref_props {
	double rho, kappa, depth, U, L;
	int use_expansions, expansion_order, infinite_depth;
};
struct ref_opto {   // OptomechanicalVariables, L/OptomechanicalVariables.h
	double detuning, gamma, G, Tau, max_intensity, initial_time, location_x0_mode, sigma_optical_mode, Beta, DampingStrength;
};
}

namespace {

enum Kind { WATER = 0, HELIUM = 1, HELIUM_INF = 2 };

ProblemProperties toProps(const ref_props* p)
{
	ProblemProperties q;
	q.rho = p->rho; q.kappa = p->kappa; q.depth = p->depth; q.U = p->U; q.L = p->L;
	q.use_expansions = p->use_expansions != 0; q.expansion_order = p->expansion_order; q.infinite_depth = p->infinite_depth != 0;
	q.y_min = -1.0; q.y_max = 1.0;
	return q;
}

OptomechanicalVariables toOpto(const ref_opto* o)
{
	OptomechanicalVariables v;
	v.detuning = o->detuning; v.gamma = o->gamma; v.G = o->G; v.Tau = o->Tau; v.max_intensity = o->max_intensity;
	v.initial_time = o->initial_time; v.location_x0_mode = o->location_x0_mode; v.sigma_optical_mode = o->sigma_optical_mode;
	v.Beta = o->Beta; v.DampingStrength = o->DampingStrength;
	return v;
}

int lastError(const char* where)
{
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) { fprintf(stderr, "ref harness: CUDA error after %s: %s\n", where, cudaGetErrorString(e)); return -2; }
	return 0;
}

// one RHS through BaseBoundaryIntegralCalculator::run, plus the intermediates the reference exposes
template<int N, template<int, size_t> class Problem>
int rhsImpl(const ref_props* rp, const double* state, double* rhs, double* a, double* velUpper, double* zp, double* zpp,
            double* phiPrime, double* energies)
{
	ProblemProperties props = toProps(rp);
	Problem<N, 1> problem(props);
	BaseBoundaryIntegralCalculator<N, 1> calc(props, problem);
	std_complex *dState = nullptr, *dRhs = nullptr;
	if (cudaMalloc(&dState, 2 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	if (cudaMalloc(&dRhs, 2 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	cudaMemcpy(dState, state, 2 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
	cudaMemset(dRhs, 0, 2 * N * sizeof(std_complex));
	calc.run(dState, dRhs);
	int rc = lastError("run");
	if (rc == 0) {
		cudaMemcpy(rhs, dRhs, 2 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		if (a) cudaMemcpy(a, calc.getDevA(), N * sizeof(double), cudaMemcpyDeviceToHost);
		if (velUpper) cudaMemcpy(velUpper, calc.devVelocitiesUpper, N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		if (zp) cudaMemcpy(zp, calc.getDevZp(), N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		if (zpp) cudaMemcpy(zpp, calc.getDevZpp(), N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		if (phiPrime) cudaMemcpy(phiPrime, calc.devPhiPrime, N * sizeof(double), cudaMemcpyDeviceToHost);
		if (energies) {
			energies[0] = problem.energyContainer.kineticEnergy->getEnergy();
			energies[1] = problem.energyContainer.potentialEnergy->getEnergy();
			energies[2] = problem.energyContainer.surfaceEnergy->getEnergy();
			energies[3] = calc.volumeFlux.getEnergy();
		}
		rc = lastError("read-back");
	}
	cudaFree(dState);
	cudaFree(dRhs);
	return rc;
}

// warmup + steps classical RK4 steps with the reference's stepper; `seconds` = host clock around the last `steps` steps with a
// device synchronisation on both sides; energies (if asked) are those of the last RHS evaluated (4th stage of the last step)
template<int N, template<int, size_t> class Problem>
int rk4Impl(const ref_props* rp, double* state, double dt, int warmup, int steps, double* seconds, double* energies)
{
	ProblemProperties props = toProps(rp);
	Problem<N, 1> problem(props);
	BaseBoundaryIntegralCalculator<N, 1> calc(props, problem);
	AutonomousRungeKuttaStepper<std_complex, 2 * N> stepper(calc, dt);
	std_complex* dState = nullptr;
	if (cudaMalloc(&dState, 2 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	cudaMemcpy(dState, state, 2 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
	stepper.initialize(dState, true);
	for (int i = 0; i < warmup; i++) stepper.runStep(i);
	int rc = lastError("warm-up steps");
	if (rc == 0) {
		auto t0 = std::chrono::steady_clock::now();
		for (int i = 0; i < steps; i++) stepper.runStep(warmup + i);
		rc = lastError("steps");
		auto t1 = std::chrono::steady_clock::now();
		if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	}
	if (rc == 0) {
		cudaMemcpy(state, dState, 2 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		if (energies) {
			energies[0] = problem.energyContainer.kineticEnergy->getEnergy();
			energies[1] = problem.energyContainer.potentialEnergy->getEnergy();
			energies[2] = problem.energyContainer.surfaceEnergy->getEnergy();
			energies[3] = calc.volumeFlux.getEnergy();
		}
		rc = lastError("read-back");
	}
	cudaFree(dState);
	return rc;
}

// the augmented autonomous system [Z | Phi | D] exactly as L/Export.cu:1136-1150 and A/kernel.cu:85-94 assemble it; steps == 0: one RHS
// into `out` (3N complex); steps > 0: warmup + steps RK4 steps of AutonomousRungeKuttaStepper<std_complex, 3N>, final state into `out`
template<int N>
int augImpl(const ref_props* rp, const ref_opto* ro, const double* state, double* out, double dt, int warmup, int steps, double* seconds)
{
	ProblemProperties props = toProps(rp);
	OptomechanicalVariables vars = toOpto(ro);
	HeliumDrivenAutonomousProblem<N, 1> problem(props, vars);
	std::unique_ptr<BaseBoundaryIntegralCalculator<N, 1>> calc = std::make_unique<BaseBoundaryIntegralCalculator<N, 1>>(props, problem);
	AugmentedBoundaryIntegrator<N, 1> integrator(std::move(calc), std::make_unique<DelayedIntensityIntegrator<N, 1>>(vars));
	std_complex *dState = nullptr, *dRhs = nullptr;
	if (cudaMalloc(&dState, 3 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	if (cudaMalloc(&dRhs, 3 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	cudaMemcpy(dState, state, 3 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
	cudaMemset(dRhs, 0, 3 * N * sizeof(std_complex));
	int rc = 0;
	if (steps <= 0) {
		integrator.run(dState, dRhs);
		rc = lastError("augmented run");
		if (rc == 0) cudaMemcpy(out, dRhs, 3 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
	} else {
		AutonomousRungeKuttaStepper<std_complex, 3 * N> stepper(integrator, dt);
		stepper.initialize(dState, true);
		for (int i = 0; i < warmup; i++) stepper.runStep(i);
		rc = lastError("augmented warm-up steps");
		if (rc == 0) {
			auto t0 = std::chrono::steady_clock::now();
			for (int i = 0; i < steps; i++) stepper.runStep(warmup + i);
			rc = lastError("augmented steps");
			auto t1 = std::chrono::steady_clock::now();
			if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
		}
		if (rc == 0) cudaMemcpy(out, dState, 3 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
	}
	if (rc == 0) rc = lastError("augmented read-back");
	cudaFree(dState);
	cudaFree(dRhs);
	return rc;
}

// the explicitly time-dependent drive: HeliumWithOptomechanicalDrivingProblem<N> + TimedBoundaryIntegrator<N,1> +
// RungeKuttaStepper<std_complex, 2N>::runEvolution(t0, t0 + steps*dt), as L/Export.cu:797-826 assembles it; state 2N complex in/out
template<int N>
int timedImpl(const ref_props* rp, const ref_opto* ro, double* state, double t0, double dt, int steps, double* seconds)
{
	ProblemProperties props = toProps(rp);
	OptomechanicalVariables vars = toOpto(ro);
	HeliumWithOptomechanicalDrivingProblem<N> problem(props, vars);
	TimedBoundaryIntegrator<N, 1> integrator(props, problem);
	RungeKuttaStepper<std_complex, 2 * N> stepper(integrator);
	RK4Options options;
	options.initial_timestep = dt;
	options.returnTrajectory = false;
	stepper.setOptions(options);
	std_complex* dState = nullptr;
	if (cudaMalloc(&dState, 2 * N * sizeof(std_complex)) != cudaSuccess) return -3;
	cudaMemcpy(dState, state, 2 * N * sizeof(std_complex), cudaMemcpyHostToDevice);
	stepper.initialize(dState, true);
	auto c0 = std::chrono::steady_clock::now();
	stepper.runEvolution(t0, t0 + (steps + 0.5) * dt);   // steps = size_t((t1 - t0) / dt): the half step keeps the truncation exact
	int rc = lastError("timed evolution");
	auto c1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(c1 - c0).count();
	if (rc == 0) {
		cudaMemcpy(state, dState, 2 * N * sizeof(std_complex), cudaMemcpyDeviceToHost);
		rc = lastError("timed read-back");
	}
	cudaFree(dState);
	return rc;
}

template<int N>
int rhsKind(int kind, const ref_props* rp, const double* state, double* rhs, double* a, double* vu, double* zp, double* zpp, double* pp, double* en)
{
	switch (kind) {
	case WATER: return rhsImpl<N, WaterBoundaryProblem>(rp, state, rhs, a, vu, zp, zpp, pp, en);
	case HELIUM: return rhsImpl<N, HeliumBoundaryProblem>(rp, state, rhs, a, vu, zp, zpp, pp, en);
	case HELIUM_INF: return rhsImpl<N, HeliumInfiniteDepthBoundaryProblem>(rp, state, rhs, a, vu, zp, zpp, pp, en);
	}
	return -1;
}

template<int N>
int rk4Kind(int kind, const ref_props* rp, double* state, double dt, int warmup, int steps, double* seconds, double* en)
{
	switch (kind) {
	case WATER: return rk4Impl<N, WaterBoundaryProblem>(rp, state, dt, warmup, steps, seconds, en);
	case HELIUM: return rk4Impl<N, HeliumBoundaryProblem>(rp, state, dt, warmup, steps, seconds, en);
	case HELIUM_INF: return rk4Impl<N, HeliumInfiniteDepthBoundaryProblem>(rp, state, dt, warmup, steps, seconds, en);
	}
	return -1;
}

}  // namespace

// the sizes instantiated (the reference fixes N at compile time: L/Export.cu:560-599 switches over a similar list)
#define REF_SIZES(X) X(64) X(256) X(1024) X(4096) X(16384)

extern "C" {

__attribute__((visibility("default"))) int ref_num_sizes(int* out, int cap)
{
	const int sizes[] = {
#define X(n) n,
		REF_SIZES(X)
#undef X
	};
	int n = (int)(sizeof(sizes) / sizeof(sizes[0]));
	for (int i = 0; i < n && i < cap; i++) out[i] = sizes[i];
	return n;
}

// state, rhs: 2N complex128 [Z | Phi] / [conj-velocity | dPhi/dt] exactly as AutonomousProblem::run sees them; optional outputs
// may be NULL.  Returns 0, -1 (N / kind not instantiated), -2 (CUDA error), -3 (allocation).
__attribute__((visibility("default"))) int ref_rhs(int kind, int N, const ref_props* props, const double* state, double* rhs,
                                                   double* a, double* velUpper, double* zp, double* zpp, double* phiPrime, double* energies)
{
	try {
		switch (N) {
#define X(n) case n: return rhsKind<n>(kind, props, state, rhs, a, velUpper, zp, zpp, phiPrime, energies);
			REF_SIZES(X)
#undef X
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_rhs: %s\n", e.what());
		return -2;
	}
	return -1;
}

__attribute__((visibility("default"))) int ref_rk4(int kind, int N, const ref_props* props, double* state, double dt, int warmup,
                                                   int steps, double* seconds, double* energies)
{
	try {
		switch (N) {
#define X(n) case n: return rk4Kind<n>(kind, props, state, dt, warmup, steps, seconds, energies);
			REF_SIZES(X)
#undef X
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_rk4: %s\n", e.what());
		return -2;
	}
	return -1;
}

// augmented optomechanical system, N in {64, 256, 1024}; steps == 0: rhs of `state` into `out`, else the state after the steps
#define REF_AUG_SIZES(X) X(64) X(256) X(1024)
__attribute__((visibility("default"))) int ref_augmented(int N, const ref_props* props, const ref_opto* opto, const double* state,
                                                         double* out, double dt, int warmup, int steps, double* seconds)
{
	try {
		switch (N) {
#define X(n) case n: return augImpl<n>(props, opto, state, out, dt, warmup, steps, seconds);
			REF_AUG_SIZES(X)
#undef X
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_augmented: %s\n", e.what());
		return -2;
	}
	return -1;
}

__attribute__((visibility("default"))) int ref_timed(int N, const ref_props* props, const ref_opto* opto, double* state, double t0,
                                                     double dt, int steps, double* seconds)
{
	try {
		switch (N) {
#define X(n) case n: return timedImpl<n>(props, opto, state, t0, dt, steps, seconds);
			REF_AUG_SIZES(X)
#undef X
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_timed: %s\n", e.what());
		return -2;
	}
	return -1;
}

}  // extern "C"
