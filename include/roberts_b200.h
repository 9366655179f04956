/* roberts_b200.h -- C ABI of libroberts_b200.so
 *
 * B200-native (sm_100a) implementation of the Roberts (1983) boundary-integral
 * RK4 time step of CuSuperHelium.  Plain pointers and sizes only; every entry
 * point names the reference interface it replaces (paths relative to the
 * reference repository; L/ = CuSuperHelium/CuSuperHelium/).
 *
 * Conventions
 *   - all functions return 0 on success, -1 on error (message on stderr and in
 *     rb_last_error()), like the reference exports (L/Export.cu: try/catch -> -1).
 *   - "dev" pointers are CUDA device pointers on the handle's device; "host"
 *     pointers are ordinary host memory.
 *   - complex numbers are interleaved (re, im) doubles, 16-byte aligned
 *     (== cuda::std::complex<double> == c_double, L/constants.cuh:12, L/ExportTypes.cuh:7).
 *   - state layout [Z_b0 .. Z_b(B-1) | Phi_b0 .. Phi_b(B-1)], 2*B*N complex
 *     (L/BaseBoundaryIntegrator.cuh:141-145); rhs layout [u+iv | dPhi/dt + 0i].
 *   - N and the batch size are runtime values (the reference fixes them as
 *     template parameters and dispatches through switch tables, L/Export.cu:560-599).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef ROBERTS_B200_H
#define ROBERTS_B200_H

#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RB_API __declspec(dllexport)
#else
#define RB_API __attribute__((visibility("default")))
#endif

/* c_double of L/ExportTypes.cuh:7: two doubles, over-aligned to 16 bytes as the reference declares it (__declspec(align(16)),
 * static_assert(alignof(c_double) == 16) at :78-79) -- the layout of cufftDoubleComplex / double2, so arrays and struct members of
 * this type land where the reference's would */
#if defined(_MSC_VER)
#define RB_ALIGN16 __declspec(align(16))
#else
#define RB_ALIGN16 __attribute__((aligned(16)))
#endif
typedef struct RB_ALIGN16 rb_complex { double re, im; } rb_complex;
typedef struct rb_solver rb_solver;                           /* opaque: one RHS assembler + work buffers */
typedef struct rb_stepper rb_stepper;                         /* opaque: RK4 stepper bound to a solver */
typedef struct rb_rk45 rb_rk45;                               /* opaque: adaptive RKF45 stepper (L/RK45.cuh) */
typedef struct rb_aug_stepper rb_aug_stepper;                 /* opaque: RK4 stepper of the optomechanically driven (augmented) system */
typedef struct rb_timed_stepper rb_timed_stepper;             /* opaque: RK4 stepper of the explicitly time-dependent drive */
typedef struct rb_jacobian rb_jacobian;                       /* opaque: finite-difference Jacobian of the real-state RHS (batch 3N) */
typedef struct rb_gl2 rb_gl2;                                 /* opaque: implicit Gauss-Legendre-2 integrator */

/* physics plugin selector == which BoundaryProblem<N,B> subclass the reference would instantiate */
enum rb_physics {
    RB_WATER = 0,        /* WaterBoundaryProblem            L/WaterBoundaryProblem.cuh:8-39   */
    RB_HELIUM = 1,       /* HeliumBoundaryProblem           L/HeliumBoundaryProblem.cuh:6-47  */
    RB_HELIUM_INF = 2    /* HeliumInfiniteDepthBoundaryProblem  L/HeliumBoundaryProblem.cuh:50-81 */
};

/* how the vortex-sheet strength system M a = Re(Phi') is solved (replaces MatrixSolver<N,B>::solve,
 * L/MatrixSolver.cuh:114-172, cuSOLVER LU) */
enum rb_solve_mode {
    RB_SOLVE_MATRIX_FREE = 0,  /* Richardson/Neumann iteration on the second-kind system, M never stored */
    RB_SOLVE_DENSE_LU = 1      /* assemble M in HBM, in-place partial-pivot LU (validation path, small N) */
};

/* how the iteration is started */
enum rb_guess_mode {
    RB_GUESS_COLD = 0,     /* a0 = omega * b (first Neumann term); results independent of call history */
    RB_GUESS_WARM = 1      /* previous solution of this solver (or the stepper's stage history) */
};

/* mirrors ProblemProperties (L/ProblemProperties.hpp:5-32), nondimensional values */
typedef struct rb_props {
    double rho;            /* density ratio upper/lower                                    */
    double U;              /* mean shear; Phi carries the linear part -(1+rho) pi U j / N  */
    double kappa;          /* surface tension                                              */
    double depth;          /* film depth h (helium)                                        */
    int use_expansions;    /* HeliumBoundaryProblem::CalculateRhsPhi switch, :35-46        */
    int expansion_order;
    int infinite_depth;
    int physics;           /* enum rb_physics                                              */
    int solve_mode;        /* enum rb_solve_mode                                           */
    int guess_mode;        /* enum rb_guess_mode                                           */
    int max_iterations;    /* cap on M*x applications per solve (0 -> default 200)         */
    int compute_energies;  /* 1: evaluate the Energies.cuh sums on every RHS (reference behaviour) */
    double tolerance;      /* relative residual ||b - M a|| / ||b|| (0 -> default 1e-13)   */
} rb_props;

/* ---- library / device ---------------------------------------------------------------- */
RB_API const char* rb_last_error(void);
RB_API int rb_version(void);
RB_API int rb_device_count(void);                 /* 0 without a GPU; never falls back to the CPU */
RB_API int rb_set_device(int device);             /* replaces setDevice(), L/utilities.cuh:15-27 (device 0 hard-wired there) */
RB_API void rb_default_props(rb_props* p);        /* ProblemProperties defaults, water, rho = 0 */

/* ---- RHS assembler == BaseBoundaryIntegralCalculator<N,B> (L/BaseBoundaryIntegrator.cuh:10-85) ---- */
RB_API rb_solver* rb_create(int N, int batch, const rb_props* props);            /* ctor  :89-108  */
RB_API int rb_destroy(rb_solver* s);                                             /* dtor  :111-135 */
RB_API int rb_set_stream(rb_solver* s, void* cuda_stream);                       /* setStream :37-40 */
RB_API void* rb_get_stream(rb_solver* s);                                        /* the cudaStream_t every kernel of s is issued on */
RB_API int rb_get_props(rb_solver* s, rb_props* out, int* N, int* batch);        /* what the solver was created with */
RB_API int rb_rhs(rb_solver* s, const rb_complex* state_dev, rb_complex* rhs_dev);   /* run / runTimeStep :138-273, 310-314 */
RB_API int rb_vorticities(rb_solver* s, const rb_complex* state_dev);            /* calculateVorticities :276-306 */
RB_API double* rb_dev_a(rb_solver* s);                                           /* getDevA   :21-24 */
RB_API rb_complex* rb_dev_zp(rb_solver* s);                                      /* getDevZp  :25-27 */
RB_API rb_complex* rb_dev_zpp(rb_solver* s);                                     /* getDevZpp :28-30 */
RB_API rb_complex* rb_dev_velocities_upper(rb_solver* s);                        /* devVelocitiesUpper :43 */
RB_API double* rb_dev_phi_prime(rb_solver* s);                                   /* devPhiPrime :46 */
RB_API int rb_synchronize(rb_solver* s);
/* energies of the last RHS: out[0] kinetic, [1] potential (gravitational or van der Waals), [2] surface,
 * [3] volume flux, [4] volume sum(Y X').  Replaces EnergyBase::getEnergy x4 (L/Energies.cuh:229-235) + VolumeFlux. */
RB_API int rb_energies(rb_solver* s, double out_host[5]);
/* statistics: out[0] M*x applications of the last solve, [1] converged flag, [2] relative residual,
 * [3] solver sweeps (M*x applications, including the combined verify+velocity sweeps) summed over all solves, [4] number of
 * solves, [5] velocity-only sweeps */
RB_API int rb_solve_stats(rb_solver* s, double out_host[6]);
/* How solves ended.  The reference's direct LU (MatrixSolver<N,B>::solve, L/MatrixSolver.cuh:114-172) cannot fail to converge; the
 * iterations that replace it can, and say so:  converged = the relative residual met the tolerance, nothing else;  stagnated = the
 * iteration stopped on the round-off floor of the residual (<= 1e-10, no longer contracting) above the tolerance -- accepted, counted
 * separately;  failed = neither (iteration cap, NaN, a peer rank that never signalled).  Inside the RK4 stepper "last" refers to
 * the last step and aggregates its four stage solves (converged: all four; stagnated: any; out[4] of rb_solve_stats' residual: the
 * largest).  out[0] last converged, [1] last stagnated, [2] stagnated solves since creation, [3] failed solves since creation,
 * [4] largest final relative residual of any accepted solve, [5] strict flag, [6..7] reserved (0). */
RB_API int rb_solve_status(rb_solver* s, double out_host[8]);
/* strict (default 1): a failed solve makes rb_rhs / rb_vorticities / rb_rk4_step / ... return -1 with rb_last_error set, and a failed
 * RK4 step leaves the state as it was before the step.  0: failures are only reported through rb_solve_status. */
RB_API int rb_set_strict(rb_solver* s, int strict);

/* ---- spectral derivatives (L/Derivatives.cuh) ---- */
/* ZPhiDerivative<N,B>::exec :311-384 */
RB_API int rb_zphi_derivative(rb_solver* s, const rb_complex* Z_dev, const rb_complex* Phi_dev,
                              rb_complex* Zp_dev, rb_complex* PhiPrime_dev, rb_complex* Zpp_dev);
/* FftDerivative<N,B>::exec :190-257 (second != 0 -> doubleDev) */
RB_API int rb_fft_derivative(rb_solver* s, const rb_complex* in_dev, rb_complex* out_dev, int second, double scaling);

/* ---- kernels the reference's tests launch by name (materialised, column-major A[k + j*n + b*n*n]) ---- */
RB_API int rb_create_M(double* A_dev, const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp,
                       double rho, int n, size_t batch, void* stream);                            /* createMKernel L/createM.cuh:43-63 */
RB_API int rb_create_finite_depth_M(double* A_dev, const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp,
                                    double h, int n, size_t batch, int infinite_depth, void* stream); /* createFiniteDepthMKernel :65-92 */
RB_API int rb_velocity_matrices(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, int n,
                                rb_complex* V1, rb_complex* V2, int lower, size_t batch, void* stream); /* createVelocityMatrices L/WaterVelocities.cuh:38-70 */
RB_API int rb_helium_velocity_matrices(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp, double h, int n,
                                       rb_complex* V1, rb_complex* V2, int lower, size_t batch, int infinite_depth,
                                       void* stream);                                             /* createHeliumVelocityMatrices :72-107 */
RB_API int rb_rhs_phi_water(const rb_complex* Z, const rb_complex* V1, const rb_complex* V2, rb_complex* result,
                            double rho, int n, void* stream);                                     /* compute_rhs_phi_expression L/createM.cuh:96-107 */
RB_API int rb_rhs_phi_helium(const rb_complex* Z, const rb_complex* V1, rb_complex* result, double h, int n,
                             void* stream);                                                       /* compute_rhs_helium_phi_expression :109-117 */
RB_API int rb_rhs_phi_helium_surface_tension(const rb_complex* Z, const rb_complex* Zp, const rb_complex* Zpp,
                                             const rb_complex* V1, rb_complex* result, double h, double kappa, int n,
                                             void* stream);                                       /* ..._with_surface_tension :200-211 */
RB_API int rb_rhs_phi_helium_expansion(const rb_complex* Z, const rb_complex* V1, rb_complex* result, double h, int n,
                                       int order, void* stream);                                  /* ..._expansion_terms :171-193 */
/* materialised operator applied by this library's own kernels: y = V1*a + V2.*aprime, conj (VelocityCalculator, L/WaterVelocities.cuh:206-242) */
RB_API int rb_cotangent_sum(rb_solver* s, const rb_complex* Z_dev, const double* x_dev, rb_complex* S_dev);
                                                  /* S_k = sum_{j!=k} cot((Z_k-Z_j)/2) x_j, the matrix-free core (needs Zp/Zpp: call after rb_vorticities or rb_zphi_derivative) */

/* ---- RK4 stepper == AutonomousRungeKuttaStepper<std_complex, 2N> (L/AutonomousRungeKuttaStepper.cuh:24-121) ---- */
RB_API rb_stepper* rb_rk4_create(rb_solver* s, double tstep);                    /* ctor :91-105 */
RB_API int rb_rk4_destroy(rb_stepper* st);
RB_API int rb_rk4_set_time_step(rb_stepper* st, double tstep);                   /* setTimeStep :31-36 */
RB_API int rb_rk4_initialize(rb_stepper* st, rb_complex* y0, int on_device);     /* initialize :310-329: on_device aliases the caller's buffer */
RB_API int rb_rk4_step(rb_stepper* st);                                          /* runStep :124-307 */
/* runEvolution :418-437: steps = size_t((t1-t0)/dt); returns the number of steps taken through *steps_out */
RB_API int rb_rk4_evolve(rb_stepper* st, double t0, double t1, size_t* steps_out);
RB_API int rb_rk4_run_steps(rb_stepper* st, size_t steps);                       /* the hot loop without the time bookkeeping */
RB_API rb_complex* rb_rk4_dev_state(rb_stepper* st);                             /* devY0 */
RB_API int rb_rk4_get_state(rb_stepper* st, rb_complex* y_host);                 /* D2H of devY0 */
RB_API double rb_rk4_current_time(rb_stepper* st);
/* out[0] CUDA-graph launches (one per step), [1] graph captures, [2] steps redone outside the graph, [3] sweeps recorded per solve */
RB_API int rb_rk4_stats(rb_stepper* st, double out_host[4]);
/* asynchronous chunks of rb_rk4_run_steps / rb_rk4_evolve (several recorded steps launched back to back, one host look per chunk; a
   chunk in which any solve ran out of recorded sweeps or failed is rolled back and redone step by step): out[0] steps per chunk
   (0: off, RB_ASYNC_STEPS), [1] chunks launched, [2] chunks rolled back, [3] steps recorded without a surplus sweep round that had to
   be redone (+ 0.5 while such a tight recording is in use) */
RB_API int rb_rk4_chunk_stats(rb_stepper* st, double out_host[4]);
/* quality of the extrapolated initial iterates: out[0..3] relative residual of the guess at RK stage 1..4 of the last step,
   [4] mask of stages that currently start with a combined (verify + velocity) sweep, [5] solves started that way so far,
   [6] of those, solves that finished in ONE O(N^2) sweep, [7] policy (0 never, 1 adaptive per stage, 2 always) */
RB_API int rb_rk4_guess_stats(rb_stepper* st, double out_host[8]);
RB_API int rb_rk4_set_optimistic(rb_stepper* st, int policy);
/* initial iterate of every solve: polynomial extrapolation of order `order` (1..6 previous steps, same RK stage) of the recorded
   solutions; predict = 1 adds one Richardson sweep whose O(N^2) row sums are themselves extrapolated in time (O(N) work),
   0 = off, -1 = automatic (on when the solve tolerance is >= 4e-13, the round-off floor of that extrapolation). Resets the history. */
RB_API int rb_rk4_set_guess(rb_stepper* st, int order, int predict);
/* trajectory logging (TrajectoryLogger<T,N>, L/TrajectoryLogger.cuh:7-73): every `every` steps into a device ring, 0 = off */
RB_API int rb_rk4_set_logging(rb_stepper* st, size_t every, size_t capacity);
RB_API int rb_rk4_copy_trajectory(rb_stepper* st, double** times_out, size_t* times_count,
                                  rb_complex** states_out, size_t* states_count);  /* copyTimesToHost/copyStatesToHost; free with rb_free */
RB_API void rb_free(void* p);
/* generic stage kernels (cublasZaxpy x4 + add_k_vectors, :349-361, L/utilities.cuh:78-83) for callers that own the loop */
RB_API int rb_rk4_stage_update(rb_complex* y_out, const rb_complex* y0, const rb_complex* k, double c, size_t n, void* stream);
RB_API int rb_rk4_final_update(rb_complex* y0, const rb_complex* k1, const rb_complex* k2, const rb_complex* k3,
                               const rb_complex* k4, double h, size_t n, void* stream);

/* ---- adaptive Runge-Kutta-Fehlberg 4(5): RK45Base<T,N> / RK45_std_complex<N> (L/RK45.cuh:103-180, 258-330; kernels
 *      L/RK45_Kernels.cuh:63-106, L/LinearAlgebra.cuh:10-112).  Same tableau, error norm sqrt(mean((|e|/(atol + rtol max(|y|,|y5|)))^2)),
 *      step-size controller (safety 0.9, factor in [0.2, 5], no growth on a retry) and k1 reuse after a rejected attempt. ---- */
typedef struct rb_rk45_options {           /* RK45_Options, L/RK45.cuh:21-27 */
    double atol;                           /* 1e-6 */
    double rtol;                           /* 1e-3 */
    double h_min;                          /* 1e-16 */
    double h_max;                          /* 1e10 */
    double initial_timestep;               /* 1e-2 */
} rb_rk45_options;
/* AutonomousProblem<T,N>::run(T* state, T* rhs) (L/AutonomousProblem.h:9-28) for problems other than the boundary integral:
   device pointers, work must be enqueued on `cuda_stream` */
typedef void (*rb_rhs_fn)(void* user, const rb_complex* state_dev, rb_complex* rhs_dev, void* cuda_stream);
RB_API rb_rk45* rb_rk45_create(rb_solver* s, const rb_rk45_options* opt);          /* RHS = rb_rhs of s; state = 2 N B complex */
RB_API rb_rk45* rb_rk45_create_generic(size_t n, rb_rhs_fn f, void* user, const rb_rk45_options* opt, void* stream);
RB_API int rb_rk45_destroy(rb_rk45* r);
RB_API int rb_rk45_set_options(rb_rk45* r, const rb_rk45_options* opt);            /* setOptions :140-145 */
RB_API int rb_rk45_set_tolerance(rb_rk45* r, double atol, double rtol);            /* setTolerance :123-127 */
RB_API int rb_rk45_set_max_rejected(rb_rk45* r, size_t max_rejected);              /* setMaxRejectedSteps :137-139 (500) */
RB_API int rb_rk45_initialize(rb_rk45* r, const rb_complex* y0, int on_device);    /* copies, :332-341 */
RB_API int rb_rk45_step(rb_rk45* r, int* accepted);                                /* runStep :258-304 */
RB_API int rb_rk45_evolve(rb_rk45* r, double t0, double t1, int* result);          /* runEvolution :194-255; 0 ReachedEndTime, 1 StiffnessDetected */
RB_API rb_complex* rb_rk45_dev_state(rb_rk45* r);                                  /* getY :130-132 */
RB_API int rb_rk45_get_state(rb_rk45* r, rb_complex* y_host);
RB_API double rb_rk45_current_time(rb_rk45* r);                                    /* getCurrentTime :133-135 */
RB_API double rb_rk45_current_timestep(rb_rk45* r);
RB_API int rb_rk45_stats(rb_rk45* r, double out_host[4]);   /* accepted steps, rejected attempts, RHS evaluations, last scaled error */

/* ---- optomechanically driven helium film: the autonomous augmented system y = [Z | Phi | D], D = delayed light intensity
 *      (what CuSuperHelium.App runs, A/kernel.cu:60-96).  Replaces
 *      HeliumDrivenAutonomousProblem<N,B> (L/HeliumDrivenAutonomousProblem.cuh:10-26, kernel L/createM.cuh:138-149),
 *      DelayedIntensityIntegrator<N,B> (L/DelayedIntensityIntegrator.cuh:9-39, kernels L/createM.cuh:151-169),
 *      AugmentedBoundaryIntegrator<N,B> (L/AugmentedBoundaryIntegrator.cuh:10-40) and LightIntensity (L/LightIntensity.cuh:11-35).
 *      state / rhs: 3 N B complex, [Z | Phi | D] -> [w | dPhi/dt | dD/dt]. ---- */
typedef struct rb_opto {          /* OptomechanicalVariables, L/OptomechanicalVariables.h:3-28 (nondimensional values) */
    double detuning, gamma, G, Tau, max_intensity, initial_time, location_x0_mode, sigma_optical_mode, Beta, DampingStrength;
    /* LightIntensity::get_current_intensity_drive_strength(variables, properties), L/LightIntensity.cuh:30-33:
       hbar / (base_energy base_time rho) G / sigma^2 -- a property of (variables, properties); see rb_opto_drive_strength */
    double drive_strength;
} rb_opto;
RB_API void rb_default_opto(rb_opto* v);                                          /* the struct's default member initialisers */
RB_API double rb_opto_drive_strength(const rb_opto* v, double base_energy, double base_time, double rho);
RB_API int rb_light_intensity(const rb_complex* Z_dev, double* intensity_dev, const rb_opto* v, size_t n, void* stream);
/* AugmentedBoundaryIntegrator::run :25-29 = BaseBoundaryIntegralCalculator::run with the driven dPhi/dt, then the delayed-intensity
   terms.  The reference derives the driven problem from HeliumBoundaryProblem; any physics of the solver is accepted here. */
RB_API int rb_augmented_rhs(rb_solver* s, const rb_opto* v, const rb_complex* state_dev, rb_complex* rhs_dev);
/* AutonomousRungeKuttaStepper<std_complex, 3N>(AugmentedBoundaryIntegrator&, dt): classical RK4 over the 3 N B state */
RB_API rb_aug_stepper* rb_aug_rk4_create(rb_solver* s, const rb_opto* v, double tstep);
RB_API int rb_aug_rk4_destroy(rb_aug_stepper* st);
RB_API int rb_aug_rk4_set_time_step(rb_aug_stepper* st, double tstep);
RB_API int rb_aug_rk4_initialize(rb_aug_stepper* st, rb_complex* y0, int on_device);  /* on_device aliases the caller's 3 N B buffer */
RB_API int rb_aug_rk4_step(rb_aug_stepper* st);
RB_API int rb_aug_rk4_run_steps(rb_aug_stepper* st, size_t steps);
RB_API int rb_aug_rk4_evolve(rb_aug_stepper* st, double t0, double t1, size_t* steps_out);   /* steps = size_t((t1-t0)/dt) */
RB_API rb_complex* rb_aug_rk4_dev_state(rb_aug_stepper* st);
RB_API int rb_aug_rk4_get_state(rb_aug_stepper* st, rb_complex* y_host);
RB_API double rb_aug_rk4_current_time(rb_aug_stepper* st);

/* ---- the same drive in its explicitly time-dependent form: HeliumWithOptomechanicalDrivingProblem<N>
 *      (L/HeliumWithDrivingBoundaryProblem.cuh:7-67, kernel add_optical_field_drive_terms L/createM.cuh:119-136), the exponential
 *      integrator DelayedIntensityTerm<N> (L/DelayedIntensityTerm.cuh:9-71), TimedBoundaryIntegrator<N,B>
 *      (L/TimedBoundaryIntegrator.cuh:8-49) and RungeKuttaStepper<std_complex, 2N>(TimedProblem&) (L/RK4_Time_Dependent.cuh:18-460;
 *      usage A/kernel.cu:281-366, L/Export.cu:797-826).  State 2 N B complex [Z | Phi]; the delayed intensity lives in the stepper
 *      and is advanced (saved) by the first RK stage of every step. ---- */
RB_API rb_timed_stepper* rb_timed_rk4_create(rb_solver* s, const rb_opto* v, double tstep);
RB_API int rb_timed_rk4_destroy(rb_timed_stepper* st);
RB_API int rb_timed_rk4_set_time_step(rb_timed_stepper* st, double tstep);
RB_API int rb_timed_rk4_initialize(rb_timed_stepper* st, rb_complex* y0, int on_device);
/* TimedBoundaryIntegrator::setStartingTime :44-48: the current time and the delayed term's reference time */
RB_API int rb_timed_rk4_set_starting_time(rb_timed_stepper* st, double time);
/* TimedBoundaryIntegrator::run at an explicit time (setCurrentTime + setSaveProgress + run, L/RK4_Time_Dependent.cuh:148-150) */
RB_API int rb_timed_rhs(rb_timed_stepper* st, double time, int save_progress, const rb_complex* state_dev, rb_complex* rhs_dev);
/* runStep :145-283: one RK4 step at the stepper's current time (stages at t, t + h/2, t + h/2, t + h; only the first saves the
   delayed intensity).  As in the reference the time is advanced by the evolution loop, not by runStep: advance_time = 0 is
   runStep verbatim, advance_time = 1 adds the `currentTime += timeStep` of runEvolution's loop body (:313-325). */
RB_API int rb_timed_rk4_step(rb_timed_stepper* st, int advance_time);
/* runEvolution :307-328: setStartingTime(t0), steps = size_t((t1 - t0)/dt) */
RB_API int rb_timed_rk4_evolve(rb_timed_stepper* st, double t0, double t1, size_t* steps_out);
/* RK4Options::returnTrajectory (setOptions :29-33) and copyTimesToHost / copyStatesToHost :66-131: with a trajectory, one (time at the
   START of the step, state after the step) pair per step of rb_timed_rk4_evolve; without, no times and the current state alone.
   Buffers are malloc'd; release with rb_free. */
RB_API int rb_timed_rk4_set_logging(rb_timed_stepper* st, int return_trajectory);
RB_API int rb_timed_rk4_copy_trajectory(rb_timed_stepper* st, double** times_out, size_t* times_count, rb_complex** states_out,
                                        size_t* states_count);
RB_API rb_complex* rb_timed_rk4_dev_state(rb_timed_stepper* st);
RB_API double* rb_timed_rk4_dev_delayed_intensity(rb_timed_stepper* st);                /* DelayedIntensityTerm::delayed_intensity, N B doubles */
RB_API int rb_timed_rk4_get_state(rb_timed_stepper* st, rb_complex* y_host);
RB_API double rb_timed_rk4_current_time(rb_timed_stepper* st);

/* ---- real-state wrapper, finite-difference Jacobian and the implicit Gauss-Legendre-2 integrator:
 *      RealBoundaryItegralCalculator<N> (L/RealBoundaryIntegralCalculator.cuh:37-89), JacobianCalculator<N>
 *      (L/JacobianCalculator.cuh:168-284, kernels :11-166), GaussLegendre2<N> (L/GaussLegendre.cuh:107-612), assembled as
 *      L/Export.cu:392-518 (calculateJacobian) and :665-739 (integrateSimulationGL2).  Real states are [x | y | phi], 3N doubles. ---- */
/* RealBoundaryItegralCalculator::run :59-70: s must have batch 1 */
RB_API int rb_real_rhs(rb_solver* s, const double* state_real_dev, double* rhs_real_dev);
/* createInitialBatchedZ :11-77: 3N copies of [Z | Phi] (2N complex) in the batched layout [Z of member 0.. | Phi of member 0..]
   (6 N^2 complex), member c N + j with coordinate c (0 x, 1 y, 2 phi) of point j moved by eps */
RB_API int rb_perturbed_states(const rb_complex* state_dev, rb_complex* batched_dev, double eps, int N, void* cuda_stream);
/* JacobianCalculator<N>(std::make_unique<BaseBoundaryIntegralCalculator<N, 3N>>(properties, problem)): owns the batch-3N solver */
RB_API rb_jacobian* rb_jacobian_create(int N, const rb_props* props);
RB_API int rb_jacobian_destroy(rb_jacobian* j);
RB_API int rb_jacobian_set_epsilon(rb_jacobian* j, double eps);                  /* setEpsilon :217-221 (default 1e-6) */
RB_API int rb_jacobian_set_stream(rb_jacobian* j, void* cuda_stream);            /* setStream :186 */
RB_API rb_solver* rb_jacobian_solver(rb_jacobian* j);                            /* the batch-3N RHS assembler (statistics, stream) */
/* calculateJacobian :226-284: central differences of the batched RHS at +-eps; jac_dev is 3N x 3N column-major,
   jac[c * 3N + r] = d f_r / d y_c (createJacobianMatrixFromPerturbedRhs :92-156) */
RB_API int rb_jacobian_calculate(rb_jacobian* j, const double* state_real_dev, double* jac_dev);

/* MatrixSolver<N,1>::solve (cusolverDnDgetrf + Dgetrs, L/MatrixSolver.cuh:114-125): A (n x n, column-major, destroyed) x = b, x
   returned in b.  blocked = 0: unblocked right-looking LU (two launches per column); 1: panels of 32 columns with the trailing
   update on the FP64 tensor path; -1: what the library itself uses (RB_LU_BLOCKED).  *info_host = 0, or 1 + the first column
   without a usable pivot (getrf's info). */
RB_API int rb_lu_solve(double* A_dev, double* b_dev, int n, int blocked, int* info_host, void* cuda_stream);

/* GaussLegendre2Options (L/GaussLegendre.cuh:70-90) */
typedef struct rb_gl2_options {
    double stepSize;              /* 0.01 */
    double newtonTolerance;       /* 1e-10 */
    size_t maxNewtonIterations;   /* 20 */
    int allowSimplifiedFallback;  /* 0 (the struct's constructor overrides the member initialiser) */
    int returnTrajectory;         /* 1 */
    double armijo_c;              /* 1e-4 */
    double backtrack;             /* 0.5 */
    double minAlpha;              /* 1e-6 */
    size_t maxStepsHalves;        /* 6 */
} rb_gl2_options;
/* StepResult (L/GaussLegendre.cuh:176-181) of the last attempted step, plus totals since create */
typedef struct rb_gl2_stats {
    size_t numberIterations;      /* Newton iterations of the last attempted step */
    int converged;
    double residualNorm;
    int simplifiedFallbackUsed;
    size_t steps_accepted, steps_halved, newton_iterations, rhs_evaluations, jacobians, linear_solves;
} rb_gl2_stats;
RB_API void rb_gl2_default_options(rb_gl2_options* o);
/* GaussLegendre2<N>(problem, jacobianCalculator, options): s is the batch-1 RHS assembler behind the real wrapper; both stay
   caller-owned and are put on s's stream */
RB_API rb_gl2* rb_gl2_create(rb_solver* s, rb_jacobian* j, const rb_gl2_options* options);
RB_API int rb_gl2_destroy(rb_gl2* g);
RB_API int rb_gl2_set_options(rb_gl2* g, const rb_gl2_options* options);
RB_API int rb_gl2_initialize(rb_gl2* g, double* state, int on_device);           /* initialize :303-321 */
/* gaussLegendreS2Step :441-560 from the current state with step h: on convergence the state is advanced (and 1 written to
   *converged); otherwise it is left as it was */
RB_API int rb_gl2_step(rb_gl2* g, double h, int* converged);
/* runEvolution :216-299: steps of min(stepSize, |t1 - t|) towards t1 (either direction), halving a failed step at most
   maxStepsHalves times (never below |t1 - t0| / 2^20) and keeping the reduced size for the rest of the call (every call starts
   again from options.stepSize, as the reference's Python statement does; its CUDA class writes the last -- possibly
   end-truncated -- size back into the options, :290); the trajectory starts with (t0, y0).
   Returns -1 with "failed to converge" in rb_last_error when a step cannot be completed (the reference throws). */
RB_API int rb_gl2_evolve(rb_gl2* g, double t0, double t1);
/* copyTimesToHost / copyStatesToHost :355-401: states_count x 3N doubles; without a trajectory no times and the current state.
   Buffers are malloc'd; release with rb_free. */
RB_API int rb_gl2_copy_trajectory(rb_gl2* g, double** times_out, size_t* times_count, double** states_out, size_t* states_count);
RB_API double* rb_gl2_dev_state(rb_gl2* g);
RB_API int rb_gl2_get_state(rb_gl2* g, double* state_host);
RB_API int rb_gl2_get_stats(rb_gl2* g, rb_gl2_stats* out);

/* ---- multi-GPU (new; the reference is single-GPU, L/utilities.cuh:20): contiguous blocks of 256-row cells of every O(N^2)
 *      sweep are owned by one rank each; all ranks keep the full state and exchange result rows by peer stores over NVLink into
 *      a per-rank arena mapped with CUDA IPC.  One process per GPU of one node; ship the handles with any out-of-band channel
 *      (torch.distributed all_gather in superfluid_dynamics_b200/api.py). ---- */
RB_API int rb_comm_handle_bytes(void);                                           /* size of one exported handle (64) */
RB_API int rb_comm_export(rb_solver* s, char* handle_out);                       /* this rank's arena handle */
RB_API int rb_comm_init(rb_solver* s, int rank, int nranks, const char* handles); /* handles: nranks x rb_comm_handle_bytes() */
RB_API int rb_comm_row_range(int N, int rank, int nranks, int out_rows[2]);      /* host-only: rows [out[0], out[1]) owned by rank */
RB_API int rb_comm_error(rb_solver* s);                                          /* 1 if a peer wait timed out */
RB_API int rb_comm_destroy(rb_solver* s);

/* ---- measurement helpers ---- */
/* restrict the sweeps of a single-GPU solver to the row cells (256 rows each) [cell0, cell0 + cells) a rank of a row-sharded run would
 * own (no peer involved; cells <= 0: whole surface again) -- times and tunes a per-rank sweep on one GPU with rb_bench_sweep */
RB_API int rb_debug_set_row_range(rb_solver* s, int cell0, int cells);
/* the sweep schedule in use: out[0] kernel (1 tiled, 2 persistent), [1] rows per thread, [2] source tile, [3] tiles per chunk,
 * [4] source chunks, [5] row cells of this rank, [6] CTAs per sweep, [7] threads per CTA */
RB_API int rb_sweep_plan(rb_solver* s, int out[8]);
RB_API unsigned long long rb_launch_count(void);                                  /* kernels of this library launched since load */
RB_API int rb_measure_fp64_peak(double* tflops_out, void* stream);               /* DFMA-only kernel: the FP64 roofline denominator */
RB_API int rb_measure_fp64_rate_3operand(double* tflops_out, void* stream);      /* DFMA with three distinct register operands */
/* FP64 tensor path (DMMA m8n8k4) beside the FP64 vector pipe: out[2i] = ms, out[2i+1] = TFLOP/s for the per-iteration mixes
   {8 mma}, {32 fma}, {8 mma + 32 fma}, {4 + 32}, {2 + 32}, {1 + 32}: tells whether the two issue concurrently */
RB_API int rb_measure_fp64_tensor_overlap(double out_host[12], void* stream);
RB_API int rb_bench_sweep(rb_solver* s, const rb_complex* state_dev, int reps, float* ms_per_sweep_out, double* pairs_per_sweep_out);

/* ---- legacy exports (same names and argument meaning as L/Export.cuh:27-70; SI inputs, nondimensionalised inside,
 *      HeliumBoundaryProblem hard-wired as in L/Export.cu:205-207) ---- */
RB_API int calculateRHSFromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy, double* rhsPhi,
                                   double L, double rho, double kappa, double depth, size_t N);       /* L/Export.cuh:31 */
RB_API int calculateRHS256FromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy,
                                      double* rhsPhi, double L, double rho, double kappa, double depth); /* :30 */
RB_API int calculateRHS2048FromVectors(const double* x, const double* y, const double* phi, double* vx, double* vy,
                                       double* rhsPhi, double L, double rho, double kappa, double depth); /* :33 */
RB_API int calculateRHS256FromVectorsBatched(const double* x, const double* y, const double* phi, double* vx, double* vy,
                                             double* rhsPhi, double L, double rho, double kappa, double depth,
                                             int batchSize);                                           /* :51 */
RB_API int calculateVorticities256FromVectors(const rb_complex* Z, const rb_complex* phi, double* a, rb_complex* Zp,
                                              rb_complex* Zpp, double L, double rho, double kappa, double depth); /* :48 */
RB_API int calculateDerivativeFFT256(const rb_complex* input, rb_complex* output);                     /* :49 */

typedef struct SimProperties {            /* L/ExportTypes.cuh:8-20 */
    double L, rho, kappa, depth;
    bool use_expansions;
    int expansion_order;
    bool infinite_depth;
} SimProperties;
typedef struct RK4SolverOptions {         /* L/ExportTypes.cuh:35-40 */
    double timeStep, t0, t1;
    bool returnTrajectory;
} RK4SolverOptions;
/* declared by the reference (L/Export.cuh:69-70) but a stub there (L/Export.cu:767-777); implemented here.
 * initialState = [x | y | phi] (3N doubles, SI time in the options, lengths already in units of L/2pi as the
 * reference's callers pass them); *statesOut = statesCount x 3N doubles, *timesOut = timesCount doubles, malloc'd;
 * release with integrateSimulationRK4_freeMemory. */
RB_API int integrateSimulationRK4(double* initialState, double** statesOut, size_t* statesCount, double** timesOut,
                                  size_t* timesCount, SimProperties* simProperties, RK4SolverOptions* rkOptions, size_t N);
RB_API int integrateSimulationRK4_freeMemory(double* statesOut, double* timesOut);
/* nondimensional variant of the same call (no SI conversion, physics selectable): the plain RK4 path used by bench.py's e2e leg */
RB_API int rb_integrate_rk4_host(const double* initialState_host, double* finalState_host, size_t N, size_t batch,
                                 const rb_props* props, double dt, size_t steps);

typedef struct GaussLegendreOptions {      /* L/ExportTypes.cuh:20-33 */
    double t0, t1, stepSize, newtonTolerance;
    size_t maxNewtonIterations;
    bool allowSimplifiedFallback;
    bool returnTrajectory;
    double armijo_c, backtrack, minAlpha;
    size_t maxStepsHalves;
} GaussLegendreOptions;
/* L/Export.cuh:50, L/Export.cu:392-493: finite-difference Jacobian of the real-state helium RHS.  state = [x | y | phi] (3N
 * doubles, host), jac = 9 N^2 doubles (host), column-major: jac[c * 3N + r] = d f_r / d y_c.  SI properties are
 * nondimensionalised inside.  Any N >= 2 (the reference: N = 32 ... 2048 by switch table, no return value outside it, L/Export.cu:495-518). */
RB_API int calculateJacobian(const double* state, double* jac, double L, double rho, double kappa, double depth, double epsilon,
                             size_t N);
/* L/Export.cuh:53, L/Export.cu:600-663: the 3N perturbed copies of the state that calculateJacobian evaluates (N = 256):
 * Zperturbed = 6 N^2 complex, [Z of member 0 .. Z of member 3N-1 | Phi of member 0 ..] */
RB_API int calculatePerturbedStates256(const double* x, const double* y, const double* phi, rb_complex* Zperturbed, double L,
                                       double rho, double kappa, double depth, double epsilon);
/* L/Export.cuh:66-67, L/Export.cu:665-739: implicit Gauss-Legendre-2 evolution of the helium film.  initialState = [x | y | phi];
 * t0, t1 and stepSize of the options are used as given (the reference does not nondimensionalise them, :696); *statesOut =
 * statesCount x 3N doubles starting with the initial state, *timesOut = timesCount doubles; returnTrajectory = false: the final
 * state, no times.  Any N >= 2 (the reference: 32 ... 1024).  Release with integrateSimulationGL2_freeMemory. */
RB_API int integrateSimulationGL2(double* initialState, double** statesOut, size_t* statesCount, double** timesOut,
                                  size_t* timesCount, SimProperties* simProperties, GaussLegendreOptions* glCOptions, size_t N);
RB_API int integrateSimulationGL2_freeMemory(double* statesOut, double* timesOut);

typedef struct COptomechanicalVariables {   /* L/ExportTypes.cuh:41-57 (SI / laboratory units) */
    double detuning, gamma, G, tau, max_intensity, initial_time, location_x0_mode, sigma_optical_mode, beta, damping_strength;
} COptomechanicalVariables;
/* L/Export.cuh:78, L/Export.cu:1108-1209: one RHS of the augmented system.  state = [x | y | phi | D] (4N doubles),
 * rhs = [vx | vy | dphi/dt | dD/dt]; SI properties and variables are nondimensionalised inside (adimensionalizeProperties,
 * adimensionalizeOptomechanicalVariables L/Export.cu:1222-1275).  Any N >= 2 (the reference: 32..8192 by switch table). */
RB_API int calculateRhsAugmentedOptomechanical(double* state, double* rhs, SimProperties* simProperties,
                                               COptomechanicalVariables* optomechanicalVariables, size_t N);
/* L/Export.cuh:75-76, L/Export.cu:980-1106: RK4 evolution of the augmented system.  initialState = [x | y | phi | D];
 * *statesOut = statesCount x 4N doubles [x | y | phi | D] (the reference writes phi a second time into the D block,
 * L/Export.cu:1046 -- here the block holds D), *timesOut = timesCount doubles; with returnTrajectory = false only the final
 * state is returned.  Release with the matching _freeMemory. */
RB_API int integrateAugmentedOptomechanicalSimulationRK4(double* initialState, double** statesOut, size_t* statesCount,
                                                         double** timesOut, size_t* timesCount, SimProperties* simProperties,
                                                         RK4SolverOptions* rkOptions,
                                                         COptomechanicalVariables* optomechanicalVariables, size_t N);
RB_API int integrateAugmentedOptomechanicalSimulationRK4_freeMemory(double* statesOut, double* timesOut);
/* L/Export.cuh:72-73, L/Export.cu:779-975: RK4 evolution with the explicitly time-dependent drive.  initialState = [x | y | phi];
 * *statesOut = statesCount x 3N doubles; the logged time of a state is the time at the START of the step that produced it
 * (L/RK4_Time_Dependent.cuh:318-324); returnTrajectory = false: the final state, no times.  Any N >= 2 (the reference returns 0
 * without doing anything for an N outside its switch table, L/Export.cu:970). */
RB_API int integrateOptomechanicalSimulationRK4(double* initialState, double** statesOut, size_t* statesCount, double** timesOut,
                                                size_t* timesCount, SimProperties* simProperties, RK4SolverOptions* rkOptions,
                                                COptomechanicalVariables* optomechanicalVariables, size_t N);
RB_API int integrateOptomechanicalSimulationRK4_freeMemory(double* statesOut, double* timesOut);
/* nondimensional variant (no SI conversion): [Z | Phi | D] as 4N doubles in, final state out */
RB_API int rb_integrate_aug_rk4_host(const double* initialState_host, double* finalState_host, size_t N, const rb_props* props,
                                     const rb_opto* v, double dt, size_t steps);

#ifdef __cplusplus
}
#endif
#endif /* ROBERTS_B200_H */
