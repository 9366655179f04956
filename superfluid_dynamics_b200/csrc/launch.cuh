// launch.cuh -- the one place a kernel launch is spelled.  RB_LAUNCH_EW: element-wise kernels (no barriers, no shuffles);
// RB_LAUNCH: kernels whose threads cooperate.  Under nvcc both are the plain <<<grid, block, 0, stream>>> launch; under g++ with
// tests/cpp/cuda_emu.h (RB_EMULATE, CPU test tier) they run the same kernel source thread for thread on the host.
#pragma once
#ifdef RB_EMULATE
#define RB_LAUNCH(kernel, grid, block, stream, ...) emu::launch_coop(kernel, dim3(grid), dim3(block), __VA_ARGS__)
#define RB_LAUNCH_EW(kernel, grid, block, stream, ...) emu::launch_seq(kernel, dim3(grid), dim3(block), __VA_ARGS__)
#else
#define RB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define RB_LAUNCH_EW(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif
