// probes.cu -- measurement entry points: FP64 pipe probes and the sweep timed alone (rb_bench_sweep), used by bench.py and tests/gpu_*.py
#include "host.cuh"

extern "C" {

// ---- measurement -------------------------------------------------------------------------------
int rb_measure_fp64_peak(double* tflops_out, void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    launch_fp64_peak(sink, 256, blocks, st);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        RB_CUDA(cudaEventRecord(e0, st));
        launch_fp64_peak(sink, iters, blocks, st);
        RB_CUDA(cudaEventRecord(e1, st));
        RB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    double flops = (double)blocks * 256.0 * iters * 64.0 * 2.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

// DMMA (m8n8k4, 512 flop per warp instruction) beside DFMA (64 flop per warp instruction): out[2*i] = ms, out[2*i+1] = TFLOP/s of
// mix i in {8 mma, 32 fma, 8+32, 4+32, 2+32, 1+32} per loop iteration
int rb_measure_fp64_tensor_overlap(double out_host[12], void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 2048;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    const int mixes[6][2] = {{8, 0}, {0, 32}, {8, 32}, {4, 32}, {2, 32}, {1, 32}};
    for (int m = 0; m < 6; ++m) {
        launch_fp64_mix(sink, 64, blocks, mixes[m][0], mixes[m][1], st);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            RB_CUDA(cudaEventRecord(e0, st));
            launch_fp64_mix(sink, iters, blocks, mixes[m][0], mixes[m][1], st);
            RB_CUDA(cudaEventRecord(e1, st));
            RB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        const double warps = (double)blocks * 8.0;
        const double flops = warps * iters * (mixes[m][0] * 512.0 + mixes[m][1] * 64.0);
        out_host[2 * m] = best;
        out_host[2 * m + 1] = flops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

int rb_measure_fp64_rate_3operand(double* tflops_out, void* stream) {
    RB_TRY
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = dmalloc<double>(1);
    int sms = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    launch_fp64_peak3(sink, 256, blocks, 1e-9, st);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        RB_CUDA(cudaEventRecord(e0, st));
        launch_fp64_peak3(sink, iters, blocks, 1e-9, st);
        RB_CUDA(cudaEventRecord(e1, st));
        RB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    *tflops_out = (double)blocks * 256.0 * iters * 64.0 * 2.0 / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RB_CATCH
}

int rb_bench_sweep(rb_solver* s, const rb_complex* state_dev, int reps, float* ms_per_sweep_out, double* pairs_per_sweep_out) {
    RB_TRY
    cudaStream_t st = s->stream;
    const double2* Z = (const double2*)state_dev;
    surface_stage(s, Z, Z + s->BN);
    launch_guess(s->b, nullptr, HistoryRing(), s->xbuf[0], s->xsum_part[0], s->bnorm_part, s->ctrl, s->omega, s->N, s->batch,
                 s->ncell, st);
    SweepArgs base = base_args(s, Z);
    base.max_iters = 1 << 30;
    base.tol2 = 0.0;
    for (int i = 0; i < 3; ++i) launch_mv(s, base, i, 0);
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0));
    RB_CUDA(cudaEventCreate(&e1));
    RB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) launch_mv(s, base, i + 1, 0);
    RB_CUDA(cudaEventRecord(e1, st));
    RB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_sweep_out) *ms_per_sweep_out = ms / reps;
    if (pairs_per_sweep_out) *pairs_per_sweep_out = (double)s->N * s->N * s->batch * (s->has_image ? 2.0 : 1.0);
    RB_CATCH
}

}  // extern "C"
