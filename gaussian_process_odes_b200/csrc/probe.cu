// Measurement utility (not part of the reference's API): sustained FP32 FMA rate of this GPU, the denominator of the
// roofline bench.py reports for the FMA-bound integrator kernels (MEASURED_PEAKS.json has no FP32 entry).
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456f) out[0] = s;  // never true; keeps the loop alive
}
}  // namespace

extern "C" int gpode_probe_fp32_fma(double* tflops_out, double* ms_out, float* scratch, void* stream) {
    GPODE_CHECK_ARG(tflops_out && ms_out && scratch, "NULL argument");
    int sms = 148, dev = 0;
    GPODE_CUDA(cudaGetDevice(&dev));
    GPODE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int iters = 1 << 15, grid = sms * 8, threads = 256;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    GPODE_CUDA(cudaEventCreate(&e0));
    GPODE_CUDA(cudaEventCreate(&e1));
    fma_probe_kernel<<<grid, threads, 0, st>>>(scratch, iters, 1.000001f, 1e-7f);  // warm-up
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        GPODE_CUDA(cudaEventRecord(e0, st));
        fma_probe_kernel<<<grid, threads, 0, st>>>(scratch, iters, 1.000001f, 1e-7f);
        GPODE_CUDA(cudaEventRecord(e1, st));
        GPODE_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        GPODE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 16.0 * (double)iters * (double)grid * (double)threads;
    *ms_out = best;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    GPODE_LAUNCH_CHECK();
    return 0;
}
