// One translation unit per state dimension (compiled with -DGPODE_D=<1..8>).
#include "dopri5_impl.cuh"

#ifndef GPODE_D
#error "compile with -DGPODE_D=<state dimension>"
#endif
#define GPODE_CAT_(a, b) a##b
#define GPODE_CAT(a, b) GPODE_CAT_(a, b)

int GPODE_CAT(gpode_dopri5_fwd_d, GPODE_D)(const float* packed, int M, int S, const float* x0, const double* t, int Tg,
                                           int64_t B, double rtol, double atol, float* xs, float* work,
                                           int32_t* stats, float* ckpt, int cap, cudaStream_t st) {
    return launch_dopri5<GPODE_D>(packed, M, S, x0, t, Tg, B, rtol, atol, xs, work, stats, ckpt, cap, st);
}

int GPODE_CAT(gpode_dopri5_bwd_d, GPODE_D)(const float* packed, int M, int S, const double* t, int Tg, int64_t B,
                                           const float* gxs, const float* ckpt, int cap, int n_acc,
                                           const int32_t* stats_dev, float* gx0, float* vrows, float* acc,
                                           cudaStream_t st) {
    return launch_dopri5_bwd<GPODE_D>(packed, M, S, t, Tg, B, gxs, ckpt, cap, n_acc, stats_dev, gx0, vrows, acc, st);
}

int GPODE_CAT(gpode_dopri5_sets_d, GPODE_D)(const float* packed, int M, int S, int n_sets, int64_t set_rows,
                                            const float* x0, const double* t, int Tg, double rtol, double atol,
                                            float* xs, float* work, int32_t* stats, cudaStream_t st) {
    return launch_dopri5_sets<GPODE_D>(packed, M, S, n_sets, set_rows, x0, t, Tg, rtol, atol, xs, work, stats, st);
}
