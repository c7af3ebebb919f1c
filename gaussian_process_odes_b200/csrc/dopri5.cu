// C-ABI entry points of the adaptive integrator; dispatch on the state dimension.
#include "common.cuh"

#define GPODE_DECL(D_)                                                                                            \
    int gpode_dopri5_fwd_d##D_(const float*, int, int, const float*, const double*, int, int64_t, double, double, \
                               float*, float*, int32_t*, float*, int, cudaStream_t);                              \
    int gpode_dopri5_bwd_d##D_(const float*, int, int, const double*, int, int64_t, const float*, const float*,   \
                               int, int, const int32_t*, float*, float*, float*, cudaStream_t);                   \
    int gpode_dopri5_sets_d##D_(const float*, int, int, int, int64_t, const float*, const double*, int, double,   \
                                double, float*, float*, int32_t*, cudaStream_t);
GPODE_DECL(1) GPODE_DECL(2) GPODE_DECL(3) GPODE_DECL(4) GPODE_DECL(5) GPODE_DECL(6) GPODE_DECL(7) GPODE_DECL(8)
#undef GPODE_DECL

extern "C" int64_t gpode_dopri5_work_floats(int D, int64_t B) {
    const int64_t plane = B * (int64_t)D;
    return 5 * plane + 1 + 2 * 3 * 2048;  // five [B,D] state planes + alignment + 3 x 2048 float64 per-CTA partial sums
}

// checkpoint block: y [cap][B][D] | k [cap][7][B][D] | dt [cap] | out_x [Tg] | out_step [Tg] (int32)
extern "C" int64_t gpode_dopri5_ckpt_floats(int D, int64_t B, int Tg, int cap) {
    const int64_t plane = B * (int64_t)D;
    return (int64_t)cap * 8 * plane + cap + 2 * (int64_t)Tg;
}

static int check(const float* packed, int D, int M, int S, int64_t B, int Tg) {
    GPODE_CHECK_ARG(packed != nullptr, "packed parameter block is NULL");
    GPODE_CHECK_ARG(D >= 1 && D <= GPODE_MAX_D, "state dimension D=%d outside 1..%d", D, GPODE_MAX_D);
    GPODE_CHECK_ARG(M >= 1 && S >= 1, "M=%d and S=%d must be positive", M, S);
    GPODE_CHECK_ARG(B >= 0 && Tg >= 1, "bad sizes B=%lld Tg=%d", (long long)B, Tg);
    return 0;
}

extern "C" int gpode_dopri5_fwd(const float* packed, int D, int M, int S, const float* x0, const double* t, int Tg,
                                int64_t B, double rtol, double atol, float* xs, float* work, int32_t* stats_out,
                                float* ckpt, int cap, void* stream) {
    if (int rc = check(packed, D, M, S, B, Tg)) return rc;
    GPODE_CHECK_ARG(rtol > 0 && atol > 0, "rtol/atol must be positive");
    GPODE_CHECK_ARG(ckpt == nullptr || cap > 0, "checkpointing needs cap > 0");
    if (B == 0) return 0;
    GPODE_CHECK_ARG(x0 && t && xs && work && stats_out, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
#define GPODE_CASE(D_) \
    case D_:           \
        return gpode_dopri5_fwd_d##D_(packed, M, S, x0, t, Tg, B, rtol, atol, xs, work, stats_out, ckpt, cap, st);
        GPODE_CASE(1) GPODE_CASE(2) GPODE_CASE(3) GPODE_CASE(4) GPODE_CASE(5) GPODE_CASE(6) GPODE_CASE(7) GPODE_CASE(8)
#undef GPODE_CASE
    }
    return -1;
}

extern "C" int gpode_dopri5_bwd(const float* packed, int D, int M, int S, const double* t, int Tg, int64_t B,
                                const float* grad_xs, const float* ckpt, int cap, int n_accepted, float* grad_x0,
                                float* vrows, float* acc, void* stream) {
    if (int rc = check(packed, D, M, S, B, Tg)) return rc;
    if (B == 0) return 0;
    GPODE_CHECK_ARG(t && grad_xs && ckpt && grad_x0 && vrows && acc, "NULL argument");
    GPODE_CHECK_ARG(cap > 0 && n_accepted >= 0 && n_accepted <= cap, "n_accepted=%d outside 0..cap=%d", n_accepted, cap);
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
#define GPODE_CASE(D_) \
    case D_:           \
        return gpode_dopri5_bwd_d##D_(packed, M, S, t, Tg, B, grad_xs, ckpt, cap, n_accepted, nullptr, grad_x0, vrows, \
                                      acc, st);
        GPODE_CASE(1) GPODE_CASE(2) GPODE_CASE(3) GPODE_CASE(4) GPODE_CASE(5) GPODE_CASE(6) GPODE_CASE(7) GPODE_CASE(8)
#undef GPODE_CASE
    }
    return -1;
}

extern "C" int gpode_dopri5_fwd_sets(const float* packed, int D, int M, int S, int n_sets, int64_t set_rows,
                                     const float* x0, const double* t, int Tg, double rtol, double atol, float* xs,
                                     float* work, int32_t* stats_out, void* stream) {
    if (int rc = check(packed, D, M, S, set_rows, Tg)) return rc;
    GPODE_CHECK_ARG(n_sets >= 1 && n_sets <= 65535, "n_sets=%d outside 1..65535", n_sets);
    GPODE_CHECK_ARG(rtol > 0 && atol > 0, "rtol/atol must be positive");
    if (set_rows == 0) return 0;
    GPODE_CHECK_ARG(x0 && t && xs && work && stats_out, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
#define GPODE_CASE(D_) \
    case D_:           \
        return gpode_dopri5_sets_d##D_(packed, M, S, n_sets, set_rows, x0, t, Tg, rtol, atol, xs, work, stats_out, st);
        GPODE_CASE(1) GPODE_CASE(2) GPODE_CASE(3) GPODE_CASE(4) GPODE_CASE(5) GPODE_CASE(6) GPODE_CASE(7) GPODE_CASE(8)
#undef GPODE_CASE
    }
    return -1;
}

extern "C" int gpode_dopri5_bwd_dev(const float* packed, int D, int M, int S, const double* t, int Tg, int64_t B,
                                    const float* grad_xs, const float* ckpt, int cap, const int32_t* stats_dev,
                                    float* grad_x0, float* vrows, float* acc, void* stream) {
    if (int rc = check(packed, D, M, S, B, Tg)) return rc;
    if (B == 0) return 0;
    GPODE_CHECK_ARG(t && grad_xs && ckpt && stats_dev && grad_x0 && vrows && acc, "NULL argument");
    GPODE_CHECK_ARG(cap > 0, "cap=%d must be positive", cap);
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
#define GPODE_CASE(D_) \
    case D_:           \
        return gpode_dopri5_bwd_d##D_(packed, M, S, t, Tg, B, grad_xs, ckpt, cap, 0, stats_dev, grad_x0, vrows, acc, st);
        GPODE_CASE(1) GPODE_CASE(2) GPODE_CASE(3) GPODE_CASE(4) GPODE_CASE(5) GPODE_CASE(6) GPODE_CASE(7) GPODE_CASE(8)
#undef GPODE_CASE
    }
    return -1;
}
