// C-ABI entry points of the integrator path; dispatch on the state dimension to the per-D translation units.
#include "common.cuh"

#define GPODE_FOR_EACH_D(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)

#define GPODE_DECL(D_)                                                                                              \
    int gpode_vf_fwd_d##D_(const float*, int, int, const float*, float*, int64_t, cudaStream_t);                    \
    int gpode_rk4_fwd_d##D_(const float*, int, int, const float*, const float*, int, int64_t, float*, float*,       \
                            cudaStream_t);                                                                          \
    int gpode_rk4_bwd_d##D_(const float*, int, int, const float*, int, int64_t, const float*, const float*,         \
                            const float*, float*, float*, float*, float*, cudaStream_t);                            \
    int gpode_vf_bwd_d##D_(const float*, int, int, const float*, const float*, const float*, float*, int64_t,       \
                           float*, cudaStream_t);                                                                   \
    int gpode_fwd_sets_d##D_(const float*, int, int, int, int64_t, const float*, const float*, int, float*,         \
                             cudaStream_t);
GPODE_FOR_EACH_D(GPODE_DECL)
#undef GPODE_DECL

int gpode_param_grad_launch(const float* packed, int D, int M, int S, const float* ys, const float* kbs, int64_t VR,
                            float* acc, cudaStream_t stream, const int32_t* stats_dev = nullptr,
                            int64_t rows_per_step = 0);  // param_grad.cu

static int check_common(const float* packed, int D, int M, int S, int64_t B) {
    GPODE_CHECK_ARG(packed != nullptr, "packed parameter block is NULL");
    GPODE_CHECK_ARG(D >= 1 && D <= GPODE_MAX_D, "state dimension D=%d outside 1..%d", D, GPODE_MAX_D);
    GPODE_CHECK_ARG(M >= 1 && S >= 1, "M=%d and S=%d must be positive", M, S);
    GPODE_CHECK_ARG(B >= 0, "negative batch size %lld", (long long)B);
    return 0;
}

#define GPODE_SWITCH_D(D_, CALL)                                   \
    switch (D_) {                                                  \
        case 1: return CALL(1);                                    \
        case 2: return CALL(2);                                    \
        case 3: return CALL(3);                                    \
        case 4: return CALL(4);                                    \
        case 5: return CALL(5);                                    \
        case 6: return CALL(6);                                    \
        case 7: return CALL(7);                                    \
        case 8: return CALL(8);                                    \
        default: break;                                            \
    }                                                              \
    gpode_set_error("state dimension D=%d outside 1..%d", D_, GPODE_MAX_D); \
    return -1;

extern "C" int gpode_vf_fwd(const float* packed, int D, int M, int S, const float* x, float* f, int64_t B,
                            void* stream) {
    if (int rc = check_common(packed, D, M, S, B)) return rc;
    if (B == 0) return 0;
    GPODE_CHECK_ARG(x && f, "x / f is NULL");
#define CALL(D_) gpode_vf_fwd_d##D_(packed, M, S, x, f, B, (cudaStream_t)stream)
    GPODE_SWITCH_D(D, CALL)
#undef CALL
}

extern "C" int gpode_rk4_fwd(const float* packed, int D, int M, int S, const float* x0, const float* t, int Tg,
                             int64_t B, float* xs, float* kstages, void* stream) {
    if (int rc = check_common(packed, D, M, S, B)) return rc;
    GPODE_CHECK_ARG(Tg >= 1, "time grid needs at least one point, got %d", Tg);
    if (B == 0) return 0;
    GPODE_CHECK_ARG(x0 && t && xs, "x0 / t / xs is NULL");
#define CALL(D_) gpode_rk4_fwd_d##D_(packed, M, S, x0, t, Tg, B, xs, kstages, (cudaStream_t)stream)
    GPODE_SWITCH_D(D, CALL)
#undef CALL
}

static int fwd_sets_dispatch(const float* packed, int D, int M, int S, int n_sets, int64_t set_rows, const float* x0,
                             const float* t, int Tg, float* out, cudaStream_t st) {
#define CALL(D_) gpode_fwd_sets_d##D_(packed, M, S, n_sets, set_rows, x0, t, Tg, out, st)
    GPODE_SWITCH_D(D, CALL)
#undef CALL
}

static int check_sets(int n_sets, int64_t set_rows) {
    GPODE_CHECK_ARG(n_sets >= 1 && n_sets <= 65535, "n_sets=%d outside 1..65535", n_sets);
    GPODE_CHECK_ARG(set_rows >= 0, "negative rows per set %lld", (long long)set_rows);
    return 0;
}

extern "C" int gpode_vf_fwd_sets(const float* packed, int D, int M, int S, int n_sets, int64_t set_rows,
                                 const float* x, float* f, void* stream) {
    if (int rc = check_common(packed, D, M, S, set_rows)) return rc;
    if (int rc = check_sets(n_sets, set_rows)) return rc;
    if (set_rows == 0) return 0;
    GPODE_CHECK_ARG(x && f, "x / f is NULL");
    return fwd_sets_dispatch(packed, D, M, S, n_sets, set_rows, x, nullptr, 0, f, (cudaStream_t)stream);
}

extern "C" int gpode_rk4_fwd_sets(const float* packed, int D, int M, int S, int n_sets, int64_t set_rows,
                                  const float* x0, const float* t, int Tg, float* xs, void* stream) {
    if (int rc = check_common(packed, D, M, S, set_rows)) return rc;
    if (int rc = check_sets(n_sets, set_rows)) return rc;
    GPODE_CHECK_ARG(Tg >= 1, "time grid needs at least one point, got %d", Tg);
    if (set_rows == 0) return 0;
    GPODE_CHECK_ARG(x0 && t && xs, "x0 / t / xs is NULL");
    return fwd_sets_dispatch(packed, D, M, S, n_sets, set_rows, x0, t, Tg, xs, (cudaStream_t)stream);
}

static int rk4_bwd_dispatch(const float* packed, int D, int M, int S, const float* t, int Tg, int64_t B,
                            const float* xs, const float* kst, const float* gxs, float* gx0, float* vy, float* vk,
                            float* acc, cudaStream_t st) {
#define CALL(D_) gpode_rk4_bwd_d##D_(packed, M, S, t, Tg, B, xs, kst, gxs, gx0, vy, vk, acc, st)
    GPODE_SWITCH_D(D, CALL)
#undef CALL
}

extern "C" int gpode_rk4_bwd(const float* packed, int D, int M, int S, const float* t, int Tg, int64_t B,
                             const float* xs, const float* kstages, const float* grad_xs, float* grad_x0,
                             float* vrows, float* acc, void* stream) {
    if (int rc = check_common(packed, D, M, S, B)) return rc;
    GPODE_CHECK_ARG(Tg >= 1, "time grid needs at least one point, got %d", Tg);
    if (B == 0) return 0;
    GPODE_CHECK_ARG(t && xs && grad_xs && grad_x0 && acc, "NULL argument");
    if (Tg == 1) {
        GPODE_CUDA(cudaMemcpyAsync(grad_x0, grad_xs, sizeof(float) * B * D, cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
        return 0;
    }
    GPODE_CHECK_ARG(kstages && vrows, "kstages / vrows is NULL");
    const int64_t VR = (int64_t)(Tg - 1) * 4 * B;
    float* vy = vrows;
    float* vk = vrows + VR * D;
    return rk4_bwd_dispatch(packed, D, M, S, t, Tg, B, xs, kstages, grad_xs, grad_x0, vy, vk, acc,
                            (cudaStream_t)stream);
}

static int vf_bwd_dispatch(const float* packed, int D, int M, int S, const float* x, const float* f, const float* gf,
                           float* gx, int64_t B, float* acc, cudaStream_t st) {
#define CALL(D_) gpode_vf_bwd_d##D_(packed, M, S, x, f, gf, gx, B, acc, st)
    GPODE_SWITCH_D(D, CALL)
#undef CALL
}

extern "C" int gpode_vf_bwd(const float* packed, int D, int M, int S, const float* x, const float* f,
                            const float* grad_f, float* grad_x, float* acc, int64_t B, void* stream) {
    if (int rc = check_common(packed, D, M, S, B)) return rc;
    if (B == 0) return 0;
    GPODE_CHECK_ARG(x && f && grad_f && grad_x && acc, "NULL argument");
    return vf_bwd_dispatch(packed, D, M, S, x, f, grad_f, grad_x, B, acc, (cudaStream_t)stream);
}

extern "C" int gpode_param_grad(const float* packed, int D, int M, int S, const float* ys, const float* kbs,
                                int64_t n_rows, float* acc, void* stream) {
    if (int rc = check_common(packed, D, M, S, n_rows)) return rc;
    if (n_rows == 0) return 0;
    GPODE_CHECK_ARG(ys && kbs && acc, "NULL argument");
    return gpode_param_grad_launch(packed, D, M, S, ys, kbs, n_rows, acc, (cudaStream_t)stream);
}

extern "C" int gpode_param_grad_dev(const float* packed, int D, int M, int S, const float* ys, const float* kbs,
                                    int64_t n_rows_max, const int32_t* stats_dev, int64_t rows_per_step, float* acc,
                                    void* stream) {
    if (int rc = check_common(packed, D, M, S, n_rows_max)) return rc;
    if (n_rows_max == 0) return 0;
    GPODE_CHECK_ARG(ys && kbs && acc && stats_dev, "NULL argument");
    return gpode_param_grad_launch(packed, D, M, S, ys, kbs, n_rows_max, acc, (cudaStream_t)stream, stats_dev,
                                   rows_per_step);
}

extern "C" int64_t gpode_vrow_floats(int D, int64_t n_virtual_rows) { return 2 * n_virtual_rows * (int64_t)D; }
