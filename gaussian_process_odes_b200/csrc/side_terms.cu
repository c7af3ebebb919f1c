// ELBO side terms either side of the integrator: the full-rank Gaussian state
// posteriors and the (projected) Gaussian observation log-likelihood.
//
// Reference arithmetic replaced:
//  * src/core/states.py:69-74,91-92,177-182,199-204 -- MultivariateNormal(mean, L L^T + 1e-5 I).rsample / .entropy
//    with L scattered from its packed lower triangle by src/misc/transforms.py:70-76,105-112: per matrix a D x D
//    Cholesky, a matrix-vector product per sample, a log-determinant; autograd through all of it. In PyTorch that is
//    ~50 batched-library launches per ELBO step (magma gemm, potrf_batch, trsm_batch, ...); here one thread owns one
//    D x D matrix in registers.
//  * src/core/likelihoods.py:27-28,38-45 -- mean over all elements of the Gaussian log-density of the decoded
//    prediction, decoder = fixed affine map (src/misc/mocap_utils.py:24-34). One warp per observation row, lanes over
//    observed dimensions; value and gradient in one pass (the mean is a scalar, so backward is a scale).
#include "common.cuh"
#include <math.h>

namespace {

// ---- D x D helpers (fully unrolled, registers) -------------------------------------------------------------------
template <int D>
struct Tri {
    static constexpr int P = D * (D + 1) / 2;
    __device__ static __forceinline__ int idx(int i, int j) { return i * (i + 1) / 2 + j; }  // row-major tril packing
};

// C = chol(L L^T + jitter I), L given packed
template <int D>
__device__ __forceinline__ void chol_from_packed(const float (&Lp)[Tri<D>::P], const float jitter, float (&C)[D][D]) {
    float A[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float s = (i == j) ? jitter : 0.f;
#pragma unroll
            for (int k = 0; k <= j; ++k) s = fmaf(Lp[Tri<D>::idx(i, k)], Lp[Tri<D>::idx(j, k)], s);
            A[i][j] = s;
        }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float s = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-C[j][k], C[j][k], s);
        const float cjj = sqrtf(s);
        C[j][j] = cjj;
        const float inv = 1.0f / cjj;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            float t = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fmaf(-C[i][k], C[j][k], t);
            C[i][j] = t * inv;
        }
#pragma unroll
        for (int i = 0; i < j; ++i) C[i][j] = 0.f;
    }
}

// Cb (lower) -> gradient w.r.t. the packed L:  Ab = sym(C^-T Phi(C^T Cb) C^-1),  Lb = tril(2 Ab L)
template <int D>
__device__ __forceinline__ void chol_backward_to_packed(const float (&Lp)[Tri<D>::P], const float (&C)[D][D],
                                                        const float (&Cb)[D][D], float (&gLp)[Tri<D>::P]) {
    float Pm[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float s = 0.f;
            if (j <= i) {
#pragma unroll
                for (int l = i; l < D; ++l) s = fmaf(C[l][i], Cb[l][j], s);  // (C^T Cb)_ij, Cb lower => l >= i >= j
                if (i == j) s *= 0.5f;
            }
            Pm[i][j] = s;
        }
    // X = C^-T P : back substitution over rows
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        const float inv = 1.0f / C[i][i];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float s = Pm[i][j];
#pragma unroll
            for (int l = i + 1; l < D; ++l) s = fmaf(-C[l][i], Pm[l][j], s);
            Pm[i][j] = s * inv;
        }
    }
    // Y = X C^-1 : columns from the right
#pragma unroll
    for (int c = D - 1; c >= 0; --c) {
        const float inv = 1.0f / C[c][c];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float s = Pm[i][c];
#pragma unroll
            for (int l = c + 1; l < D; ++l) s = fmaf(-Pm[i][l], C[l][c], s);
            Pm[i][c] = s * inv;
        }
    }
    // Lb_ij = sum_k (Y + Y^T)_ik L_kj, i >= j
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = j; k < D; ++k) s = fmaf(Pm[i][k] + Pm[k][i], Lp[Tri<D>::idx(k, j)], s);
            gLp[Tri<D>::idx(i, j)] = s;
        }
}

// mode bit 0: samples, bit 1: entropy
template <int D>
__global__ void __launch_bounds__(128)
state_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ Lpk, const float* __restrict__ eps,
                 const int S, const int64_t R, const float jitter, float* __restrict__ samples,
                 float* __restrict__ entropy) {
    constexpr int P = Tri<D>::P;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float Lp[P], C[D][D];
#pragma unroll
    for (int i = 0; i < P; ++i) Lp[i] = __ldg(Lpk + r * P + i);
    chol_from_packed<D>(Lp, jitter, C);
    if (entropy != nullptr) {
        float h = 0.5f * D * (1.0f + 1.8378770664093453f);  // D/2 (1 + log 2 pi)
#pragma unroll
        for (int i = 0; i < D; ++i) h += logf(C[i][i]);
        entropy[r] = h;
    }
    if (samples != nullptr) {
        float m[D];
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = __ldg(mean + r * D + i);
        for (int s = 0; s < S; ++s) {
            float e[D];
#pragma unroll
            for (int i = 0; i < D; ++i) e[i] = __ldg(eps + ((int64_t)s * R + r) * D + i);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float y = m[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) y = fmaf(C[i][j], e[j], y);
                samples[((int64_t)s * R + r) * D + i] = y;
            }
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128)
state_bwd_kernel(const float* __restrict__ Lpk, const float* __restrict__ eps, const int S, const int64_t R,
                 const float jitter, const float* __restrict__ g_samples, const float* __restrict__ g_entropy,
                 float* __restrict__ g_mean, float* __restrict__ g_Lpk) {
    constexpr int P = Tri<D>::P;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float Lp[P], C[D][D], Cb[D][D], gm[D];
#pragma unroll
    for (int i = 0; i < P; ++i) Lp[i] = __ldg(Lpk + r * P + i);
    chol_from_packed<D>(Lp, jitter, C);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        gm[i] = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) Cb[i][j] = 0.f;
    }
    if (g_samples != nullptr) {
        for (int s = 0; s < S; ++s) {
            float e[D], gy[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                e[i] = __ldg(eps + ((int64_t)s * R + r) * D + i);
                gy[i] = __ldg(g_samples + ((int64_t)s * R + r) * D + i);
                gm[i] += gy[i];
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) Cb[i][j] = fmaf(gy[i], e[j], Cb[i][j]);
        }
    }
    if (g_entropy != nullptr) {
        const float gh = __ldg(g_entropy + r);
#pragma unroll
        for (int i = 0; i < D; ++i) Cb[i][i] += gh / C[i][i];
    }
    float gL[P];
    chol_backward_to_packed<D>(Lp, C, Cb, gL);
#pragma unroll
    for (int i = 0; i < P; ++i) g_Lpk[r * P + i] = gL[i];
    if (g_mean != nullptr) {
#pragma unroll
        for (int i = 0; i < D; ++i) g_mean[r * D + i] = gm[i];
    }
}

// ---- mean Gaussian log-likelihood of decoded predictions, value + gradients in one pass ------------------------------
// pred [S, R, D], ys [R, Dobs], W [D, Dobs], bias [Dobs] (may be NULL), var [Dobs].
// out[0] += sum log N(y | pred W + b, var);  g_pred [S,R,D] = d(sum)/d pred;  g_var [Dobs] += d(sum)/d var.
constexpr int kLlMaxD = GPODE_MAX_D;
constexpr int kLlPasses = 4;  // Dobs <= 128

__global__ void __launch_bounds__(256)
loglik_kernel(const float* __restrict__ pred, const float* __restrict__ ys, const float* __restrict__ W,
              const float* __restrict__ bias, const float* __restrict__ var, const int S, const int64_t R, const int D,
              const int Dobs, double* __restrict__ work, float* __restrict__ g_pred) {
    __shared__ double s_sum[8];
    __shared__ float s_gv[8][32 * kLlPasses];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float w[kLlPasses][kLlMaxD], b[kLlPasses], iv[kLlPasses], lv[kLlPasses], gv[kLlPasses];
#pragma unroll
    for (int p = 0; p < kLlPasses; ++p) {
        const int d = p * 32 + lane;
        const bool ok = d < Dobs;
        b[p] = (ok && bias) ? bias[d] : 0.f;
        const float v = ok ? var[d] : 1.f;
        iv[p] = 1.0f / v;
        lv[p] = ok ? (1.8378770664093453f + logf(v)) : 0.f;  // log(2 pi) + log var
        gv[p] = 0.f;
#pragma unroll
        for (int l = 0; l < kLlMaxD; ++l) w[p][l] = (ok && l < D) ? W[l * Dobs + d] : 0.f;
    }
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < R; r += (int64_t)gridDim.x * nwarps) {
        float y[kLlPasses];
#pragma unroll
        for (int p = 0; p < kLlPasses; ++p) {
            const int d = p * 32 + lane;
            y[p] = d < Dobs ? __ldg(ys + r * Dobs + d) : 0.f;
        }
        float local = 0.f;
        constexpr int kSC = 4;  // samples per chunk: their (independent) loads are all issued before any use
        for (int s0 = 0; s0 < S; s0 += kSC) {
            float x[kSC][kLlMaxD];
#pragma unroll
            for (int c = 0; c < kSC; ++c) {
                const int sc = s0 + c < S ? s0 + c : S - 1;
                const float* xr = pred + ((int64_t)sc * R + r) * D;
#pragma unroll
                for (int l = 0; l < kLlMaxD; ++l) x[c][l] = l < D ? __ldg(xr + l) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < kSC; ++c) {
                if (s0 + c < S) {
                    const int s = s0 + c;
                    float gx[kLlMaxD];
#pragma unroll
                    for (int l = 0; l < kLlMaxD; ++l) gx[l] = 0.f;
#pragma unroll
                    for (int p = 0; p < kLlPasses; ++p) {
                        if (p * 32 < Dobs) {
                            float f = b[p];
#pragma unroll
                            for (int l = 0; l < kLlMaxD; ++l) f = fmaf(x[c][l], w[p][l], f);
                            const bool ok = p * 32 + lane < Dobs;
                            const float diff = ok ? f - y[p] : 0.f;
                            const float q = diff * iv[p];
                            local += ok ? -0.5f * (lv[p] + diff * q) : 0.f;
                            gv[p] += ok ? -0.5f * (iv[p] - q * q) : 0.f;
#pragma unroll
                            for (int l = 0; l < kLlMaxD; ++l) gx[l] = fmaf(-q, w[p][l], gx[l]);
                        }
                    }
                    if (g_pred != nullptr) {
#pragma unroll
                        for (int l = 0; l < kLlMaxD; ++l) {
                            if (l < D) {
                                const float v = gpode_warp_sum(gx[l]);
                                if (lane == 0) g_pred[((int64_t)s * R + r) * D + l] = v;
                            }
                        }
                    }
                }
            }
        }
        acc += (double)local;
    }
    // this CTA's partial sums -> row blockIdx.x of `work` ([1 + Dobs] float64: log-density | variance gradients); the
    // rows are added in row order by side_sum_kernel (no atomics: bitwise reproducible)
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_sum[warp] = acc;
#pragma unroll
    for (int p = 0; p < kLlPasses; ++p) s_gv[warp][p * 32 + lane] = gv[p];
    __syncthreads();
    double* __restrict__ row = work + (size_t)blockIdx.x * (1 + Dobs);
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < nwarps; ++i) t += s_sum[i];
        row[0] = t;
    }
    for (int d = threadIdx.x; d < Dobs; d += blockDim.x) {
        double t = 0.0;
        for (int i = 0; i < nwarps; ++i) t += (double)s_gv[i][d];
        row[1 + d] = t;
    }
}

// rows of float64 partial sums -> out[0] (float64) and, optionally, g[0..cols-1) (float32); one CTA
constexpr int kSideCap = 2048;                  // most CTAs (= rows of `work`) any side-term kernel launches
constexpr int kSideCols = 1 + 32 * kLlPasses;   // widest row
__global__ void __launch_bounds__(256)
side_sum_kernel(const double* __restrict__ work, const int n_rows, const int cols, double* __restrict__ out,
                float* __restrict__ g) {
    __shared__ double part[8 * kSideCols];
    __shared__ double tot[kSideCols];
    gpode_sum_rows_ordered<(kSideCols + 31) / 32>(work, (size_t)cols, n_rows, cols, kSideCols, part, tot);
    if (threadIdx.x == 0) out[0] = tot[0];
    if (g != nullptr)
        for (int i = threadIdx.x; i + 1 < cols; i += blockDim.x) g[i] = (float)tot[1 + i];
}

// ---- shooting-constraint term: sum over (s, n, t < T-1, d) of log N(ss[s,n,t+1,d] | pred[s,n,t,d], scale) ----------
// ss, pred: [SN, T, D] (SN = S*N sequences). One thread per element of the (SN, T-1, D) index space, grid-stride;
// value and both gradients in one pass (the result is a scalar, so backward is a scale).
__global__ void __launch_bounds__(256)
constraint_kernel(const float* __restrict__ ss, const float* __restrict__ pred, const float* __restrict__ scale_p,
                  const int64_t SN, const int T, const int D, const int laplace, double* __restrict__ work,
                  float* __restrict__ g_ss, float* __restrict__ g_pred) {
    __shared__ double s_sum[8];
    const float scale = scale_p[0];
    const float inv = 1.0f / scale, inv2 = inv * inv;
    const float c0 = laplace ? -logf(2.0f * scale) : -logf(scale) - 0.9189385332046727f;  // -log s - 0.5 log 2 pi
    const int64_t per_seq = (int64_t)(T - 1) * D;
    const int64_t total = SN * per_seq;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t seq = i / per_seq, rem = i - seq * per_seq;  // rem = t*D + d, t < T-1
        const int64_t ip = seq * (int64_t)T * D + rem;             // pred[seq, t, d]
        const int64_t is = ip + D;                                 // ss[seq, t+1, d]
        const float diff = __ldg(ss + is) - __ldg(pred + ip);
        float lp, g;
        if (laplace) {
            lp = c0 - fabsf(diff) * inv;
            g = diff > 0.f ? -inv : (diff < 0.f ? inv : 0.f);      // d lp / d ss
        } else {
            lp = c0 - 0.5f * diff * diff * inv2;
            g = -diff * inv2;
        }
        acc += (double)lp;
        if (g_ss != nullptr) g_ss[is] = g;
        if (g_pred != nullptr) g_pred[ip] = -g;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_sum[i];
        work[blockIdx.x] = t;  // added in row order by side_sum_kernel
    }
}

int check_states(int D, int S, int64_t R) {
    GPODE_CHECK_ARG(D >= 1 && D <= GPODE_MAX_D, "state dimension D=%d outside 1..%d", D, GPODE_MAX_D);
    GPODE_CHECK_ARG(S >= 0 && R >= 0, "negative sizes S=%d R=%lld", S, (long long)R);
    return 0;
}

}  // namespace

#define GPODE_STATE_SWITCH(D_, KERNEL, ...)                                      \
    switch (D_) {                                                                \
        case 1: KERNEL<1><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 2: KERNEL<2><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 3: KERNEL<3><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 4: KERNEL<4><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 5: KERNEL<5><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 6: KERNEL<6><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 7: KERNEL<7><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
        case 8: KERNEL<8><<<grid, 128, 0, st>>>(__VA_ARGS__); break;             \
    }

extern "C" int gpode_state_fwd(const float* mean, const float* L_packed, const float* eps, int S, int64_t R, int D,
                               float jitter, float* samples_out, float* entropy_out, void* stream) {
    if (int rc = check_states(D, S, R)) return rc;
    if (R == 0) return 0;
    GPODE_CHECK_ARG(L_packed != nullptr, "L_packed is NULL");
    GPODE_CHECK_ARG(samples_out == nullptr || (mean && eps), "sampling needs mean and eps");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((R + 127) / 128);
    GPODE_STATE_SWITCH(D, state_fwd_kernel, mean, L_packed, eps, S, R, jitter, samples_out, entropy_out)
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int gpode_state_bwd(const float* L_packed, const float* eps, int S, int64_t R, int D, float jitter,
                               const float* grad_samples, const float* grad_entropy, float* grad_mean,
                               float* grad_L_packed, void* stream) {
    if (int rc = check_states(D, S, R)) return rc;
    if (R == 0) return 0;
    GPODE_CHECK_ARG(L_packed && grad_L_packed, "NULL argument");
    GPODE_CHECK_ARG(grad_samples == nullptr || eps != nullptr, "sample gradient needs eps");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((R + 127) / 128);
    GPODE_STATE_SWITCH(D, state_bwd_kernel, L_packed, eps, S, R, jitter, grad_samples, grad_entropy, grad_mean,
                       grad_L_packed)
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int gpode_loglik_sum(const float* pred, const float* ys, const float* W, const float* bias,
                                const float* var, int S, int64_t R, int D, int D_obs, double* sum_out,
                                float* grad_pred, float* grad_var, double* work, void* stream) {
    GPODE_CHECK_ARG(D >= 1 && D <= GPODE_MAX_D, "latent dimension D=%d outside 1..%d", D, GPODE_MAX_D);
    GPODE_CHECK_ARG(D_obs >= 1 && D_obs <= 32 * kLlPasses, "observed dimension %d outside 1..%d", D_obs, 32 * kLlPasses);
    GPODE_CHECK_ARG(S >= 1 && R >= 0, "bad sizes S=%d R=%lld", S, (long long)R);
    GPODE_CHECK_ARG(pred && ys && W && var && sum_out && work, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (R == 0) {
        GPODE_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double), st));
        if (grad_var) GPODE_CUDA(cudaMemsetAsync(grad_var, 0, sizeof(float) * D_obs, st));
        return 0;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (R + 7) / 8;
    int64_t cap = (int64_t)sms * 8;
    if (cap > kSideCap) cap = kSideCap;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    loglik_kernel<<<grid, 256, 0, st>>>(pred, ys, W, bias, var, S, R, D, D_obs, work, grad_pred);
    side_sum_kernel<<<1, 256, 0, st>>>(work, (int)grid, 1 + D_obs, sum_out, grad_var);
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int gpode_constraint_sum(const float* ss, const float* pred, const float* scale, int64_t SN, int T, int D,
                                    int laplace, double* sum_out, float* grad_ss, float* grad_pred, double* work,
                                    void* stream) {
    GPODE_CHECK_ARG(ss && pred && scale && sum_out && work, "NULL argument");
    GPODE_CHECK_ARG(SN >= 0 && T >= 1 && D >= 1, "bad sizes SN=%lld T=%d D=%d", (long long)SN, T, D);
    cudaStream_t st = (cudaStream_t)stream;
    GPODE_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double), st));
    const int64_t n_all = SN * (int64_t)T * D;
    if (grad_ss) GPODE_CUDA(cudaMemsetAsync(grad_ss, 0, sizeof(float) * n_all, st));      // t = 0 slots stay zero
    if (grad_pred) GPODE_CUDA(cudaMemsetAsync(grad_pred, 0, sizeof(float) * n_all, st));  // t = T-1 slots stay zero
    const int64_t total = SN * (int64_t)(T - 1) * D;
    if (total == 0) return 0;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (total + 255) / 256;
    int64_t cap = (int64_t)sms * 8;
    if (cap > kSideCap) cap = kSideCap;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    constraint_kernel<<<grid, 256, 0, st>>>(ss, pred, scale, SN, T, D, laplace, work, grad_ss, grad_pred);
    side_sum_kernel<<<1, 256, 0, st>>>(work, (int)grid, 1, sum_out, nullptr);
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t gpode_side_work_doubles(void) { return (int64_t)kSideCap * kSideCols; }
