// Per-inducing-point gradient contraction ("phase B" of the backward).
//
// For the RBF pathwise-update term f_upd,k(x) = var_k sum_m nu_km K_km(x) the gradients w.r.t. nu and Z contract over
// ROWS, not over inducing points, so they cannot be reduced inside the row-per-thread adjoint kernel without M*D
// cross-thread reductions per row. Instead the adjoint kernel writes (stage input y, cotangent kb) pairs -- "virtual
// rows" -- and this kernel maps one thread to one (output dim k, inducing point m) pair and streams the virtual rows
// through shared memory (broadcast reads), accumulating in registers
//     T[k][m]    += kb_k K_km                      -> grad nu_km = var_k T
//     W[k][m][j] += kb_k K_km (y_j - Z_mj)         -> grad Z_mj  = sum_k c_km W / ell_kj^2
// ("kernel" VJP lines; reference arithmetic: autograd through src/core/kernels.py:53-99
// and src/core/dsvgp.py:192).
#include "vf.cuh"

namespace {

constexpr int kPgThreads = 256;
constexpr int kPgTile = 128;  // virtual rows staged per pass

template <int D>
__global__ void __launch_bounds__(kPgThreads)
param_grad_kernel(const float* __restrict__ packed, const int M, const int S, const float* __restrict__ ys,
                  const float* __restrict__ kbs, const int64_t VR, const int64_t rows_per_cta,
                  float* __restrict__ acc) {
    constexpr int RS = VfShape<D>::RS, KS = VfShape<D>::KS, DP = VfShape<D>::DP;
    __shared__ float sy[kPgTile * D];
    __shared__ float sk[kPgTile * D];

    const float* __restrict__ kern = packed + D * S * RS;
    const float* __restrict__ ilp = kern + M * KS;
    const int P = D * M;
    // pair assignment: blockIdx.y selects a block of kPgThreads pairs; if the whole problem has fewer pairs than
    // threads, the CTA is split into G row-groups that each take every G-th row of the tile.
    int pair, group, G;
    if (P >= kPgThreads) {
        pair = blockIdx.y * kPgThreads + threadIdx.x;
        group = 0;
        G = 1;
    } else {
        G = kPgThreads / P;
        group = threadIdx.x / P;
        pair = threadIdx.x - group * P;
        if (group >= G) pair = P;  // idle tail threads
    }
    const bool active = pair < P;
    const int k = active ? pair / M : 0;
    const int m = active ? pair - k * M : 0;

    float z[D], il[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        z[j] = __ldg(kern + m * KS + j);
        il[j] = __ldg(ilp + k * DP + j);
    }
    float T = 0.f, W[D];
#pragma unroll
    for (int j = 0; j < D; ++j) W[j] = 0.f;

    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = r_begin + rows_per_cta < VR ? r_begin + rows_per_cta : VR;
    for (int64_t base = r_begin; base < r_end; base += kPgTile) {
        const int n = (int)((r_end - base) < kPgTile ? (r_end - base) : kPgTile);
        __syncthreads();
        for (int i = threadIdx.x; i < n * D; i += kPgThreads) {
            sy[i] = __ldg(ys + base * D + i);
            sk[i] = __ldg(kbs + base * D + i);
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int r = group; r < n; r += G) {
                float d[D], e = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    d[j] = sy[r * D + j] - z[j];
                    e = fmaf(d[j] * d[j], il[j], e);
                }
                const float p = sk[r * D + k] * gpode_ex2(-e);
                T += p;
#pragma unroll
                for (int j = 0; j < D; ++j) W[j] = fmaf(p, d[j], W[j]);
            }
        }
    }
    if (active) {
        const GpodeAcc a = gpode_acc_layout(D, M);
        atomicAdd(acc + a.off_T + k * M + m, T);
#pragma unroll
        for (int j = 0; j < D; ++j) atomicAdd(acc + a.off_W + (k * M + m) * D + j, W[j]);
    }
}

// acc -> parameter gradients (tiny; one CTA)
__global__ void grads_finalize_kernel(const int D, const int M, const float* __restrict__ Z,
                                      const float* __restrict__ nu, const float* __restrict__ ell,
                                      const float* __restrict__ var, const float* __restrict__ acc,
                                      float* __restrict__ g_ell, float* __restrict__ g_var, float* __restrict__ g_Z,
                                      float* __restrict__ g_nu) {
    const GpodeAcc a = gpode_acc_layout(D, M);
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) g_ell[i] = -acc[a.off_A + i] / ell[i];
    for (int k = threadIdx.x; k < D; k += blockDim.x) g_var[k] = 0.5f * acc[a.off_V + k] / var[k];
    for (int i = threadIdx.x; i < D * M; i += blockDim.x) g_nu[i] = var[i / M] * acc[a.off_T + i];
    for (int i = threadIdx.x; i < M * D; i += blockDim.x) {
        const int m = i / D, j = i - m * D;
        float s = 0.f;
        for (int k = 0; k < D; ++k) {
            const float l = ell[k * D + j];
            s += var[k] * nu[k * M + m] * acc[a.off_W + (k * M + m) * D + j] / (l * l);
        }
        g_Z[i] = s;
    }
}

}  // namespace

int gpode_param_grad_launch(const float* packed, int D, int M, int S, const float* ys, const float* kbs, int64_t VR,
                            float* acc, cudaStream_t stream) {
    if (VR == 0) return 0;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int P = D * M;
    const int gy = P >= kPgThreads ? (P + kPgThreads - 1) / kPgThreads : 1;
    // rows per CTA: enough CTAs to fill the machine (~4 per SM across gy), but at least one full tile each
    int64_t want_ctas = (int64_t)sms * 4 / gy;
    if (want_ctas < 1) want_ctas = 1;
    int64_t rows_per_cta = (VR + want_ctas - 1) / want_ctas;
    rows_per_cta = ((rows_per_cta + kPgTile - 1) / kPgTile) * kPgTile;
    const int64_t gx = (VR + rows_per_cta - 1) / rows_per_cta;
    dim3 grid((unsigned)gx, (unsigned)gy);
    switch (D) {
#define GPODE_PG_CASE(D_)                                                                                        \
    case D_:                                                                                                     \
        param_grad_kernel<D_><<<grid, kPgThreads, 0, stream>>>(packed, M, S, ys, kbs, VR, rows_per_cta, acc);    \
        break;
        GPODE_PG_CASE(1) GPODE_PG_CASE(2) GPODE_PG_CASE(3) GPODE_PG_CASE(4)
        GPODE_PG_CASE(5) GPODE_PG_CASE(6) GPODE_PG_CASE(7) GPODE_PG_CASE(8)
#undef GPODE_PG_CASE
        default:
            gpode_set_error("state dimension D=%d outside 1..%d", D, GPODE_MAX_D);
            return -1;
    }
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t gpode_acc_floats(int D, int M) { return gpode_acc_layout(D, M).total; }

extern "C" int gpode_grads_finalize(const gpode_cache_t* c, const float* acc, float* grad_ell, float* grad_var,
                                    float* grad_Z, float* grad_nu, void* stream) {
    GPODE_CHECK_ARG(c != nullptr && acc != nullptr, "cache / acc is NULL");
    GPODE_CHECK_ARG(c->Z && c->nu && c->ell && c->var, "cache needs Z, nu, ell, var");
    GPODE_CHECK_ARG(grad_ell && grad_var && grad_Z && grad_nu, "NULL output");
    grads_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c->D, c->M, c->Z, c->nu, c->ell, c->var, acc, grad_ell,
                                                              grad_var, grad_Z, grad_nu);
    GPODE_LAUNCH_CHECK();
    return 0;
}
