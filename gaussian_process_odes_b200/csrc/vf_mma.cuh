// Tensor-core variant of the vector field: the random-Fourier-feature projection theta = x Omega (and, in the VJP,
// the back-projection G = g Omega^T) runs on the tensor cores as error-compensated 3xTF32 mma.sync m16n8k8
// (SASS HMMA.1688.F32.TF32), everything else (cos / sin / exp2, the RBF term, the stage algebra) stays FP32.
// north_star: "tensor cores (tf32 mma) for the Phi_rff projection only if ncu shows the kernel is FMA-bound and the
// stated tolerance still holds" -- profiles/r01_summary.md: the FFMA2 adjoint kernel is FMA-pipe bound (75 %), the
// forward kernel co-limited by FMA (66 %) and MUFU (69 %); 3xTF32 keeps theta to ~2^-21 relative.
//
// Warp layout ("quad layout"): lane = (g = lane / 4, t = lane % 4). A warp owns 32 rows as two 16-row MMA tiles;
// lane (g,t) keeps the state of the four rows g, g+8, g+16, g+24 of the warp's block in registers (replicated over
// t), owns the feature columns 2t, 2t+1 of every 8-feature tile of the theta accumulator and every fourth inducing
// point of the RBF term; partial sums are combined with a 2-step xor shuffle inside the quad.
#pragma once
#include "vf.cuh"

template <int D>
__device__ __forceinline__ float pick_dim(const float (&v)[D], const int j) {
    float r = 0.f;
#pragma unroll
    for (int jj = 0; jj < D; ++jj) r = (j == jj) ? v[jj] : r;
    return r;
}

// A-operand fragments (hi / lo) of the two 16-row tiles from the lane's four rows
template <int D>
__device__ __forceinline__ void make_a_frags(const float (&x)[4][D], const int t, uint32_t (&ah)[2][4],
                                             uint32_t (&al)[2][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float v[4] = {pick_dim<D>(x[2 * mt], t), pick_dim<D>(x[2 * mt + 1], t),
                            D > 4 ? pick_dim<D>(x[2 * mt], t + 4) : 0.f, D > 4 ? pick_dim<D>(x[2 * mt + 1], t + 4) : 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) gpode_split_tf32(v[i], ah[mt][i], al[mt][i]);
    }
}

// RBF (pathwise update) partial sums for R rows over the inducing points i0, i0+istep, ...; adds into f
template <int D, int R>
__device__ __forceinline__ void rbf_eval_partial(const float* __restrict__ kern, const float* __restrict__ wnp,
                                                 const int M, const float (&x)[R][D], float (&f)[R][D], const int i0,
                                                 const int istep) {
    constexpr int KS = VfShape<D>::KS, WP = VfShape<D>::WP, KP = VfShape<D>::KP;
    float2 wn[D][KP];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float tt[WP];
        lds_vec<WP>(tt, wnp + j * WP);
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) wn[j][kp] = make_float2(tt[2 * kp], tt[2 * kp + 1]);
    }
    float2 fk[R][KP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) fk[r][kp] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int m = i0; m < M; m += istep) {
        float kp_[KS];
        lds_vec<KS>(kp_, kern + m * KS);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float dd[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float d = x[r][j] - kp_[j];
                dd[j] = d * d;
            }
#pragma unroll
            for (int kp = 0; kp < KP; ++kp) {
                float2 e = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < D; ++j) e = ffma2(dd[j], wn[j][kp], e);
                float2 K;
                K.x = gpode_ex2(e.x);
                K.y = (2 * kp + 1 < D) ? gpode_ex2(e.y) : 0.f;
                fk[r][kp] = ffma2(make_float2(kp_[D + 2 * kp], kp_[D + 2 * kp + 1]), K, fk[r][kp]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) f[r][k] += (k & 1) ? fk[r][k >> 1].y : fk[r][k >> 1].x;
}

// f[r][k] (r = the lane's four rows) for the quad layout. sp points at the staged [kern | il | mma] region.
template <int D>
__device__ __forceinline__ void vf_eval_mma(const float* __restrict__ sp, const int M, const int S,
                                            const float (&x)[4][D], float (&f)[4][D], const int lane) {
    constexpr int KS = VfShape<D>::KS, WP = VfShape<D>::WP;
    const int S8 = (S + 7) >> 3;
    const float* __restrict__ kern = sp;
    const float* __restrict__ wnp = kern + M * KS;
    const float* __restrict__ mma = wnp + D * WP;
    const int t = lane & 3;

    uint32_t ah[2][4], al[2][4];
    make_a_frags<D>(x, t, ah, al);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) f[r][k] = 0.f;

#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float* __restrict__ rec = mma + k * S8 * GPODE_MMA_REC;
#pragma unroll 4
        for (int ft = 0; ft < S8; ++ft, rec += GPODE_MMA_REC) {
            const float2 b = *reinterpret_cast<const float2*>(rec + lane * 2);
            const float4 pa = *reinterpret_cast<const float4*>(rec + 64 + t * 4);
            uint32_t b0h, b0l, b1h, b1l;
            gpode_split_tf32(b.x, b0h, b0l);
            gpode_split_tf32(b.y, b1h, b1l);
            float c[2][4] = {{pa.x, pa.y, pa.x, pa.y}, {pa.x, pa.y, pa.x, pa.y}};  // theta starts at the phase
            // the two row tiles are independent accumulator chains: interleave them
            gpode_mma_tf32(c[0], al[0], b0h, b1h);
            gpode_mma_tf32(c[1], al[1], b0h, b1h);
            gpode_mma_tf32(c[0], ah[0], b0l, b1l);
            gpode_mma_tf32(c[1], ah[1], b0l, b1l);
            gpode_mma_tf32(c[0], ah[0], b0h, b1h);
            gpode_mma_tf32(c[1], ah[1], b0h, b1h);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                f[2 * mt][k] = fmaf(pa.z, __cosf(c[mt][0]), f[2 * mt][k]);
                f[2 * mt][k] = fmaf(pa.w, __cosf(c[mt][1]), f[2 * mt][k]);
                f[2 * mt + 1][k] = fmaf(pa.z, __cosf(c[mt][2]), f[2 * mt + 1][k]);
                f[2 * mt + 1][k] = fmaf(pa.w, __cosf(c[mt][3]), f[2 * mt + 1][k]);
            }
        }
    }
    rbf_eval_partial<D, 4>(kern, wnp, M, x, f, t, 4);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float v = f[r][k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            f[r][k] = v;
        }
}

// stage [kern | il | mma] (contiguous in the packed buffer) into shared memory; layout as stage_params
__device__ __forceinline__ const float* stage_params_mma(unsigned char* smem_raw, const float* __restrict__ packed,
                                                         const int off_kern, const int total_all) {  // end offset
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* sp = reinterpret_cast<float*>(smem_raw + 16);
    if (threadIdx.x == 0) {
        gpode_mbar_init(mbar, 1);
        gpode_bulk_g2s(sp, packed + off_kern, (uint32_t)(total_all - off_kern) * 4u, mbar);
    }
    __syncthreads();
    gpode_mbar_wait(mbar, 0);
    return sp;
}
