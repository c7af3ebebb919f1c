// The GP vector field f(x) = Phi_rff(x) w + K(x,Z) nu and its vector-Jacobian product, as register-resident device
// templates: each thread owns R rows (trajectories) of state dimension D and walks the packed parameter block that
// lives in shared memory (every lane reads the same address -> shared-memory broadcast, LDS.128).
//
// Arithmetic replaced: DSVGP_Layer.forward (reference src/core/dsvgp.py:172-197), i.e. rff_forward (:124-137),
// RBF.K / square_dist_dimwise (src/core/kernels.py:53-68,87-99) and the einsum of :192; closed form in SURVEY.md
// section 8(a) row A1, VJP formulas in row A7.
#pragma once
#include "common.cuh"

template <int D>
struct VfShape {
    static constexpr int RS = (D + 2 + 3) & ~3;
    static constexpr int KS = (2 * D + 3) & ~3;
    static constexpr int DP = (D + 3) & ~3;
};

template <int N>
__device__ __forceinline__ void lds_vec(float (&dst)[N], const float* __restrict__ src) {
    static_assert(N % 4 == 0, "packed records are padded to 16 bytes");
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
}

// f[r][k] = sum_s a_sk cos(sum_j x_j Omega_jsk + phase_sk) + sum_m c_km 2^(-sum_j (x_j - Z_mj)^2 w_kj)
// with w_kj = 0.5 log2(e) / ell_kj^2 (the `il` block of the packed cache)
// (i0, istep): which features / inducing points this thread sums -- (0,1) when the thread owns whole rows, (lane,32) in
// the warp-per-row kernels, where the partial sums are then combined with a warp all-reduce.
template <int D, int R>
__device__ __forceinline__ void vf_eval(const float* __restrict__ sp, const int M, const int S,
                                        const float (&x)[R][D], float (&f)[R][D], const int i0 = 0,
                                        const int istep = 1) {
    constexpr int RS = VfShape<D>::RS, KS = VfShape<D>::KS, DP = VfShape<D>::DP;
    const float* __restrict__ rff = sp;
    const float* __restrict__ kern = sp + D * S * RS;
    const float* __restrict__ ilp = kern + M * KS;

#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) f[r][k] = 0.f;

    // ---- random-Fourier-feature prior sample: feature s outer, output k inner (D*R independent chains) ----
#pragma unroll 2
    for (int s = i0; s < S; s += istep) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float prm[RS];
            lds_vec<RS>(prm, rff + (k * S + s) * RS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float th = prm[D];
#pragma unroll
                for (int j = 0; j < D; ++j) th = fmaf(x[r][j], prm[j], th);
                f[r][k] = fmaf(prm[D + 1], __cosf(th), f[r][k]);
            }
        }
    }

    // ---- pathwise update: inducing point m outer (x - Z_m shared by all k), output k inner ----
    float il[D][DP];
#pragma unroll
    for (int k = 0; k < D; ++k) lds_vec<DP>(il[k], ilp + k * DP);

#pragma unroll 2
    for (int m = i0; m < M; m += istep) {
        float kp[KS];
        lds_vec<KS>(kp, kern + m * KS);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float dd[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float d = x[r][j] - kp[j];
                dd[j] = d * d;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                float e = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) e = fmaf(dd[j], il[k][j], e);
                f[r][k] = fmaf(kp[D + k], gpode_ex2(-e), f[r][k]);
            }
        }
    }
}

// VJP at x with cotangent kb: xb = J(x)^T kb, and the per-thread partial sums of the shared-parameter gradients that
// do not need a cross-row contraction per inducing point:
//   A[k][j] += x_j G_kj + sum_m q' w_kj d_j^2    (lengthscale gradient = -A/ell, RFF path through omega = eps/ell + RBF)
//   V[k]    += kb_k (f_k + f_upd_k)          (variance gradient = V / (2 var))
// fst = f(x) from the forward pass (so f_rff = fst - f_upd needs no cosine here).
template <int D, int R>
__device__ __forceinline__ void vf_vjp(const float* __restrict__ sp, const int M, const int S,
                                       const float (&x)[R][D], const float (&kb)[R][D], const float (&fst)[R][D],
                                       float (&xb)[R][D], float (&A)[D][D], float (&V)[D], const int i0 = 0,
                                       const int istep = 1) {
    constexpr int RS = VfShape<D>::RS, KS = VfShape<D>::KS, DP = VfShape<D>::DP;
    const float* __restrict__ rff = sp;
    const float* __restrict__ kern = sp + D * S * RS;
    const float* __restrict__ ilp = kern + M * KS;

#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) xb[r][j] = 0.f;

    // ---- RFF part: output k outer (only R*D partial-Jacobian registers live), feature s inner ----
#pragma unroll
    for (int k = 0; k < D; ++k) {
        float G[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) G[r][j] = 0.f;
#pragma unroll 4
        for (int s = i0; s < S; s += istep) {
            float prm[RS];
            lds_vec<RS>(prm, rff + (k * S + s) * RS);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float th = prm[D];
#pragma unroll
                for (int j = 0; j < D; ++j) th = fmaf(x[r][j], prm[j], th);
                const float g = -(kb[r][k] * prm[D + 1]) * __sinf(th);
#pragma unroll
                for (int j = 0; j < D; ++j) G[r][j] = fmaf(g, prm[j], G[r][j]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                xb[r][j] += G[r][j];
                A[k][j] = fmaf(x[r][j], G[r][j], A[k][j]);
            }
    }

    // ---- RBF part ----
    float il[D][DP];
#pragma unroll
    for (int k = 0; k < D; ++k) lds_vec<DP>(il[k], ilp + k * DP);
    float fu[R][D];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) fu[r][k] = 0.f;

#pragma unroll 2
    for (int m = i0; m < M; m += istep) {
        float kp[KS];
        lds_vec<KS>(kp, kern + m * KS);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float d[D], dd[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                d[j] = x[r][j] - kp[j];
                dd[j] = d[j] * d[j];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                float e = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) e = fmaf(dd[j], il[k][j], e);
                const float cK = kp[D + k] * gpode_ex2(-e);
                fu[r][k] += cK;
                const float q = kb[r][k] * cK * GPODE_NEG_2LN2;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float u = q * il[k][j];
                    xb[r][j] = fmaf(u, d[j], xb[r][j]);
                    A[k][j] = fmaf(u, dd[j], A[k][j]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) V[k] = fmaf(kb[r][k], fst[r][k] + fu[r][k], V[k]);
}

// ---- staging of the packed block into shared memory (bulk async copy + mbarrier) -------------------------------
// dynamic shared memory layout: [0,16) mbarrier, [16, 16 + 4*total) parameters
__device__ __forceinline__ const float* stage_params(unsigned char* smem_raw, const float* __restrict__ packed,
                                                     const int total_floats) {
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* sp = reinterpret_cast<float*>(smem_raw + 16);
    if (threadIdx.x == 0) {
        gpode_mbar_init(mbar, 1);
        gpode_bulk_g2s(sp, packed, (uint32_t)total_floats * 4u, mbar);
    }
    __syncthreads();  // mbarrier init visible to the waiters
    gpode_mbar_wait(mbar, 0);
    return sp;
}

// A[D][D] | V[D] per-thread partials -> block reduction -> one atomicAdd per value per CTA
template <int D>
__device__ __forceinline__ void reduce_AV(float (&A)[D][D], float (&V)[D], float* __restrict__ acc, float* red_smem) {
    // red_smem: at least (D*D + D) floats, zero-initialised by the caller before the barrier below
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float v = gpode_warp_sum(A[k][j]);
            if (lane == 0) atomicAdd(&red_smem[k * D + j], v);
        }
        const float v = gpode_warp_sum(V[k]);
        if (lane == 0) atomicAdd(&red_smem[D * D + k], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * D + D; i += blockDim.x) atomicAdd(&acc[i], red_smem[i]);
}
