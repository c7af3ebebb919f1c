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
    static constexpr int KP = (D + 1) / 2;              // output-dimension pairs
    static constexpr int RP = (2 * D + 4 + 3) & ~3;     // floats per (k, feature pair) record
    static constexpr int KS = (D + 2 * KP + 3) & ~3;    // floats per inducing-point record
    static constexpr int WP = (2 * KP + 3) & ~3;        // floats per input-dimension row of -w
};

// one RFF record (see common.cuh: chunk-major inside groups of 32 records)
template <int RP>
__device__ __forceinline__ void lds_rff(float (&dst)[RP], const float* __restrict__ src) {
#pragma unroll
    for (int c = 0; c < RP / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(src + c * 128);
        dst[4 * c + 0] = v.x; dst[4 * c + 1] = v.y; dst[4 * c + 2] = v.z; dst[4 * c + 3] = v.w;
    }
}

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
// Two features (RFF term) / two output dimensions (RBF term) ride in one FFMA2.
// (i0, istep): which feature pairs / inducing points this thread sums -- (0,1) when the thread owns whole rows,
// (lane,32) in the warp-per-row kernels, where the partial sums are then combined with a warp all-reduce.
template <int D, int R, bool kWarp = false>
__device__ __forceinline__ void vf_eval(const float* __restrict__ sp, const int M, const int S,
                                        const float (&x)[R][D], float (&f)[R][D], const int i0 = 0,
                                        const int istep = 1) {
    constexpr int RP = VfShape<D>::RP, KS = VfShape<D>::KS, WP = VfShape<D>::WP;
    const int S2 = (S + 1) >> 1, S2P = (S2 + 31) & ~31;
    const float* __restrict__ rff = sp;
    const float* __restrict__ kern = sp + D * S2P * RP;
    const float* __restrict__ wnp = kern + M * KS;

    float2 fr[R][D];   // RFF partial sums, one half per feature of the pair
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) fr[r][k] = make_float2(0.f, 0.f);

    // records come in groups of 32 (chunk-major inside a group)
    auto rff_body = [&](const float* __restrict__ rec) {   // rec: chunk 0 of the k = 0 record of one feature pair
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float prm[RP];
            lds_rff<RP>(prm, rec + k * S2P * RP);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 th = make_float2(prm[2 * D], prm[2 * D + 1]);
#pragma unroll
                for (int j = 0; j < D; ++j) th = ffma2(x[r][j], make_float2(prm[2 * j], prm[2 * j + 1]), th);
                // small batches (warp-per-row kernels) are latency-bound: the range-reduced polynomial cosine costs
                // nothing there and is good to 1e-7 at any |theta|; the wide kernels use MUFU.COS (abs err ~1e-6)
                const float2 c = kWarp ? make_float2(gpode_cos_cw(th.x), gpode_cos_cw(th.y)) : make_float2(__cosf(th.x), __cosf(th.y));
                fr[r][k] = ffma2(c, make_float2(prm[2 * D + 2], prm[2 * D + 3]), fr[r][k]);
            }
        }
    };
    if constexpr (kWarp) {  // lane i0 owns slot i0 of every group: stride one whole group
#pragma unroll 2
        for (int s2 = i0; s2 < S2; s2 += 32) rff_body(rff + (s2 - i0) * RP + i0 * 4);
    } else {                // walk the groups, then the slots of a group
        for (int g0 = 0; g0 < S2; g0 += 32) {
            const int lim = S2 - g0 < 32 ? S2 - g0 : 32;
#pragma unroll 2
            for (int i = i0; i < lim; i += istep) rff_body(rff + g0 * RP + i * 4);
        }
    }

    // Full output pairs ride in FFMA2; the odd last output (D = 1, 3, 5, 7) takes scalar FMAs instead of a half-empty
    // pair, which would cost the FMA pipe as much as a full one (the pipe bounds this part: ncu math_pipe_throttle).
    constexpr int KF = D / 2;
    constexpr bool kOdd = (D & 1) != 0;
    float2 wn[D][KF > 0 ? KF : 1];  // -w, output pairs
    float wl[D];                    // -w of the last output
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float t[WP];
        lds_vec<WP>(t, wnp + j * WP);
#pragma unroll
        for (int kp = 0; kp < KF; ++kp) wn[j][kp] = make_float2(t[2 * kp], t[2 * kp + 1]);
        wl[j] = t[D - 1];
    }
    float2 fk[R][KF > 0 ? KF : 1];
    float fl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl[r] = 0.f;
#pragma unroll
        for (int kp = 0; kp < KF; ++kp) fk[r][kp] = make_float2(0.f, 0.f);
    }

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
            for (int kp = 0; kp < KF; ++kp) {
                float2 e = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < D; ++j) e = ffma2(dd[j], wn[j][kp], e);
                const float2 K = make_float2(gpode_ex2(e.x), gpode_ex2(e.y));
                fk[r][kp] = ffma2(make_float2(kp_[D + 2 * kp], kp_[D + 2 * kp + 1]), K, fk[r][kp]);
            }
            if constexpr (kOdd) {
                float e = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) e = fmaf(dd[j], wl[j], e);
                fl[r] = fmaf(kp_[2 * D - 1], gpode_ex2(e), fl[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float u = (kOdd && k == D - 1) ? fl[r] : ((k & 1) ? fk[r][(k >> 1) < KF ? (k >> 1) : 0].y
                                                                   : fk[r][(k >> 1) < KF ? (k >> 1) : 0].x);
            f[r][k] = (fr[r][k].x + fr[r][k].y) + u;
        }
}

// VJP at x with cotangent kb: xb = J(x)^T kb, and the per-thread partial sums of the shared-parameter gradients that
// do not need a cross-row contraction per inducing point:
//   A[k][j] += x_j G_kj + sum_m q'_km w_kj d_j^2     (lengthscale gradient = -A/ell: RFF path via omega = eps/ell + RBF)
//   V[k]    += kb_k (f_k + f_upd_k)                   (variance gradient = V / (2 var))
// fst = f(x) from the forward pass (so f_rff = fst - f_upd needs no cosine here).
template <int D, int R, bool kWarp = false>
__device__ __forceinline__ void vf_vjp(const float* __restrict__ sp, const int M, const int S,
                                       const float (&x)[R][D], const float (&kb)[R][D], const float (&fst)[R][D],
                                       float (&xb)[R][D], float (&A)[D][D], float (&V)[D], const int i0 = 0,
                                       const int istep = 1) {
    constexpr int RP = VfShape<D>::RP, KS = VfShape<D>::KS, WP = VfShape<D>::WP;
    const int S2 = (S + 1) >> 1, S2P = (S2 + 31) & ~31;
    const float* __restrict__ rff = sp;
    const float* __restrict__ kern = sp + D * S2P * RP;
    const float* __restrict__ wnp = kern + M * KS;

    float2 xb2[R][D];  // two partial sums per component, folded at the end
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) xb2[r][j] = make_float2(0.f, 0.f);

    // ---- RFF part: output k outer, feature pair inner ----
#pragma unroll
    for (int k = 0; k < D; ++k) {
        float2 G[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) G[r][j] = make_float2(0.f, 0.f);
        auto rff_body = [&](const float* __restrict__ rec) {
            float prm[RP];
            lds_rff<RP>(prm, rec);
            const float2 a2 = make_float2(prm[2 * D + 2], prm[2 * D + 3]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 th = make_float2(prm[2 * D], prm[2 * D + 1]);
#pragma unroll
                for (int j = 0; j < D; ++j) th = ffma2(x[r][j], make_float2(prm[2 * j], prm[2 * j + 1]), th);
                const float2 sn = kWarp ? make_float2(gpode_sin_cw(th.x), gpode_sin_cw(th.y)) : make_float2(__sinf(th.x), __sinf(th.y));
                const float2 g = fmul2(a2, sn);   // the row's factor -kb_k is applied once, after the feature loop
#pragma unroll
                for (int j = 0; j < D; ++j) G[r][j] = ffma2(g, make_float2(prm[2 * j], prm[2 * j + 1]), G[r][j]);
            }
        };
        const float* __restrict__ rk = rff + k * S2P * RP;
        if constexpr (kWarp) {
#pragma unroll 2
            for (int s2 = i0; s2 < S2; s2 += 32) rff_body(rk + (s2 - i0) * RP + i0 * 4);
        } else {
            for (int g0 = 0; g0 < S2; g0 += 32) {
                const int lim = S2 - g0 < 32 ? S2 - g0 : 32;
#pragma unroll 2
                for (int i = i0; i < lim; i += istep) rff_body(rk + g0 * RP + i * 4);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float gj = -kb[r][k] * (G[r][j].x + G[r][j].y);
                xb2[r][j].x += gj;
                A[k][j] = fmaf(x[r][j], gj, A[k][j]);
            }
    }

    // ---- RBF part: inducing point m outer, output pairs inner (odd last output: scalar FMAs, see vf_eval) ----
    constexpr int KF = D / 2;
    constexpr bool kOdd = (D & 1) != 0;
    float2 wn[D][KF > 0 ? KF : 1];
    float wl[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float t[WP];
        lds_vec<WP>(t, wnp + j * WP);
#pragma unroll
        for (int kp = 0; kp < KF; ++kp) wn[j][kp] = make_float2(t[2 * kp], t[2 * kp + 1]);
        wl[j] = t[D - 1];
    }
    float2 fu[R][KF > 0 ? KF : 1], A2[KF > 0 ? KF : 1][D];
    float ful[R], A2l[D];
#pragma unroll
    for (int j = 0; j < D; ++j) A2l[j] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) ful[r] = 0.f;
#pragma unroll
    for (int kp = 0; kp < KF; ++kp) {
#pragma unroll
        for (int r = 0; r < R; ++r) fu[r][kp] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < D; ++j) A2[kp][j] = make_float2(0.f, 0.f);
    }
    // 2 ln2 * kb  (q' = -2 ln2 kb c K and u = q' w = (2 ln2 kb c K)(-w))
    float2 kbn[R][KF > 0 ? KF : 1];
    float kbl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int kp = 0; kp < KF; ++kp)
            kbn[r][kp] = make_float2(-GPODE_NEG_2LN2 * kb[r][2 * kp], -GPODE_NEG_2LN2 * kb[r][2 * kp + 1]);
        kbl[r] = -GPODE_NEG_2LN2 * kb[r][D - 1];
    }

#pragma unroll 2
    for (int m = i0; m < M; m += istep) {
        float kp_[KS];
        lds_vec<KS>(kp_, kern + m * KS);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float d[D], dd[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                d[j] = x[r][j] - kp_[j];
                dd[j] = d[j] * d[j];
            }
            float2 tq[D];  // t_j = sum_k q'_k (-w_kj), one partial per output of the pair
            float tl[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                tq[j] = make_float2(0.f, 0.f);
                tl[j] = 0.f;
            }
#pragma unroll
            for (int kp = 0; kp < KF; ++kp) {
                float2 e = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < D; ++j) e = ffma2(dd[j], wn[j][kp], e);
                const float2 K = make_float2(gpode_ex2(e.x), gpode_ex2(e.y));
                const float2 cK = fmul2(make_float2(kp_[D + 2 * kp], kp_[D + 2 * kp + 1]), K);
                fu[r][kp] = fadd2(fu[r][kp], cK);
                const float2 q = fmul2(kbn[r][kp], cK);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    tq[j] = ffma2(q, wn[j][kp], tq[j]);
                    A2[kp][j] = ffma2(dd[j], q, A2[kp][j]);   // the factor -w_kj is applied once, at the very end
                }
            }
            if constexpr (kOdd) {
                float e = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) e = fmaf(dd[j], wl[j], e);
                const float cK = kp_[2 * D - 1] * gpode_ex2(e);
                ful[r] += cK;
                const float q = kbl[r] * cK;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    tl[j] = q * wl[j];
                    A2l[j] = fmaf(dd[j], q, A2l[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < D; ++j) xb2[r][j].x = fmaf(d[j], (tq[j].x + tq[j].y) + tl[j], xb2[r][j].x);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int j = 0; j < D; ++j) xb[r][j] = xb2[r][j].x + xb2[r][j].y;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int kp = (k >> 1) < KF ? (k >> 1) : 0;
            const float fuk = (kOdd && k == D - 1) ? ful[r] : ((k & 1) ? fu[r][kp].y : fu[r][kp].x);
            V[k] = fmaf(kb[r][k], fst[r][k] + fuk, V[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int kp = (k >> 1) < KF ? (k >> 1) : 0;
            const float w = (kOdd && k == D - 1) ? wl[j] : ((k & 1) ? wn[j][kp].y : wn[j][kp].x);
            const float a = (kOdd && k == D - 1) ? A2l[j] : ((k & 1) ? A2[kp][j].y : A2[kp][j].x);
            A[k][j] = fmaf(w, a, A[k][j]);
        }
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

// A[D][D] | V[D] per-thread partials -> fixed-order block reduction -> this CTA's row of the accumulator block
// (common.cuh, GpodeAcc). No atomics: warp shuffles, then one thread per value adds the warps in warp order, so the
// result does not depend on scheduling; gpode_grads_finalize adds the rows in row order.
// red_smem: at least kRedWarps * (D*D + D) floats.
constexpr int kRedWarps = 16;  // widest CTA of any adjoint kernel is 12 warps
template <int D>
constexpr int kRedFloats = (kRedWarps * (D * D + D) + 3) & ~3;
template <int D>
__device__ __forceinline__ void reduce_AV(float (&A)[D][D], float (&V)[D], float* __restrict__ acc, float* red_smem) {
    constexpr int N = D * D + D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float v = gpode_warp_sum(A[k][j]);
            if (lane == 0) red_smem[warp * N + k * D + j] = v;
        }
        const float v = gpode_warp_sum(V[k]);
        if (lane == 0) red_smem[warp * N + D * D + k] = v;
    }
    __syncthreads();
    float* __restrict__ row = acc + GPODE_ACC_HDR + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red_smem[w * N + i];
        row[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<int*>(acc)[0] = (int)gridDim.x;
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-per-row helpers (small batches: one warp owns one trajectory, its lanes split the Fourier features and the
// inducing points). The lane partial sums of the pathwise update are individually ~|nu| = 1e2 times larger than their
// total (they cancel across inducing points), so everything that crosses lanes or accumulates over time is done in
// FLOAT64 here: the cross-lane sums of f and J^T kb, and the lengthscale / variance partial sums, which live in a
// shared-memory slab [D*D + D][32 lanes] of doubles per warp for the whole kernel. These kernels are latency-bound;
// the float64 work is a few percent of an evaluation. (Measured on 20 seeds of the VDP plain-GPODE ELBO: see
// profiles/r02_seed_table_vdp_gpode_rk4.md.)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gpode_warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int D>
__device__ __forceinline__ void vf_eval_warp(const float* sp, int M, int S, const float (&x)[1][D], float (&f)[1][D],
                                             int lane) {
    vf_eval<D, 1, true>(sp, M, S, x, f, lane, 32);
#pragma unroll
    for (int j = 0; j < D; ++j) f[0][j] = (float)gpode_warp_sum_f64((double)f[0][j]);
}

template <int D>
struct WarpAcc64 {
    static constexpr int N = D * D + D;
    static constexpr int kSlabDoubles = N * 32;  // per warp
    double* w;                                    // this lane's column of the warp's slab
    __device__ __forceinline__ void init(double* slabs) {
        w = slabs + (size_t)(threadIdx.x >> 5) * kSlabDoubles + (threadIdx.x & 31);
#pragma unroll
        for (int i = 0; i < N; ++i) w[i * 32] = 0.0;
    }
    __device__ __forceinline__ void add(const float (&A)[D][D], const float (&V)[D]) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int j = 0; j < D; ++j) w[(k * D + j) * 32] += (double)A[k][j];
            w[(D * D + k) * 32] += (double)V[k];
        }
    }
};
// shared-memory doubles a warp-per-row adjoint kernel with `nwarps` warps needs for WarpAcc64 + its block reduction
template <int D>
constexpr int kWarpAccDoubles(int nwarps) { return nwarps * (WarpAcc64<D>::kSlabDoubles + WarpAcc64<D>::N); }

// xb = J(x)^T kb for the warp's row (all lanes get the full sum); parameter partial sums -> wa. fst = f(x).
template <int D>
__device__ __forceinline__ void vf_vjp_warp(const float* sp, int M, int S, const float (&x)[1][D],
                                            const float (&kb)[1][D], const float (&fst)[1][D], float (&xb)[1][D],
                                            WarpAcc64<D>& wa, int lane) {
    float fm[1][D], A[D][D], V[D];  // f(x) enters the variance partial sum once per row, not once per lane
#pragma unroll
    for (int k = 0; k < D; ++k) {
        fm[0][k] = lane == 0 ? fst[0][k] : 0.f;
        V[k] = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) A[k][j] = 0.f;
    }
    vf_vjp<D, 1, true>(sp, M, S, x, kb, fm, xb, A, V, lane, 32);
    wa.add(A, V);
#pragma unroll
    for (int j = 0; j < D; ++j) xb[0][j] = (float)gpode_warp_sum_f64((double)xb[0][j]);
}

// WarpAcc64 slabs -> this CTA's row of the accumulator block: lanes by shuffle, warps in warp order, all in float64
template <int D>
__device__ __forceinline__ void reduce_AV64(const WarpAcc64<D>& wa, float* __restrict__ acc, double* red64) {
    constexpr int N = D * D + D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double v = gpode_warp_sum_f64(wa.w[i * 32]);
        if (lane == 0) red64[warp * N + i] = v;
    }
    __syncthreads();
    float* __restrict__ row = acc + GPODE_ACC_HDR + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red64[w * N + i];
        row[i] = (float)s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<int*>(acc)[0] = (int)gridDim.x;
}
