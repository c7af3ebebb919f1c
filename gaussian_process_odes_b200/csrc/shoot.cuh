// Fused multiple-shooting ELBO epilogue.
//
// In UniformSequenceModel.build_lowerbound_terms (reference src/gpode_shooting/models.py:108-146) every row of the
// (S_mc, N, T) segment batch is integrated over ONE interval and its end point x(t+1; s_t) is used twice:
//   * observation term   log N(ys[n,t] | decode(pred), var)                  (src/core/likelihoods.py:27-45,
//                                                                             decoder src/misc/mocap_utils.py:24-34)
//   * shooting constraint log p(ss[s,n,t+1] | loc = pred, scale), t < T-1    (src/core/constraints.py:26-36,56-66;
//                                                                             models.py:134-135)
// The reference (and round 1 of this library) writes the (B, 2, D) trajectory tensor, re-reads it in two more passes
// and again in their backward. Here the thread that owns a row in the RK4 kernel evaluates both terms on the end point
// while it is still in registers, and keeps two "seed" vectors per row -- d(loglik sum)/d pred and
// d(constraint sum)/d pred -- from which the adjoint kernel starts (lambda_T = g_ll seed_ll + g_cons seed_c with the
// two upstream scalar gradients read from device memory), so the trajectory tensor is never materialised.
//
// Sums are bitwise reproducible: per-thread float64 partials, warp shuffles, warps in warp order, one row of `work`
// per CTA, rows added in row order by shoot_sum_kernel (common.cuh gpode_sum_rows_ordered).
#pragma once
#include "common.cuh"

struct ShootArgs {
    const float* ys;          // [N, T, Dobs]
    const float* W;           // [D, Dobs]
    const float* bias;        // [Dobs] or NULL
    const float* lik_var;     // [Dobs]
    const float* cons_scale;  // 1 float
    const float* ss;          // [S_mc, N, T, D] all sampled states (rows of the segment batch)
    int N, T, Dobs, laplace;
    int64_t row_lo;           // first row of this launch in the (S_mc, N, T) batch
    int64_t n_total;          // S_mc * N * T
    float* pred_out;          // [B_local, D] or NULL
    float* seeds;             // [2, B_local, D] or NULL (forward without gradients)
    double* work;             // [gridDim.x][2 + Dobs]
};

// shared-memory region of the epilogue: floats [W D*Dobs | bias Dobs | 1/var Dobs | log(2 pi var) Dobs] (padded to an
// even count), then float64 [nwarps][2 + Dobs] accumulators (loglik, constraint, d loglik / d var_d)
__host__ __device__ inline size_t shoot_smem_floats(int D, int Dobs) { return ((size_t)(D + 3) * Dobs + 1) & ~(size_t)1; }
__host__ __device__ inline size_t shoot_smem_bytes(int D, int Dobs, int nwarps) {
    return shoot_smem_floats(D, Dobs) * 4 + (size_t)nwarps * (2 + Dobs) * 8;
}
#define GPODE_SHOOT_MAX_DOBS 128

#ifdef __CUDACC__
template <int D>
struct ShootSmem {
    const float* W;
    const float* bias;
    const float* iv;
    const float* lv;
    double* wacc;  // this warp's [2 + Dobs]
    float inv, inv2, c0;

    // all threads of the CTA; `base` is 8-byte aligned. Ends with a __syncthreads().
    __device__ __forceinline__ void init(unsigned char* base, const ShootArgs& a) {
        float* f = reinterpret_cast<float*>(base);
        const int Dobs = a.Dobs;
        float* sW = f;
        float* sb = f + D * Dobs;
        float* siv = sb + Dobs;
        float* slv = siv + Dobs;
        for (int i = threadIdx.x; i < D * Dobs; i += blockDim.x) sW[i] = a.W[i];
        for (int d = threadIdx.x; d < Dobs; d += blockDim.x) {
            const float v = a.lik_var[d];
            sb[d] = a.bias ? a.bias[d] : 0.f;
            siv[d] = 1.0f / v;
            slv[d] = 1.8378770664093453f + logf(v);  // log(2 pi) + log var
        }
        double* acc = reinterpret_cast<double*>(base + shoot_smem_floats(D, Dobs) * 4);
        const int nwarps = (blockDim.x + 31) >> 5;
        for (int i = threadIdx.x; i < nwarps * (2 + Dobs); i += blockDim.x) acc[i] = 0.0;
        W = sW; bias = sb; iv = siv; lv = slv;
        wacc = acc + (threadIdx.x >> 5) * (2 + Dobs);
        const float scale = a.cons_scale[0];
        inv = 1.0f / scale;
        inv2 = inv * inv;
        c0 = a.laplace ? -logf(2.0f * scale) : -logf(scale) - 0.9189385332046727f;  // -log s - 0.5 log 2 pi
        __syncthreads();
    }
};

// One row's contribution. kLanes: the 32 lanes of the warp hold 32 different rows (warp-collective: every lane must
// call it; `valid` masks rows past the end) -- otherwise exactly one lane of the warp calls it.
// ll / cs: the calling thread's float64 running sums.
template <int D, bool kLanes>
__device__ __forceinline__ void shoot_epilogue(const ShootArgs& a, const ShootSmem<D>& sm, const bool valid,
                                               const int64_t row_local, const int64_t n_local, const float (&pred)[D],
                                               double& ll, double& cs) {
    const int Dobs = a.Dobs;
    const int64_t g = a.row_lo + (valid ? row_local : 0);
    const int64_t NT = (int64_t)a.N * a.T;
    const float* __restrict__ y = a.ys + (g % NT) * Dobs;
    float sl[D];
#pragma unroll
    for (int l = 0; l < D; ++l) sl[l] = 0.f;
    float lsum = 0.f;
    for (int d = 0; d < Dobs; ++d) {
        float f = sm.bias[d];
#pragma unroll
        for (int l = 0; l < D; ++l) f = fmaf(pred[l], sm.W[l * Dobs + d], f);
        const float diff = valid ? f - __ldg(y + d) : 0.f;
        const float q = diff * sm.iv[d];
        lsum += -0.5f * (sm.lv[d] + diff * q);
        float gvd = valid ? -0.5f * (sm.iv[d] - q * q) : 0.f;
        if constexpr (kLanes) {
            gvd = gpode_warp_sum(gvd);
            if ((threadIdx.x & 31) == 0) sm.wacc[2 + d] += (double)gvd;
        } else {
            sm.wacc[2 + d] += (double)gvd;
        }
#pragma unroll
        for (int l = 0; l < D; ++l) sl[l] = fmaf(-q, sm.W[l * Dobs + d], sl[l]);
    }
    // shooting constraint: this row's end point against the NEXT sampled state of the same sequence
    float sc[D];
    float csum = 0.f;
    const int t = (int)(g % a.T);
    const bool has_next = valid && t < a.T - 1;
    const float* __restrict__ nxt = a.ss + (g + (has_next ? 1 : 0)) * D;
#pragma unroll
    for (int l = 0; l < D; ++l) {
        const float diff = has_next ? __ldg(nxt + l) - pred[l] : 0.f;
        if (a.laplace) {
            csum += has_next ? sm.c0 - fabsf(diff) * sm.inv : 0.f;
            sc[l] = diff > 0.f ? sm.inv : (diff < 0.f ? -sm.inv : 0.f);   // d lp / d pred
        } else {
            csum += has_next ? sm.c0 - 0.5f * diff * diff * sm.inv2 : 0.f;
            sc[l] = diff * sm.inv2;
        }
    }
    if (valid) {
        ll += (double)lsum;
        cs += (double)csum;
        if (a.seeds != nullptr) {
#pragma unroll
            for (int l = 0; l < D; ++l) {
                a.seeds[row_local * D + l] = sl[l];
                a.seeds[(n_local + row_local) * D + l] = sc[l];
            }
        }
        if (a.pred_out != nullptr) {
#pragma unroll
            for (int l = 0; l < D; ++l) a.pred_out[row_local * D + l] = pred[l];
        }
    }
}

// end of the kernel: thread sums -> warp slots -> this CTA's row of `work` (all in float64, fixed order)
template <int D>
__device__ __forceinline__ void shoot_finish(const ShootArgs& a, const ShootSmem<D>& sm, double ll, double cs) {
    const int lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5, Dobs = a.Dobs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ll += __shfl_xor_sync(0xffffffffu, ll, o);
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
    }
    if (lane == 0) {
        sm.wacc[0] = ll;
        sm.wacc[1] = cs;
    }
    __syncthreads();
    const double* acc0 = sm.wacc - (threadIdx.x >> 5) * (2 + Dobs);
    double* __restrict__ row = a.work + (size_t)blockIdx.x * (2 + Dobs);
    for (int i = threadIdx.x; i < 2 + Dobs; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += acc0[w * (2 + Dobs) + i];
        row[i] = s;
    }
}
#endif  // __CUDACC__
