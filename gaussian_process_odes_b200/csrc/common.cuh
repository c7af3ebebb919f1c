// Shared helpers of the gpode_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gpode_b200.h"

// ------------------------------------------------------------------------------------------------------------------
// error plumbing (per-thread last error string, int return codes as documented in include/gpode_b200.h)
// ------------------------------------------------------------------------------------------------------------------
void gpode_set_error(const char* fmt, ...);

#define GPODE_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            gpode_set_error(__VA_ARGS__);          \
            return -1;                             \
        }                                          \
    } while (0)

#define GPODE_CUDA(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            gpode_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                       \
        }                                                                                         \
    } while (0)

#define GPODE_LAUNCH_CHECK()                                                                      \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            gpode_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                       \
        }                                                                                         \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// packed parameter block (what every integrator CTA stages into shared memory with one bulk copy)
//   rff  : [k][s][RS]  = Omega_{0,s,k} .. Omega_{D-1,s,k}, phase_{s,k}, a_{s,k} = w_{s,k} sqrt(var_k/S), pad
//   kern : [m][KS]     = Z_{m,0..D-1}, c_{0,m} .. c_{D-1,m} (c_{k,m} = var_k nu_{k,m}), pad
//   il   : [k][DP]     = w_{k,j} = 0.5 log2(e) / ell_{k,j}^2   (so that exp(-0.5 r^2) = 2^(-sum d_j^2 w_kj))
// ------------------------------------------------------------------------------------------------------------------
struct GpodeLayout {
    int D, M, S, RS, KS, DP;
    int off_rff, off_kern, off_il, total;  // in floats; every offset and `total` is a multiple of 4 (16 bytes)
};

__host__ __device__ inline int gpode_round_up4(int x) { return (x + 3) & ~3; }

__host__ __device__ inline GpodeLayout gpode_layout(int D, int M, int S) {
    GpodeLayout L;
    L.D = D; L.M = M; L.S = S;
    L.RS = gpode_round_up4(D + 2);
    L.KS = gpode_round_up4(2 * D);
    L.DP = gpode_round_up4(D);
    L.off_rff = 0;
    L.off_kern = L.off_rff + D * S * L.RS;
    L.off_il = L.off_kern + M * L.KS;
    L.total = L.off_il + D * L.DP;
    return L;
}

// accumulator block of the backward kernels: A[D,D] | V[D] | T[D,M] | W[D,M,D]
struct GpodeAcc {
    int off_A, off_V, off_T, off_W, total;
};
__host__ __device__ inline GpodeAcc gpode_acc_layout(int D, int M) {
    GpodeAcc a;
    a.off_A = 0;
    a.off_V = D * D;
    a.off_T = a.off_V + D;
    a.off_W = a.off_T + D * M;
    a.total = a.off_W + D * M * D;
    return a;
}

#define GPODE_HALF_LOG2E 0.72134752044448170368f  // 0.5 * log2(e)
#define GPODE_NEG_2LN2 (-1.3862943611198906f)          // -2 ln 2

// ------------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float gpode_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t gpode_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// One thread: arm an mbarrier and pull `bytes` (multiple of 16, both sides 16B aligned) from global into shared
// memory with the bulk async-copy engine (TMA, non-tensor form: SASS UBLKCP).
__device__ __forceinline__ void gpode_mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gpode_smem_u32(mbar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void gpode_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    const uint32_t kChunk = 32768;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gpode_smem_u32(mbar)), "r"(bytes)
                 : "memory");
    for (uint32_t off = 0; off < bytes; off += kChunk) {
        uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                gpode_smem_u32((const char*)smem_dst + off)),
            "l"((const char*)gmem_src + off), "r"(n), "r"(gpode_smem_u32(mbar))
            : "memory");
    }
}

__device__ __forceinline__ void gpode_mbar_wait(uint64_t* mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(gpode_smem_u32(mbar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ float gpode_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__
