// Adjoint (vector-Jacobian product) of the GP vector field with BOTH random-Fourier-feature projections on the
// 5th-generation tensor cores (tcgen05, accumulators and the second GEMM's A operand in TMEM), D = 4, 5.
//
//   theta[row, s] = sum_j x[row, j] Omega[j, s, k]                      GEMM 1: [rows x 16] x [16 x features]
//   G[row, j]     = sum_s sin(theta[row, s] + phase[s]) a_s Omega[j, s, k]   GEMM 2: [rows x features] x [features x 16]
//
// GEMM 1 is kind::f16 with the error-compensated split of vjp_mma.cuh (contraction slots x_hi | x_hi | x_lo against
// Omega_hi ; Omega_lo ; Omega_hi, K = 16, one MMA per 32-feature chunk), operands in shared memory (SS form), fp32
// accumulator in TMEM. The row threads (thread = row = TMEM lane) read theta with tcgen05.ld, add the phase, take the sine
// (FMUL.RZ + MUFU.SIN), split it into fp16 hi / lo (hi = 11 significant bits, exact in fp16; lo = the rest), pack pairs
// and write both back to TMEM with tcgen05.st: they are the A operand of GEMM 2 (TS form, kind::f16, K = 16 per
// instruction, two features per 32-bit TMEM column). Its B operand holds [256 a Omega^T hi | 256 a Omega^T lo] in its 16
// rows, so TWO MMAs per 16 features give sin_hi B_hi, sin_hi B_lo (columns 0-7 / 8-15 of G) and sin_lo B_hi -- the three
// terms of the split product (the fourth, sin_lo B_lo ~ 2^-22, rides along). A first version ran GEMM 2 as kind::tf32
// (K = 8: twice the MMA count, single-buffered sine): the row threads spent 22 % of their samples waiting for the tensor
// pipe (ncu), 1.27 ms per 1e6-row VJP against 0.87 for the mma.sync kernel. Per feature-output the CUDA cores issue
// FADD (phase), FMUL.RZ, MUFU.SIN, LOP3 (hi), FADD (lo), F2FP + 2/32 tcgen05.ld / st: ~6 slots against the 8 cycles the MUFU
// pipe needs, where the mma.sync kernel (vjp_mma.cuh) spends ~10 (fragment bookkeeping: two LOP3, two F2FP, the operand
// LDS) and is dispatch-bound. The RBF term, the parameter partial sums and the row algebra are those of vjp_mma.cuh.
//
// MEASURED OUTCOME (B200, D = 5, M = 100, S = 256, 1e6 rows, tools/time_vjp_umma.py): parity with the FFMA2 adjoint to
// 1e-6 on grad_x and the parameter sums, but 1.15 ms per VJP against 0.87 ms for the mma.sync kernel (FFMA2: 1.08 ms), so
// this kernel is NOT dispatched. ncu: 8.3e8 warp instructions per VJP against 5.6e8 -- per feature-output the row threads
// still issue ~8.5 instructions (phase add, FMUL.RZ, MUFU.SIN, mask, subtract, half a F2FP pair twice, plus ~1.3 of
// mbarrier / tcgen05.ld / st / fence bookkeeping per chunk of 32), barely fewer than the ~10 of the mma.sync kernel,
// and with three row warps per scheduler the two TMEM round trips per chunk are exposed (16 % of the samples on the
// theta-ready wait). The adjoint is bound by what surrounds the sine (its range-reduction multiply, the split, the
// packing), not by where the two GEMMs run; putting them on tcgen05 frees the tensor-op issue slots (0.16 HMMA per
// value), which were never the limiter.
//
// CTA = kUbTiles warpgroups of 128 row threads + one issuer warp per warpgroup (lane 0 issues the MMAs of its tile and
// commits them to mbarriers). Per warpgroup TMEM: theta ping-pong 2 x 32 columns, sin hi / lo 32 + 32, G ping-pong
// 2 x 16 = 160 columns.
#include "umma.cuh"
#include "vf.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int kUbTiles = 3;                  // 128-row tiles (warpgroups) per CTA
constexpr int kUbRowThreads = 128 * kUbTiles;
constexpr int kUbThreads = kUbRowThreads + 32 * kUbTiles;
constexpr int kUbChunk = 32;                 // features per theta accumulator buffer
constexpr int kUbCols = 160;                 // TMEM columns per warpgroup
constexpr int C_TH = 0, C_S = 64, C_G = 128;   // sine buffers: [2][hi 16 | lo 16] columns at C_S

struct UbLayout {
    int D, S, SU, NCH;
    int64_t off_b1, off_b2, off_ph, total;   // floats
};
__host__ __device__ inline UbLayout ub_layout(int D, int S) {
    UbLayout L;
    L.D = D; L.S = S;
    L.SU = (S + kUbChunk - 1) / kUbChunk * kUbChunk;
    L.NCH = L.SU / kUbChunk;
    L.off_b1 = 0;                                   // per k: [2 K-chunks][SU/8 groups][8 features][8 halfs]  (fp16)
    L.off_b2 = L.off_b1 + (int64_t)D * L.SU * 8;    //   = SU * 16 halfs = SU * 8 floats per k
    L.off_ph = L.off_b2 + (int64_t)D * L.SU * 8;    // per k: [SU/8 K-chunks][2 groups][8 rows][8 halfs]  (fp16)
    L.total = L.off_ph + (int64_t)D * L.SU;         // per k: phase [SU]
    return L;
}

__global__ void pack_ub_kernel(const UbLayout L, const float* __restrict__ omega, const float* __restrict__ phase,
                               const float* __restrict__ w, const float* __restrict__ var, float* __restrict__ out) {
    const int D = L.D, S = L.S, SU = L.SU;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto om = [&](int j, int s, int k) -> float { return (j < D && s < S) ? omega[((size_t)j * S + s) * D + k] : 0.f; };
    // GEMM 1 B operand: feature s (N), contraction slot q (K = 16): Omega_hi | Omega_lo | Omega_hi | 0
    __half* b1 = reinterpret_cast<__half*>(out + L.off_b1);
    for (int64_t i = i0; i < (int64_t)D * SU * 16; i += stride) {
        const int q = (int)(i % 16), s = (int)((i / 16) % SU), k = (int)(i / (16 * (int64_t)SU));
        float v = 0.f;
        if (q < 3 * D) {
            const float o = om(q % D, s, k);
            const __half hi = __float2half_rn(o);
            v = (q >= D && q < 2 * D) ? o - __half2float(hi) : __half2float(hi);
        }
        b1[(int64_t)k * SU * 16 + (q >> 3) * (SU * 8) + (s >> 3) * 64 + (s & 7) * 8 + (q & 7)] = __float2half_rn(v);
    }
    // GEMM 2 B operand: row n (N = 16): n < 8 -> hi of 256 a_s Omega_{n,s,k}, n >= 8 -> lo of 256 a_s Omega_{n-8,s,k};
    // K = feature s (the scale keeps the low parts out of the fp16 subnormals, as GPODE_MMAH_SCALE in vjp_mma.cuh)
    __half* b2 = reinterpret_cast<__half*>(out + L.off_b2);
    for (int64_t i = i0; i < (int64_t)D * SU * 16; i += stride) {
        const int n = (int)(i % 16), s = (int)((i / 16) % SU), k = (int)(i / (16 * (int64_t)SU));
        const float a = s < S ? w[s * D + k] * sqrtf(var[k] / (float)S) * GPODE_MMAH_SCALE : 0.f;
        const float v = a * om(n & 7, s, k);
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        b2[(int64_t)k * SU * 16 + (s >> 3) * 128 + (n >> 3) * 64 + (n & 7) * 8 + (s & 7)] = n < 8 ? hi : lo;
    }
    for (int64_t i = i0; i < (int64_t)D * SU; i += stride) {
        const int s = (int)(i % SU), k = (int)(i / SU);
        out[L.off_ph + i] = s < S ? phase[s * D + k] : 0.f;
    }
}

// kind::f16 (A, B fp16), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row, two fp16 elements per 32-bit column), B in shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
                 "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
                 "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t* mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(gpode_smem_u32(mbar)), "r"(parity)
        : "memory");
    return ok != 0;
}

struct UbBars {  // mbarriers of one warpgroup
    uint64_t a_full, th_full[2], th_free[2], s_full[2], s_free[2], g_full[2], g_free[2];
};

// dynamic shared memory: [bars kUbTiles][tmem_ptr][pad to 128] | small (kern | il) | ub block (b1 | b2 | phase) |
// A tiles (kUbTiles x 4 KB, fp16) | block-reduction floats
template <int D>
__global__ void __launch_bounds__(kUbThreads, 1)
vf_bwd_umma_kernel(const float* __restrict__ packed, const float* __restrict__ ub, const int M, const int S,
                   const int off_kern, const int n_small, const float* __restrict__ x, const float* __restrict__ f,
                   const float* __restrict__ gf, float* __restrict__ gx, const int64_t B, float* __restrict__ acc,
                   const int flags) {
    extern __shared__ __align__(128) unsigned char smem[];
    const UbLayout L = ub_layout(D, S);
    UbBars* bars = reinterpret_cast<UbBars*>(smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kUbTiles * sizeof(UbBars));
    uint64_t* bar_stage = reinterpret_cast<uint64_t*>(smem + kUbTiles * sizeof(UbBars) + 8);
    constexpr int kHdr = 512;
    static_assert(kUbTiles * sizeof(UbBars) + 16 <= kHdr, "header");
    float* small = reinterpret_cast<float*>(smem + kHdr);
    float* ubs = small + n_small;                                   // 16-byte aligned (n_small is a multiple of 4)
    const int n_ub = (int)L.total;
    unsigned char* a_tiles = reinterpret_cast<unsigned char*>(ubs + n_ub);   // kUbTiles x 4096 bytes
    float* red = reinterpret_cast<float*>(a_tiles + kUbTiles * 4096);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_row = tid < kUbRowThreads;
    const int wg = is_row ? tid / 128 : (tid - kUbRowThreads) / 32;
    UbBars& bar = bars[wg];

    if (tid == 0) {
        for (int i = 0; i < kUbTiles; ++i) {
            gpode_mbar_init(&bars[i].a_full, 128);
            for (int b = 0; b < 2; ++b) {
                gpode_mbar_init(&bars[i].s_full[b], 128);
                gpode_mbar_init(&bars[i].s_free[b], 1);
                gpode_mbar_init(&bars[i].th_full[b], 1);
                gpode_mbar_init(&bars[i].th_free[b], 128);
                gpode_mbar_init(&bars[i].g_full[b], 1);
                gpode_mbar_init(&bars[i].g_free[b], 128);
            }
        }
        gpode_mbar_init(bar_stage, 1);
        const uint32_t bytes = (uint32_t)(n_small + n_ub) * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gpode_smem_u32(bar_stage)), "r"(bytes)
                     : "memory");
        auto bulk = [&](void* dst, const void* src, uint32_t nbytes) {
            for (uint32_t off = 0; off < nbytes; off += 32768u) {
                const uint32_t n = nbytes - off < 32768u ? nbytes - off : 32768u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 gpode_smem_u32((const char*)dst + off)),
                             "l"((const char*)src + off), "r"(n), "r"(gpode_smem_u32(bar_stage))
                             : "memory");
            }
        };
        bulk(small, packed + off_kern, (uint32_t)n_small * 4u);
        bulk(ubs, ub, (uint32_t)n_ub * 4u);
    }
    if (warp == kUbRowThreads / 32) {   // first issuer warp allocates the CTA's tensor memory
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_wg = *tmem_ptr + (uint32_t)(wg * kUbCols);
    mbar_wait_bounded(bar_stage, 0);

    const __half* b1 = reinterpret_cast<const __half*>(ubs + L.off_b1);
    const __half* b2 = reinterpret_cast<const __half*>(ubs + L.off_b2);
    const float* phs = ubs + L.off_ph;
    const int SU = L.SU, NCH = L.NCH;
    unsigned char* a_tile = a_tiles + wg * 4096;

    float A[D][D], V[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        V[k] = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) A[k][j] = 0.f;
    }

    // every warpgroup walks its own tiles: tile index = (blockIdx.x * kUbTiles + wg) + i * gridDim.x * kUbTiles
    const int64_t n_tiles = (B + 127) / 128;
    uint32_t tile_it = 0;   // tiles done by this warpgroup
    uint32_t gi = 0;        // chunks done by this warpgroup (theta buffer = gi & 1)
    uint32_t ki = 0;        // outputs done by this warpgroup (G buffer = ki & 1)
    for (int64_t tile = (int64_t)blockIdx.x * kUbTiles + wg; tile < n_tiles;
         tile += (int64_t)gridDim.x * kUbTiles, ++tile_it) {
        if (is_row) {
            const int t = tid - wg * 128;
            const int64_t row = tile * 128 + t;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            float xr[D], kb[D], fr[D], xb[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                xr[j] = row < B ? __ldg(x + row * D + j) : 0.f;
                kb[j] = row < B ? __ldg(gf + row * D + j) : 0.f;
                fr[j] = row < B ? __ldg(f + row * D + j) : 0.f;
                xb[j] = 0.f;
            }
            {   // this row of the GEMM-1 A tile: x_hi | x_hi | x_lo | 0 as 16 halfs (two 16-byte core-matrix lines)
                __align__(16) __half hv[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float v = 0.f;
                    if (q < 3 * D) {
                        const float xv = xr[q % D];
                        const float hi = __half2float(__float2half_rn(xv));
                        v = q >= 2 * D ? xv - hi : hi;
                    }
                    hv[q] = __float2half_rn(v);
                }
                uint4* dst0 = reinterpret_cast<uint4*>(a_tile + (t >> 3) * 128 + (t & 7) * 16);
                uint4* dst1 = reinterpret_cast<uint4*>(a_tile + 2048 + (t >> 3) * 128 + (t & 7) * 16);
                *dst0 = *reinterpret_cast<const uint4*>(&hv[0]);
                *dst1 = *reinterpret_cast<const uint4*>(&hv[8]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar.a_full);

            // G_k of this row is read one step LATE (after the first chunk of output k + 1, the last one after the RBF
            // part of a warpgroup that runs RFF first), so the G MMAs behind the last sine chunk are never waited for
            auto consume_G = [&](const int k, const uint32_t kidx) {
                const int kbuf = kidx & 1;
                mbar_wait_bounded(&bar.g_full[kbuf], (kidx >> 1) & 1);
                tc_fence_after_sync();
                uint32_t g[16];
                tmem_ld16(tmem_wg + lane_base + (uint32_t)(C_G + kbuf * 16), g);
                tc_fence_before_sync();
                mbar_arrive(&bar.g_free[kbuf]);
                float kbk = 0.f;
#pragma unroll
                for (int kk = 0; kk < D; ++kk) kbk = kk == k ? kb[kk] : kbk;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float gj = (-1.f / GPODE_MMAH_SCALE) * kbk * (__uint_as_float(g[j]) + __uint_as_float(g[8 + j]));
                    xb[j] += gj;
                    const float aj = xr[j] * gj;
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) A[kk][j] += kk == k ? aj : 0.f;
                }
            };
            auto rff_part = [&]() {
#pragma unroll 1
            for (int k = 0; k < D; ++k, ++ki) {
                const float* __restrict__ php = phs + (size_t)k * SU;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c, ++gi) {
                    const int buf = gi & 1;
                    mbar_wait_bounded(&bar.th_full[buf], (gi >> 1) & 1);
                    tc_fence_after_sync();
                    uint32_t r[32];
                    tmem_ld32_issue(tmem_wg + lane_base + (uint32_t)(C_TH + buf * kUbChunk), r);
                    tmem_ld_wait(r);
                    tc_fence_before_sync();
                    mbar_arrive(&bar.th_free[buf]);
                    uint32_t hi[16], lo[16];   // two features per 32-bit word: low half = even feature
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 p4 = *reinterpret_cast<const float4*>(php + c * kUbChunk + i);
                        const float ph[4] = {p4.x, p4.y, p4.z, p4.w};
                        float h[4], l[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float sv = __sinf(__uint_as_float(r[i + u]) + ph[u]);
                            h[u] = __uint_as_float(__float_as_uint(sv) & 0xffffe000u);
                            l[u] = sv - h[u];
                        }
                        const __half2 h0 = __floats2half2_rn(h[0], h[1]), h1 = __floats2half2_rn(h[2], h[3]);
                        const __half2 l0 = __floats2half2_rn(l[0], l[1]), l1 = __floats2half2_rn(l[2], l[3]);
                        hi[i / 2] = *reinterpret_cast<const uint32_t*>(&h0);
                        hi[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                        lo[i / 2] = *reinterpret_cast<const uint32_t*>(&l0);
                        lo[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&l1);
                    }
                    if (gi >= 2) {   // the G MMAs two chunks back have consumed this sine buffer
                        mbar_wait_bounded(&bar.s_free[buf], ((gi >> 1) - 1) & 1);
                        tc_fence_after_sync();
                    }
                    tmem_st16(tmem_wg + lane_base + (uint32_t)(C_S + buf * 32), hi);
                    tmem_st16(tmem_wg + lane_base + (uint32_t)(C_S + buf * 32 + 16), lo);
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bar.s_full[buf]);
                    if ((flags & 4) && c == 0 && k > 0) consume_G(k - 1, ki - 1);
                }
                if (!(flags & 4)) consume_G(k, ki);
            }
            };
            auto rbf_part = [&]() {
            {
                constexpr int KS = VfShape<D>::KS, WP = VfShape<D>::WP;
                const float* __restrict__ kern = small;
                const float* __restrict__ wnp = kern + M * KS;
                constexpr int KF = D / 2;
                constexpr bool kOdd = (D & 1) != 0;
                float2 wn[D][KF > 0 ? KF : 1];
                float wl[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    float tw[WP];
                    lds_vec<WP>(tw, wnp + j * WP);
#pragma unroll
                    for (int kp = 0; kp < KF; ++kp) wn[j][kp] = make_float2(tw[2 * kp], tw[2 * kp + 1]);
                    wl[j] = tw[D - 1];
                }
                float2 fu[KF > 0 ? KF : 1], A2[KF > 0 ? KF : 1][D], kbn[KF > 0 ? KF : 1];
                float ful = 0.f, A2l[D];
                const float kbl = -GPODE_NEG_2LN2 * kb[D - 1];
#pragma unroll
                for (int j = 0; j < D; ++j) A2l[j] = 0.f;
#pragma unroll
                for (int kp = 0; kp < KF; ++kp) {
                    fu[kp] = make_float2(0.f, 0.f);
                    kbn[kp] = make_float2(-GPODE_NEG_2LN2 * kb[2 * kp], -GPODE_NEG_2LN2 * kb[2 * kp + 1]);
#pragma unroll
                    for (int j = 0; j < D; ++j) A2[kp][j] = make_float2(0.f, 0.f);
                }
#pragma unroll 2
                for (int m = 0; m < M; ++m) {
                    float kp_[KS];
                    lds_vec<KS>(kp_, kern + m * KS);
                    float d[D], dd[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        d[j] = xr[j] - kp_[j];
                        dd[j] = d[j] * d[j];
                    }
                    float2 tq[D];
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
                        fu[kp] = fadd2(fu[kp], cK);
                        const float2 q = fmul2(kbn[kp], cK);
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            tq[j] = ffma2(q, wn[j][kp], tq[j]);
                            A2[kp][j] = ffma2(dd[j], q, A2[kp][j]);
                        }
                    }
                    if constexpr (kOdd) {
                        float e = 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) e = fmaf(dd[j], wl[j], e);
                        const float cK = kp_[2 * D - 1] * gpode_ex2(e);
                        ful += cK;
                        const float q = kbl * cK;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            tl[j] = q * wl[j];
                            A2l[j] = fmaf(dd[j], q, A2l[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < D; ++j) xb[j] = fmaf(d[j], (tq[j].x + tq[j].y) + tl[j], xb[j]);
                }
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int kp = (k >> 1) < KF ? (k >> 1) : 0;
                    const float fuk = (kOdd && k == D - 1) ? ful : ((k & 1) ? fu[kp].y : fu[kp].x);
                    V[k] = fmaf(kb[k], fr[k] + fuk, V[k]);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const float w_ = (kOdd && k == D - 1) ? wl[j] : ((k & 1) ? wn[j][kp].y : wn[j][kp].x);
                        const float a_ = (kOdd && k == D - 1) ? A2l[j] : ((k & 1) ? A2[kp][j].y : A2[kp][j].x);
                        A[k][j] = fmaf(w_, a_, A[k][j]);
                    }
                }
            }
            };
            // Warpgroups alternate the order of the two parts: one is in its MUFU-bound part (and its issuer runs ahead)
            // while its neighbour is in the FMA-bound one.
            if ((wg & 1) && (flags & 1)) {
                rbf_part();
                rff_part();
                if (flags & 4) consume_G(D - 1, ki - 1);
            } else {
                rff_part();
                rbf_part();
                if (flags & 4) consume_G(D - 1, ki - 1);
            }
            if (row < B) {
#pragma unroll
                for (int j = 0; j < D; ++j) gx[row * D + j] = xb[j];
            }
        } else if (lane == 0) {
            // ---------------- issuer: the two GEMMs of this warpgroup's tile ----------------
            mbar_wait_bounded(&bar.a_full, tile_it & 1);
            tc_fence_after_sync();
            const uint64_t a_desc = umma_smem_desc(gpode_smem_u32(a_tile), 2048, 128);
            const uint32_t idesc1 = umma_idesc_f16(128, kUbChunk);
            const uint32_t idesc2 = umma_idesc_f16(128, 16);
            const int n_it = D * NCH;
            auto issue_theta = [&](int it) {
                const int k = it / NCH, c = it - k * NCH;
                const uint32_t g_ = gi + (uint32_t)it;
                const int buf = g_ & 1;
                if (g_ >= 2) {
                    mbar_wait_bounded(&bar.th_free[buf], ((g_ >> 1) - 1) & 1);
                    tc_fence_after_sync();
                }
                const __half* bb = b1 + (size_t)k * SU * 16 + (size_t)c * kUbChunk * 8;   // 32 features = 4 row groups
                const uint64_t b_desc = umma_smem_desc(gpode_smem_u32(bb), (uint32_t)SU * 16u, 128);
                umma_f16_ss(tmem_wg + (uint32_t)(C_TH + buf * kUbChunk), a_desc, b_desc, idesc1, 0);
                umma_commit(&bar.th_full[buf]);
            };
            auto issue_G = [&](int it) {
                const int k = it / NCH, c = it - k * NCH;
                const uint32_t g_ = gi + (uint32_t)it;
                const uint32_t kk = ki + (uint32_t)k;
                const int kbuf = kk & 1;
                const int sbuf = g_ & 1;
                mbar_wait_bounded(&bar.s_full[sbuf], (g_ >> 1) & 1);
                tc_fence_after_sync();
                if (c == 0 && kk >= 2) {
                    mbar_wait_bounded(&bar.g_free[kbuf], ((kk >> 1) - 1) & 1);
                    tc_fence_after_sync();
                }
                const uint32_t dG = tmem_wg + (uint32_t)(C_G + kbuf * 16);
                const uint32_t aS = tmem_wg + (uint32_t)(C_S + sbuf * 32);
#pragma unroll
                for (int ks = 0; ks < kUbChunk / 16; ++ks) {
                    // 16 features = 2 K-chunks of 8 halfs; each K-chunk holds the two 8-row groups (hi | lo rows)
                    const __half* bb = b2 + (size_t)k * SU * 16 + (size_t)(c * kUbChunk + ks * 16) * 16;
                    const uint64_t b_desc = umma_smem_desc(gpode_smem_u32(bb), 256, 128);
                    umma_f16_ts(dG, aS + (uint32_t)(ks * 8), b_desc, idesc2, (c > 0 || ks > 0) ? 1u : 0u);
                    umma_f16_ts(dG, aS + (uint32_t)(16 + ks * 8), b_desc, idesc2, 1u);
                }
                umma_commit(&bar.s_free[sbuf]);
                if (c == NCH - 1) umma_commit(&bar.g_full[kbuf]);
            };
            // theta runs as far ahead as its two buffers allow, G follows the rows: both event streams are polled
            // (mbarrier.test_wait) so that neither blocks the other
            int nt = 0, ng = 0;
            uint32_t spins = 0;
            if (!(flags & 2)) {   // blocking order: theta one chunk ahead
                for (int it = 0; it < n_it; ++it) {
                    issue_theta(it);
                    if (it > 0) issue_G(it - 1);
                }
                issue_G(n_it - 1);
                ng = n_it;
            }
            while (ng < n_it) {
                bool progress = false;
                if (nt < n_it) {
                    const uint32_t g_ = gi + (uint32_t)nt;
                    if (g_ < 2 || mbar_test(&bar.th_free[g_ & 1], ((g_ >> 1) - 1) & 1)) {
                        issue_theta(nt);
                        ++nt;
                        progress = true;
                    }
                }
                if (ng < nt) {
                    const uint32_t g_ = gi + (uint32_t)ng;
                    const int k = ng / NCH, c = ng - k * NCH;
                    const uint32_t kk = ki + (uint32_t)k;
                    bool ready = mbar_test(&bar.s_full[g_ & 1], (g_ >> 1) & 1);
                    if (ready && c == 0 && kk >= 2) ready = mbar_test(&bar.g_free[kk & 1], ((kk >> 1) - 1) & 1);
                    if (ready) {
                        issue_G(ng);
                        ++ng;
                        progress = true;
                    }
                }
                if (!progress) {
                    if (++spins > (1u << 24)) __trap();
                    __nanosleep(32);
                } else {
                    spins = 0;
                }
            }
            gi += (uint32_t)n_it;
            ki += (uint32_t)D;
        } else {
            gi += (uint32_t)(D * NCH);
            ki += (uint32_t)D;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kUbRowThreads / 32) tmem_dealloc(*tmem_ptr, 512);
    reduce_AV<D>(A, V, acc, red);
}

template <int D>
int launch_vf_bwd_umma(const float* packed, const float* ub, int M, int S, const float* x, const float* f,
                       const float* gf, float* gx, int64_t B, float* acc, cudaStream_t st) {
    const GpodeLayout L = gpode_layout(D, M, S);
    const UbLayout U = ub_layout(D, S);
    const int n_small = L.total - L.off_kern;
    const size_t smem = 512 + (size_t)(n_small + U.total) * 4 + kUbTiles * 4096 + (size_t)kRedFloats<D> * 4;
    if (smem > 227u * 1024u) {
        gpode_set_error("tcgen05 adjoint needs %zu bytes of shared memory", smem);
        return -2;
    }
    GPODE_CUDA(cudaFuncSetAttribute(vf_bwd_umma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 148, dev = 0;
    GPODE_CUDA(cudaGetDevice(&dev));
    GPODE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = ((B + 127) / 128 + kUbTiles - 1) / kUbTiles;
    const int grid = (int)(want < sms ? want : sms);
    // schedule variants measured on B200 (1e6 rows, D = 5): 0 = 1.17 ms, 1 (warpgroups alternate RFF / RBF order) = 1.17,
    // 2 (polling issuer, theta two chunks ahead) = 1.26, 4 (G read one step late) = 1.15, 6 = 1.22, 7 = 1.21
    static const int flags = getenv("GPODE_UMMA_FLAGS") ? atoi(getenv("GPODE_UMMA_FLAGS")) : 4;
    vf_bwd_umma_kernel<D><<<grid, kUbThreads, smem, st>>>(packed, ub, M, S, L.off_kern, n_small, x, f, gf, gx, B, acc, flags);
    GPODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int64_t gpode_packed_ubwd_floats(int D, int S) {
    if (D < 4 || D > 5 || S < 1) return -1;
    return ub_layout(D, S).total;
}

extern "C" int gpode_pack_cache_ubwd(const gpode_cache_t* c, float* packed_ubwd, void* stream) {
    GPODE_CHECK_ARG(c != nullptr && packed_ubwd != nullptr, "cache / packed_ubwd is NULL");
    GPODE_CHECK_ARG(c->D >= 4 && c->D <= 5, "the tcgen05 adjoint covers D = 4, 5, got %d", c->D);
    GPODE_CHECK_ARG(c->S >= 1 && c->omega && c->phase && c->w && c->var, "cache tensor is NULL");
    pack_ub_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(ub_layout(c->D, c->S), c->omega, c->phase, c->w, c->var,
                                                         packed_ubwd);
    GPODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int gpode_vf_bwd_umma(const float* packed, const float* packed_ubwd, int D, int M, int S, const float* x,
                                 const float* f, const float* grad_f, float* grad_x, float* acc, int64_t B,
                                 void* stream) {
    GPODE_CHECK_ARG(packed && packed_ubwd && x && f && grad_f && grad_x && acc, "NULL argument");
    GPODE_CHECK_ARG(D >= 4 && D <= 5, "the tcgen05 adjoint covers D = 4, 5, got %d", D);
    GPODE_CHECK_ARG(M >= 1 && S >= 1 && B >= 0, "bad sizes M=%d S=%d B=%lld", M, S, (long long)B);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 4) return launch_vf_bwd_umma<4>(packed, packed_ubwd, M, S, x, f, grad_f, grad_x, B, acc, st);
    return launch_vf_bwd_umma<5>(packed, packed_ubwd, M, S, x, f, grad_f, grad_x, B, acc, st);
}
