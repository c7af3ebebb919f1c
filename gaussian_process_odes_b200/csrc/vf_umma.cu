// EXPERIMENTAL: the GP vector field with the random-Fourier-feature projection on the 5th-generation tensor cores.
//
//   theta[row, s] = sum_j x[row, j] Omega[j, s, k] + phase[s, k]     as     [x | 1 | 0] (128 x 8)  x  B_k (8 x N)
//
// tcgen05.mma kind::tf32 (M = 128 rows, N = 128-feature chunks, K = 8), operands in shared memory, accumulators in
// TMEM, error-compensated 3xTF32. The phase rides in the spare K slot D (the state tile holds a constant 1 there), so
// the accumulator IS theta. Thread = row = TMEM lane: after tcgen05.ld each thread owns its row's thetas in registers
// and does cos + weighted sum (1 MUFU + 1 FFMA per feature instead of D+1 FMAs); the RBF term stays on FFMA2.
//
// CTA = 160 threads: warps 0-3 own the 128 rows of a tile (warp w = TMEM lanes 32w..32w+31), warp 4 allocates TMEM
// and its lane 0 issues the MMAs. Two 128-column accumulator buffers ping-pong between the tensor core and the rows.
#include <stdlib.h>
#include "umma.cuh"
#include "vf_mma.cuh"

namespace {

constexpr int kUmmaThreads = 160;
constexpr int kRows = 128;
constexpr int kChunk = 128;   // features (TMEM columns) per accumulator buffer
constexpr int kTmemCols = 256;

struct UmmaSmem {  // byte offsets inside dynamic shared memory
    static constexpr int bar_afull = 0, bar_full = 8, bar_empty = 24, bar_stage = 40, tmem_ptr = 48;
    static constexpr int a_hi = 64, a_lo = 64 + 4096, params = 64 + 8192;
};

template <int D>
__global__ void __launch_bounds__(kUmmaThreads)
vf_fwd_umma_kernel(const float* __restrict__ packed, const int M, const int S, const int off_kern,
                   const int n_small, const int off_umma, const int n_umma, const float* __restrict__ x,
                   float* __restrict__ f, const int64_t B) {
    static_assert(D <= 7, "the phase needs a spare K slot");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar_afull = reinterpret_cast<uint64_t*>(smem + UmmaSmem::bar_afull);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + UmmaSmem::bar_full);    // [2]
    uint64_t* bar_empty = reinterpret_cast<uint64_t*>(smem + UmmaSmem::bar_empty);  // [2]
    uint64_t* bar_stage = reinterpret_cast<uint64_t*>(smem + UmmaSmem::bar_stage);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + UmmaSmem::tmem_ptr);
    float* a_hi = reinterpret_cast<float*>(smem + UmmaSmem::a_hi);
    float* a_lo = reinterpret_cast<float*>(smem + UmmaSmem::a_lo);
    float* sp = reinterpret_cast<float*>(smem + UmmaSmem::params);  // [kern | il] then the umma records
    const float* um = sp + n_small;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int SU = (S + 31) & ~31;
    const int n_chunks_k = (SU + kChunk - 1) / kChunk;  // chunks per output dimension

    if (tid == 128) {
        gpode_mbar_init(bar_afull, kRows);
        gpode_mbar_init(bar_full, 1);
        gpode_mbar_init(bar_full + 1, 1);
        gpode_mbar_init(bar_empty, kRows);
        gpode_mbar_init(bar_empty + 1, kRows);
        gpode_mbar_init(bar_stage, 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gpode_smem_u32(bar_stage)),
                     "r"((uint32_t)(n_small + n_umma) * 4u)
                     : "memory");
        for (uint32_t off = 0; off < (uint32_t)n_small * 4u; off += 32768u) {
            const uint32_t n = (uint32_t)n_small * 4u - off < 32768u ? (uint32_t)n_small * 4u - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             gpode_smem_u32((const char*)sp + off)),
                         "l"((const char*)(packed + off_kern) + off), "r"(n), "r"(gpode_smem_u32(bar_stage))
                         : "memory");
        }
        for (uint32_t off = 0; off < (uint32_t)n_umma * 4u; off += 32768u) {
            const uint32_t n = (uint32_t)n_umma * 4u - off < 32768u ? (uint32_t)n_umma * 4u - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             gpode_smem_u32((const char*)um + off)),
                         "l"((const char*)(packed + off_umma) + off), "r"(n), "r"(gpode_smem_u32(bar_stage))
                         : "memory");
        }
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(tmem_ptr, kTmemCols);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    mbar_wait_bounded(bar_stage, 0);

    const int64_t n_tiles = (B + kRows - 1) / kRows;
    uint32_t g = 0;        // running chunk counter (buffer = g & 1, use number = g >> 1), identical in every thread
    uint32_t tile_it = 0;  // tiles done by this CTA
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
        if (warp < 4) {
            // ---- rows: publish the state tile, then consume theta chunk by chunk ----
            const int64_t row = tile * kRows + tid;
            float xr[1][D], fr[1][D];
#pragma unroll
            for (int j = 0; j < D; ++j) xr[0][j] = row < B ? x[row * D + j] : 0.f;
            {
                float hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float v = q < D ? xr[0][q < D ? q : 0] : (q == D ? 1.f : 0.f);
                    gpode_split_tf32_rn(v, hi[q], lo[q]);
                }
                const int o = (tid >> 3) * 64 + (tid & 7) * 4;  // core matrix of the row's group, then its 16-byte line
                *reinterpret_cast<float4*>(a_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a_hi + o + 32) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<float4*>(a_lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(a_lo + o + 32) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_afull);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float* aw = um + (size_t)k * GPODE_UMMA_REC(SU) + 16 * SU;
                float acc0 = 0.f, acc1 = 0.f;
                for (int c = 0; c < n_chunks_k; ++c, ++g) {
                    const int buf = g & 1;
                    const int ncols = SU - c * kChunk < kChunk ? SU - c * kChunk : kChunk;
                    mbar_wait_bounded(bar_full + buf, (g >> 1) & 1);
                    tc_fence_after_sync();
                    // software pipeline over the 32-column blocks: the next block's TMEM load is in flight while the
                    // current block goes through the MUFU
                    const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * kChunk);
                    const float* a = aw + c * kChunk;
                    uint32_t ra[32], rb[32];
                    auto consume = [&](const uint32_t (&r)[32], const float* __restrict__ wgt) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wgt + i);
                            acc0 = fmaf(w4.x, __cosf(__uint_as_float(r[i])), acc0);
                            acc1 = fmaf(w4.y, __cosf(__uint_as_float(r[i + 1])), acc1);
                            acc0 = fmaf(w4.z, __cosf(__uint_as_float(r[i + 2])), acc0);
                            acc1 = fmaf(w4.w, __cosf(__uint_as_float(r[i + 3])), acc1);
                        }
                    };
                    tmem_ld32_issue(t0, ra);
                    for (int cc = 0; cc < ncols; cc += 64) {
                        tmem_ld_wait(ra);
                        if (cc + 32 < ncols) tmem_ld32_issue(t0 + cc + 32, rb);
                        consume(ra, a + cc);
                        if (cc + 32 < ncols) {
                            tmem_ld_wait(rb);
                            if (cc + 64 < ncols) tmem_ld32_issue(t0 + cc + 64, ra);
                            consume(rb, a + cc + 32);
                        }
                    }
                    tc_fence_before_sync();
                    mbar_arrive(bar_empty + buf);
                }
                fr[0][k] = acc0 + acc1;
            }
            // ---- RBF (pathwise update) term on FFMA2, as in the register kernels ----
            constexpr int KS = VfShape<D>::KS;
            rbf_eval_partial<D, 1>(sp, sp + M * KS, M, xr, fr, 0, 1);
            if (row < B) {
#pragma unroll
                for (int j = 0; j < D; ++j) f[row * D + j] = fr[0][j];
            }
        } else {
            // ---- warp 4: lane 0 issues the MMAs ----
            if (tid == 128) {
                mbar_wait_bounded(bar_afull, tile_it & 1);
                tc_fence_after_sync();
                const uint64_t adesc_hi = umma_smem_desc(gpode_smem_u32(a_hi), 128, 256);
                const uint64_t adesc_lo = umma_smem_desc(gpode_smem_u32(a_lo), 128, 256);
                for (int k = 0; k < D; ++k) {
                    const float* rec = um + (size_t)k * GPODE_UMMA_REC(SU);
                    for (int c = 0; c < n_chunks_k; ++c, ++g) {
                        const int buf = g & 1;
                        const int ncols = SU - c * kChunk < kChunk ? SU - c * kChunk : kChunk;
                        if (g >= 2) {
                            mbar_wait_bounded(bar_empty + buf, ((g >> 1) - 1) & 1);
                            tc_fence_after_sync();
                        }
                        const uint32_t d = tmem_base + (uint32_t)(buf * kChunk);
                        const uint64_t b_hi = umma_smem_desc(gpode_smem_u32(rec + (size_t)c * kChunk * 8), 128, 256);
                        const uint64_t b_lo =
                            umma_smem_desc(gpode_smem_u32(rec + 8 * SU + (size_t)c * kChunk * 8), 128, 256);
                        const uint32_t idesc = umma_idesc_tf32(kRows, ncols);
                        umma_tf32_ss(d, adesc_hi, b_hi, idesc, 0);
                        umma_tf32_ss(d, adesc_lo, b_hi, idesc, 1);
                        umma_tf32_ss(d, adesc_hi, b_lo, idesc, 1);
                        umma_commit(bar_full + buf);
                    }
                }
            } else {
                g += (uint32_t)(D * n_chunks_k);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

template <int D>
int launch_umma(const float* packed, int M, int S, const float* x, float* f, int64_t B, cudaStream_t st) {
    const GpodeLayout L = gpode_layout(D, M, S);
    const int n_small = L.total - L.off_kern, n_umma = D * GPODE_UMMA_REC(L.SU);
    const size_t smem = UmmaSmem::params + (size_t)(n_small + n_umma) * 4;
    GPODE_CUDA(cudaFuncSetAttribute(vf_fwd_umma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 148, dev = 0;
    GPODE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vf_fwd_umma_kernel<D>, kUmmaThreads, smem));
    if (occ < 1) {
        gpode_set_error("umma kernel does not fit on an SM (smem %zu bytes)", smem);
        return -2;
    }
    // the occupancy query answers 1 for this kernel although two CTAs co-reside (measured: 1.7x throughput with a
    // grid of two CTAs per SM); the real limits are TMEM (512 columns per SM) and shared memory
    occ = (int)((227u * 1024u) / (smem + 1024u));
    if (occ > 512 / kTmemCols) occ = 512 / kTmemCols;
    if (occ < 1) occ = 1;
    GPODE_CUDA(cudaGetDevice(&dev));
    GPODE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (B + kRows - 1) / kRows, cap = (int64_t)sms * occ;
    vf_fwd_umma_kernel<D><<<(unsigned)(tiles < cap ? tiles : cap), kUmmaThreads, smem, st>>>(
        packed, M, S, L.off_kern, n_small, L.off_umma, n_umma, x, f, B);
    GPODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int gpode_vf_fwd_umma(const float* packed, int D, int M, int S, const float* x, float* f, int64_t B,
                                 void* stream) {
    GPODE_CHECK_ARG(packed && x && f, "NULL argument");
    GPODE_CHECK_ARG(D >= 2 && D <= 7, "the tcgen05 path covers state dimensions 2..7, got %d", D);
    GPODE_CHECK_ARG(M >= 1 && S >= 1 && B >= 0, "bad sizes M=%d S=%d B=%lld", M, S, (long long)B);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
        case 2: return launch_umma<2>(packed, M, S, x, f, B, st);
        case 3: return launch_umma<3>(packed, M, S, x, f, B, st);
        case 4: return launch_umma<4>(packed, M, S, x, f, B, st);
        case 5: return launch_umma<5>(packed, M, S, x, f, B, st);
        case 6: return launch_umma<6>(packed, M, S, x, f, B, st);
        case 7: return launch_umma<7>(packed, M, S, x, f, B, st);
    }
    return -1;
}
