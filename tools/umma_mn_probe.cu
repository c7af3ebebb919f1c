// Probe: does tcgen05.mma kind::tf32 accept an MN-MAJOR shared-memory B operand with the no-swizzle canonical layout, and
// which of LBO / SBO is which? (large_rffb.cu would read Omega^T for its second GEMM from the first GEMM's tile.)
//   nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O2 -I include -I gaussian_process_odes_b200/csrc \
//        -o tools/_build/umma_mn_probe tools/umma_mn_probe.cu && tools/_build/umma_mn_probe
// D[128 x 16] = A[128 x 8] B^T, A K-major (canonical), B[n][k] stored (1) K-major = baseline, (2) MN-major: core matrix =
// 8 K-rows x 16 bytes (4 consecutive n), K-rows 16 bytes apart; blocks of 4 n `mn_stride` apart, blocks of 8 k `k_stride`
// apart. Variants: which stride goes into the LBO field and which into SBO.
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"

static char g_err[256];
void gpode_set_error(const char* fmt, ...) { (void)fmt; }
const char* gpode_last_error(void) { return g_err; }

constexpr int M = 128, N = 32, K = 16;   // two K-steps of 8

__global__ void __launch_bounds__(160) probe(const float* A, const float* B, float* D, int variant) {
    __shared__ __align__(128) float sa[M * K];
    __shared__ __align__(1024) float sb[N * K * 2];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A canonical K-major: [K/4 chunks][M/8 groups][8 rows][4 floats]
    for (int i = tid; i < M * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        sa[(k >> 2) * (M * 4) + (r >> 3) * 32 + (r & 7) * 4 + (k & 3)] = A[i];
    }
    for (int i = tid; i < N * K * 2; i += blockDim.x) sb[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        int o;
        if (variant == 0) o = (k >> 2) * (N * 4) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);             // K-major
        else if (variant < 3) o = (n >> 2) * 32 + (k >> 3) * (N / 4 * 32) + (k & 7) * 4 + (n & 3);      // MN-major
        else {   // MN-major, 128-byte swizzle: 32 consecutive n (128 bytes) per k row, 8 k rows per 1 KB atom
            uint32_t byte = (uint32_t)((k >> 3) * 1024 + (k & 7) * 128 + n * 4);
            byte ^= ((byte >> 7) & 7u) << 4;
            o = (int)(byte >> 2);
        }
        sb[o] = B[i];
    }
    if (tid == 128) gpode_mbar_init(&bar, 1);
    if (warp == 4) { __syncwarp(); tmem_alloc(&tmem_ptr, 32u); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = tmem_ptr;
    if (tid == 128) {
        uint32_t idesc = umma_idesc_tf32(M, N);
        const uint64_t ad = umma_smem_desc(gpode_smem_u32(sa), M * 16, 128);
        uint64_t bd;
        uint64_t step_b;
        const uint32_t mn_stride = 128, k_stride = N / 4 * 128;   // bytes
        if (variant == 0) { bd = umma_smem_desc(gpode_smem_u32(sb), N * 16, 128); step_b = (2u * N * 16) >> 4; }
        else {
            idesc |= 1u << 16;
            if (variant == 1) { bd = umma_smem_desc(gpode_smem_u32(sb), k_stride, mn_stride); step_b = k_stride >> 4; }  // LBO = K-block stride
            else if (variant == 2) { bd = umma_smem_desc(gpode_smem_u32(sb), mn_stride, k_stride); step_b = k_stride >> 4; }
            else {   // SWIZZLE_128B (layout type 2, bits 61-63): LBO = stride between 32-n groups, SBO = stride between 8-k groups
                bd = umma_smem_desc(gpode_smem_u32(sb), variant == 3 ? 1024 : 2048, variant == 3 ? 1024 : 1024) | ((uint64_t)2 << 61);
                if (variant == 5) bd = umma_smem_desc(gpode_smem_u32(sb), 1024, 2048) | ((uint64_t)2 << 61);
                step_b = 1024 >> 4;
            }
        }
        const uint64_t step_a = (2u * M * 16) >> 4;
        umma_tf32_ss(tb, ad, bd, idesc, 0u);
        umma_tf32_ss(tb, ad + step_a, bd + step_b, idesc, 1u);
        umma_commit(&bar);
    }
    if (warp < 4) {
        mbar_wait_bounded(&bar, 0);
        tc_fence_after_sync();
        uint32_t r[32];
        tmem_ld32_issue(tb + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait(r);
        for (int n = 0; n < N; ++n) D[tid * N + n] = __uint_as_float(r[n]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tb, 32u);
}

int main() {
    float hA[M * K], hB[N * K], hD[M * N], ref[M * N];
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = (float)((r * 7 + k * 3) % 11) - 5.f;
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = (float)((n * 5 + k * 2) % 13) - 6.f;
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += hA[r * K + k] * hB[n * K + k];
        ref[r * N + n] = s;
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    for (int v = 0; v < 6; ++v) {
        cudaMemset(dD, 0, sizeof(hD));
        probe<<<1, 160>>>(dA, dB, dD, v);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double worst = 0, mag = 0;
        int nz = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs((double)hD[i] - ref[i]); if (d > worst) worst = d; if (fabs(ref[i]) > mag) mag = fabs(ref[i]); if (hD[i] != 0.f) ++nz; }
        printf("{\"variant\": %d, \"what\": \"%s\", \"cuda\": \"%s\", \"max_abs_err\": %.4g, \"max_abs_ref\": %.4g, \"nonzero\": %d, \"D00\": %.3f, \"ref00\": %.3f, \"D[5][3]\": %.3f, \"ref[5][3]\": %.3f}\n", v,
               v == 0 ? "B K-major (baseline)" : (v == 1 ? "B MN-major, LBO = K-block stride, SBO = MN-block stride" : (v == 2 ? "B MN-major, LBO = MN-block stride, SBO = K-block stride" : "B MN-major SWIZZLE_128B")),
               cudaGetErrorString(e), worst, mag, nz, hD[0], ref[0], hD[5 * N + 3], ref[5 * N + 3]);
    }
    return 0;
}
