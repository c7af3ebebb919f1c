// Error plumbing, ABI version, and the cache repacking kernel.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void gpode_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gpode_last_error(void) { return g_err; }
extern "C" int gpode_abi_version(void) { return GPODE_B200_ABI_VERSION; }

extern "C" int64_t gpode_packed_floats(int D, int M, int S) {
    if (D < 1 || M < 1 || S < 1) return -1;
    return gpode_layout(D, M, S).total;
}

namespace {

// One thread per packed record. Source layouts are the reference's cache tensors (src/core/dsvgp.py:100-103,122):
// omega (j,s,k), phase (s,k), w (s,k), Z (m,j), nu (k,m), ell (k,j), var (k).
__global__ void pack_kernel(const GpodeLayout L, const float* __restrict__ omega, const float* __restrict__ phase,
                            const float* __restrict__ w, const float* __restrict__ Z, const float* __restrict__ nu,
                            const float* __restrict__ ell, const float* __restrict__ var, float* __restrict__ out) {
    const int D = L.D, M = L.M, S = L.S, S2 = L.S2;
    const int n_rff = D * S2, n_kern = M, n_il = D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rff + n_kern + n_il; i += gridDim.x * blockDim.x) {
        if (i < n_rff) {
            const int k = i / S2, s2 = i - k * S2;
            // element e of the record lives at chunk e/4, slot e%4 of the chunk-major group layout
            float* base = out + L.off_rff + ((size_t)k * L.S2P + (s2 & ~31)) * L.RP + (s2 & 31) * 4;
#define GPODE_REC(e) base[((e) >> 2) * 128 + ((e) & 3)]
            for (int h = 0; h < 2; ++h) {
                const int s = 2 * s2 + h;
                const bool ok = s < S;
                for (int j = 0; j < D; ++j) GPODE_REC(2 * j + h) = ok ? omega[((size_t)j * S + s) * D + k] : 0.f;
                GPODE_REC(2 * D + h) = ok ? phase[s * D + k] : 0.f;
                GPODE_REC(2 * D + 2 + h) = ok ? w[s * D + k] * sqrtf(var[k] / (float)S) : 0.f;
            }
            for (int j = 2 * D + 4; j < L.RP; ++j) GPODE_REC(j) = 0.f;
#undef GPODE_REC
        } else if (i < n_rff + n_kern) {
            const int m = i - n_rff;
            float* o = out + L.off_kern + (size_t)m * L.KS;
            for (int j = 0; j < D; ++j) o[j] = Z[m * D + j];
            for (int k = 0; k < L.KS - D; ++k) o[D + k] = (nu && k < D) ? var[k] * nu[k * M + m] : 0.f;
        } else {
            const int j = i - n_rff - n_kern;
            float* o = out + L.off_il + (size_t)j * L.WP;
            for (int k = 0; k < L.WP; ++k)
                o[k] = k < D ? -GPODE_HALF_LOG2E / (ell[k * D + j] * ell[k * D + j]) : 0.f;
        }
    }
}

}  // namespace

extern "C" int gpode_pack_cache(const gpode_cache_t* c, float* packed, void* stream) {
    GPODE_CHECK_ARG(c != nullptr && packed != nullptr, "cache / packed is NULL");
    GPODE_CHECK_ARG(c->D >= 1 && c->D <= GPODE_MAX_D, "state dimension D=%d outside 1..%d", c->D, GPODE_MAX_D);
    GPODE_CHECK_ARG(c->M >= 1 && c->S >= 1, "M=%d and S=%d must be positive", c->M, c->S);
    GPODE_CHECK_ARG(c->omega && c->phase && c->w && c->Z && c->ell && c->var, "cache tensor is NULL");
    const GpodeLayout L = gpode_layout(c->D, c->M, c->S);
    const int n = c->D * L.S2 + c->M + c->D;
    pack_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(L, c->omega, c->phase, c->w, c->Z, c->nu, c->ell,
                                                                   c->var, packed);
    GPODE_LAUNCH_CHECK();
    return 0;
}
