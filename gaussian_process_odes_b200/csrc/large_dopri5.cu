// Adaptive dopri5 for 8 < D <= 64 with the controller ON THE DEVICE (forward only).
//
// torchdiffeq 0.2.0's dopri5 as called by Flow.forward (reference src/core/flow.py:84-90; restated in
// oracle/torchdiffeq_shim): Dormand-Prince 5(4), FSAL, whole-batch RMS error norm, float64 time, Hairer initial step,
// quartic dense output. The reference synchronises the host once per attempt (accept / reject is a Python `if`); round 1
// of this library did the same at these state dimensions (an eager host loop around the tiled vector-field kernel).
// Here one attempt is a short chain of launches -- six stage evaluations on the tcgen05 vector-field kernels
// (large_umma.cu), element-wise stage / error kernels, a one-CTA controller kernel that decides accept / reject, updates
// (t, dt) in float64 and names the outputs to interpolate, and an element-wise accept kernel -- and the chain is the body
// of a CUDA-graph WHILE node whose condition the controller kernel sets with cudaGraphSetConditional: the loop runs on
// the device until the last output time is covered, the host enqueues one graph launch and never reads a value.
// Arithmetic (operation order, float32 state, float64 time) follows dopri5_impl.cuh, which serves D <= 8.
#include "common.cuh"
#include "../../include/gpode_b200.h"
#include <deque>
#include <mutex>

int gpode_vf_large_eval(const float* packed_large, const gpode_cache_t* c, const float* x, float* tmp, float* f,
                        int64_t B, cudaStream_t st);  // large_umma.cu

namespace {

__device__ __constant__ float ldBeta[6][6] = {
    {1.f / 5, 0, 0, 0, 0, 0},
    {3.f / 40, 9.f / 40, 0, 0, 0, 0},
    {44.f / 45, -56.f / 15, 32.f / 9, 0, 0, 0},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0, 0},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176),
     (float)(-5103.0 / 18656), 0},
    {(float)(35.0 / 384), 0.f, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784),
     (float)(11.0 / 84)}};
__device__ __constant__ float ldCErr[7] = {
    (float)(35.0 / 384 - 1951.0 / 21600), 0.f, (float)(500.0 / 1113 - 22642.0 / 50085),
    (float)(125.0 / 192 - 451.0 / 720), (float)(-2187.0 / 6784 + 12231.0 / 42400), (float)(11.0 / 84 - 649.0 / 6300),
    (float)(-1.0 / 60)};
__device__ __constant__ float ldCMid[7] = {
    (float)(6025192743.0 / 30085553152.0 / 2), 0.f, (float)(51252292925.0 / 65400821598.0 / 2),
    (float)(-2691868925.0 / 45128329728.0 / 2), (float)(187940372067.0 / 1594534317056.0 / 2),
    (float)(-1776094331.0 / 19743644256.0 / 2), (float)(11237099.0 / 235043384.0 / 2)};

struct LdCtrl {
    double t1, dt, dir;        // current time (already multiplied by dir), step size, +1 / -1
    double t1_used, t1n_used;  // the interval of the attempt just decided (for the accept kernel)
    float dts_used, h0, d1;
    int jout, jlo, jhi, accept, done;
    int nfe, n_acc, n_rej, status, attempt, max_attempts;
};

constexpr int kLdThreads = 256;
constexpr int kLdMaxBlocks = 2368;

inline unsigned ld_grid(int64_t n) {
    const int64_t g = (n + kLdThreads - 1) / kLdThreads;
    return (unsigned)(g < kLdMaxBlocks ? (g < 1 ? 1 : g) : kLdMaxBlocks);
}

// block sum of up to two float64 values -> partial[blockIdx.x (+ gridDim.x)]
__device__ __forceinline__ void ld_block_sums(double a, double b, double* __restrict__ partial, const bool two) {
    __shared__ double sa[kLdThreads / 32], sb[kLdThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sa[threadIdx.x >> 5] = a;
        sb[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int i = 0; i < kLdThreads / 32; ++i) {
            ta += sa[i];
            tb += sb[i];
        }
        partial[blockIdx.x] = ta;
        if (two) partial[gridDim.x + blockIdx.x] = tb;
    }
}

// fixed-order sum of n partials by one CTA (thread-strided, then warps and lanes in index order)
__device__ double ld_total(const double* __restrict__ partial, const int n) {
    __shared__ double s[kLdThreads];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += kLdThreads) a += partial[i];
    s[threadIdx.x] = a;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kLdThreads; ++i) t += s[i];
    __syncthreads();
    if (threadIdx.x == 0) s[0] = t;
    __syncthreads();
    t = s[0];
    __syncthreads();
    return t;
}

__global__ void ld_start_kernel(LdCtrl* c, const double* __restrict__ t, const int Tg, const int max_attempts) {
    const double dir = (Tg > 1 && t[Tg - 1] < t[0]) ? -1.0 : 1.0;
    c->dir = dir;
    c->t1 = dir * t[0];
    c->dt = 0.0;
    c->jout = 1; c->jlo = c->jhi = 0; c->accept = 0; c->done = Tg <= 1;
    c->nfe = 2; c->n_acc = 0; c->n_rej = 0; c->status = 0; c->attempt = 0; c->max_attempts = max_attempts;
}

// s0 = sum (y / sc)^2, s1 = sum (f / sc)^2, sc = atol + |y| rtol
__global__ void __launch_bounds__(kLdThreads)
ld_norm01_kernel(const float* __restrict__ y, const float* __restrict__ f, const float rtol, const float atol,
                 const int64_t n, double* __restrict__ partial) {
    double s0 = 0.0, s1 = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float sc = atol + fabsf(y[e]) * rtol;
        const float q0 = y[e] / sc, q1 = f[e] / sc;
        s0 += (double)(q0 * q0);
        s1 += (double)(q1 * q1);
    }
    ld_block_sums(s0, s1, partial, true);
}

__global__ void __launch_bounds__(kLdThreads)
ld_init1_kernel(LdCtrl* c, const double* __restrict__ partial, const int nb, const double n_elem) {
    const double t0 = ld_total(partial, nb), t1 = ld_total(partial + nb, nb);
    if (threadIdx.x == 0) {
        const float d0 = (float)sqrt(t0 / n_elem), d1 = (float)sqrt(t1 / n_elem);
        c->h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
        c->d1 = d1;
    }
}

// out = y + h0 * fsign * f
__global__ void ld_euler_kernel(const LdCtrl* __restrict__ c, const float* __restrict__ y, const float* __restrict__ f,
                                float* __restrict__ out, const int64_t n) {
    const float h0 = c->h0, fs = (float)c->dir;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = y[e] + h0 * (fs * f[e]);
}

__global__ void __launch_bounds__(kLdThreads)
ld_norm2_kernel(const float* __restrict__ y, const float* __restrict__ f0, const float* __restrict__ f1, const float rtol,
                const float atol, const int64_t n, double* __restrict__ partial) {
    double s2 = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float sc = atol + fabsf(y[e]) * rtol;
        const float q = (f1[e] - f0[e]) / sc;   // the direction sign cancels in the square
        s2 += (double)(q * q);
    }
    ld_block_sums(s2, 0.0, partial, false);
}

// the part of the controller that decides whether another attempt is needed (top of torchdiffeq's while loop)
__device__ void ld_loop_head(LdCtrl* c, const double* __restrict__ t, const int Tg) {
    while (c->jout < Tg && !(c->dir * t[c->jout] > c->t1)) ++c->jout;
    if (c->jout >= Tg) {
        c->done = 1;
    } else if (c->attempt >= c->max_attempts || !(c->t1 + c->dt > c->t1)) {
        c->status = c->attempt >= c->max_attempts ? 1 : 2;   // attempt limit / step-size underflow
        c->done = 1;
    }
}

__global__ void __launch_bounds__(kLdThreads)
ld_init2_kernel(LdCtrl* c, const double* __restrict__ partial, const int nb, const double n_elem,
                const double* __restrict__ t, const int Tg) {
    const double t2 = ld_total(partial, nb);
    if (threadIdx.x == 0) {
        const float h0 = c->h0, d1 = c->d1;
        const float d2 = (float)sqrt(t2 / n_elem) / h0;
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
        else h1 = powf(0.01f / fmaxf(d1, d2), 1.0f / 5.0f);
        c->dt = (double)fminf(100.f * h0, h1);
        if (!c->done) ld_loop_head(c, t, Tg);
    }
}

struct LdBufs {
    float* Y;      // current state
    float* K[7];   // raw vector-field values of the attempt's stages; K[0] = f(Y) (FSAL)
    float* Ys;     // stage input
    float* Y1;     // 7th stage input = the proposed state
    float* YM;     // dense-output midpoint
    float* tmp;    // scratch of the vector-field evaluation
};

// stage input i (1..6): Y_i = Y + sum_{l < i} (fsign K_l) (beta_{i-1,l} dt)
__global__ void ld_stage_kernel(const LdCtrl* __restrict__ c, const int i, const LdBufs b, float* __restrict__ out,
                                const int64_t n) {
    if (c->done) return;
    const float dts = (float)c->dt, fs = (float)c->dir;
    float bd[6];
#pragma unroll
    for (int l = 0; l < 6; ++l) bd[l] = __fmul_rn(ldBeta[i - 1][l], dts);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < 6; ++l)
            if (l < i) s = fmaf(fs * b.K[l][e], bd[l], s);
        out[e] = b.Y[e] + s;
    }
}

__global__ void __launch_bounds__(kLdThreads)
ld_err_kernel(const LdCtrl* __restrict__ c, const LdBufs b, const float rtol, const float atol, const int64_t n,
              double* __restrict__ partial) {
    double se = 0.0;
    if (!c->done) {
        const float dts = (float)c->dt, fs = (float)c->dir;
        float ce[7], cm[7];
#pragma unroll
        for (int l = 0; l < 7; ++l) {
            ce[l] = __fmul_rn(dts, ldCErr[l]);
            cm[l] = __fmul_rn(dts, ldCMid[l]);
        }
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
            float er = 0.f, m = 0.f;
#pragma unroll
            for (int l = 0; l < 7; ++l) {
                const float k = fs * b.K[l][e];
                er = fmaf(k, ce[l], er);
                m = fmaf(k, cm[l], m);
            }
            const float y = b.Y[e];
            b.YM[e] = y + m;
            const float tol = atol + rtol * fmaxf(fabsf(y), fabsf(b.Y1[e]));
            const float q = er / tol;
            se += (double)(q * q);
        }
    }
    ld_block_sums(se, 0.0, partial, false);
}

// accept / reject, step-size update (_optimal_step_size: safety 0.9, ifactor 10, dfactor 0.2, order 5), loop condition
__global__ void __launch_bounds__(kLdThreads)
ld_ctrl_kernel(LdCtrl* c, const double* __restrict__ partial, const int nb, const double n_elem,
               const double* __restrict__ t, const int Tg, const cudaGraphConditionalHandle handle) {
    const double tot = ld_total(partial, nb);
    if (threadIdx.x != 0) return;
    if (!c->done) {
        const float ratio = (float)sqrt(tot / n_elem);
        const bool accept = ratio <= 1.0f;
        c->nfe += 6;
        c->accept = accept ? 1 : 0;
        c->dts_used = (float)c->dt;
        if (accept) {
            const double t1n = c->t1 + c->dt;
            int jend = c->jout;
            while (jend < Tg && !(c->dir * t[jend] > t1n)) ++jend;
            c->jlo = c->jout;
            c->jhi = jend;
            c->t1_used = c->t1;
            c->t1n_used = t1n;
            c->jout = jend;
            c->t1 = t1n;
            ++c->n_acc;
        } else {
            ++c->n_rej;
        }
        if (ratio == 0.f) {
            c->dt = c->dt * 10.0;
        } else {
            const double dfac = ratio < 1.f ? 1.0 : 0.2;
            c->dt = c->dt * fmin(10.0, fmax(0.9 / pow((double)ratio, 0.2), dfac));
        }
        ++c->attempt;
        ld_loop_head(c, t, Tg);
    } else {
        c->accept = 0;
    }
    cudaGraphSetConditional(handle, c->done ? 0u : 1u);
}

// accepted step: dense output at the requested times inside it, then Y <- Y1, K0 <- K6 (FSAL)
__global__ void ld_accept_kernel(const LdCtrl* __restrict__ c, const LdBufs b, const double* __restrict__ t,
                                 float* __restrict__ xs, const int64_t n) {
    if (!c->accept) return;
    const float dts = c->dts_used, fs = (float)c->dir;
    const int jlo = c->jlo, jhi = c->jhi;
    const double t1 = c->t1_used, t1n = c->t1n_used, dir = c->dir;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float yn = b.Y1[e], kn = b.K[6][e];
        if (jhi > jlo) {
            const float y0 = b.Y[e], f0 = fs * b.K[0][e], fn = fs * kn, ymj = b.YM[e];
            const float ca = 2 * dts * (fn - f0) - 8 * (yn + y0) + 16 * ymj;
            const float cb = dts * (5 * f0 - 3 * fn) + 18 * y0 + 14 * yn - 32 * ymj;
            const float cc = dts * (fn - 4 * f0) - 11 * y0 - 5 * yn + 16 * ymj;
            const float cd = dts * f0;
            for (int jo = jlo; jo < jhi; ++jo) {
                const float x = (float)((dir * t[jo] - t1) / (t1n - t1));
                float total = y0 + x * cd;
                float xp = x * x;
                total = total + xp * cc;
                xp = xp * x;
                total = total + xp * cb;
                xp = xp * x;
                total = total + xp * ca;
                xs[(int64_t)jo * n + e] = total;
            }
        }
        b.Y[e] = yn;
        b.K[0][e] = kn;
    }
}

__global__ void ld_stats_kernel(const LdCtrl* __restrict__ c, int32_t* __restrict__ stats) {
    stats[0] = c->nfe;
    stats[1] = c->n_acc;
    stats[2] = c->n_rej;
    stats[3] = c->status;
}

// executable graphs of earlier calls: destroyed once the event recorded behind their launch has completed
struct LdPending {
    cudaGraphExec_t exec;
    cudaGraph_t graph;
    cudaEvent_t ev;
};
std::mutex g_ld_mutex;
std::deque<LdPending> g_ld_pending;
cudaStream_t g_ld_capture_stream = nullptr;

void ld_reap(bool all) {
    while (!g_ld_pending.empty()) {
        LdPending& p = g_ld_pending.front();
        if (!all && cudaEventQuery(p.ev) != cudaSuccess) {
            (void)cudaGetLastError();
            break;
        }
        if (all) cudaEventSynchronize(p.ev);
        cudaGraphExecDestroy(p.exec);
        cudaGraphDestroy(p.graph);
        cudaEventDestroy(p.ev);
        g_ld_pending.pop_front();
    }
}

}  // namespace

// work layout (floats): Y | K0..K6 | Ys | Y1 | YM | tmp  (12 B D), then 8-byte aligned: partial sums
// (2 * kLdMaxBlocks float64) and the controller block
extern "C" int64_t gpode_dopri5_large_work_floats(int D, int64_t B) {
    return 12 * B * (int64_t)D + 2 + 2 * (2 * (int64_t)kLdMaxBlocks) + (int64_t)(sizeof(LdCtrl) + 3) / 4 + 2;
}

extern "C" int gpode_dopri5_fwd_large(const float* packed_large, const gpode_cache_t* c, const float* x0, const double* t,
                                      int Tg, int64_t B, double rtol, double atol, float* xs, float* work,
                                      int32_t* stats_out, int max_attempts, void* stream) {
    GPODE_CHECK_ARG(c != nullptr && c->D > GPODE_MAX_D && c->D <= GPODE_MAX_D_LARGE, "large-D path needs %d < D <= %d",
                    GPODE_MAX_D, GPODE_MAX_D_LARGE);
    GPODE_CHECK_ARG(Tg >= 1 && B >= 0 && max_attempts >= 1, "bad sizes Tg=%d B=%lld", Tg, (long long)B);
    GPODE_CHECK_ARG(packed_large && x0 && t && xs && work && stats_out, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    GPODE_CUDA(cudaStreamIsCapturing(st, &cap));
    GPODE_CHECK_ARG(cap == cudaStreamCaptureStatusNone,
                    "gpode_dopri5_fwd_large builds its own CUDA graph (device-side while loop) and cannot run inside a "
                    "stream capture");
    const int D = c->D;
    const int64_t n = B * D;
    if (B == 0) {
        GPODE_CUDA(cudaMemsetAsync(stats_out, 0, 4 * sizeof(int32_t), st));
        return 0;
    }
    LdBufs b;
    b.Y = work;
    for (int l = 0; l < 7; ++l) b.K[l] = work + (1 + l) * n;
    b.Ys = work + 8 * n;
    b.Y1 = work + 9 * n;
    b.YM = work + 10 * n;
    b.tmp = work + 11 * n;
    uintptr_t p = reinterpret_cast<uintptr_t>(work + 12 * n);
    p = (p + 7) & ~(uintptr_t)7;
    double* partial = reinterpret_cast<double*>(p);
    LdCtrl* ctrl = reinterpret_cast<LdCtrl*>(partial + 2 * kLdMaxBlocks);
    const unsigned g = ld_grid(n);
    const double n_elem = (double)n;
    const float rt = (float)rtol, at = (float)atol;

    // ---- start: y = x0, f0, initial step (two evaluations) ----
    GPODE_CUDA(cudaMemcpyAsync(b.Y, x0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    GPODE_CUDA(cudaMemcpyAsync(xs, x0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    ld_start_kernel<<<1, 1, 0, st>>>(ctrl, t, Tg, max_attempts);
    if (int rc = gpode_vf_large_eval(packed_large, c, b.Y, b.tmp, b.K[0], B, st)) return rc;
    ld_norm01_kernel<<<g, kLdThreads, 0, st>>>(b.Y, b.K[0], rt, at, n, partial);
    ld_init1_kernel<<<1, kLdThreads, 0, st>>>(ctrl, partial, (int)g, n_elem);
    ld_euler_kernel<<<g, kLdThreads, 0, st>>>(ctrl, b.Y, b.K[0], b.Ys, n);
    if (int rc = gpode_vf_large_eval(packed_large, c, b.Ys, b.tmp, b.K[1], B, st)) return rc;
    ld_norm2_kernel<<<g, kLdThreads, 0, st>>>(b.Y, b.K[0], b.K[1], rt, at, n, partial);
    ld_init2_kernel<<<1, kLdThreads, 0, st>>>(ctrl, partial, (int)g, n_elem, t, Tg);
    GPODE_LAUNCH_CHECK();

    // ---- the attempt loop as a WHILE node (condition: "not done", set by the controller kernel) ----
    std::lock_guard<std::mutex> lock(g_ld_mutex);
    ld_reap(false);
    if (g_ld_capture_stream == nullptr) GPODE_CUDA(cudaStreamCreateWithFlags(&g_ld_capture_stream, cudaStreamNonBlocking));
    cudaGraph_t graph;
    GPODE_CUDA(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle handle;
    // default value 1: the body's first kernels return at once when the start kernels already set `done`
    GPODE_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    GPODE_CUDA(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    cudaStream_t cs = g_ld_capture_stream;
    GPODE_CUDA(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    int rc = 0;
    for (int i = 1; i <= 6 && rc == 0; ++i) {
        float* out = i == 6 ? b.Y1 : b.Ys;
        ld_stage_kernel<<<g, kLdThreads, 0, cs>>>(ctrl, i, b, out, n);
        rc = gpode_vf_large_eval(packed_large, c, out, b.tmp, b.K[i], B, cs);
    }
    if (rc == 0) {
        ld_err_kernel<<<g, kLdThreads, 0, cs>>>(ctrl, b, rt, at, n, partial);
        ld_ctrl_kernel<<<1, kLdThreads, 0, cs>>>(ctrl, partial, (int)g, n_elem, t, Tg, handle);
        ld_accept_kernel<<<g, kLdThreads, 0, cs>>>(ctrl, b, t, xs, n);
    }
    cudaGraph_t captured = nullptr;
    cudaError_t ce = cudaStreamEndCapture(cs, &captured);
    if (rc != 0 || ce != cudaSuccess) {
        cudaGraphDestroy(graph);
        if (rc == 0) {
            gpode_set_error("capturing the dopri5 attempt failed: %s", cudaGetErrorString(ce));
            rc = (int)ce;
        }
        return rc;
    }
    cudaGraphExec_t exec;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) {
        cudaGraphDestroy(graph);
        gpode_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        return (int)ce;
    }
    GPODE_CUDA(cudaGraphLaunch(exec, st));
    ld_stats_kernel<<<1, 1, 0, st>>>(ctrl, stats_out);
    LdPending pend;
    pend.exec = exec;
    pend.graph = graph;
    GPODE_CUDA(cudaEventCreateWithFlags(&pend.ev, cudaEventDisableTiming));
    GPODE_CUDA(cudaEventRecord(pend.ev, st));
    g_ld_pending.push_back(pend);
    GPODE_LAUNCH_CHECK();
    return 0;
}
