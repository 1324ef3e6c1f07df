// One persistent kernel per eval-mode forward of the dense denoisers -- the latent MLP "U-Net" (latent_diffusion.py:107-128)
// and the length-1-sequence "DiT" (diffusion_transformer.py:81-109) -- at the reference batch sizes.
//
// At batch 128 these models are 0.36 / 0.69 GFLOP per forward: one launch per Linear / LayerNorm / BatchNorm1d made a reverse
// step a chain of ~45 latency-bound launches (229 / 329 us).  Here the host compiles the model into a TAPE of ops (device
// table); a grid of <= 128 co-resident CTAs walks the tape, every op split into (32 rows x 16 features) tiles over the grid,
// with a sense-reversing grid barrier only where an op reads what an earlier op of the same phase wrote.  Weights are read
// from the L2 once per row tile (4 times per forward at batch 128), activations stay in the L2 between ops, eval-mode
// BatchNorm1d + ReLU, bias, activation, residual and embedding gather are folded into the Linear epilogue.
#include <algorithm>

#include "common.cuh"

namespace td {

struct TapeOp {                  // mirrored by tinydiff/dense.py (ctypes)
    int kind;                    // 0 linear, 1 layernorm, 2 add (out (+)= x), 3 time features
    int N, K;
    int act;                     // TD_ACT_* (linear)
    int accumulate;              // add: out += x
    int tmode;                   // time features: 0 raw t, 1 t / 1000, 2 sinusoidal
    int barrier_before;          // grid barrier before this op (it depends on an op since the previous barrier)
    int bn_relu;
    const float* x; long long ldx;
    const float* w;              // linear: [N][K] row-major; layernorm: gamma
    const float* bias;           // linear: [N] or NULL; layernorm: beta
    float* out; long long ldo;
    const float* res; long long ldr;
    const long long* gidx; const float* gtab; long long ldt;
    const float* bn_mean; const float* bn_var; const float* bn_gamma; const float* bn_beta;   // fused eval BatchNorm1d
    float bn_eps, ln_eps;
    const long long* t; const int* t_dev;
};

constexpr int TP_THREADS = 256, TP_RT = 32, TP_NT = 16, TP_KMAX = 1024;
constexpr int TP_PAD = 8;
constexpr int TP_LN_CACHE = 8;            // LayerNorm rows of up to 256 features are held in registers
constexpr int TP_MAX_SMEM_OPS = 64;       // tapes up to this length are walked from shared memory
constexpr int TP_SMEM = (TP_RT + TP_NT) * (TP_KMAX + TP_PAD) * 4;

__device__ inline float tape_act(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return fmaxf(v, 0.f);
        case TD_ACT_SILU: return v / (1.f + expf(-v));
        case TD_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case TD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

// Grid barrier on one monotonically growing 64-bit arrival counter (never reset, reusable across launches).  The k-th barrier
// of a launch is passed when the counter reaches base + k * gridDim.x, base = the counter at kernel start rounded DOWN to a
// multiple of gridDim.x (a CTA that starts late may already see arrivals of the first barrier, but fewer than gridDim.x of
// them).  Arrival is a fire-and-forget reduction after a fence; only the wait is a round trip.
__device__ inline void grid_barrier(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
        unsigned long long seen;
        unsigned int spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
            if (++spins > (1u << 28)) { printf("tinydiff: dense tape grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
        } while (seen < target);
    }
    __syncthreads();
}

__device__ inline void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ inline void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stage `rows` rows of `K` floats (row r at src + r*ld, rows >= valid are zero) into smem rows of stride K + TP_PAD
__device__ inline void stage_rows(float* dst, const float* __restrict__ src, long long ld, int rows, int valid, int K, bool vec) {
    const int stride = K + TP_PAD;
    if (vec) {
        const int k4 = K >> 2;
        for (int e = threadIdx.x; e < rows * k4; e += TP_THREADS) {
            const int r = e / k4, q = e - r * k4;
            if (r < valid) cp_async16(dst + r * stride + q * 4, src + (long long)r * ld + q * 4);
            else *reinterpret_cast<float4*>(dst + r * stride + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int e = threadIdx.x; e < rows * K; e += TP_THREADS) {
            const int r = e / K, k = e - r * K;
            dst[r * stride + k] = r < valid ? __ldcg(src + (long long)r * ld + k) : 0.f;
        }
    }
}

__global__ void __launch_bounds__(TP_THREADS, 1)
dense_tape_kernel(const TapeOp* __restrict__ ops, int n_ops, int M, unsigned long long* bar) {
    extern __shared__ __align__(16) float tp_smem[];
    float* const xs = tp_smem;                                   // [TP_RT][K + TP_PAD]
    // The tape itself is host-written and constant: copy it to shared memory BEFORE the dependency wait (it overlaps the tail of the
    // preceding kernel) so that fetching op i is not one more dependent global round trip in front of every op.
    __shared__ __align__(16) unsigned char s_ops_raw[TP_MAX_SMEM_OPS * sizeof(TapeOp)];
    const int tid = threadIdx.x;
    const bool ops_in_smem = n_ops <= TP_MAX_SMEM_OPS;
    if (ops_in_smem) {
        static_assert(sizeof(TapeOp) % 16 == 0, "TapeOp is copied as 16-byte vectors");
        const int4* src = reinterpret_cast<const int4*>(ops);
        int4* dst = reinterpret_cast<int4*>(s_ops_raw);
        for (int i = tid; i < n_ops * (int)(sizeof(TapeOp) / 16); i += TP_THREADS) dst[i] = __ldg(src + i);
    }
    td::pdl_sync();
    __syncthreads();
    unsigned long long bar_target;
    {
        unsigned long long c;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(bar) : "memory");
        bar_target = c / gridDim.x * gridDim.x;
    }
    for (int oi = 0; oi < n_ops; ++oi) {
        const TapeOp op = ops_in_smem ? reinterpret_cast<const TapeOp*>(s_ops_raw)[oi] : ops[oi];
        if (op.kind == 0) {
            // ---- Linear: tiles of 32 rows x 16 features over the grid; the whole K extent of a tile's operands is staged in
            // shared memory with asynchronous copies (one memory latency per op); the WEIGHT slice does not depend on earlier
            // ops and is requested BEFORE the grid barrier.  thread = (row, features c and c + 8).
            const int row_tiles = (M + TP_RT - 1) / TP_RT, col_tiles = (op.N + TP_NT - 1) / TP_NT;
            const int n_tiles = row_tiles * col_tiles;
            const int stride = op.K + TP_PAD;
            float* const ws = xs + TP_RT * stride;               // [TP_NT][K + TP_PAD]
            const bool vecw = (op.K & 3) == 0 && ((uintptr_t)op.w & 15) == 0;
            const bool vecx = (op.K & 3) == 0 && ((uintptr_t)op.x & 15) == 0 && (op.ldx & 3) == 0;
            const int r = tid >> 3, c = tid & 7;
            __syncthreads();                                     // the previous op's tile is no longer read
            bool w_staged = false;
            if ((int)blockIdx.x < n_tiles && vecw) {
                const int n0 = ((int)blockIdx.x / row_tiles) * TP_NT;
                stage_rows(ws, op.w + (long long)n0 * op.K, op.K, TP_NT, min(TP_NT, op.N - n0), op.K, true);
                w_staged = true;
            }
            if (op.barrier_before) { bar_target += gridDim.x; grid_barrier(bar, bar_target); }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int rt = tile % row_tiles, ct = tile / row_tiles;
                const int m0 = rt * TP_RT, n0 = ct * TP_NT;
                if (!w_staged) {
                    __syncthreads();
                    stage_rows(ws, op.w + (long long)n0 * op.K, op.K, TP_NT, min(TP_NT, op.N - n0), op.K, vecw);
                }
                w_staged = false;
                stage_rows(xs, op.x + (long long)m0 * op.ldx, op.ldx, TP_RT, min(TP_RT, M - m0), op.K, vecx);
                // the epilogue's operands (bias, folded BatchNorm, residual, gathered row) are requested now, beside the operand
                // copies, instead of as a dependent round trip after the reduction
                float e_bias[2] = {0.f, 0.f}, e_scale[2] = {1.f, 1.f}, e_shift[2] = {0.f, 0.f}, e_add[2] = {0.f, 0.f};
                float e_gamma[2] = {1.f, 1.f}, e_beta[2] = {0.f, 0.f}, e_gat[2] = {0.f, 0.f};
                {
                    const int m = m0 + r;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int n = n0 + c + 8 * j;
                        if (m >= M || n >= op.N) continue;
                        if (op.bias) e_bias[j] = __ldg(op.bias + n);
                        if (op.bn_mean) {         // eval-mode BatchNorm1d: (x - mean) * invstd * gamma + beta
                            const float invstd = 1.f / sqrtf(__ldg(op.bn_var + n) + op.bn_eps);
                            e_scale[j] = invstd;
                            e_shift[j] = __ldg(op.bn_mean + n);
                            e_gamma[j] = __ldg(op.bn_gamma + n);
                            e_beta[j] = __ldg(op.bn_beta + n);
                        }
                        if (op.res) e_add[j] = __ldcg(op.res + (long long)m * op.ldr + n);
                        if (op.gidx) e_gat[j] = __ldg(op.gtab + (long long)op.gidx[m] * op.ldt + n);
                    }
                }
                cp_async_wait_all();
                __syncthreads();
                // 16 k-slices x (4 row groups x 4 column groups): a thread owns an 8 x 4 register tile over its slice of K (k = 4*ks +
                // 64*j), so one 16-byte shared-memory read feeds 8-16 FMAs (a thread per output pair re-read both operands for
                // every FMA pair and the loop was shared-memory-bandwidth bound: 12 k cycles per 512-deep tile).  Rows tr + 4i and
                // columns tc + 4j are interleaved and the row stride is K + 8 floats: conflict-free for the two slices of a warp.
                const int ks = tid >> 4, tr = (tid >> 2) & 3, tc = tid & 3;
                float acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                if ((op.K & 3) == 0) {
                    for (int k = ks * 4; k < op.K; k += 64) {
                        float4 wv[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(ws + (tc + 4 * j) * stride + k);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 xv = *reinterpret_cast<const float4*>(xs + (tr + 4 * i) * stride + k);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[i][j] = fmaf(xv.x, wv[j].x, acc[i][j]);
                                acc[i][j] = fmaf(xv.y, wv[j].y, acc[i][j]);
                                acc[i][j] = fmaf(xv.z, wv[j].z, acc[i][j]);
                                acc[i][j] = fmaf(xv.w, wv[j].w, acc[i][j]);
                            }
                        }
                    }
                } else {
                    for (int k = ks; k < op.K; k += 16) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float xv = xs[(tr + 4 * i) * stride + k];
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv, ws[(tc + 4 * j) * stride + k], acc[i][j]);
                        }
                    }
                }
                // fixed-order sum of the 16 k-slices through shared memory (the operand tiles are dead by now)
                __syncthreads();
                float* red = xs;                                   // [16 slices][512 outputs]
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) red[ks * 512 + (tr + 4 * i) * TP_NT + tc + 4 * j] = acc[i][j];
                __syncthreads();
                const int m = m0 + r;
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int cl = c + 8 * j, n = n0 + cl;
                        if (n >= op.N) continue;
                        float v = 0.f;
#pragma unroll
                        for (int q = 0; q < 16; ++q) v += red[q * 512 + r * TP_NT + cl];
                        v += e_bias[j];
                        if (op.bn_mean) {         // same operation order as before: (v - mean) * invstd * gamma + beta
                            v = (v - e_shift[j]) * e_scale[j] * e_gamma[j] + e_beta[j];
                            if (op.bn_relu) v = fmaxf(v, 0.f);
                        }
                        v = tape_act(v, op.act);
                        if (op.res) v += e_add[j];
                        if (op.gidx) v += e_gat[j];
                        op.out[(long long)m * op.ldo + n] = v;
                    }
                }
            }
            continue;
        }
        if (op.barrier_before) { bar_target += gridDim.x; grid_barrier(bar, bar_target); }
        if (op.kind == 1) {
            // ---- LayerNorm over N features: one warp per row, two-pass (mean, then centred variance) like ATen
            const int warp = tid >> 5, lane = tid & 31;
            for (int m = blockIdx.x * (TP_THREADS / 32) + warp; m < M; m += gridDim.x * (TP_THREADS / 32)) {
                const float* xr = op.x + (long long)m * op.ldx;
                if (op.N <= 32 * TP_LN_CACHE) {
                    // the row lives in registers: one L2 round trip instead of three (same lane-strided summation order)
                    float xv[TP_LN_CACHE], gv[TP_LN_CACHE], bv[TP_LN_CACHE];
#pragma unroll
                    for (int i = 0; i < TP_LN_CACHE; ++i) {
                        const int n = lane + 32 * i;
                        const bool ok = n < op.N;
                        xv[i] = ok ? __ldcg(xr + n) : 0.f;
                        gv[i] = ok ? __ldg(op.w + n) : 0.f;
                        bv[i] = ok ? __ldg(op.bias + n) : 0.f;
                    }
                    float s1 = 0.f;
#pragma unroll
                    for (int i = 0; i < TP_LN_CACHE; ++i) if (lane + 32 * i < op.N) s1 += xv[i];
                    const float mean1 = warp_sum(s1) / (float)op.N;
                    float q1 = 0.f;
#pragma unroll
                    for (int i = 0; i < TP_LN_CACHE; ++i)
                        if (lane + 32 * i < op.N) { const float d = xv[i] - mean1; q1 = fmaf(d, d, q1); }
                    const float rstd1 = 1.f / sqrtf(warp_sum(q1) / (float)op.N + op.ln_eps);
#pragma unroll
                    for (int i = 0; i < TP_LN_CACHE; ++i) {
                        const int n = lane + 32 * i;
                        if (n < op.N) op.out[(long long)m * op.ldo + n] = (xv[i] - mean1) * rstd1 * gv[i] + bv[i];
                    }
                    continue;
                }
                float s = 0.f;
                for (int n = lane; n < op.N; n += 32) s += __ldcg(xr + n);
                const float mean = warp_sum(s) / (float)op.N;
                float q = 0.f;
                for (int n = lane; n < op.N; n += 32) { const float d = __ldcg(xr + n) - mean; q = fmaf(d, d, q); }
                const float rstd = 1.f / sqrtf(warp_sum(q) / (float)op.N + op.ln_eps);
                for (int n = lane; n < op.N; n += 32)
                    op.out[(long long)m * op.ldo + n] = (__ldcg(xr + n) - mean) * rstd * __ldg(op.w + n) + __ldg(op.bias + n);
            }
        } else if (op.kind == 2) {
            const long long total = (long long)M * op.N;
            for (long long e = (long long)blockIdx.x * TP_THREADS + tid; e < total; e += (long long)gridDim.x * TP_THREADS) {
                const long long m = e / op.N, n = e - m * op.N;
                const float v = __ldcg(op.x + m * op.ldx + n);
                float* o = op.out + m * op.ldo + n;
                *o = op.accumulate ? __ldcg(o) + v : v;
            }
        } else {
            const int width = op.tmode == 2 ? op.N : 1;
            const long long total = (long long)M * width;
            for (long long e = (long long)blockIdx.x * TP_THREADS + tid; e < total; e += (long long)gridDim.x * TP_THREADS) {
                const int b = (int)(e / width), j = (int)(e % width);
                float tv = op.t ? (float)op.t[b] : (float)op.t_dev[0];
                if (op.tmode == 1) tv = tv / 1000.0f;
                float v = tv;
                if (op.tmode == 2) {
                    const int half = op.N / 2;
                    v = 0.f;
                    if (j < 2 * half) {
                        const int jj = (j < half) ? j : j - half;
                        const float arg = tv * expf(-logf(10000.0f) * (float)jj / (float)(half - 1));
                        v = (j < half) ? sinf(arg) : cosf(arg);
                    }
                }
                op.out[(long long)b * op.ldo + j] = v;
            }
        }
    }
}

}  // namespace td

using namespace td;

extern "C" int td_dense_tape_op_bytes(void) { return (int)sizeof(TapeOp); }

extern "C" int td_dense_tape_run(const void* ops, int n_ops, int batch, unsigned long long* barrier, int max_ctas, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(ops && n_ops > 0 && batch > 0 && barrier, "td_dense_tape_run: bad args");
    static bool configured = false;
    if (!configured) {
        TD_CUDA(cudaFuncSetAttribute(dense_tape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
        configured = true;
    }
    // the grid barrier needs every CTA resident at once: never more CTAs than SMs (one CTA per SM).  The grid size is part of the
    // barrier arithmetic: keep it the same for every launch that shares `barrier`.
    int grid = std::min(kNumSMs - 20, 128);
    if (max_ctas > 0) grid = std::min(grid, max_ctas);
    TD_CHECK_ARG((((uintptr_t)barrier) & 7) == 0, "td_dense_tape_run: barrier must be 8-byte aligned");
    td::launch(dense_tape_kernel, td::LaunchCfg(grid, TP_THREADS, TP_SMEM, (cudaStream_t)stream), reinterpret_cast<const TapeOp*>(ops), n_ops,
               batch, barrier);
    return launch_status("dense_tape");
}
