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

constexpr int TP_THREADS = 256, TP_RT = 32, TP_NT = 16, TP_KC = 64;

__device__ inline float tape_act(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return fmaxf(v, 0.f);
        case TD_ACT_SILU: return v / (1.f + expf(-v));
        case TD_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case TD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

// sense-reversing grid barrier (reusable across launches: `arrive` returns to 0, `gen` only grows)
__device__ inline void grid_barrier(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* gen = bar + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            unsigned int spins = 0;
            while (*gen == g) {
                if (++spins > (1u << 28)) { printf("tinydiff: dense tape grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TP_THREADS)
dense_tape_kernel(const TapeOp* __restrict__ ops, int n_ops, int M, unsigned int* bar) {
    td::pdl_sync();
    __shared__ float xs[TP_RT][TP_KC + 1];
    __shared__ float ws[TP_NT][TP_KC + 1];
    const int tid = threadIdx.x;
    for (int oi = 0; oi < n_ops; ++oi) {
        const TapeOp op = ops[oi];
        if (op.barrier_before) grid_barrier(bar);
        if (op.kind == 0) {
            // ---- Linear: tiles of 32 rows x 16 features; thread = (row, feature pair)
            const int row_tiles = (M + TP_RT - 1) / TP_RT, col_tiles = (op.N + TP_NT - 1) / TP_NT;
            const int r = tid >> 3, cp = (tid & 7) * 2;
            for (int tile = blockIdx.x; tile < row_tiles * col_tiles; tile += gridDim.x) {
                const int rt = tile % row_tiles, ct = tile / row_tiles;
                const int m0 = rt * TP_RT, n0 = ct * TP_NT;
                float a0 = 0.f, a1 = 0.f;
                for (int k0 = 0; k0 < op.K; k0 += TP_KC) {
                    __syncthreads();
                    for (int e = tid; e < TP_RT * TP_KC; e += TP_THREADS) {
                        const int rr = e / TP_KC, kk = e - rr * TP_KC;
                        const int m = m0 + rr, k = k0 + kk;
                        xs[rr][kk] = (m < M && k < op.K) ? __ldcg(op.x + (long long)m * op.ldx + k) : 0.f;
                    }
                    for (int e = tid; e < TP_NT * TP_KC; e += TP_THREADS) {
                        const int nn = e / TP_KC, kk = e - nn * TP_KC;
                        const int n = n0 + nn, k = k0 + kk;
                        ws[nn][kk] = (n < op.N && k < op.K) ? __ldg(op.w + (long long)n * op.K + k) : 0.f;
                    }
                    __syncthreads();
#pragma unroll 16
                    for (int kk = 0; kk < TP_KC; ++kk) {
                        const float xv = xs[r][kk];
                        a0 = fmaf(xv, ws[cp][kk], a0);
                        a1 = fmaf(xv, ws[cp + 1][kk], a1);
                    }
                }
                const int m = m0 + r;
                if (m < M) {
                    float acc[2] = {a0, a1};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int n = n0 + cp + j;
                        if (n >= op.N) continue;
                        float v = acc[j] + (op.bias ? __ldg(op.bias + n) : 0.f);
                        if (op.bn_mean) {         // eval-mode BatchNorm1d (+ ReLU): (x - mean) * invstd * gamma + beta
                            const float invstd = 1.f / sqrtf(__ldg(op.bn_var + n) + op.bn_eps);
                            v = (v - __ldg(op.bn_mean + n)) * invstd * __ldg(op.bn_gamma + n) + __ldg(op.bn_beta + n);
                            if (op.bn_relu) v = fmaxf(v, 0.f);
                        }
                        v = tape_act(v, op.act);
                        if (op.res) v += __ldcg(op.res + (long long)m * op.ldr + n);
                        if (op.gidx) v += __ldg(op.gtab + (long long)op.gidx[m] * op.ldt + n);
                        op.out[(long long)m * op.ldo + n] = v;
                    }
                }
            }
        } else if (op.kind == 1) {
            // ---- LayerNorm over N features: one warp per row, two-pass (mean, then centred variance) like ATen
            const int warp = tid >> 5, lane = tid & 31;
            for (int m = blockIdx.x * (TP_THREADS / 32) + warp; m < M; m += gridDim.x * (TP_THREADS / 32)) {
                const float* xr = op.x + (long long)m * op.ldx;
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

extern "C" int td_dense_tape_run(const void* ops, int n_ops, int batch, unsigned int* barrier, int max_ctas, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(ops && n_ops > 0 && batch > 0 && barrier, "td_dense_tape_run: bad args");
    // the grid barrier needs every CTA resident at once: never more CTAs than SMs (one small CTA per SM)
    int grid = std::min(kNumSMs - 20, 128);
    if (max_ctas > 0) grid = std::min(grid, max_ctas);
    td::launch(dense_tape_kernel, td::LaunchCfg(grid, TP_THREADS, 0, (cudaStream_t)stream), reinterpret_cast<const TapeOp*>(ops), n_ops,
               batch, barrier);
    return launch_status("dense_tape");
}
