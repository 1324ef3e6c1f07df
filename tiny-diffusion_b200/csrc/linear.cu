// fp32 dense layers of the conditioning head, the latent-MLP denoiser and the DiT parity path:
// a strided SIMT GEMM with a fused epilogue (bias, activation, residual, embedding-row gather),
// column sums (bias gradients), activation backward, LayerNorm and BatchNorm1d.
// These are small (<= 0.1 GFLOP at the reference batch sizes) and latency-bound; the tensor-core
// (kind::tf32) path for batches of 2048 and more lives in linear_tc.cu.
#include <algorithm>

#include "common.cuh"

namespace td {

bool gemm_tf32_try(const td_gemm_args* a, cudaStream_t s, int* status, int* nz_out);   // linear_tc.cu
int gemm_tf32_splits(int M, int N, int K);                                             // linear_tc.cu

constexpr int G_BM = 64, G_BN = 64, G_BK = 16, G_THREADS = 256;

__device__ inline float act_apply(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return fmaxf(v, 0.f);
        case TD_ACT_SILU: return v / (1.0f + expf(-v));
        case TD_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        case TD_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
        default: return v;
    }
}
// d act(v) / dv
__device__ inline float act_grad(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return v > 0.f ? 1.f : 0.f;
        case TD_ACT_SILU: {
            const float s = 1.0f / (1.0f + expf(-v));
            return s * (1.0f + v * (1.0f - s));
        }
        case TD_ACT_GELU: {
            const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
            return cdf + v * pdf;
        }
        case TD_ACT_SIGMOID: {
            const float sg = 1.0f / (1.0f + expf(-v));
            return sg * (1.0f - sg);
        }
        default: return 1.f;
    }
}

__global__ void __launch_bounds__(G_THREADS)
gemm_f32_kernel(const td_gemm_args g) {
    td::pdl_sync();
    __shared__ float As[2][G_BK][G_BM + 4];
    __shared__ float Bs[2][G_BK][G_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
    const int ty = tid >> 4, tx = tid & 15;
    // split-K over blockIdx.z (only used with the deterministic two-pass reduction below)
    const int kz = blockIdx.z, nz = gridDim.z;
    const int kchunk = (int)ceil_div(ceil_div(g.K, G_BK), nz) * G_BK;
    const int kbeg = kz * kchunk, kend = min(g.K, kbeg + kchunk);

    const bool a_kc = (g.a_cs == 1), b_kc = (g.b_rs == 1);
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int e = tid + l * G_THREADS;
            int i, k;
            if (a_kc) { k = e & 15; i = e >> 4; } else { i = e & 63; k = e >> 6; }
            const int gi = m0 + i, gk = k0 + k;
            ra[l] = (gi < g.M && gk < kend) ? __ldg(g.A + (int64_t)gi * g.a_rs + (int64_t)gk * g.a_cs) : 0.f;
            int j;
            if (b_kc) { k = e & 15; j = e >> 4; } else { j = e & 63; k = e >> 6; }
            const int gj = n0 + j;
            const int gk2 = k0 + k;
            rb[l] = (gj < g.N && gk2 < kend) ? __ldg(g.B + (int64_t)gk2 * g.b_rs + (int64_t)gj * g.b_cs) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int e = tid + l * G_THREADS;
            int i, k;
            if (a_kc) { k = e & 15; i = e >> 4; } else { i = e & 63; k = e >> 6; }
            As[buf][k][i] = ra[l];
            int j;
            if (b_kc) { k = e & 15; j = e >> 4; } else { j = e & 63; k = e >> 6; }
            Bs[buf][k][j] = rb[l];
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int buf = 0;
    if (kbeg < kend) {
        fetch(kbeg);
        stash(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += G_BK) {
        const bool more = (k0 + G_BK) < kend;
        if (more) fetch(k0 + G_BK);
#pragma unroll
        for (int k = 0; k < G_BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    if (nz > 1) {       // split-K partial: raw accumulators to the workspace slice kz
        float* ws = g.splitk_ws + (int64_t)kz * g.M * g.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gi = m0 + ty * 4 + i;
            if (gi >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gj = n0 + tx * 4 + j;
                if (gj < g.N) ws[(int64_t)gi * g.N + gj] = acc[i][j];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = m0 + ty * 4 + i;
        if (gi >= g.M) continue;
        const float* grow = (g.gather_idx && g.gather_table) ? g.gather_table + g.gather_idx[gi] * g.ld_table : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = n0 + tx * 4 + j;
            if (gj >= g.N) continue;
            float v = acc[i][j] * g.alpha;
            if (g.bias) v += __ldg(g.bias + gj);
            if (g.pre_out) g.pre_out[(int64_t)gi * g.ld_pre + gj] = v;
            v = act_apply(v, g.act);
            if (g.residual) v += g.residual[(int64_t)gi * g.ldr + gj];
            if (grow) v += grow[gj];
            float* c = g.C + (int64_t)gi * g.ldc + gj;
            if (g.accumulate) v += *c;
            *c = v;
        }
    }
}

// epilogue shared by the GEMM variants
__device__ inline void gemm_store(const td_gemm_args& g, int gi, int gj, float acc) {
    float v = acc * g.alpha;
    if (g.bias) v += __ldg(g.bias + gj);
    if (g.pre_out) g.pre_out[(int64_t)gi * g.ld_pre + gj] = v;
    v = act_apply(v, g.act);
    if (g.residual) v += g.residual[(int64_t)gi * g.ldr + gj];
    if (g.gather_idx && g.gather_table) v += g.gather_table[g.gather_idx[gi] * g.ld_table + gj];
    float* c = g.C + (int64_t)gi * g.ldc + gj;
    if (g.accumulate) v += *c;
    *c = v;
}

// Small-output variant: 32x32 tiles, BK = 32, 2x2 outputs per thread.  Used when the 64x64 grid would
// leave most SMs idle (the conditioning head and the latent denoisers: M = batch = 128, N, K <= 1024):
// 4x the CTAs and half the k-steps of the big-tile kernel.
constexpr int S_BM = 32, S_BN = 32, S_BK = 32;
__global__ void __launch_bounds__(256)
gemm_f32_small_kernel(const td_gemm_args g) {
    td::pdl_sync();
    __shared__ float As[S_BK][S_BM + 1];
    __shared__ float Bs[S_BK][S_BN + 1];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * S_BM, n0 = blockIdx.x * S_BN;
    const int ty = tid >> 4, tx = tid & 15;
    const bool a_kc = (g.a_cs == 1), b_kc = (g.b_rs == 1);
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < g.K; k0 += S_BK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int e = tid + l * 256;
            int i, k;
            if (a_kc) { k = e & 31; i = e >> 5; } else { i = e & 31; k = e >> 5; }
            int gi = m0 + i, gk = k0 + k;
            As[k][i] = (gi < g.M && gk < g.K) ? __ldg(g.A + (int64_t)gi * g.a_rs + (int64_t)gk * g.a_cs) : 0.f;
            int j;
            if (b_kc) { k = e & 31; j = e >> 5; } else { j = e & 31; k = e >> 5; }
            const int gj = n0 + j;
            gk = k0 + k;
            Bs[k][j] = (gj < g.N && gk < g.K) ? __ldg(g.B + (int64_t)gk * g.b_rs + (int64_t)gj * g.b_cs) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < S_BK; ++k) {
            const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
            const float b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = m0 + ty * 2 + i, gj = n0 + tx * 2 + j;
            if (gi < g.M && gj < g.N) gemm_store(g, gi, gj, acc[i][j]);
        }
}

// Row-vector variant (M <= 4, both operands K-contiguous): one warp per output column, lanes stride K.
// The sampler's per-step conditioning head is this shape (every sample shares t).
__global__ void __launch_bounds__(256)
gemv_f32_kernel(const td_gemm_args g) {
    td::pdl_sync();
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= g.N) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* w = g.B + (int64_t)j * g.b_cs;
    for (int k = lane; k < g.K; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < g.M) acc[i] = fmaf(__ldg(g.A + (int64_t)i * g.a_rs + k), wv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0)
        for (int i = 0; i < g.M; ++i) gemm_store(g, i, j, acc[i]);
}

// second pass of the deterministic split-K: fixed-order sum over slices + the same epilogue
__global__ void __launch_bounds__(256)
gemm_splitk_reduce_kernel(const td_gemm_args g, int nz) {
    td::pdl_sync();
    const int64_t total = (int64_t)g.M * g.N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int gi = (int)(e / g.N), gj = (int)(e % g.N);
        float v = 0.f;
        for (int z = 0; z < nz; ++z) v += g.splitk_ws[(int64_t)z * total + e];
        v *= g.alpha;
        if (g.bias) v += __ldg(g.bias + gj);
        if (g.pre_out) g.pre_out[(int64_t)gi * g.ld_pre + gj] = v;
        v = act_apply(v, g.act);
        if (g.residual) v += g.residual[(int64_t)gi * g.ldr + gj];
        if (g.gather_idx && g.gather_table) v += g.gather_table[g.gather_idx[gi] * g.ld_table + gj];
        float* c = g.C + (int64_t)gi * g.ldc + gj;
        if (g.accumulate) v += *c;
        *c = v;
    }
}

// out[j] (+)= sum_i x[i, j]      (bias gradients); one warp-row per 32 columns, fixed order
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int M, int N, int accumulate) {
    td::pdl_sync();
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (j < N)
        for (int i = w; i < M; i += 8) s += x[(int64_t)i * ldx + j];
    part[w][lane] = s;
    __syncthreads();
    if (w == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][lane];
        out[j] = accumulate ? out[j] + t : t;
    }
}

// dx = dy * act'(pre)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx, int64_t n, int act) {
    td::pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * act_grad(pre[i], act);
}

// table_grad[c, :] (+)= sum_{i : idx[i] == c} g[i, :]    (nn.Embedding backward, deterministic)
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ idx, float* __restrict__ out,
                     int M, int D, int accumulate) {
    td::pdl_sync();
    const int c = blockIdx.x;
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < M; ++i)
            if (idx[i] == c) s += g[(int64_t)i * ldg + j];
        float* o = out + (int64_t)c * D + j;
        *o = accumulate ? *o + s : s;
    }
}

// Large batches: the kernel above walks all M rows with one CTA per table row (10 CTAs, or ONE for the single-row positional
// table).  Here grid.x row chunks accumulate table-shaped partial sums in shared memory -- thread j owns column j, so the order
// inside a chunk is the row order and there are no conflicts --, written to partials[chunk][num_rows][D]; the second kernel sums
// the chunks in fixed order.  Deterministic.
__global__ void __launch_bounds__(256)
embedding_bwd_partial_kernel(const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ idx, float* __restrict__ partials,
                             int M, int D, int num_rows) {
    td::pdl_sync();
    extern __shared__ float acc[];                   // [num_rows][D]
    for (int e = threadIdx.x; e < num_rows * D; e += blockDim.x) acc[e] = 0.f;
    __syncthreads();
    const int i0 = (int)((int64_t)M * blockIdx.x / gridDim.x), i1 = (int)((int64_t)M * (blockIdx.x + 1) / gridDim.x);
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const int c = (int)idx[i];
            if (c >= 0 && c < num_rows) acc[c * D + j] += g[(int64_t)i * ldg + j];
        }
    }
    __syncthreads();
    float* out = partials + (int64_t)blockIdx.x * num_rows * D;
    for (int e = threadIdx.x; e < num_rows * D; e += blockDim.x) out[e] = acc[e];
}
__global__ void __launch_bounds__(256)
embedding_bwd_finalize_kernel(const float* __restrict__ partials, int chunks, int n, float* __restrict__ out, int accumulate) {
    td::pdl_sync();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= chunks; k += 4) {
        s0 += partials[(int64_t)k * n + e]; s1 += partials[(int64_t)(k + 1) * n + e];
        s2 += partials[(int64_t)(k + 2) * n + e]; s3 += partials[(int64_t)(k + 3) * n + e];
    }
    for (; k < chunks; ++k) s0 += partials[(int64_t)k * n + e];
    const float s = (s0 + s1) + (s2 + s3);
    out[e] = accumulate ? out[e] + s : s;
}

// First layer input of the conditioning head: raw t, t/1000, or the sinusoidal embedding
// (conditional_diffusion_laion.py:223-232: [sin | cos], divisor half-1).
__global__ void __launch_bounds__(256)
time_features_kernel(const int64_t* __restrict__ t, const int32_t* __restrict__ t_dev, float* __restrict__ out, int B,
                     int D, int mode) {
    td::pdl_sync();
    const int width = (mode == 2) ? D : 1;
    const int64_t total = (int64_t)B * width;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / width), j = (int)(e % width);
        float tv = t ? (float)t[b] : (float)t_dev[0];
        if (mode == 1) tv = tv / 1000.0f;
        if (mode == 2) {
            const int half = D / 2;
            float v = 0.f;
            if (j < 2 * half) {
                const int jj = (j < half) ? j : j - half;
                const float fr = expf(-logf(10000.0f) * (float)jj / (float)(half - 1));
                const float arg = tv * fr;
                v = (j < half) ? sinf(arg) : cosf(arg);
            }
            out[e] = v;
        } else {
            out[e] = tv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dimension (diffusion_transformer.py:22,33,69): one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
                     float eps) {
    td::pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s += xr[j];
    const float mean = warp_sum(s) / (float)D;
    float v = 0.f;
    for (int j = lane; j < D; j += 32) { const float d = xr[j] - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / (float)D + eps);
    for (int j = lane; j < D; j += 32) y[(int64_t)row * D + j] = (xr[j] - mean) * rstd * gamma[j] + beta[j];
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ; xhat recomputed from x
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx, int M, int D) {
    td::pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + (int64_t)row * D;
    const float* gr = dy + (int64_t)row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < D; j += 32) {
        const float g = gr[j] * gamma[j], xh = (xr[j] - mu) * rs;
        s1 += g; s2 += g * xh;
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    for (int j = lane; j < D; j += 32) {
        const float g = gr[j] * gamma[j], xh = (xr[j] - mu) * rs;
        dx[(int64_t)row * D + j] = rs * (g - s1 - xh * s2);
    }
}

// dgamma[j] = sum_i dy[i,j]*xhat[i,j], dbeta[j] = sum_i dy[i,j]  (fixed order: 8 row-slices then sum)
__global__ void __launch_bounds__(256)
layernorm_bwd_params_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                            const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                            int D) {
    td::pdl_sync();
    __shared__ float pg[8][33], pb[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float sg = 0.f, sb = 0.f;
    if (j < D)
        for (int i = w; i < M; i += 8) {
            const float g = dy[(int64_t)i * D + j];
            sg += g * (x[(int64_t)i * D + j] - mean[i]) * rstd[i];
            sb += g;
        }
    pg[w][lane] = sg; pb[w][lane] = sb;
    __syncthreads();
    if (w == 0 && j < D) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += pg[k][lane]; b += pb[k][lane]; }
        dgamma[j] = a; dbeta[j] = b;
    }
}

// Large batches: grid (D / 32, row chunks) partial sums in the [chunk][2][D] layout of the BatchNorm partials (td_partial_sum
// sums the chunks in fixed order): the kernel above is one CTA per 32 columns walking all M rows (8 CTAs for D = 256).
__global__ void __launch_bounds__(256)
layernorm_bwd_params_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, float* __restrict__ partials, int M, int D) {
    td::pdl_sync();
    __shared__ float pg[8][33], pb[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int i0 = (int)((int64_t)M * blockIdx.y / gridDim.y), i1 = (int)((int64_t)M * (blockIdx.y + 1) / gridDim.y);
    float sg = 0.f, sb = 0.f;
    if (j < D) {
#pragma unroll 4
        for (int i = i0 + w; i < i1; i += 8) {
            const float g = dy[(int64_t)i * D + j];
            sg += g * (x[(int64_t)i * D + j] - mean[i]) * rstd[i];
            sb += g;
        }
    }
    pg[w][lane] = sg; pb[w][lane] = sb;
    __syncthreads();
    if (w == 0 && j < D) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += pg[k][lane]; b += pb[k][lane]; }
        partials[((int64_t)blockIdx.y * 2 + 0) * D + j] = a;
        partials[((int64_t)blockIdx.y * 2 + 1) * D + j] = b;
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d (+ReLU) over a [M, N] matrix (latent_diffusion.py:35-36 ...): one CTA per 32
// columns; train mode computes batch statistics (two-pass, fixed order) and updates the running
// statistics exactly like torch (momentum, unbiased running variance).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn1d_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                float* __restrict__ save_rstd, float* __restrict__ y, int64_t ldy, int M, int N, float eps,
                float momentum, int training, int relu) {
    td::pdl_sync();
    __shared__ float part[8][33];
    __shared__ float s_mean[32], s_rstd[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    if (training) {
        float s = 0.f;
        if (j < N) for (int i = w; i < M; i += 8) s += x[(int64_t)i * ldx + j];
        part[w][lane] = s;
        __syncthreads();
        if (w == 0) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[k][lane];
            s_mean[lane] = t / (float)M;
        }
        __syncthreads();
        const float mu = s_mean[lane];
        float v = 0.f;
        if (j < N) for (int i = w; i < M; i += 8) { const float d = x[(int64_t)i * ldx + j] - mu; v += d * d; }
        __syncthreads();
        part[w][lane] = v;
        __syncthreads();
        if (w == 0) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[k][lane];
            const float var = t / (float)M;
            s_rstd[lane] = rsqrtf(var + eps);
            if (j < N) {
                if (save_mean) save_mean[j] = mu;
                if (save_rstd) save_rstd[j] = s_rstd[lane];
                if (running_mean) running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * mu;
                if (running_var) {
                    const float unb = (M > 1) ? t / (float)(M - 1) : var;
                    running_var[j] = (1.f - momentum) * running_var[j] + momentum * unb;
                }
            }
        }
        __syncthreads();
    } else {
        if (w == 0 && j < N) {
            s_mean[lane] = running_mean[j];
            s_rstd[lane] = rsqrtf(running_var[j] + eps);
        }
        __syncthreads();
    }
    if (j < N) {
        const float mu = s_mean[lane], rs = s_rstd[lane], ga = gamma[j], be = beta[j];
        // eval mode is purely elementwise: gridDim.y splits the rows (a 65536-row batch on N / 32 CTAs used 2-16 SMs)
        const int rows_per = (M + (int)gridDim.y - 1) / (int)gridDim.y;
        const int i_lo = blockIdx.y * rows_per, i_hi = min(M, i_lo + rows_per);
        for (int i = i_lo + w; i < i_hi; i += 8) {
            float v = (x[(int64_t)i * ldx + j] - mu) * rs * ga + be;
            if (relu) v = fmaxf(v, 0.f);
            y[(int64_t)i * ldy + j] = v;
        }
    }
}

// backward of y = relu?(bn(x)) in train mode; y_out is the saved forward output (ReLU mask)
__global__ void __launch_bounds__(256)
bn1d_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                const float* __restrict__ y_out, int64_t ldy, const float* __restrict__ gamma,
                const float* __restrict__ save_mean, const float* __restrict__ save_rstd, float* __restrict__ dx,
                int64_t lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int N, int relu) {
    td::pdl_sync();
    __shared__ float p1[8][33], p2[8][33];
    __shared__ float s1[32], s2[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const float mu = (j < N) ? save_mean[j] : 0.f, rs = (j < N) ? save_rstd[j] : 0.f;
    float a = 0.f, b = 0.f;
    if (j < N)
        for (int i = w; i < M; i += 8) {
            float g = dy[(int64_t)i * lddy + j];
            if (relu && !(y_out[(int64_t)i * ldy + j] > 0.f)) g = 0.f;
            a += g;
            b += g * (x[(int64_t)i * ldx + j] - mu) * rs;
        }
    p1[w][lane] = a; p2[w][lane] = b;
    __syncthreads();
    if (w == 0) {
        float ta = 0.f, tb = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ta += p1[k][lane]; tb += p2[k][lane]; }
        s1[lane] = ta; s2[lane] = tb;
        if (j < N) { dbeta[j] = ta; dgamma[j] = tb; }
    }
    __syncthreads();
    if (j < N) {
        const float ga = gamma[j], m1 = s1[lane] / (float)M, m2 = s2[lane] / (float)M;
        for (int i = w; i < M; i += 8) {
            float g = dy[(int64_t)i * lddy + j];
            if (relu && !(y_out[(int64_t)i * ldy + j] > 0.f)) g = 0.f;
            const float xh = (x[(int64_t)i * ldx + j] - mu) * rs;
            dx[(int64_t)i * lddx + j] = ga * rs * (g - m1 - xh * m2);
        }
    }
}

// dst[i, j] (+)= src[i, j]  (strided fp32 matrices: residual-branch gradients, column-slice copies)
__global__ void __launch_bounds__(256)
add2d_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols,
             int accumulate) {
    td::pdl_sync();
    const int64_t total = (int64_t)rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / cols), j = (int)(e % cols);
        const float v = src[(int64_t)i * lds + j];
        float* d = dst + (int64_t)i * ldd + j;
        *d = accumulate ? *d + v : v;
    }
}

// dst[j, i] = src[i, j]: 32 x 32 tiles through shared memory, coalesced on both sides.  Feeds the tensor-core weight
// gradient of a Linear layer at large batch: dW = dY^T X reduces over the batch, and kind::tf32 wants that dimension
// contiguous in both operands (linear_tc.cu).
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols) {
    td::pdl_sync();
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < rows && j0 + tx < cols) tile[r][tx] = src[(int64_t)(i0 + r) * lds + j0 + tx];
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < cols && i0 + tx < rows) dst[(int64_t)(j0 + r) * ldd + i0 + tx] = tile[tx][r];
}

// nn.Dropout / the attention-weight dropout of nn.MultiheadAttention at sequence length 1
// (diffusion_transformer.py:19,27,29): out = x * keep / (1 - p), one Bernoulli(1-p) draw per `group`
// consecutive columns of a row (group 1: elementwise; group = head_dim: per (sample, head)).
// keep = Philox4x32-10(seed; counter = draw index, subsequence = seed_ptr[1]); the same call with the
// same seed applied to a gradient is the backward.
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t ldo, int rows, int cols,
               int group, float p, const uint64_t* __restrict__ seed_ptr) {
    td::pdl_sync();
    const Philox rng(seed_ptr[0]);
    const uint64_t sub = seed_ptr[1];
    const int gcols = cols / group;
    const float scale = 1.0f / (1.0f - p);
    const int64_t total = (int64_t)rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / cols), j = (int)(e % cols);
        const uint64_t draw = (uint64_t)i * gcols + j / group;
        uint32_t r[4];
        rng.gen(draw >> 2, sub, r);
        const float u = ((float)r[draw & 3] + 0.5f) * 2.3283064365386963e-10f;
        const float keep = (u >= p) ? scale : 0.f;
        out[(int64_t)i * ldo + j] = x[(int64_t)i * ldx + j] * keep;
    }
}

static inline int grid1d(int64_t items) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), (int64_t)kNumSMs * 8));
}

}  // namespace td

using namespace td;

extern "C" int64_t td_gemm_f32_workspace(int M, int N, int K) {
    // split-K slices are used when the output grid alone cannot fill the machine
    const int64_t tiles = ceil_div(M, G_BM) * ceil_div(N, G_BN);
    int nz = 1;
    if (tiles < kNumSMs && K >= 1024) nz = (int)std::min<int64_t>(std::min<int64_t>(kNumSMs / tiles, K / 256), 64);
    const int nt = td::gemm_tf32_splits(M, N, K);            // the tcgen05 path's own split (linear_tc.cu)
    return std::max<int64_t>(nz > 1 ? (int64_t)nz * M * N : 0, nt > 1 ? (int64_t)nt * M * N : 0);
}

extern "C" int td_gemm_f32(const td_gemm_args* a, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(a && a->A && a->B && a->C, "td_gemm_f32: null pointer");
    TD_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "td_gemm_f32: bad shape %d %d %d", a->M, a->N, a->K);
    TD_CHECK_ARG(a->act >= 0 && a->act <= TD_ACT_SIGMOID, "td_gemm_f32: bad activation %d", a->act);
    cudaStream_t s = (cudaStream_t)stream;
    {
        int st = TD_OK, nzt = 1;
        if (td::gemm_tf32_try(a, s, &st, &nzt)) {          // large batches, K-major operands: tcgen05 kind::tf32 (linear_tc.cu)
            if (st != TD_OK || nzt == 1) return st;
            td::launch(gemm_splitk_reduce_kernel, td::LaunchCfg(grid1d((int64_t)a->M * a->N), 256, 0, s), *a, nzt);
            return launch_status("gemm_splitk_reduce");
        }
    }
    int nz = 1;
    {   // FFMA split-K slices (the workspace may be larger: it also covers the tcgen05 path's split)
        const int64_t tiles = ceil_div(a->M, G_BM) * ceil_div(a->N, G_BN);
        if (a->splitk_ws && tiles < kNumSMs && a->K >= 1024)
            nz = (int)std::min<int64_t>(std::min<int64_t>(kNumSMs / tiles, a->K / 256), 64);
    }
    if (nz == 1 && a->M <= 4 && a->a_cs == 1 && a->b_rs == 1) {
        td::launch(gemv_f32_kernel, td::LaunchCfg((unsigned)ceil_div(a->N, 8), 256, 0, s), *a);
        return launch_status("gemv_f32");
    }
    if (nz == 1 && ceil_div(a->N, G_BN) * ceil_div(a->M, G_BM) < kNumSMs) {
        dim3 sgrid((unsigned)ceil_div(a->N, S_BN), (unsigned)ceil_div(a->M, S_BM));
        td::launch(gemm_f32_small_kernel, td::LaunchCfg(sgrid, 256, 0, s), *a);
        return launch_status("gemm_f32_small");
    }
    dim3 grid((unsigned)ceil_div(a->N, G_BN), (unsigned)ceil_div(a->M, G_BM), (unsigned)nz);
    td::launch(gemm_f32_kernel, td::LaunchCfg(grid, G_THREADS, 0, s), *a);
    int st = launch_status("gemm_f32");
    if (st != TD_OK || nz == 1) return st;
    td::launch(gemm_splitk_reduce_kernel, td::LaunchCfg(grid1d((int64_t)a->M * a->N), 256, 0, s), *a, nz);
    return launch_status("gemm_splitk_reduce");
}

extern "C" int td_colsum_f32(const float* x, int64_t ldx, float* out, int M, int N, int accumulate, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && out && M > 0 && N > 0, "td_colsum_f32: bad args");
    td::launch(colsum_kernel, td::LaunchCfg((N + 31) / 32, 256, 0, (cudaStream_t)stream), x, ldx, out, M, N, accumulate);
    return launch_status("colsum");
}

extern "C" int td_act_bwd_f32(const float* dy, const float* pre, float* dx, int64_t n, int act, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dy && pre && dx && n > 0, "td_act_bwd_f32: bad args");
    td::launch(act_bwd_kernel, td::LaunchCfg(grid1d(n), 256, 0, (cudaStream_t)stream), dy, pre, dx, n, act);
    return launch_status("act_bwd");
}

extern "C" int td_add2d_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int accumulate,
                            void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(src && dst && rows > 0 && cols > 0, "td_add2d_f32: bad args");
    td::launch(add2d_kernel, td::LaunchCfg(grid1d((int64_t)rows * cols), 256, 0, (cudaStream_t)stream), src, lds, dst, ldd, rows, cols, accumulate);
    return launch_status("add2d");
}

extern "C" int td_transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= rows, "td_transpose_f32: bad args");
    const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    TD_CHECK_ARG(grid.y <= 65535, "td_transpose_f32: more than 2^21 rows");
    td::launch(transpose_kernel, td::LaunchCfg(grid, 256, 0, (cudaStream_t)stream), src, lds, dst, ldd, rows, cols);
    return launch_status("transpose");
}

extern "C" int td_dropout_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int rows, int cols, int group,
                              float p, const uint64_t* seed_ptr, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && out && seed_ptr && rows > 0 && cols > 0 && group > 0 && cols % group == 0, "td_dropout_f32: bad args");
    TD_CHECK_ARG(p >= 0.f && p < 1.f, "td_dropout_f32: p must be in [0, 1)");
    td::launch(dropout_kernel, td::LaunchCfg(grid1d((int64_t)rows * cols), 256, 0, (cudaStream_t)stream), x, ldx, out, ldo, rows, cols, group, p,
                                                                                 seed_ptr);
    return launch_status("dropout");
}

extern "C" int td_embedding_bwd(const float* g, int64_t ldg, const int64_t* idx, float* table_grad, int M, int D,
                                int num_rows, int accumulate, float* workspace, int64_t workspace_floats, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(g && idx && table_grad && M > 0 && D > 0 && num_rows > 0, "td_embedding_bwd: bad args");
    const int64_t n = (int64_t)num_rows * D;
    if (workspace && M >= 4096 && n * 4 <= 96 * 1024 && workspace_floats >= 2 * n) {
        const int chunks = (int)std::min<int64_t>(std::min<int64_t>(workspace_floats / n, 2 * kNumSMs), M / 256);
        static size_t configured = 0;
        if ((size_t)n * 4 > 48 * 1024 && configured < 96 * 1024) {
            TD_CUDA(cudaFuncSetAttribute(embedding_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = 96 * 1024;
        }
        cudaStream_t s = (cudaStream_t)stream;
        td::launch(embedding_bwd_partial_kernel, td::LaunchCfg(chunks, 256, (size_t)n * 4, s), g, ldg, idx, workspace, M, D, num_rows);
        td::launch(embedding_bwd_finalize_kernel, td::LaunchCfg((unsigned)ceil_div(n, 256), 256, 0, s), (const float*)workspace, chunks, (int)n,
                   table_grad, accumulate);
        return launch_status("embedding_bwd");
    }
    td::launch(embedding_bwd_kernel, td::LaunchCfg(num_rows, 256, 0, (cudaStream_t)stream), g, ldg, idx, table_grad, M, D, accumulate);
    return launch_status("embedding_bwd");
}

extern "C" int td_time_features(const int64_t* t, const int32_t* t_dev, float* out, int batch, int dim, int mode,
                                void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG((t || t_dev) && out && batch > 0 && dim > 0 && mode >= 0 && mode <= 2, "td_time_features: bad args");
    td::launch(time_features_kernel, td::LaunchCfg(grid1d((int64_t)batch * (mode == 2 ? dim : 1)), 256, 0, (cudaStream_t)stream), 
        t, t_dev, out, batch, dim, mode);
    return launch_status("time_features");
}

extern "C" int td_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                float* rstd, int M, int D, float eps, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && gamma && beta && y && M > 0 && D > 0, "td_layernorm_fwd: bad args");
    td::launch(layernorm_fwd_kernel, td::LaunchCfg((M + 7) / 8, 256, 0, (cudaStream_t)stream), x, gamma, beta, y, mean, rstd, M, D, eps);
    return launch_status("layernorm_fwd");
}

extern "C" int td_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                const float* rstd, float* dx, float* dgamma, float* dbeta, int M, int D, float* workspace,
                                int64_t workspace_floats, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dy && x && gamma && mean && rstd && dx && M > 0 && D > 0, "td_layernorm_bwd: bad args");
    cudaStream_t s = (cudaStream_t)stream;
    td::launch(layernorm_bwd_kernel, td::LaunchCfg((M + 7) / 8, 256, 0, s), dy, x, gamma, mean, rstd, dx, M, D);
    int st = launch_status("layernorm_bwd");
    if (st != TD_OK || !dgamma || !dbeta) return st;
    if (workspace && M >= 4096 && workspace_floats >= 4 * (int64_t)D) {
        const int gx = (D + 31) / 32;
        const int chunks = (int)std::min<int64_t>(std::min<int64_t>(workspace_floats / (2 * (int64_t)D), std::max(1, 4 * kNumSMs / gx)), M / 64);
        td::launch(layernorm_bwd_params_partial_kernel, td::LaunchCfg(dim3((unsigned)gx, (unsigned)chunks), 256, 0, s), dy, x, mean, rstd,
                   workspace, M, D);
        st = launch_status("layernorm_bwd_params_partial");
        if (st != TD_OK) return st;
        st = td_partial_sum(workspace, chunks, D, 0, dgamma, stream);
        if (st != TD_OK) return st;
        return td_partial_sum(workspace, chunks, D, 1, dbeta, stream);
    }
    td::launch(layernorm_bwd_params_kernel, td::LaunchCfg((D + 31) / 32, 256, 0, s), dy, x, mean, rstd, dgamma, dbeta, M, D);
    return launch_status("layernorm_bwd_params");
}

extern "C" int td_bn1d_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float* save_mean, float* save_rstd, float* y, int64_t ldy, int M, int N,
                           float eps, float momentum, int training, int relu, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && gamma && beta && y && M > 0 && N > 0, "td_bn1d_fwd: bad args");
    TD_CHECK_ARG(training || (running_mean && running_var), "td_bn1d_fwd: eval mode needs running statistics");
    // train mode needs the whole column in one CTA (two-pass statistics in fixed order); eval mode splits the rows
    const int gx = (N + 31) / 32;
    const int gy = training ? 1 : std::max(1, std::min((M + 255) / 256, (4 * kNumSMs + gx - 1) / gx));
    td::launch(bn1d_fwd_kernel, td::LaunchCfg(dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream), x, ldx, gamma, beta, running_mean, running_var,
                                                                     save_mean, save_rstd, y, ldy, M, N, eps, momentum,
                                                                     training, relu);
    return launch_status("bn1d_fwd");
}

extern "C" int td_bn1d_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* y_out, int64_t ldy,
                           const float* gamma, const float* save_mean, const float* save_rstd, float* dx, int64_t lddx,
                           float* dgamma, float* dbeta, int M, int N, int relu, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dy && x && gamma && save_mean && save_rstd && dx && dgamma && dbeta && M > 0 && N > 0,
                 "td_bn1d_bwd: bad args");
    TD_CHECK_ARG(!relu || y_out, "td_bn1d_bwd: relu needs the saved forward output");
    td::launch(bn1d_bwd_kernel, td::LaunchCfg((N + 31) / 32, 256, 0, (cudaStream_t)stream), dy, lddy, x, ldx, y_out, ldy, gamma, save_mean,
                                                                     save_rstd, dx, lddx, dgamma, dbeta, M, N, relu);
    return launch_status("bn1d_bwd");
}

// ---------------------------------------------------------------------------------------------
// Conditioning head = time features -> Linear+SiLU -> Linear (+class row / +text) -> the three
// time_proj 1x1 convs as one [proj_out, D] Linear.  A chain of the GEMM above (4 launches); every
// intermediate is kept in `saved` for the backward.
// ---------------------------------------------------------------------------------------------
namespace {
struct EmbedLayout { int din; float *feat, *h_pre, *h, *emb; };
EmbedLayout embed_layout(const td_embed_args* a) {
    EmbedLayout l;
    l.din = a->in_mode == 2 ? a->dim : 1;
    l.feat = a->saved;
    l.h_pre = l.feat + (int64_t)a->batch * l.din;
    l.h = l.h_pre + (int64_t)a->batch * a->dim;
    l.emb = l.h + (int64_t)a->batch * a->dim;
    return l;
}
td_gemm_args gemm_init(int M, int N, int K) {
    td_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.alpha = 1.f;
    return g;
}
}  // namespace

extern "C" int64_t td_embed_head_saved_floats(int batch, int dim, int in_mode) {
    return (int64_t)batch * ((in_mode == 2 ? dim : 1) + 3 * (int64_t)dim);
}

extern "C" int td_embed_head_fwd(const td_embed_args* a, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(a, "td_embed_head_fwd: null args");
    TD_CHECK_ARG(a->batch > 0 && a->dim > 0, "td_embed_head_fwd: bad batch/dim");
    TD_CHECK_ARG(a->t || a->t_dev, "td_embed_head_fwd: need t or t_dev");
    TD_CHECK_ARG(a->w0 && a->b0 && a->w2 && a->b2 && a->saved, "td_embed_head_fwd: null weights / saved buffer");
    TD_CHECK_ARG(a->proj_out == 0 || (a->proj_w && a->proj_b && a->proj_out_ptr), "td_embed_head_fwd: null projection");
    TD_CHECK_ARG(a->in_mode >= 0 && a->in_mode <= 2, "td_embed_head_fwd: bad in_mode");
    const EmbedLayout l = embed_layout(a);
    const int B = a->batch, D = a->dim;
    int st = td_time_features(a->t, a->t_dev, l.feat, B, D, a->in_mode, stream);
    if (st != TD_OK) return st;
    {   // h = silu(feat W0^T + b0)
        td_gemm_args g = gemm_init(B, D, l.din);
        g.A = l.feat; g.a_rs = l.din; g.a_cs = 1;
        g.B = a->w0; g.b_rs = 1; g.b_cs = l.din;
        g.C = l.h; g.ldc = D; g.bias = a->b0; g.act = TD_ACT_SILU; g.pre_out = l.h_pre; g.ld_pre = D;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
    }
    {   // emb = h W2^T + b2 (+ class_table[y]) (+ text)
        td_gemm_args g = gemm_init(B, D, D);
        g.A = l.h; g.a_rs = D; g.a_cs = 1;
        g.B = a->w2; g.b_rs = 1; g.b_cs = D;
        g.C = l.emb; g.ldc = D; g.bias = a->b2;
        if (a->text) { g.residual = a->text; g.ldr = D; }
        if (a->y && a->class_table) { g.gather_idx = a->y; g.gather_table = a->class_table; g.ld_table = D; }
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
    }
    if (a->proj_out > 0) {   // proj = emb P^T + pb
        td_gemm_args g = gemm_init(B, a->proj_out, D);
        g.A = l.emb; g.a_rs = D; g.a_cs = 1;
        g.B = a->proj_w; g.b_rs = 1; g.b_cs = D;
        g.C = a->proj_out_ptr; g.ldc = a->proj_out; g.bias = a->proj_b;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
    }
    return TD_OK;
}

extern "C" int td_embed_head_bwd(const td_embed_args* a, const td_embed_grads* gr, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(a && gr && a->saved && gr->d_proj && gr->scratch, "td_embed_head_bwd: null args");
    TD_CHECK_ARG(gr->d_w0 && gr->d_b0 && gr->d_w2 && gr->d_b2 && gr->d_proj_w && gr->d_proj_b, "td_embed_head_bwd: null grads");
    const EmbedLayout l = embed_layout(a);
    const int B = a->batch, D = a->dim, P = a->proj_out;
    float* demb = gr->scratch;
    float* dh = gr->scratch + (int64_t)B * D;
    int st;
    {   // d_proj_w [P, D] = d_proj^T emb
        td_gemm_args g = gemm_init(P, D, B);
        g.A = gr->d_proj; g.a_rs = 1; g.a_cs = P;
        g.B = l.emb; g.b_rs = D; g.b_cs = 1;
        g.C = gr->d_proj_w; g.ldc = D;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
        if ((st = td_colsum_f32(gr->d_proj, P, gr->d_proj_b, B, P, 0, stream)) != TD_OK) return st;
    }
    {   // demb [B, D] = d_proj proj_w
        td_gemm_args g = gemm_init(B, D, P);
        g.A = gr->d_proj; g.a_rs = P; g.a_cs = 1;
        g.B = a->proj_w; g.b_rs = D; g.b_cs = 1;
        g.C = demb; g.ldc = D;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
    }
    if (a->y && gr->d_class_table) {
        if ((st = td_embedding_bwd(demb, D, a->y, gr->d_class_table, B, D, gr->num_classes, 0, nullptr, 0, stream)) != TD_OK) return st;
    }
    {   // d_w2 [D, D] = demb^T h ; d_b2 = colsum(demb)
        td_gemm_args g = gemm_init(D, D, B);
        g.A = demb; g.a_rs = 1; g.a_cs = D;
        g.B = l.h; g.b_rs = D; g.b_cs = 1;
        g.C = gr->d_w2; g.ldc = D;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
        if ((st = td_colsum_f32(demb, D, gr->d_b2, B, D, 0, stream)) != TD_OK) return st;
    }
    {   // dh = demb W2 ; dpre = dh * silu'(h_pre)
        td_gemm_args g = gemm_init(B, D, D);
        g.A = demb; g.a_rs = D; g.a_cs = 1;
        g.B = a->w2; g.b_rs = D; g.b_cs = 1;
        g.C = dh; g.ldc = D;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
        if ((st = td_act_bwd_f32(dh, l.h_pre, dh, (int64_t)B * D, TD_ACT_SILU, stream)) != TD_OK) return st;
    }
    {   // d_w0 [D, Din] = dpre^T feat ; d_b0 = colsum(dpre)
        td_gemm_args g = gemm_init(D, l.din, B);
        g.A = dh; g.a_rs = 1; g.a_cs = D;
        g.B = l.feat; g.b_rs = l.din; g.b_cs = 1;
        g.C = gr->d_w0; g.ldc = l.din;
        if ((st = td_gemm_f32(&g, stream)) != TD_OK) return st;
        if ((st = td_colsum_f32(dh, D, gr->d_b0, B, D, 0, stream)) != TD_OK) return st;
    }
    return TD_OK;
}
