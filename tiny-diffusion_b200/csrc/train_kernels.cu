// Train-mode kernels of the conv UNet (all HBM-bound, NHWC, 16-byte vector accesses):
//   BatchNorm2d batch statistics / finalize / apply+ReLU and its two-pass backward
//   (diffusion.py:34-35 in train mode), MaxPool2d backward, the transposes of the bilinear
//   resizes (decoder concat assembly, final 32->28 resize) and per-channel sums (bias gradients).
// Reductions are deterministic: every CTA writes its partial to its own row, a finalize kernel
// sums the rows in fixed order (in double).
#include <algorithm>

#include "common.cuh"

namespace td {

constexpr int kT = 256;

struct Bil { int i0, i1; float l0, l1; };
// identical to the forward helper in glue.cu (align_corners=True)
__device__ inline Bil bil_t(int dst, int in, int out) {
    Bil b;
    const float scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float src = scale * (float)dst;
    b.i0 = min((int)floorf(src), in - 1);
    b.l1 = fminf(fmaxf(src - (float)b.i0, 0.f), 1.f);
    b.l0 = 1.f - b.l1;
    b.i1 = b.i0 + ((b.i0 < in - 1) ? 1 : 0);
    return b;
}

// ---------------------------------------------------------------------------------------------
// per-channel partial sums over pixels (T = activation type; the raw conv output y is always fp32,
// see DESIGN.md: a bf16 y loses the signal under the large per-sample time-embedding offsets).
//   MODE 0: s1 = sum (x-K), s2 = sum (x-K)^2, K[c] = x[pixel 0, c] or 0   (shifted-data statistics)
//   MODE 1: s1 = sum g,     s2 = sum g*(y-mean),  g = da * [y*scale+shift > 0]   (BatchNorm backward)
// thread layout: lanesC = C / V channel-vector lanes, rows = kT / lanesC pixel rows per CTA pass.
// partials: [gridDim.x][2][C] followed by K[C]
// ---------------------------------------------------------------------------------------------
template <int V> __device__ inline void load_f32(const float* p, float* f) {
#pragma unroll
    for (int k = 0; k < V; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + k);
        f[k] = v.x; f[k + 1] = v.y; f[k + 2] = v.z; f[k + 3] = v.w;
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kT)
chan_reduce_kernel(const T* __restrict__ a, int64_t lda, int a_coff, const float* __restrict__ y, int64_t P, int C,
                   const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                   int shifted, float* __restrict__ partials) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    extern __shared__ float red[];               // [rows][2*C]
    const int lanesC = C / V;
    const int rows = kT / lanesC;
    const int lane = threadIdx.x % lanesC, row = threadIdx.x / lanesC;
    const int c0 = lane * V;
    float s1[V], s2[V], sc[V], sh[V], mu[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { s1[k] = 0.f; s2[k] = 0.f; sc[k] = 1.f; sh[k] = 0.f; mu[k] = 0.f; }
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < V; ++k) { sc[k] = scale[c0 + k]; sh[k] = shift[c0 + k]; mu[k] = mean[c0 + k]; }
    } else if (shifted) {
        Vec<T>::load(a + a_coff + c0).unpack(mu);            // K = the first pixel of every channel
    }
    if (row < rows) {
#pragma unroll 4
        for (int64_t p = (int64_t)blockIdx.x * rows + row; p < P; p += (int64_t)gridDim.x * rows) {
            float f[V];
            Vec<T>::load(a + p * lda + a_coff + c0).unpack(f);
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < V; ++k) { const float d = f[k] - mu[k]; s1[k] += d; s2[k] = fmaf(d, d, s2[k]); }
            } else {
                float yy[V];
                load_f32<V>(y + p * C + c0, yy);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const float g = (fmaf(yy[k], sc[k], sh[k]) > 0.f) ? f[k] : 0.f;
                    s1[k] += g;
                    s2[k] = fmaf(g, yy[k] - mu[k], s2[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            red[(size_t)row * 2 * C + c0 + k] = s1[k];
            red[(size_t)row * 2 * C + C + c0 + k] = s2[k];
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * C; j += kT) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += red[(size_t)r * 2 * C + j];
        partials[(size_t)blockIdx.x * 2 * C + j] = t;
    }
    if (MODE == 0 && blockIdx.x == 0 && row == 0) {
#pragma unroll
        for (int k = 0; k < V; ++k) partials[(size_t)gridDim.x * 2 * C + c0 + k] = mu[k];
    }
}

// BatchNorm2d train-mode finalize (diffusion.py:34): batch mean / biased variance -> the affine
// used by the apply pass; running statistics updated like torch (momentum, unbiased variance,
// conv bias folded into the mean only).
// CTA = 32 channels (threadIdx.x, coalesced 128-byte reads of a partial row) x 32 row groups (threadIdx.y); the
// row groups are combined through shared memory in fixed order, in double.  Result valid on the threadIdx.y == 0 row.
// (One warp per channel with lanes striding the rows read one 4-byte element per 32-byte sector and took ~10 us.)
__device__ inline void block_sum_partials(const float* __restrict__ partials, int nrows, int C, int c, double& s1,
                                          double& s2) {
    __shared__ double red[32][2][33];
    double a = 0.0, b = 0.0;
    if (c < C) {
        const float* p = partials + c;
        // batches of 16 rows per thread with every load of a batch issued before the first use: the kernel is a pure
        // latency chain (4..16 CTAs), so the loads in flight decide its duration (896 rows: 2 round trips instead of 7)
        for (int r0 = threadIdx.y; r0 < nrows; r0 += 32 * 16) {
            float va[16], vb[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int r = r0 + 32 * k;
                const bool ok = r < nrows;
                va[k] = ok ? __ldg(p + (size_t)r * 2 * C) : 0.f;
                vb[k] = ok ? __ldg(p + (size_t)r * 2 * C + C) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) { a += (double)va[k]; b += (double)vb[k]; }
        }
    }
    red[threadIdx.y][0][threadIdx.x] = a;
    red[threadIdx.y][1][threadIdx.x] = b;
    __syncthreads();
    s1 = 0.0; s2 = 0.0;
    if (threadIdx.y == 0) {
        for (int g = 0; g < 32; ++g) {
            s1 += red[g][0][threadIdx.x];
            s2 += red[g][1][threadIdx.x];
        }
    }
}

__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partials, int nrows, int C, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, int64_t* __restrict__ nbt,
                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd) {
    td::pdl_sync();
    const int c = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && nbt) nbt[0] += 1;
    // per-channel inputs of the tail are fetched before the reduction (one more round trip off the chain)
    const bool tail = threadIdx.y == 0 && c < C;
    float k_c = 0.f, g_c = 0.f, b_c = 0.f, cb_c = 0.f, rm_c = 0.f, rv_c = 0.f;
    if (tail) {
        k_c = partials[(size_t)nrows * 2 * C + c];
        g_c = gamma[c];
        b_c = beta[c];
        if (conv_bias) cb_c = conv_bias[c];
        if (running_mean) rm_c = running_mean[c];
        if (running_var) rv_c = running_var[c];
    }
    double s1, s2;
    block_sum_partials(partials, nrows, C, c, s1, s2);
    if (!tail) return;
    const double dm = s1 / count;                         // mean of (x - K)
    const double mean = (double)k_c + dm;
    double var = s2 / count - dm * dm;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = g_c * invstd;
    scale[c] = sc;
    shift[c] = b_c - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean) {
        const float mb = (float)mean + cb_c;
        running_mean[c] = (1.f - momentum) * rm_c + momentum * mb;
    }
    if (running_var) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.f - momentum) * rv_c + momentum * (float)unb;
    }
}

// a = relu(y * scale + shift)
template <typename T>
__global__ void __launch_bounds__(kT)
bn_relu_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                     T* __restrict__ a, int64_t lda, int a_coff, int64_t P, int C, int relu) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    const int lanesC = C / V;
    const int64_t total = P * lanesC;
    const uint32_t stride = gridDim.x * kT;               // multiple of lanesC -> channel group is loop invariant
    uint32_t i = blockIdx.x * kT + threadIdx.x;
    const int c0 = (int)(i % lanesC) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = scale[c0 + k]; sh[k] = shift[c0 + k]; }
    const int lsh = __ffs(lanesC) - 1;                    // lanesC is a power of two (checked on the host)
#pragma unroll 4
    for (; i < (uint32_t)total; i += stride) {
        const int64_t p = i >> lsh;
        float f[V];
        load_f32<V>(y + p * C + c0, f);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            f[k] = fmaf(f[k], sc[k], sh[k]);
            if (relu) f[k] = fmaxf(f[k], 0.f);
        }
        Vec<T>::pack(f).store(a + p * lda + a_coff + c0);
    }
}

// BatchNorm backward coefficients: dy = cA*g + cB*y + cC (see bn_relu_bwd_apply_kernel)
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int nrows, int C, double count,
                       const float* __restrict__ scale, const float* __restrict__ save_mean,
                       const float* __restrict__ save_invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ coef) {
    td::pdl_sync();
    const int c = blockIdx.x * 32 + threadIdx.x;
    const bool tail = threadIdx.y == 0 && c < C;
    float mean_c = 0.f, invstd_c = 0.f, sc_c = 0.f;
    if (tail) { mean_c = save_mean[c]; invstd_c = save_invstd[c]; sc_c = scale[c]; }
    double s1, s2;
    block_sum_partials(partials, nrows, C, c, s1, s2);
    if (!tail) return;
    const double mean = mean_c, invstd = invstd_c, sc = sc_c;
    const double dg = s2 * invstd;                     // sum g * xhat   (s2 = sum g * (y - mean))
    dgamma[c] = (float)dg;
    dbeta[c] = (float)s1;
    const double cB = -sc * invstd * dg / count;
    coef[c] = (float)sc;
    coef[C + c] = (float)cB;
    coef[2 * C + c] = (float)(-sc * s1 / count - cB * mean);
}

// dy = scale*(g - mean(g) - xhat*mean(g*xhat)) = cA*g + cB*y + cC,  g = da * [y*scale+shift > 0]
template <typename T>
__global__ void __launch_bounds__(kT)
bn_relu_bwd_apply_kernel(const T* __restrict__ da, int64_t ldda, int da_coff, const float* __restrict__ y,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ coef, T* __restrict__ dy, int64_t P, int C) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    const int lanesC = C / V;
    const int64_t total = P * lanesC;
    const uint32_t stride = gridDim.x * kT;
    uint32_t i = blockIdx.x * kT + threadIdx.x;
    const int c0 = (int)(i % lanesC) * V;
    float sc[V], sh[V], cA[V], cB[V], cC[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        sc[k] = scale[c0 + k]; sh[k] = shift[c0 + k];
        cA[k] = coef[c0 + k]; cB[k] = coef[C + c0 + k]; cC[k] = coef[2 * C + c0 + k];
    }
    const int lsh = __ffs(lanesC) - 1;                    // lanesC is a power of two (checked on the host)
#pragma unroll 4
    for (; i < (uint32_t)total; i += stride) {
        const int64_t p = i >> lsh;
        float g[V], yy[V];
        Vec<T>::load(da + p * ldda + da_coff + c0).unpack(g);
        load_f32<V>(y + p * C + c0, yy);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float gm = (fmaf(yy[k], sc[k], sh[k]) > 0.f) ? g[k] : 0.f;
            g[k] = fmaf(cA[k], gm, fmaf(cB[k], yy[k], cC[k]));
        }
        Vec<T>::pack(g).store(dy + p * C + c0);
    }
}

// V consecutive values of the raw conv output y (fp32, or bf16 on the layers without large per-sample offsets) as floats
template <int V> __device__ inline void load_y(const float* p, float* f) { load_f32<V>(p, f); }
template <int V> __device__ inline void load_y(const __nv_bfloat16* p, float* f) {
    static_assert(V == 8, "bf16 y is paired with bf16 activations (8-element vectors)");
    Vec<__nv_bfloat16>::load(p).unpack(f);
}

// ---------------------------------------------------------------------------------------------
// Fused BatchNorm passes: the finalize (sum of the partial rows -> per-channel coefficients) runs in the PROLOGUE of the
// streaming kernel that consumes it, so a BatchNorm is conv -> apply (forward) and reduce -> apply (backward) with no
// finalize launch in between (the finalize kernels were 4..16-CTA latency chains of ~8 us each, 26 per train step).
// A CTA owns a slice of 32 channels and a contiguous pixel range: its prologue sums nrows x 2 x 32 partials (<= 148 rows:
// the producers write one row per CTA) in fixed order, in double -- every CTA of a slice derives bit-identical coefficients --
// and the CTAs of pixel chunk 0 publish the per-channel results (running statistics, saved mean / invstd, dgamma / dbeta).
// Body: 32 / V threads per pixel (V = 8 bf16 / 4 fp32 channels per 16-byte vector), 128-byte fp32 and 64-byte bf16 segments.
// ---------------------------------------------------------------------------------------------
constexpr int kSlice = 32;

// sum of partials[r][which][c_base + ch] over r in fixed order; valid on threads 0..31 (ch = threadIdx.x) after the call
__device__ inline void slice_sum_partials(const float* __restrict__ partials, int nrows, int C, int c_base, double& s1, double& s2) {
    __shared__ double red[8][2][kSlice];
    const int ch = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const float* p = partials + c_base + ch;
    double a = 0.0, b = 0.0;
    for (int r0 = rg; r0 < nrows; r0 += 8 * 8) {
        float va[8], vb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r0 + 8 * k;
            const bool ok = r < nrows;
            va[k] = ok ? __ldg(p + (size_t)r * 2 * C) : 0.f;
            vb[k] = ok ? __ldg(p + (size_t)r * 2 * C + C) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += (double)va[k]; b += (double)vb[k]; }
    }
    red[rg][0][ch] = a;
    red[rg][1][ch] = b;
    __syncthreads();
    s1 = 0.0; s2 = 0.0;
    if (rg == 0) {
#pragma unroll
        for (int g = 0; g < 8; ++g) { s1 += red[g][0][ch]; s2 += red[g][1][ch]; }
    }
}

template <typename T, typename Ty>
__global__ void __launch_bounds__(kT)
bn_apply_fused_kernel(const Ty* __restrict__ y, const float* __restrict__ partials, int nrows, double count,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias,
                      float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                      int64_t* __restrict__ nbt, float* __restrict__ scale_out, float* __restrict__ shift_out,
                      float* __restrict__ save_mean, float* __restrict__ save_invstd, T* __restrict__ a, int64_t lda,
                      int a_coff, int64_t P, int C, int relu) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    constexpr int TPP = kSlice / V;                 // threads per pixel
    constexpr int PPP = kT / TPP;                   // pixels per pass
    __shared__ float s_sc[kSlice], s_sh[kSlice], s_mu[kSlice];      // scale, beta, mean
    const int c_base = blockIdx.y * kSlice;
    {
        const int c = c_base + (threadIdx.x & 31);
        const bool tail = threadIdx.x < 32;
        float k_c = 0.f, g_c = 0.f, b_c = 0.f, cb_c = 0.f, rm_c = 0.f, rv_c = 0.f;
        if (tail) {                                 // per-channel inputs fetched before the reduction (off the latency chain)
            k_c = partials[(size_t)nrows * 2 * C + c];
            g_c = gamma[c];
            b_c = beta[c];
            if (blockIdx.x == 0) {
                if (conv_bias) cb_c = conv_bias[c];
                if (running_mean) rm_c = running_mean[c];
                if (running_var) rv_c = running_var[c];
            }
        }
        double s1, s2;
        slice_sum_partials(partials, nrows, C, c_base, s1, s2);
        if (tail) {
            const double dm = s1 / count;                         // mean of (x - K)
            const double mean = (double)k_c + dm;
            double var = s2 / count - dm * dm;
            if (var < 0.0) var = 0.0;
            const float invstd = (float)(1.0 / sqrt(var + (double)eps));
            const float sc = g_c * invstd;
            const float sh = b_c - (float)mean * sc;
            s_sc[threadIdx.x] = sc;
            s_sh[threadIdx.x] = b_c;
            s_mu[threadIdx.x] = (float)mean;
            if (blockIdx.x == 0) {
                scale_out[c] = sc;
                shift_out[c] = sh;
                save_mean[c] = (float)mean;
                save_invstd[c] = invstd;
                if (running_mean) running_mean[c] = (1.f - momentum) * rm_c + momentum * ((float)mean + cb_c);
                if (running_var) {
                    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
                    running_var[c] = (1.f - momentum) * rv_c + momentum * (float)unb;
                }
                if (blockIdx.y == 0 && threadIdx.x == 0 && nbt) nbt[0] += 1;
            }
        }
        __syncthreads();
    }
    // a = relu(scale * (y - mean) + beta): y carries per-channel offsets of O(1e2..1e3) under the raw-t time embedding, and
    // y * scale + (beta - mean * scale) would cancel in fp32 (ReLU masks flip at pre-activations of ~1e-5)
    const int lane_c = (threadIdx.x % TPP) * V;
    float sc[V], sh[V], mu[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = s_sc[lane_c + k]; sh[k] = s_sh[lane_c + k]; mu[k] = s_mu[lane_c + k]; }
    const int64_t p_lo = P * blockIdx.x / gridDim.x, p_hi = P * (blockIdx.x + 1) / gridDim.x;
    const int c0 = c_base + lane_c;
#pragma unroll 4
    for (int64_t p = p_lo + threadIdx.x / TPP; p < p_hi; p += PPP) {
        float f[V];
        load_y<V>(y + p * C + c0, f);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            f[k] = fmaf(f[k] - mu[k], sc[k], sh[k]);
            if (relu) f[k] = fmaxf(f[k], 0.f);
        }
        Vec<T>::pack(f).store(a + p * lda + a_coff + c0);
    }
}

// BatchNorm backward partial sums, channel-sliced: s1 = sum g, s2 = sum g*(y-mean), g = da * [y*scale+shift > 0].
// grid (chunks, C / 32); partials [chunks][2][C] (one row per pixel chunk).
template <typename T, typename Ty>
__global__ void __launch_bounds__(kT)
bn_bwd_reduce_sliced_kernel(const T* __restrict__ da, int64_t ldda, int da_coff, const Ty* __restrict__ y, int64_t P, int C,
                            const float* __restrict__ scale, const float* __restrict__ beta, const float* __restrict__ mean,
                            float* __restrict__ partials) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    constexpr int TPP = kSlice / V;
    constexpr int PPP = kT / TPP;
    __shared__ float red[PPP][2 * kSlice + 1];
    const int c_base = blockIdx.y * kSlice;
    const int lane_c = (threadIdx.x % TPP) * V;
    const int c0 = c_base + lane_c;
    float s1[V], s2[V], sc[V], sh[V], mu[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { s1[k] = 0.f; s2[k] = 0.f; sc[k] = scale[c0 + k]; sh[k] = beta[c0 + k]; mu[k] = mean[c0 + k]; }
    const int64_t p_lo = P * blockIdx.x / gridDim.x, p_hi = P * (blockIdx.x + 1) / gridDim.x;
#pragma unroll 4
    for (int64_t p = p_lo + threadIdx.x / TPP; p < p_hi; p += PPP) {
        float f[V], yy[V];
        Vec<T>::load(da + p * ldda + da_coff + c0).unpack(f);
        load_y<V>(y + p * C + c0, yy);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float d = yy[k] - mu[k];
            const float g = (fmaf(d, sc[k], sh[k]) > 0.f) ? f[k] : 0.f;      // the forward's pre-activation, same op order
            s1[k] += g;
            s2[k] = fmaf(g, d, s2[k]);
        }
    }
    const int prow = threadIdx.x / TPP;
#pragma unroll
    for (int k = 0; k < V; ++k) { red[prow][lane_c + k] = s1[k]; red[prow][kSlice + lane_c + k] = s2[k]; }
    __syncthreads();
    if (threadIdx.x < 2 * kSlice) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;                 // fixed order: four interleaved chains, then a tree
#pragma unroll 4
        for (int r = 0; r < PPP; r += 4) {
            t0 += red[r][threadIdx.x]; t1 += red[r + 1][threadIdx.x]; t2 += red[r + 2][threadIdx.x]; t3 += red[r + 3][threadIdx.x];
        }
        const int which = threadIdx.x / kSlice, cc = threadIdx.x % kSlice;
        partials[(size_t)blockIdx.x * 2 * C + (size_t)which * C + c_base + cc] = (t0 + t1) + (t2 + t3);
    }
}

// dy = scale*(g - mean(g) - xhat*mean(g*xhat)) = cA*g + cB*y + cC with the coefficients derived in the prologue from the
// partial rows of bn_bwd_reduce_sliced_kernel (or of a data-gradient convolution's epilogue); chunk 0 writes dgamma / dbeta.
template <typename T, typename Ty>
__global__ void __launch_bounds__(kT)
bn_bwd_apply_fused_kernel(const T* __restrict__ da, int64_t ldda, int da_coff, const Ty* __restrict__ y,
                          const float* __restrict__ partials, int nrows, double count, const float* __restrict__ scale,
                          const float* __restrict__ beta, const float* __restrict__ save_mean,
                          const float* __restrict__ save_invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          T* __restrict__ dy, int64_t P, int C) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    constexpr int TPP = kSlice / V;
    constexpr int PPP = kT / TPP;
    __shared__ float s_co[6][kSlice];               // scale, beta, cA, cB, cC, mean
    const int c_base = blockIdx.y * kSlice;
    {
        const int c = c_base + (threadIdx.x & 31);
        const bool tail = threadIdx.x < 32;
        float mean_c = 0.f, invstd_c = 0.f, sc_c = 0.f, sh_c = 0.f;
        if (tail) { mean_c = save_mean[c]; invstd_c = save_invstd[c]; sc_c = scale[c]; sh_c = beta[c]; }
        double s1, s2;
        slice_sum_partials(partials, nrows, C, c_base, s1, s2);
        if (tail) {
            const double mean = mean_c, invstd = invstd_c, sc = sc_c;
            const double dg = s2 * invstd;                     // sum g * xhat   (s2 = sum g * (y - mean))
            const double cB = -sc * invstd * dg / count;
            s_co[0][threadIdx.x] = sc_c;
            s_co[1][threadIdx.x] = sh_c;
            s_co[2][threadIdx.x] = (float)sc;
            s_co[3][threadIdx.x] = (float)cB;
            // dy = cA*g + cB*(y - mean) + cC with the mean subtracted per element: y carries per-channel offsets of O(1e2..1e3)
            // (raw-t time embedding), and cB*y + (cC - cB*mean) would cancel in fp32
            s_co[4][threadIdx.x] = (float)(-sc * s1 / count);
            s_co[5][threadIdx.x] = mean_c;
            if (blockIdx.x == 0) {
                dgamma[c] = (float)dg;
                dbeta[c] = (float)s1;
            }
        }
        __syncthreads();
    }
    const int lane_c = (threadIdx.x % TPP) * V;
    float sc[V], sh[V], cA[V], cB[V], cC[V], mu[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        sc[k] = s_co[0][lane_c + k]; sh[k] = s_co[1][lane_c + k];
        cA[k] = s_co[2][lane_c + k]; cB[k] = s_co[3][lane_c + k]; cC[k] = s_co[4][lane_c + k]; mu[k] = s_co[5][lane_c + k];
    }
    const int64_t p_lo = P * blockIdx.x / gridDim.x, p_hi = P * (blockIdx.x + 1) / gridDim.x;
    const int c0 = c_base + lane_c;
#pragma unroll 4
    for (int64_t p = p_lo + threadIdx.x / TPP; p < p_hi; p += PPP) {
        float g[V], yy[V];
        Vec<T>::load(da + p * ldda + da_coff + c0).unpack(g);
        load_y<V>(y + p * C + c0, yy);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float d = yy[k] - mu[k];
            const float gm = (fmaf(d, sc[k], sh[k]) > 0.f) ? g[k] : 0.f;
            g[k] = fmaf(cA[k], gm, fmaf(cB[k], d, cC[k]));
        }
        Vec<T>::pack(g).store(dy + p * C + c0);
    }
}

// pixel chunks of the channel-sliced kernels: ~4 CTAs per SM in total, never more rows than SMs, >= 2 passes per CTA
static inline int sliced_chunks(int64_t P, int C, int ppp) {
    const int slices = C / kSlice;
    int64_t n = std::max<int64_t>(1, (int64_t)kNumSMs * 4 / slices);
    n = std::min<int64_t>(n, kNumSMs);
    n = std::min<int64_t>(n, std::max<int64_t>(1, P / (2 * ppp)));
    return (int)n;
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2, ceil_mode) backward, gather form: an input pixel receives the window's gradient
// iff it is the first maximum in (h, w) scan order (ATen's tie rule).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kT)
maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int C,
                    int Ho, int Wo, int accumulate) {
    td::pdl_sync();
    // one CTA per input row (b, h); threadIdx.x = channel vector, threadIdx.y = pixel of the row
    constexpr int V = Vec<T>::N;
    const int cv = C / V;
    const int b = blockIdx.x / H, h = blockIdx.x - b * H;
    const int ho = h >> 1;
    const T* xb = x + (int64_t)b * H * W * C;
    for (int cvi = threadIdx.x; cvi < cv; cvi += blockDim.x) {
        const int c = cvi * V;
        for (int w = threadIdx.y; w < W; w += blockDim.y) {
            const int wo = w >> 1;
            float out[V];
#pragma unroll
            for (int k = 0; k < V; ++k) out[k] = 0.f;
            if (ho < Ho && wo < Wo) {
                float best[V];
                int arg[V];
#pragma unroll
                for (int k = 0; k < V; ++k) { best[k] = -INFINITY; arg[k] = -1; }
#pragma unroll
                for (int dyy = 0; dyy < 2; ++dyy) {
                    const int hh = 2 * ho + dyy;
                    if (hh >= H) continue;
#pragma unroll
                    for (int dxx = 0; dxx < 2; ++dxx) {
                        const int ww = 2 * wo + dxx;
                        if (ww >= W) continue;
                        float f[V];
                        Vec<T>::load(xb + ((int64_t)hh * W + ww) * C + c).unpack(f);
#pragma unroll
                        for (int k = 0; k < V; ++k)
                            if (f[k] > best[k] || arg[k] < 0) { best[k] = f[k]; arg[k] = dyy * 2 + dxx; }
                    }
                }
                const int me = (h & 1) * 2 + (w & 1);
                float g[V];
                Vec<T>::load(dy + (((int64_t)b * Ho + ho) * Wo + wo) * C + c).unpack(g);
#pragma unroll
                for (int k = 0; k < V; ++k) out[k] = (arg[k] == me) ? g[k] : 0.f;
            }
            T* dst = dx + (((int64_t)b * H + h) * W + w) * C + c;
            if (accumulate) {
                float prev[V];
                Vec<T>::load(dst).unpack(prev);
#pragma unroll
                for (int k = 0; k < V; ++k) out[k] += prev[k];
            }
            Vec<T>::pack(out).store(dst);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Transposed bilinear resize (gather form, deterministic): dx[b,hi,wi,:] = sum over the output
// pixels that read (hi,wi) of weight * dy.  dy: [B,Ho,Wo,ld] at channel offset coff.
// ---------------------------------------------------------------------------------------------
// weights with which input index `i` enters the outputs o_lo .. o_lo+n-1 of one dimension (n <= 8)
struct BilT { int o_lo, n; float w[8]; };
__device__ inline BilT bil_transpose(int i, int in, int out) {
    BilT t;
    t.n = 0; t.o_lo = 0;
    const float sc = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
    int lo = 0, hi = out - 1;
    if (sc > 0.f) { lo = max(0, (int)floorf((i - 1) / sc) - 1); hi = min(out - 1, (int)ceilf((i + 1) / sc) + 1); }
    bool started = false;
    for (int o = lo; o <= hi && t.n < 8; ++o) {
        const Bil b = bil_t(o, in, out);
        const float wgt = (b.i0 == i ? b.l0 : 0.f) + (b.i1 == i ? b.l1 : 0.f);
        if (!started) {
            if (wgt == 0.f) continue;
            started = true;
            t.o_lo = o;
        }
        t.w[t.n++] = wgt;            // zeros inside the run are kept (harmless), trailing zeros trimmed below
    }
    while (t.n > 0 && t.w[t.n - 1] == 0.f) --t.n;
    return t;
}

constexpr int kMaxRowW = 64;      // widest output row the column table is sized for

// Transposed resize, the mirror image of glue.cu's resample_row: one thread owns one 16-byte channel vector of one
// INPUT row and walks the OUTPUT columns left to right.  Per column it gathers the (<= 8) output rows that read this
// input row, then adds the result into two register accumulators -- the input columns i0 and i0+1 the output column
// interpolated between -- and stores an accumulator when the walk has moved past its column.  Deterministic
// (fixed summation order), no per-pixel table search (the previous per-input-pixel gather spent most of its
// instructions in bil_transpose).
template <typename T>
__global__ void __launch_bounds__(kT)
resize_bwd_kernel(const T* __restrict__ dy, int64_t ld, int coff, T* __restrict__ dx, int B, int Hi, int Wi, int Ho,
                  int Wo, int C, int nseg) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    __shared__ Bil col[kMaxRowW];
    const bool same = (Hi == Ho && Wi == Wo);
    if (!same)
        for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < Wo; i += blockDim.x * blockDim.y) col[i] = bil_t(i, Wi, Wo);
    __syncthreads();
    const int cvi = blockIdx.y * blockDim.x + threadIdx.x;
    const int seg = blockIdx.x % nseg;                            // segment of INPUT columns owned by this thread
    const int row = (blockIdx.x / nseg) * blockDim.y + threadIdx.y;        // (b, hi) flattened
    if (cvi * V >= C || row >= B * Hi) return;
    const int sw = (Wi + nseg - 1) / nseg;
    const int ca = seg * sw, cb = min(Wi, ca + sw);               // input columns [ca, cb)
    if (ca >= cb) return;
    const int c = cvi * V;
    const int b = row / Hi, hi = row - b * Hi;
    T* xrow = dx + (int64_t)row * Wi * C + c;
    if (same) {
        const T* g = dy + ((int64_t)row * Wo) * ld + coff + c;
        for (int w = ca; w < cb; ++w) Vec<T>::load(g + (int64_t)w * ld).store(xrow + (int64_t)w * C);
        return;
    }
    const BilT th = bil_transpose(hi, Hi, Ho);
    const T* g0 = dy + (((int64_t)b * Ho + th.o_lo) * Wo) * ld + coff + c;
    // output columns that can touch [ca, cb): conservative bounds from the inverse scale, the walk filters exactly
    const float inv = (Wi > 1) ? (float)(Wo - 1) / (float)(Wi - 1) : 0.f;
    const int w_begin = Wi > 1 ? max(0, (int)floorf((float)(ca - 1) * inv) - 1) : 0;
    const int w_end = Wi > 1 ? min(Wo, (int)ceilf((float)cb * inv) + 2) : Wo;
    float acc0[V], acc1[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc0[k] = acc1[k] = 0.f;
    int ia = ca;                                                  // input column held by acc0 (acc1: ia + 1)
    for (int w = w_begin; w < w_end; ++w) {
        const Bil bw = col[w];
        if (bw.i1 < ca) continue;
        if (bw.i0 >= cb) break;
        float gv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) gv[k] = 0.f;
#pragma unroll
        for (int a = 0; a < 8; ++a) {                 // static indices keep th.w in registers
            if (a < th.n) {
                float f[V];
                Vec<T>::load(g0 + ((int64_t)a * Wo + w) * ld).unpack(f);
#pragma unroll
                for (int k = 0; k < V; ++k) gv[k] = fmaf(th.w[a], f[k], gv[k]);
            }
        }
        while (ia < bw.i0) {                          // the walk has moved past column ia: it is final
            Vec<T>::pack(acc0).store(xrow + (int64_t)ia * C);
#pragma unroll
            for (int k = 0; k < V; ++k) { acc0[k] = acc1[k]; acc1[k] = 0.f; }
            ++ia;
        }
        // contributions (i0, l0) and (i1, l1); columns outside [ca, cb) belong to the neighbouring segments
        const float l0 = bw.i1 == bw.i0 ? bw.l0 + bw.l1 : bw.l0;
        if (bw.i0 == ia) {
#pragma unroll
            for (int k = 0; k < V; ++k) acc0[k] = fmaf(l0, gv[k], acc0[k]);
            if (bw.i1 != bw.i0) {
#pragma unroll
                for (int k = 0; k < V; ++k) acc1[k] = fmaf(bw.l1, gv[k], acc1[k]);
            }
        } else if (bw.i1 == ia && bw.i1 != bw.i0) {   // i0 == ia - 1 lies in the previous segment
#pragma unroll
            for (int k = 0; k < V; ++k) acc0[k] = fmaf(bw.l1, gv[k], acc0[k]);
        }
    }
    for (; ia < cb; ++ia) {
        Vec<T>::pack(acc0).store(xrow + (int64_t)ia * C);
#pragma unroll
        for (int k = 0; k < V; ++k) { acc0[k] = acc1[k]; acc1[k] = 0.f; }
    }
}

// MaxPool2d(2, ceil_mode) backward, one thread per WINDOW (output pixel) and channel vector: the window's <= 4 inputs are read
// once (the per-input-pixel form above reads every window four times), the first maximum in (h, w) scan order takes the
// gradient (ATen's tie rule), the others get zero.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_bwd_window_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int C,
                           int Ho, int Wo, int accumulate) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    const int cvt = C / V;
    const int b = blockIdx.x / Ho, ho = blockIdx.x - b * Ho;
    const int64_t img = (int64_t)b * H * W * C;
    for (int idx = threadIdx.x; idx < Wo * cvt; idx += blockDim.x) {
        const int wo = idx / cvt, c = (idx - wo * cvt) * V;
        float f[4][V], g[V];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int hh = 2 * ho + (j >> 1), ww = 2 * wo + (j & 1);
            ok[j] = hh < H && ww < W;
            if (ok[j]) Vec<T>::load(x + img + ((int64_t)hh * W + ww) * C + c).unpack(f[j]);
        }
        Vec<T>::load(dy + (((int64_t)b * Ho + ho) * Wo + wo) * C + c).unpack(g);
        int arg[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            float best = -INFINITY;
            arg[k] = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ok[j] && (f[j][k] > best || arg[k] < 0)) { best = f[j][k]; arg[k] = j; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!ok[j]) continue;
            const int hh = 2 * ho + (j >> 1), ww = 2 * wo + (j & 1);
            T* dst = dx + img + ((int64_t)hh * W + ww) * C + c;
            float o[V];
#pragma unroll
            for (int k = 0; k < V; ++k) o[k] = (arg[k] == j) ? g[k] : 0.f;
            if (accumulate) {
                float prev[V];
                Vec<T>::load(dst).unpack(prev);
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] += prev[k];
            }
            Vec<T>::pack(o).store(dst);
        }
    }
}

// one CTA per row, at most 256 threads looping over the row (several CTAs resident per SM)
static inline int row_threads(int64_t items) { return (int)std::min<int64_t>(256, (items + 31) / 32 * 32); }

// block (channel vectors, rows) for the row-walking kernels: up to 32 channel vectors wide, 256 threads
static inline dim3 walk_block(int cv) {
    int bx = 1;
    while (bx < cv && bx < 32) bx <<= 1;
    return dim3((unsigned)bx, (unsigned)(kT / bx), 1);
}
// input-row segments per thread walk: enough threads in flight (~1500 per SM) to cover the load latency
static inline int walk_segments(int64_t threads, int width) {
    int64_t n = ceil_div((int64_t)kNumSMs * 1536, std::max<int64_t>(threads, 1));
    n = std::max<int64_t>(1, std::min<int64_t>(n, std::max(1, width / 2)));
    const int sw = (int)ceil_div(width, n);
    return (int)ceil_div(width, sw);
}

template <typename T>
static int launch_resize_bwd(const void* dy, int64_t ld, int coff, void* dx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                             cudaStream_t s) {
    constexpr int V = Vec<T>::N;
    const dim3 blk = walk_block(C / V);
    const int nseg = walk_segments((int64_t)B * Hi * (C / V), Wi);
    const dim3 grd((unsigned)(ceil_div((int64_t)B * Hi, blk.y) * nseg), (unsigned)ceil_div(C / V, blk.x), 1);
    td::launch(resize_bwd_kernel<T>, td::LaunchCfg(grd, blk, 0, s), (const T*)dy, ld, coff, (T*)dx, B, Hi, Wi, Ho, Wo, C, nseg);
    return TD_OK;
}

// d_temb[b, off + c] = sum over the pixels of src[b, :, :, coff + c] (src: the skip gradient, see td_upcat_bwd).  grid (samples, channel slices of
// 8 vectors); 8 lanes x 32 pixel rows per CTA, fixed-order tree over the rows.
template <typename T>
__global__ void __launch_bounds__(kT)
temb_bwd_kernel(const T* __restrict__ dout, int64_t ld, int coff, float* __restrict__ dtemb, int ld_temb, int temb_off,
                int HW, int Cs) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    __shared__ float red[32][8 * V + 1];
    const int lane = threadIdx.x & 7, prow = threadIdx.x >> 3;
    const int b = blockIdx.x;
    const int c = (blockIdx.y * 8 + lane) * V;
    float s[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s[k] = 0.f;
    if (c < Cs) {
        const T* src = dout + (int64_t)b * HW * ld + coff + c;
#pragma unroll 8
        for (int p = prow; p < HW; p += 32) {
            float f[V];
            Vec<T>::load(src + (int64_t)p * ld).unpack(f);
#pragma unroll
            for (int k = 0; k < V; ++k) s[k] += f[k];
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) red[prow][lane * V + k] = s[k];
    __syncthreads();
    if (threadIdx.x < 8 * V) {
        const int cc = blockIdx.y * 8 * V + threadIdx.x;
        if (cc < Cs) {
            float t = 0.f;
            for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
            dtemb[(int64_t)b * ld_temb + temb_off + cc] = t;
        }
    }
}

// out[c] = sum_r partials[r][which][c]    (bias gradients from chan_reduce partials)
__global__ void __launch_bounds__(1024)
partial_sum_kernel(const float* __restrict__ partials, int nrows, int C, int which, float* __restrict__ out) {
    td::pdl_sync();
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s1, s2;
    block_sum_partials(partials, nrows, C, c, s1, s2);
    if (threadIdx.y == 0 && c < C) out[c] = (float)(which ? s2 : s1);
}

// per-channel sum of an NCHW fp32 tensor (final_conv bias gradient): grid (chunks, C) partial sums,
// then a fixed-order finalize
constexpr int kChanSumChunks = 128;
__global__ void __launch_bounds__(kT)
nchw_chansum_kernel(const float* __restrict__ x, int B, int C, int HW, float* __restrict__ ws) {
    td::pdl_sync();
    __shared__ float red[kT];
    const int c = blockIdx.y;
    const int64_t total = (int64_t)B * HW;
    float s = 0.f;
    for (uint32_t i = blockIdx.x * kT + threadIdx.x; i < (uint32_t)total; i += gridDim.x * kT) {
        const int b = (int)(i / HW), p = (int)(i % HW);
        s += x[((int64_t)b * C + c) * HW + p];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) ws[(size_t)c * gridDim.x + blockIdx.x] = red[0];
}
__global__ void nchw_chansum_finalize_kernel(const float* __restrict__ ws, int chunks, int C, float* __restrict__ out) {
    td::pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int i = 0; i < chunks; ++i) s += (double)ws[(size_t)c * chunks + i];
    out[c] = (float)s;
}

static inline int reduce_grid(int64_t P, int rows) {
    const int64_t want = ceil_div(P, (int64_t)rows * 4);     // >= 4 pixel rows per thread: one batch of loads in flight
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kNumSMs * 4));
}
static inline int stream_grid(int64_t items, int lanesC) {
    // grid * kT must stay a multiple of lanesC (kT is, for every power-of-two lanesC <= kT)
    (void)lanesC;
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, kT), (int64_t)kNumSMs * 16));
}
static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline dim3 row_block(int cv) {
    const int bx = cv < kT ? cv : kT;
    int by = kT / bx;
    if (by < 1) by = 1;
    return dim3((unsigned)bx, (unsigned)by, 1);
}

}  // namespace td

using namespace td;

#define TD_DISPATCH_T(dtype, ...)                                       \
    if ((dtype) == TD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else if ((dtype) == TD_F32) { using T = float; __VA_ARGS__; }       \
    else { TD_CHECK_ARG(false, "unknown dtype %d", (int)(dtype)); }

static int check_lanes(const char* what, int dtype, int C) {
    const int V = dtype == TD_BF16 ? 8 : 4;
    TD_CHECK_ARG(C % V == 0 && pow2(C / V) && C / V <= kT, "%s: channel count %d unsupported", what, C);
    return TD_OK;
}

extern "C" int td_chan_reduce_rows(int dtype, int64_t pixels, int channels) {
    const int V = dtype == TD_BF16 ? 8 : 4;
    const int lanesC = std::max(1, channels / V);
    return reduce_grid(pixels, std::max(1, kT / lanesC));
}

extern "C" int td_bn_stats(const void* x, int dtype, int64_t ldx, int x_coff, int64_t pixels, int channels,
                           float* partials, int shifted, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && partials && pixels > 0, "td_bn_stats: bad args");
    if (int st = check_lanes("td_bn_stats", dtype, channels)) return st;
    const int grid = td_chan_reduce_rows(dtype, pixels, channels);
    const int V = dtype == TD_BF16 ? 8 : 4;
    const size_t smem = (size_t)(kT / (channels / V)) * 2 * channels * sizeof(float);
    TD_DISPATCH_T(dtype, (td::launch(chan_reduce_kernel<T, 0>, td::LaunchCfg(grid, kT, smem, (cudaStream_t)stream), 
                             (const T*)x, ldx, x_coff, nullptr, pixels, channels, nullptr, nullptr, nullptr, shifted,
                             partials)));
    return launch_status("bn_stats");
}

extern "C" int td_bn_finalize(const float* partials, int nrows, int channels, int64_t count, const float* gamma,
                              const float* beta, const float* conv_bias, float eps, float momentum,
                              float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scale,
                              float* shift, float* save_mean, float* save_invstd, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(partials && nrows > 0 && channels > 0 && count > 0 && gamma && beta && scale && shift && save_mean &&
                     save_invstd, "td_bn_finalize: bad args");
    td::launch(bn_finalize_kernel, td::LaunchCfg((channels + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream), 
        partials, nrows, channels, (double)count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var,
        num_batches_tracked, scale, shift, save_mean, save_invstd);
    return launch_status("bn_finalize");
}

extern "C" int td_bn_relu_apply(const float* y, const float* scale, const float* shift, void* a, int dtype, int64_t lda,
                                int a_coff, int64_t pixels, int channels, int relu, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(y && scale && shift && a && pixels > 0, "td_bn_relu_apply: bad args");
    if (int st = check_lanes("td_bn_relu_apply", dtype, channels)) return st;
    const int V = dtype == TD_BF16 ? 8 : 4;
    const int lanesC = channels / V;
    const int grid = stream_grid(pixels * lanesC, lanesC);
    TD_DISPATCH_T(dtype, (td::launch(bn_relu_apply_kernel<T>, td::LaunchCfg(grid, kT, 0, (cudaStream_t)stream), 
                             y, scale, shift, (T*)a, lda, a_coff, pixels, channels, relu)));
    return launch_status("bn_relu_apply");
}

extern "C" int td_bn_relu_bwd_reduce(const void* da, int64_t ldda, int da_coff, const float* y, int dtype,
                                     const float* scale, const float* shift, const float* save_mean, int64_t pixels,
                                     int channels, float* partials, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(da && y && scale && shift && save_mean && partials && pixels > 0, "td_bn_relu_bwd_reduce: bad args");
    if (int st = check_lanes("td_bn_relu_bwd_reduce", dtype, channels)) return st;
    const int grid = td_chan_reduce_rows(dtype, pixels, channels);
    const int V = dtype == TD_BF16 ? 8 : 4;
    const size_t smem = (size_t)(kT / (channels / V)) * 2 * channels * sizeof(float);
    TD_DISPATCH_T(dtype, (td::launch(chan_reduce_kernel<T, 1>, td::LaunchCfg(grid, kT, smem, (cudaStream_t)stream), 
                             (const T*)da, ldda, da_coff, y, pixels, channels, scale, shift, save_mean, 0, partials)));
    return launch_status("bn_relu_bwd_reduce");
}

extern "C" int td_bn_bwd_finalize(const float* partials, int nrows, int channels, int64_t count, const float* scale,
                                  const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                                  float* coef, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(partials && nrows > 0 && channels > 0 && count > 0 && scale && save_mean && save_invstd && dgamma &&
                     dbeta && coef, "td_bn_bwd_finalize: bad args");
    td::launch(bn_bwd_finalize_kernel, td::LaunchCfg((channels + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream), 
        partials, nrows, channels, (double)count, scale, save_mean, save_invstd, dgamma, dbeta, coef);
    return launch_status("bn_bwd_finalize");
}

extern "C" int td_bn_relu_bwd_apply(const void* da, int64_t ldda, int da_coff, const float* y, int dtype,
                                    const float* scale, const float* shift, const float* coef, void* dy,
                                    int64_t pixels, int channels, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(da && y && scale && shift && coef && dy && pixels > 0, "td_bn_relu_bwd_apply: bad args");
    if (int st = check_lanes("td_bn_relu_bwd_apply", dtype, channels)) return st;
    const int V = dtype == TD_BF16 ? 8 : 4;
    const int lanesC = channels / V;
    const int grid = stream_grid(pixels * lanesC, lanesC);
    TD_DISPATCH_T(dtype, (td::launch(bn_relu_bwd_apply_kernel<T>, td::LaunchCfg(grid, kT, 0, (cudaStream_t)stream), 
                             (const T*)da, ldda, da_coff, y, scale, shift, coef, (T*)dy, pixels, channels)));
    return launch_status("bn_relu_bwd_apply");
}

#define TD_DISPATCH_TY(dtype, y_dtype, ...)                                                               \
    if ((y_dtype) == TD_F32) { using Ty = float; TD_DISPATCH_T(dtype, __VA_ARGS__) }                      \
    else if ((y_dtype) == TD_BF16 && (dtype) == TD_BF16) { using Ty = __nv_bfloat16; using T = __nv_bfloat16; __VA_ARGS__; } \
    else { TD_CHECK_ARG(false, "bf16 y needs bf16 activations (dtype %d, y_dtype %d)", (int)(dtype), (int)(y_dtype)); }

extern "C" int td_bn_apply_fused(const void* y, int y_dtype, const float* partials, int nrows, int64_t count, const float* gamma,
                                 const float* beta, const float* conv_bias, float eps, float momentum, float* running_mean,
                                 float* running_var, int64_t* num_batches_tracked, float* scale, float* shift,
                                 float* save_mean, float* save_invstd, void* a, int dtype, int64_t lda, int a_coff,
                                 int64_t pixels, int channels, int relu, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(y && partials && nrows > 0 && count > 0 && gamma && beta && scale && shift && save_mean && save_invstd && a &&
                     pixels > 0, "td_bn_apply_fused: bad args");
    TD_CHECK_ARG(channels % kSlice == 0 && lda % 8 == 0 && a_coff % 8 == 0, "td_bn_apply_fused: channels must be a multiple of 32");
    const int ppp = dtype == TD_BF16 ? kT / 4 : kT / 8;
    const dim3 grid((unsigned)sliced_chunks(pixels, channels, ppp), (unsigned)(channels / kSlice));
    TD_DISPATCH_TY(dtype, y_dtype, (td::launch(bn_apply_fused_kernel<T, Ty>, td::LaunchCfg(grid, kT, 0, (cudaStream_t)stream), (const Ty*)y, partials, nrows,
                             (double)count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var,
                             num_batches_tracked, scale, shift, save_mean, save_invstd, (T*)a, lda, a_coff, pixels, channels, relu)));
    return launch_status("bn_apply_fused");
}

extern "C" int td_bn_bwd_reduce_rows(int dtype, int64_t pixels, int channels) {
    if (channels % kSlice != 0) return 0;
    return sliced_chunks(pixels, channels, dtype == TD_BF16 ? kT / 4 : kT / 8);
}

extern "C" int td_bn_bwd_reduce(const void* da, int64_t ldda, int da_coff, const void* y, int y_dtype, int dtype, const float* scale,
                                const float* beta, const float* save_mean, int64_t pixels, int channels, float* partials,
                                void* stream) {
    TD_REQUIRE_ARCH();
    const float* shift = beta;
    TD_CHECK_ARG(da && y && scale && shift && save_mean && partials && pixels > 0, "td_bn_bwd_reduce: bad args");
    TD_CHECK_ARG(channels % kSlice == 0 && ldda % 8 == 0 && da_coff % 8 == 0, "td_bn_bwd_reduce: channels must be a multiple of 32");
    const dim3 grid((unsigned)td_bn_bwd_reduce_rows(dtype, pixels, channels), (unsigned)(channels / kSlice));
    TD_DISPATCH_TY(dtype, y_dtype, (td::launch(bn_bwd_reduce_sliced_kernel<T, Ty>, td::LaunchCfg(grid, kT, 0, (cudaStream_t)stream), (const T*)da,
                             ldda, da_coff, (const Ty*)y, pixels, channels, scale, shift, save_mean, partials)));
    return launch_status("bn_bwd_reduce");
}

extern "C" int td_bn_bwd_apply_fused(const void* da, int64_t ldda, int da_coff, const void* y, int y_dtype, int dtype,
                                     const float* partials, int nrows, int64_t count, const float* scale, const float* beta,
                                     const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, void* dy,
                                     int64_t pixels, int channels, void* stream) {
    TD_REQUIRE_ARCH();
    const float* shift = beta;
    TD_CHECK_ARG(da && y && partials && nrows > 0 && count > 0 && scale && shift && save_mean && save_invstd && dgamma && dbeta &&
                     dy && pixels > 0, "td_bn_bwd_apply_fused: bad args");
    TD_CHECK_ARG(channels % kSlice == 0 && ldda % 8 == 0 && da_coff % 8 == 0, "td_bn_bwd_apply_fused: channels must be a multiple of 32");
    const int ppp = dtype == TD_BF16 ? kT / 4 : kT / 8;
    const dim3 grid((unsigned)sliced_chunks(pixels, channels, ppp), (unsigned)(channels / kSlice));
    TD_DISPATCH_TY(dtype, y_dtype, (td::launch(bn_bwd_apply_fused_kernel<T, Ty>, td::LaunchCfg(grid, kT, 0, (cudaStream_t)stream), (const T*)da,
                             ldda, da_coff, (const Ty*)y, partials, nrows, (double)count, scale, shift, save_mean, save_invstd, dgamma,
                             dbeta, (T*)dy, pixels, channels)));
    return launch_status("bn_bwd_apply_fused");
}

extern "C" int td_maxpool2_bwd(const void* x, const void* dy, void* dx, int dtype, int batch, int h, int w, int c,
                               int ceil_mode, int accumulate, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && dy && dx && batch > 0 && h > 0 && w > 0 && c > 0, "td_maxpool2_bwd: bad args");
    const int V = dtype == TD_BF16 ? 8 : 4;
    TD_CHECK_ARG(c % V == 0, "td_maxpool2_bwd: channels must be a multiple of %d", V);
    const int ho = ceil_mode ? (h + 1) / 2 : h / 2, wo = ceil_mode ? (w + 1) / 2 : w / 2;
    if (2 * ho >= h && 2 * wo >= w) {            // every input pixel lies in a window (always true for ceil mode / even sizes)
        TD_DISPATCH_T(dtype, (td::launch(maxpool2_bwd_window_kernel<T>, td::LaunchCfg(batch * ho, row_threads((int64_t)wo * c / V), 0,
                                 (cudaStream_t)stream), (const T*)x, (const T*)dy, (T*)dx, batch, h, w, c, ho, wo, accumulate)));
        return launch_status("maxpool2_bwd");
    }
    TD_DISPATCH_T(dtype, (td::launch(maxpool2_bwd_kernel<T>, td::LaunchCfg(batch * h, row_block(c / V), 0, (cudaStream_t)stream), 
                             (const T*)x, (const T*)dy, (T*)dx, batch, h, w, c, ho, wo, accumulate)));
    return launch_status("maxpool2_bwd");
}

extern "C" int td_resize_bilinear_bwd(const void* dy, int64_t ld_dy, int dy_coff, void* dx, int dtype, int batch,
                                      int hi, int wi, int ho, int wo, int c, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dy && dx && batch > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "td_resize_bilinear_bwd: bad args");
    const int V = dtype == TD_BF16 ? 8 : 4;
    TD_CHECK_ARG(c % V == 0 && ld_dy % V == 0 && dy_coff % V == 0, "td_resize_bilinear_bwd: channels must be a multiple of %d", V);
    TD_CHECK_ARG(wo <= kMaxRowW, "td_resize_bilinear_bwd: output rows wider than %d", kMaxRowW);
    TD_DISPATCH_T(dtype, { if (int st = launch_resize_bwd<T>(dy, ld_dy, dy_coff, dx, batch, hi, wi, ho, wo, c, (cudaStream_t)stream)) return st; });
    return launch_status("resize_bilinear_bwd");
}

extern "C" int td_upcat_bwd(const void* dout, void* dlow, void* dskip, float* dtemb, int ld_temb, int temb_off,
                            int dtype, int batch, int ho, int wo, int cu, int hs, int ws, int cs, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dout && dlow && dskip && dtemb, "td_upcat_bwd: null pointer");
    TD_CHECK_ARG(batch > 0 && ho > 0 && wo > 0 && ho % 2 == 0 && wo % 2 == 0, "td_upcat_bwd: bad output size");
    const int V = dtype == TD_BF16 ? 8 : 4;
    TD_CHECK_ARG(cu % V == 0 && cs % V == 0, "td_upcat_bwd: channel counts must be multiples of %d", V);
    if (int st = check_lanes("td_upcat_bwd", dtype, cs)) return st;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ld = cu + cs;
    TD_CHECK_ARG(wo <= kMaxRowW, "td_upcat_bwd: output rows wider than %d", kMaxRowW);
    TD_DISPATCH_T(dtype, { if (int st = launch_resize_bwd<T>(dout, ld, 0, dlow, batch, ho / 2, wo / 2, ho, wo, cu, s)) return st; });
    TD_DISPATCH_T(dtype, { if (int st = launch_resize_bwd<T>(dout, ld, cu, dskip, batch, hs, ws, ho, wo, cs, s)) return st; });
    {
        const dim3 grd((unsigned)batch, (unsigned)ceil_div(cs / V, 8), 1);
        // bilinear weights sum to one per output pixel, so the transposed resize preserves the per-channel total:
        // sum over the output pixels of d_out == sum over the input pixels of d_skip -- read the smaller, dense tensor
        // that the launch above has just written (L2-resident) instead of the strided skip half of d_out a second time
        TD_DISPATCH_T(dtype, (td::launch(temb_bwd_kernel<T>, td::LaunchCfg(grd, kT, 0, s), (const T*)dskip, (int64_t)cs, 0, dtemb, ld_temb, temb_off, hs * ws, cs)));
    }
    return launch_status("upcat_bwd");
}

extern "C" int td_partial_sum(const float* partials, int nrows, int channels, int which, float* out, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(partials && out && nrows > 0 && channels > 0 && (which == 0 || which == 1), "td_partial_sum: bad args");
    td::launch(partial_sum_kernel, td::LaunchCfg((channels + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream), partials, nrows, channels, which, out);
    return launch_status("partial_sum");
}

extern "C" int td_nchw_chansum(const float* x, int batch, int channels, int hw, float* out, float* workspace,
                               void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && out && workspace && batch > 0 && channels > 0 && hw > 0, "td_nchw_chansum: bad args");
    cudaStream_t s = (cudaStream_t)stream;
    td::launch(nchw_chansum_kernel, td::LaunchCfg(dim3(kChanSumChunks, channels), kT, 0, s), x, batch, channels, hw, workspace);
    td::launch(nchw_chansum_finalize_kernel, td::LaunchCfg((channels + 63) / 64, 64, 0, s), workspace, kChanSumChunks, channels, out);
    return launch_status("nchw_chansum");
}
