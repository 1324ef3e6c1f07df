// fp32-accumulate CUDA-core convolutions:
//   * conv3x3_simt_kernel  : generic implicit GEMM, fp32 FFMA.  This is the fp32 parity path
//                            (north_star: eps within 1e-4 of the reference in fp32) and covers any
//                            channel count.
//   * conv3x3_first_kernel : tiny-Cin direct conv (network input, 1 or 4 channels) - bandwidth bound.
//   * conv3x3_last_kernel  : tiny-Cout direct conv (network output) - bandwidth bound.
// The tensor-core engine lives in conv_tc.cu.
#include "conv_plan.h"

namespace td {

// generic element loads with optional NCHW addressing (network boundary tensors only)
template <typename T>
__device__ inline float load_in(const T* __restrict__ x, const td_conv3x3_desc& d, int b, int h, int w, int c) {
    if (d.x_nchw) return to_f32(x[(((int64_t)b * d.cin + c) * d.height + h) * d.width + w]);
    return to_f32(x[(((int64_t)b * d.height + h) * d.width + w) * d.ldx + d.x_coff + c]);
}

// ---------------------------------------------------------------------------------------------
// Generic implicit GEMM: M = B*H*W pixels, N = Cout, K = 9*Cin (k = tap*Cin + c).
// 64x64 tile, BK = 16, 256 threads, 4x4 register tile per thread.
// ---------------------------------------------------------------------------------------------
constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16, SM_THREADS = 256;

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SM_THREADS)
conv3x3_simt_kernel(const td_conv3x3_desc d) {
    td::pdl_sync();
    __shared__ float As[SM_BK][SM_BM + 4];
    __shared__ float Bs[SM_BK][SM_BN + 4];
    const Tin* __restrict__ x = reinterpret_cast<const Tin*>(d.x);
    const float* __restrict__ wt = reinterpret_cast<const float*>(d.w);
    const int64_t M = (int64_t)d.batch * d.height * d.width;
    const int K = 9 * d.cin;
    const int64_t m0 = (int64_t)blockIdx.x * SM_BM;
    const int n0 = blockIdx.y * SM_BN;
    const int tid = threadIdx.x;

    // loader mapping: row = tid / 4, 4 consecutive k per thread
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int64_t pm = m0 + lrow;
    int pb = 0, ph = 0, pw = 0;
    const bool pvalid = pm < M;
    if (pvalid) {
        pw = (int)(pm % d.width);
        int64_t r = pm / d.width;
        ph = (int)(r % d.height);
        pb = (int)(r / d.height);
    }
    const int wn = n0 + lrow;   // weight row (cout) this thread loads
    const bool wvalid = wn < d.cout;

    const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SM_BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float av = 0.f, bv = 0.f;
            if (k < K) {
                const int tap = k / d.cin, c = k - tap * d.cin;
                const int hh = ph + tap / 3 - 1, ww = pw + tap % 3 - 1;
                if (pvalid && hh >= 0 && hh < d.height && ww >= 0 && ww < d.width) av = load_in<Tin>(x, d, pb, hh, ww, c);
                if (wvalid) bv = wt[(int64_t)wn * K + k];
            }
            As[lk + q][lrow] = av;
            Bs[lk + q][lrow] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SM_BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    Tout* __restrict__ y = reinterpret_cast<Tout*>(d.y);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int w_ = (int)(m % d.width);
        const int64_t r = m / d.width;
        const int h_ = (int)(r % d.height);
        const int b_ = (int)(r / d.height);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= d.cout) continue;
            float v = acc[i][j];
            if (d.scale) v *= d.scale[n];
            if (d.shift) v += d.shift[n];
            if (d.relu) v = fmaxf(v, 0.f);
            if (d.y_nchw) y[(((int64_t)b_ * d.cout + n) * d.height + h_) * d.width + w_] = from_f32<Tout>(v);
            else y[m * d.ldy + d.y_coff + n] = from_f32<Tout>(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tiny-Cin direct conv: one thread produces 8 consecutive output channels of one pixel.
// Weights (cout*9*cin fp32, OHWI) are staged in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const td_conv3x3_desc d) {
    td::pdl_sync();
    extern __shared__ float wsm[];          // [cout][9*cin]
    const int K = 9 * d.cin;
    for (int i = threadIdx.x; i < d.cout * K; i += blockDim.x) wsm[i] = reinterpret_cast<const float*>(d.w)[i];
    __syncthreads();
    const Tin* __restrict__ x = reinterpret_cast<const Tin*>(d.x);
    Tout* __restrict__ y = reinterpret_cast<Tout*>(d.y);
    const int groups = d.cout / 8;
    const int64_t total = (int64_t)d.batch * d.height * d.width * groups;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total; i += gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const uint32_t m = i / groups;
        const int w_ = (int)(m % d.width);
        const uint32_t r = m / d.width;
        const int h_ = (int)(r % d.height);
        const int b_ = (int)(r / d.height);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int hh = h_ + tap / 3 - 1, ww = w_ + tap % 3 - 1;
            if (hh < 0 || hh >= d.height || ww < 0 || ww >= d.width) continue;
            for (int c = 0; c < d.cin; ++c) {
                const float xv = load_in<Tin>(x, d, b_, hh, ww, c);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wsm[(g * 8 + j) * K + tap * d.cin + c], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = g * 8 + j;
            float v = acc[j];
            if (d.scale) v *= d.scale[n];
            if (d.shift) v += d.shift[n];
            if (d.relu) v = fmaxf(v, 0.f);
            acc[j] = v;
        }
        Tout* dst = y + (int64_t)m * d.ldy + d.y_coff + g * 8;
        if constexpr (sizeof(Tout) == 2) {
            Vec<__nv_bfloat16>::pack(acc).store(reinterpret_cast<__nv_bfloat16*>(dst));
        } else {
            Vec<float>::pack(acc).store(reinterpret_cast<float*>(dst));
            Vec<float>::pack(acc + 4).store(reinterpret_cast<float*>(dst) + 4);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tiny-Cout direct conv: L = cin / V lanes cooperate on one pixel (each lane owns one 16-byte
// channel vector per tap, so a pixel row is one coalesced read), shuffle-reduce, lane 0 stores.
// ---------------------------------------------------------------------------------------------
template <typename Tin, int COUT>
__global__ void __launch_bounds__(256)
conv3x3_last_kernel(const td_conv3x3_desc d, int lanes_per_pixel) {
    td::pdl_sync();
    extern __shared__ float wsm[];          // [COUT][9*cin]
    constexpr int V = Vec<Tin>::N;
    const int K = 9 * d.cin;
    for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) wsm[i] = reinterpret_cast<const float*>(d.w)[i];
    __syncthreads();
    const Tin* __restrict__ x = reinterpret_cast<const Tin*>(d.x);
    float* __restrict__ y = reinterpret_cast<float*>(d.y);
    const int L = lanes_per_pixel;
    const int64_t M = (int64_t)d.batch * d.height * d.width;
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    // iterate so that whole warps stay converged for the shuffles
    const int64_t iters = ceil_div(M * L, gthreads);
    for (int64_t it = 0; it < iters; ++it) {
        const uint32_t i = (uint32_t)(it * gthreads) + blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t m = i / (uint32_t)L;
        const int lane = (int)(i % (uint32_t)L);
        const bool valid = m < M;
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
        if (valid) {
            const int w_ = (int)(m % d.width);
            const uint32_t r = m / d.width;
            const int h_ = (int)(r % d.height);
            const int b_ = (int)(r / d.height);
            for (int tap = 0; tap < 9; ++tap) {
                const int hh = h_ + tap / 3 - 1, ww = w_ + tap % 3 - 1;
                if (hh < 0 || hh >= d.height || ww < 0 || ww >= d.width) continue;
                const Tin* px = x + (((int64_t)b_ * d.height + hh) * d.width + ww) * d.ldx + d.x_coff;
                for (int cv = lane; cv * V < d.cin; cv += L) {
                    float f[V];
                    Vec<Tin>::load(px + cv * V).unpack(f);
#pragma unroll
                    for (int j = 0; j < COUT; ++j) {
                        const float* wr = wsm + j * K + tap * d.cin + cv * V;
#pragma unroll
                        for (int k = 0; k < V; ++k) acc[j] = fmaf(f[k], wr[k], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < COUT; ++j)
            for (int o = L >> 1; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (valid && lane == 0) {
            const int w_ = (int)(m % d.width);
            const uint32_t r = m / d.width;
            const int h_ = (int)(r % d.height);
            const int b_ = (int)(r / d.height);
#pragma unroll
            for (int j = 0; j < COUT; ++j) {
                float v = acc[j];
                if (d.scale) v *= d.scale[j];
                if (d.shift) v += d.shift[j];
                if (d.relu) v = fmaxf(v, 0.f);
                if (d.y_nchw) y[(((int64_t)b_ * COUT + j) * d.height + h_) * d.width + w_] = v;
                else y[(int64_t)m * d.ldy + d.y_coff + j] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Boundary layers of the latent UNet (4 latent channels, diffusion_laion.py): the two generic kernels above spend their
// time on shared-memory weight loads (one per FMA) and on re-reading every pixel nine times.
//
// conv3x3_firstn_kernel: NCHW fp32 input with CIN channels -> NHWC channels, COUT of them per CTA (blockIdx.y selects the
// slice; also the data gradient of final_conv, 4 -> 64).  One thread owns two horizontally adjacent pixels and COUT outputs: a broadcast 16-byte weight load feeds 8 FMAs, the 3x4 input window of one
// channel lives in registers.
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT, typename Tout>
__global__ void __launch_bounds__(256, 2)
conv3x3_firstn_kernel(const td_conv3x3_desc d) {
    td::pdl_sync();
    __shared__ __align__(16) float wsm[9 * CIN][COUT];       // [tap * CIN + c][cout]
    __shared__ float ssc[COUT], ssh[COUT];
    const int n0 = blockIdx.y * COUT;                          // this CTA's slice of the output channels
    for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) {
        const int k = i / COUT, n = i - k * COUT;              // source: [cout][9][cin] (4.6 KB, L1/L2 resident)
        wsm[k][n] = __ldg(reinterpret_cast<const float*>(d.w) + (n0 + n) * (9 * CIN) + k);
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
        ssc[i] = d.scale ? d.scale[n0 + i] : 1.f;
        ssh[i] = d.shift ? d.shift[n0 + i] : 0.f;
    }
    __syncthreads();
    const float* __restrict__ x = reinterpret_cast<const float*>(d.x);
    const int H = d.height, W = d.width, pairs = (W + 1) / 2;
    const int items = d.batch * H * pairs;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int row = it / pairs, x0 = (it - row * pairs) * 2;
        const int b = row / H, h = row - b * H;
        float acc[2][COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[0][j] = acc[1][j] = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float* xc = x + ((int64_t)b * CIN + c) * H * W;
            float xv[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int hh = h + r - 1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ww = x0 + q - 1;
                    xv[r][q] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xc + hh * W + ww) : 0.f;
                }
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float a0 = xv[t / 3][t % 3], a1 = xv[t / 3][t % 3 + 1];
#pragma unroll
                for (int j = 0; j < COUT; j += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&wsm[t * CIN + c][j]);
                    acc[0][j] = fmaf(a0, wv.x, acc[0][j]);         acc[1][j] = fmaf(a1, wv.x, acc[1][j]);
                    acc[0][j + 1] = fmaf(a0, wv.y, acc[0][j + 1]); acc[1][j + 1] = fmaf(a1, wv.y, acc[1][j + 1]);
                    acc[0][j + 2] = fmaf(a0, wv.z, acc[0][j + 2]); acc[1][j + 2] = fmaf(a1, wv.z, acc[1][j + 2]);
                    acc[0][j + 3] = fmaf(a0, wv.w, acc[0][j + 3]); acc[1][j + 3] = fmaf(a1, wv.w, acc[1][j + 3]);
                }
            }
        }
        Tout* yrow = reinterpret_cast<Tout*>(d.y) + ((int64_t)row * W + x0) * d.ldy + d.y_coff + n0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            if (x0 + p >= W) break;
#pragma unroll
            for (int j = 0; j < COUT; ++j) {
                const float v = fmaf(acc[p][j], ssc[j], ssh[j]);
                acc[p][j] = d.relu ? fmaxf(v, 0.f) : v;
            }
            Tout* dst = yrow + (int64_t)p * d.ldy;
#pragma unroll
            for (int j = 0; j < COUT; j += 8) {
                if constexpr (sizeof(Tout) == 2) {
                    Vec<__nv_bfloat16>::pack(acc[p] + j).store(reinterpret_cast<__nv_bfloat16*>(dst) + j);
                } else {
                    Vec<float>::pack(acc[p] + j).store(reinterpret_cast<float*>(dst) + j);
                    Vec<float>::pack(acc[p] + j + 4).store(reinterpret_cast<float*>(dst) + j + 4);
                }
            }
        }
    }
}

// Second half of the contract-then-stencil kernels: out[j][r][w] = sum_t e[j * 9 + t][(r + dy_t, w + dx_t)] over the band's rows
// [r0, r1); plane row 0 is image row q0.  One thread per pixel and pass: the index arithmetic and the border masks are shared
// by the 9 * COUT shared-memory reads.
template <int COUT>
__device__ inline void lastn_stencil(const td_conv3x3_desc& d, const float* __restrict__ e, int plane, int b, int r0, int r1, int q0) {
    const int H = d.height, W = d.width, nr = r1 - r0;
    float* __restrict__ y = reinterpret_cast<float*>(d.y);
    float sc[COUT], sh[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        sc[j] = d.scale ? __ldg(d.scale + j) : 1.f;
        sh[j] = d.shift ? __ldg(d.shift + j) : 0.f;
    }
    for (int o = threadIdx.x; o < nr * W; o += blockDim.x) {
        const int rr = o / W, w_ = o - rr * W, r = r0 + rr;
        const float* ec = e + (r - q0) * W + w_;
        bool ok[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hh = r + t / 3 - 1, ww = w_ + t % 3 - 1;
            ok[t] = hh >= 0 && hh < H && ww >= 0 && ww < W;
        }
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (ok[t]) acc += ec[(j * 9 + t) * plane + (t / 3 - 1) * W + (t % 3 - 1)];
            acc = fmaf(acc, sc[j], sh[j]);
            if (d.relu) acc = fmaxf(acc, 0.f);
            if (d.y_nchw) y[(((int64_t)b * COUT + j) * H + r) * W + w_] = acc;
            else y[(((int64_t)b * H + r) * W + w_) * d.ldy + d.y_coff + j] = acc;
        }
    }
}

// conv3x3_lastn_kernel: NHWC input with cin channels -> COUT fp32 channels, contract-then-stencil.  The channel
// contraction commutes with the spatial shift: e[j][t][s] = sum_c x[s, c] w[j, t, c] is computed once per source pixel
// (the cin-channel vector is read exactly once; one thread owns two pixels and all 9 * COUT partials, weights are
// broadcast 16-byte shared-memory loads feeding 8 FMAs each), parked in shared memory for a band of rows plus its
// one-row halo, and the 3x3 stencil then adds nine scalars per output.
// ---------------------------------------------------------------------------------------------
template <typename Tin, int COUT>
__global__ void __launch_bounds__(512)
conv3x3_lastn_kernel(const td_conv3x3_desc d, int BH, int bands) {
    td::pdl_sync();
    constexpr int V = Vec<Tin>::N, Q = 9 * COUT;
    extern __shared__ __align__(16) float lsm[];
    const int H = d.height, W = d.width, cin = d.cin;
    float* wsm = lsm;                                    // [Q][cin]
    float* e = lsm + Q * cin;                            // [Q][(BH + 2) * W]
    const int plane = (BH + 2) * W;
    for (int i = threadIdx.x; i < Q * cin; i += blockDim.x) wsm[i] = reinterpret_cast<const float*>(d.w)[i];
    __syncthreads();
    const int b = blockIdx.x / bands, band = blockIdx.x - b * bands;
    const int r0 = band * BH, r1 = min(r0 + BH, H);
    const int q0 = max(r0 - 1, 0), q1 = min(r1 + 1, H);
    const int npix = (q1 - q0) * W;
    const Tin* __restrict__ xb = reinterpret_cast<const Tin*>(d.x) + ((int64_t)b * H + q0) * W * d.ldx + d.x_coff;
    for (int item = threadIdx.x; item * 2 < npix; item += blockDim.x) {
        const int p0 = item * 2;
        const bool has1 = p0 + 1 < npix;
        const Tin* x0 = xb + (int64_t)p0 * d.ldx;
        const Tin* x1 = has1 ? x0 + d.ldx : x0;
        float a0[Q], a1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) a0[q] = a1[q] = 0.f;
        for (int c = 0; c < cin; c += V) {
            float f0[V], f1[V];
            Vec<Tin>::load(x0 + c).unpack(f0);
            Vec<Tin>::load(x1 + c).unpack(f1);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int k = 0; k < V; k += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wsm + q * cin + c + k);
                    a0[q] = fmaf(f0[k], wv.x, a0[q]);     a1[q] = fmaf(f1[k], wv.x, a1[q]);
                    a0[q] = fmaf(f0[k + 1], wv.y, a0[q]); a1[q] = fmaf(f1[k + 1], wv.y, a1[q]);
                    a0[q] = fmaf(f0[k + 2], wv.z, a0[q]); a1[q] = fmaf(f1[k + 2], wv.z, a1[q]);
                    a0[q] = fmaf(f0[k + 3], wv.w, a0[q]); a1[q] = fmaf(f1[k + 3], wv.w, a1[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            e[q * plane + p0] = a0[q];
            if (has1) e[q * plane + p0 + 1] = a1[q];
        }
    }
    __syncthreads();
    lastn_stencil<COUT>(d, e, plane, b, r0, r1, q0);
}

// bf16 activations: the same contract-then-stencil with the contraction on the tensor cores.  It is a [pixels x cin] x [cin x 36]
// product -- far too narrow for a tcgen05 tile (N = 36) but a natural warp-level mma.sync.m16n8k16: one warp owns 16 pixels, A
// fragments come straight from global memory (every 128-byte channel vector is read exactly once), the weights sit in shared
// memory as a bf16 hi + lo pair (w = hi + lo to 2^-17, so the layer keeps its fp32 weights: two MMAs per fragment, still
// 40x fewer issue slots than the FFMA loop), fp32 accumulators go to the shared-memory planes of the stencil.
__device__ inline void mma_bf16_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int COUT, int KS>          // KS = cin / 16
__global__ void __launch_bounds__(512)
conv3x3_lastn_mma_kernel(const td_conv3x3_desc d, int BH, int bands, int plane) {
    td::pdl_sync();
    constexpr int Q = 9 * COUT, NT = (Q + 7) / 8, CIN = KS * 16, WS = CIN + 8;      // WS: padded row (conflict-free fragment loads)
    extern __shared__ __align__(16) unsigned char lraw[];
    __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(lraw);      // [NT * 8][WS]
    __nv_bfloat16* wlo = whi + NT * 8 * WS;
    float* e = reinterpret_cast<float*>(wlo + NT * 8 * WS);           // [Q][plane]
    const int H = d.height, W = d.width;
    // The contraction does not care which physical channel sits at which logical k, so the channels are permuted to make the A
    // fragments of TWO k-steps one 16-byte load per thread: thread t4 of a pixel row reads channels 32 kp + 8 t4 + [0, 8); they are
    // the logical k = 2 t4 + {0, 1} and 8 + 2 t4 + {0, 1} of k-steps 2 kp and 2 kp + 1.  The weights are staged in that order.
    for (int i = threadIdx.x; i < NT * 8 * CIN; i += blockDim.x) {
        const int q = i / CIN, l = i - q * CIN;                 // l = logical position ks * 16 + k
        const int ks = l >> 4, k = l & 15;
        const int c = 32 * (ks >> 1) + ((k >> 1) & 3) * 8 + (ks & 1) * 4 + (k >> 3) * 2 + (k & 1);
        const float w = q < Q ? __ldg(reinterpret_cast<const float*>(d.w) + q * CIN + c) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        whi[q * WS + l] = hi;
        wlo[q * WS + l] = __float2bfloat16_rn(w - __bfloat162float(hi));
    }
    __syncthreads();
    const int b = blockIdx.x / bands, band = blockIdx.x - b * bands;
    const int r0 = band * BH, r1 = min(r0 + BH, H);
    const int q0 = max(r0 - 1, 0), q1 = min(r1 + 1, H);
    const int npix = (q1 - q0) * W;
    const __nv_bfloat16* __restrict__ xb =
        reinterpret_cast<const __nv_bfloat16*>(d.x) + ((int64_t)b * H + q0) * W * d.ldx + d.x_coff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, t4 = lane & 3, nwarps = blockDim.x >> 5;
    const int mtiles = (npix + 15) >> 4;
    constexpr int MT = 8 / KS;                     // 16-pixel tiles per pass: all of their loads are in flight before the first MMA
    // ldmatrix.x4 row address of this lane: matrices {hi k0.., hi k0+8.., lo k0.., lo k0+8..} of one (n-tile, k-step)
    const uint32_t bfrag = (uint32_t)__cvta_generic_to_shared(((lane & 16) ? wlo : whi) + (lane & 7) * WS + ((lane >> 3) & 1) * 8);
    for (int mt0 = warp * MT; mt0 < mtiles; mt0 += nwarps * MT) {
        uint32_t a[MT][KS][4];
#pragma unroll
        for (int u = 0; u < MT; ++u) {
            const int p_lo = (mt0 + u) * 16 + gid, p_hi = p_lo + 8;
            const __nv_bfloat16* xl = xb + (int64_t)min(p_lo, npix - 1) * d.ldx + t4 * 8;
            const __nv_bfloat16* xh = xb + (int64_t)min(p_hi, npix - 1) * d.ldx + t4 * 8;
#pragma unroll
            for (int kp = 0; kp < KS / 2; ++kp) {
                const uint4 vl = __ldg(reinterpret_cast<const uint4*>(xl + kp * 32));
                const uint4 vh = __ldg(reinterpret_cast<const uint4*>(xh + kp * 32));
                a[u][2 * kp][0] = vl.x; a[u][2 * kp][2] = vl.y; a[u][2 * kp + 1][0] = vl.z; a[u][2 * kp + 1][2] = vl.w;
                a[u][2 * kp][1] = vh.x; a[u][2 * kp][3] = vh.y; a[u][2 * kp + 1][1] = vh.z; a[u][2 * kp + 1][3] = vh.w;
            }
        }
        float acc[MT][NT][4];
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[u][nt][0] = acc[u][nt][1] = acc[u][nt][2] = acc[u][nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                uint32_t bh0, bh1, bl0, bl1;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(bh0), "=r"(bh1), "=r"(bl0), "=r"(bl1) : "r"(bfrag + (uint32_t)((nt * 8 * WS + ks * 16) * 2)));
#pragma unroll
                for (int u = 0; u < MT; ++u) {
                    mma_bf16_m16n8k16(acc[u][nt], a[u][ks], bh0, bh1);
                    mma_bf16_m16n8k16(acc[u][nt], a[u][ks], bl0, bl1);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < MT; ++u) {
            if (mt0 + u >= mtiles) break;
            const int p_lo = (mt0 + u) * 16 + gid, p_hi = p_lo + 8;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int q = nt * 8 + t4 * 2;             // Q is even: q < Q implies q + 1 < Q
                if (q < Q) {
                    if (p_lo < npix) { e[q * plane + p_lo] = acc[u][nt][0]; e[(q + 1) * plane + p_lo] = acc[u][nt][1]; }
                    if (p_hi < npix) { e[q * plane + p_hi] = acc[u][nt][2]; e[(q + 1) * plane + p_hi] = acc[u][nt][3]; }
                }
            }
        }
    }
    __syncthreads();
    lastn_stencil<COUT>(d, e, plane, b, r0, r1, q0);
}

// Rows per band of conv3x3_lastn_kernel (0: the geometry does not fit shared memory -> generic kernel).  A whole image
// per CTA when it fits (no halo rows computed twice): 32 x 32 x 36 partials = 144 KB for the latent UNet.
constexpr int LASTN_SMEM = 200 * 1024;
static int lastn_band_rows(const td_conv3x3_desc& d) {
    const int64_t wbytes = (int64_t)d.cout * 9 * d.cin * 4, row = (int64_t)d.width * 9 * d.cout * 4;
    const int64_t max_rows = (LASTN_SMEM - wbytes) / row;
    if (max_rows < 3) return 0;
    const int bh = (int)std::min<int64_t>(d.height, max_rows - 2);
    const int bands = (int)ceil_div(d.height, bh);
    return (int)ceil_div(d.height, bands);
}

// ---------------------------------------------------------------------------------------------
// Fast paths of the two network-boundary layers of the MNIST UNets (the generic kernels above were
// instruction-bound: a shared-memory weight load per FMA and 64-bit index math per element).
// One CTA per image row (b, h); weights live in registers; no integer division per element.
// ---------------------------------------------------------------------------------------------
// initial_conv with Cin == 1: threadIdx.x = group of 8 output channels, threadIdx.y = chunk of FP pixels of
// a row; each thread keeps its 8x9 weights in registers and produces FP consecutive pixels per pass
// (3 x (FP+2) input loads shared by the FP pixels).
constexpr int FP = 4;
template <typename Tout>
__global__ void __launch_bounds__(256, 2)
conv3x3_first1_kernel(const td_conv3x3_desc d) {
    td::pdl_sync();
    const int g = threadIdx.x;
    const float* __restrict__ wt = reinterpret_cast<const float*>(d.w);          // [cout][9]
    float w[8][9], sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int t = 0; t < 9; ++t) w[j][t] = __ldg(wt + (g * 8 + j) * 9 + t);
        sc[j] = d.scale ? d.scale[g * 8 + j] : 1.f;
        sh[j] = d.shift ? d.shift[g * 8 + j] : 0.f;
    }
    const int chunks = (d.width + FP - 1) / FP;
    const int items = d.batch * d.height * chunks;           // (row, chunk) work items
    // NCHW with one channel == NHWC with one channel
    for (int it = blockIdx.x * blockDim.y + threadIdx.y; it < items; it += gridDim.x * blockDim.y) {
        const int row = it / chunks, x0 = (it - row * chunks) * FP;
        const int b = row / d.height, h = row - b * d.height;
        const float* xb = reinterpret_cast<const float*>(d.x) + (int64_t)b * d.height * d.width;
        float xv[3][FP + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hh = h + r - 1;
#pragma unroll
            for (int q = 0; q < FP + 2; ++q) {
                const int ww = x0 + q - 1;
                xv[r][q] = (hh >= 0 && hh < d.height && ww >= 0 && ww < d.width) ? __ldg(xb + hh * d.width + ww) : 0.f;
            }
        }
        Tout* yrow = reinterpret_cast<Tout*>(d.y) + ((int64_t)b * d.height + h) * d.width * d.ldy + d.y_coff + g * 8;
#pragma unroll
        for (int q = 0; q < FP; ++q) {
            if (x0 + q >= d.width) break;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) a = fmaf(xv[t / 3][q + t % 3], w[j][t], a);
                a = fmaf(a, sc[j], sh[j]);
                acc[j] = d.relu ? fmaxf(a, 0.f) : a;
            }
            Tout* dst = yrow + (int64_t)(x0 + q) * d.ldy;
            if constexpr (sizeof(Tout) == 2) {
                Vec<__nv_bfloat16>::pack(acc).store(reinterpret_cast<__nv_bfloat16*>(dst));
            } else {
                Vec<float>::pack(acc).store(reinterpret_cast<float*>(dst));
                Vec<float>::pack(acc + 4).store(reinterpret_cast<float*>(dst) + 4);
            }
        }
    }
}

// final_conv with Cout == 1: threadIdx.x = lane owning V input channels (L lanes per pixel, L a power of
// two <= 32), threadIdx.y = pixel slot; every thread handles LP pixels per pass (independent load chains),
// shuffle-reduce over the L lanes.
constexpr int LP = 2;
template <typename Tin>
__global__ void __launch_bounds__(256)
conv3x3_last1_kernel(const td_conv3x3_desc d) {
    td::pdl_sync();
    constexpr int V = Vec<Tin>::N;
    const int L = blockDim.x, lane = threadIdx.x;
    const float* __restrict__ wt = reinterpret_cast<const float*>(d.w);          // [1][9][cin]
    float w[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < V; ++k) w[t][k] = __ldg(wt + t * d.cin + lane * V + k);
    const float sc = d.scale ? d.scale[0] : 1.f, sh = d.shift ? d.shift[0] : 0.f;
    const int HW = d.height * d.width;
    const int total = d.batch * HW;
    const int per_pass = gridDim.x * blockDim.y * LP;
    const Tin* xall = reinterpret_cast<const Tin*>(d.x) + d.x_coff + lane * V;
    float* yall = reinterpret_cast<float*>(d.y);
    for (int base = 0; base < total; base += per_pass) {           // uniform trip count: shuffles stay converged
        float acc[LP];
        int pix[LP];
#pragma unroll
        for (int u = 0; u < LP; ++u) {
            pix[u] = base + (blockIdx.x * blockDim.y + threadIdx.y) * LP + u;
            acc[u] = 0.f;
            if (pix[u] < total) {
                const int b = pix[u] / HW, rem = pix[u] - b * HW;
                const int h = rem / d.width, x0 = rem - h * d.width;
                const Tin* xb = xall + (int64_t)b * HW * d.ldx;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int hh = h + t / 3 - 1, ww = x0 + t % 3 - 1;
                    if (hh < 0 || hh >= d.height || ww < 0 || ww >= d.width) continue;
                    float f[V];
                    Vec<Tin>::load(xb + ((int64_t)hh * d.width + ww) * d.ldx).unpack(f);
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[u] = fmaf(f[k], w[t][k], acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < LP; ++u) {
            for (int o = L >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            if (pix[u] < total && lane == 0) {
                float v = fmaf(acc[u], sc, sh);
                if (d.relu) v = fmaxf(v, 0.f);
                // NCHW with one channel: flat pixel index; NHWC: pixel * ldy + coff
                yall[d.y_nchw ? (int64_t)pix[u] : (int64_t)pix[u] * d.ldy + d.y_coff] = v;
            }
        }
    }
}

static int run_simt(const td_conv_plan* p, cudaStream_t s) {
    const td_conv3x3_desc& d = p->d;
    const int64_t M = (int64_t)d.batch * d.height * d.width;
    dim3 grid((unsigned)ceil_div(M, SM_BM), (unsigned)ceil_div(d.cout, SM_BN));
    if (d.x_dtype == TD_F32 && d.y_dtype == TD_F32) td::launch(conv3x3_simt_kernel<float, float>, td::LaunchCfg(grid, SM_THREADS, 0, s), d);
    else if (d.x_dtype == TD_BF16 && d.y_dtype == TD_BF16) td::launch(conv3x3_simt_kernel<__nv_bfloat16, __nv_bfloat16>, td::LaunchCfg(grid, SM_THREADS, 0, s), d);
    else if (d.x_dtype == TD_F32 && d.y_dtype == TD_BF16) td::launch(conv3x3_simt_kernel<float, __nv_bfloat16>, td::LaunchCfg(grid, SM_THREADS, 0, s), d);
    else td::launch(conv3x3_simt_kernel<__nv_bfloat16, float>, td::LaunchCfg(grid, SM_THREADS, 0, s), d);
    return launch_status("conv3x3_simt");
}

static int run_direct(const td_conv_plan* p, cudaStream_t s) {
    const td_conv3x3_desc& d = p->d;
    const int64_t M = (int64_t)d.batch * d.height * d.width;
    if (d.cin == 1 && d.x_dtype == TD_F32 && d.cout / 8 <= 256 && d.width * (int64_t)d.height < (1 << 30)) {
        const int groups = d.cout / 8;
        dim3 block((unsigned)groups, (unsigned)std::max(1, 256 / groups));
        const int items = d.batch * d.height * ((d.width + FP - 1) / FP);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, block.y), kNumSMs * 4));
        if (d.y_dtype == TD_BF16) td::launch(conv3x3_first1_kernel<__nv_bfloat16>, td::LaunchCfg(grid, block, 0, s), d);
        else td::launch(conv3x3_first1_kernel<float>, td::LaunchCfg(grid, block, 0, s), d);
        return launch_status("conv3x3_first1");
    }
    if (d.cin == 4 && d.cout % 32 == 0 && d.cout <= 32 * 65535 && d.x_nchw && d.x_dtype == TD_F32 &&
        (int64_t)d.batch * d.height * d.width < (1 << 30)) {
        const int items = d.batch * d.height * ((d.width + 1) / 2);
        const dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(items, 256)), (unsigned)(d.cout / 32));
        if (d.y_dtype == TD_BF16) td::launch(conv3x3_firstn_kernel<4, 32, __nv_bfloat16>, td::LaunchCfg(grid, 256, 0, s), d);
        else td::launch(conv3x3_firstn_kernel<4, 32, float>, td::LaunchCfg(grid, 256, 0, s), d);
        return launch_status("conv3x3_firstn");
    }
    if (d.cin <= 8) {
        const size_t smem = (size_t)d.cout * 9 * d.cin * sizeof(float);
        const int64_t items = M * (d.cout / 8);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), kNumSMs * 8));
        if (d.x_dtype == TD_F32 && d.y_dtype == TD_BF16) td::launch(conv3x3_first_kernel<float, __nv_bfloat16>, td::LaunchCfg(grid, 256, smem, s), d);
        else if (d.x_dtype == TD_F32 && d.y_dtype == TD_F32) td::launch(conv3x3_first_kernel<float, float>, td::LaunchCfg(grid, 256, smem, s), d);
        else { set_error("direct first conv: unsupported dtype combination"); return TD_ERR_UNSUPPORTED; }
        return launch_status("conv3x3_first");
    }
    // tiny cout
    const int V = d.x_dtype == TD_BF16 ? 8 : 4;
    if (d.cout == 1 && d.cin / V <= 32) {
        const int Ln = d.cin / V;            // power of two (checked at plan creation)
        dim3 block((unsigned)Ln, (unsigned)(256 / Ln));
        const int64_t total = (int64_t)d.batch * d.height * d.width;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, (int64_t)block.y * LP), kNumSMs * 8));
        if (d.x_dtype == TD_BF16) td::launch(conv3x3_last1_kernel<__nv_bfloat16>, td::LaunchCfg(grid, block, 0, s), d);
        else td::launch(conv3x3_last1_kernel<float>, td::LaunchCfg(grid, block, 0, s), d);
        return launch_status("conv3x3_last1");
    }
    if (d.cout == 4 && d.cin % 8 == 0) {
        const int bh = lastn_band_rows(d);
        if (bh > 0 && d.x_dtype == TD_BF16 && (d.cin == 64 || d.cin == 128)) {
            const int bands = (int)ceil_div(d.height, bh);
            const int npix = std::min(bh + 2, d.height) * d.width;
            const int plane = (npix + 15) / 16 * 16 + 4;                // == 4 mod 16: conflict-free accumulator stores
            const size_t smem = (size_t)2 * 40 * (d.cin + 8) * 2 + (size_t)36 * plane * sizeof(float);
            const int threads = std::min(512, std::max(64, (npix + 15) / 16 * 32));      // at most one warp per 16 pixels
            static bool attr_mma = false;
            if (!attr_mma) {
                cudaFuncSetAttribute(conv3x3_lastn_mma_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                cudaFuncSetAttribute(conv3x3_lastn_mma_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                attr_mma = true;
            }
            if (smem <= (size_t)227 * 1024) {
                if (d.cin == 64) td::launch(conv3x3_lastn_mma_kernel<4, 4>, td::LaunchCfg(d.batch * bands, threads, smem, s), d, bh, bands, plane);
                else td::launch(conv3x3_lastn_mma_kernel<4, 8>, td::LaunchCfg(d.batch * bands, threads, smem, s), d, bh, bands, plane);
                return launch_status("conv3x3_lastn_mma");
            }
        }
        if (bh > 0) {
            const int bands = (int)ceil_div(d.height, bh);
            const size_t smem = ((size_t)36 * d.cin + (size_t)36 * (bh + 2) * d.width) * sizeof(float);
            const int grid = d.batch * bands;
            // one pass over the band when its pixel pairs fit one CTA (a second pass would run on a few warps only)
            const int items = (int)ceil_div((int64_t)std::min(bh + 2, d.height) * d.width, 2);
            const int threads = std::min(512, std::max(64, (items + 31) / 32 * 32));
            static bool attr_done = false;
            if (!attr_done) {
                cudaFuncSetAttribute(conv3x3_lastn_kernel<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, LASTN_SMEM);
                cudaFuncSetAttribute(conv3x3_lastn_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, LASTN_SMEM);
                attr_done = true;
            }
            if (d.x_dtype == TD_BF16) td::launch(conv3x3_lastn_kernel<__nv_bfloat16, 4>, td::LaunchCfg(grid, threads, smem, s), d, bh, bands);
            else td::launch(conv3x3_lastn_kernel<float, 4>, td::LaunchCfg(grid, threads, smem, s), d, bh, bands);
            return launch_status("conv3x3_lastn");
        }
    }
    int L = d.cin / V;
    if (L > 32) L = 32;
    const size_t smem = (size_t)d.cout * 9 * d.cin * sizeof(float);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(M * L, 256), kNumSMs * 8));
#define TD_LAST(T, CO) td::launch(conv3x3_last_kernel<T, CO>, td::LaunchCfg(grid, 256, smem, s), d, L)
    if (d.x_dtype == TD_BF16) {
        if (d.cout == 1) TD_LAST(__nv_bfloat16, 1);
        else if (d.cout == 4) TD_LAST(__nv_bfloat16, 4);
        else { set_error("direct last conv: cout must be 1 or 4"); return TD_ERR_UNSUPPORTED; }
    } else {
        if (d.cout == 1) TD_LAST(float, 1);
        else if (d.cout == 4) TD_LAST(float, 4);
        else { set_error("direct last conv: cout must be 1 or 4"); return TD_ERR_UNSUPPORTED; }
    }
#undef TD_LAST
    return launch_status("conv3x3_last");
}

}  // namespace td

using namespace td;

extern "C" int td_conv3x3_plan_create(td_conv_plan** plan, const td_conv3x3_desc* desc, int engine) {
    TD_CHECK_ARG(plan && desc, "td_conv3x3_plan_create: null pointer");
    const td_conv3x3_desc& d = *desc;
    TD_CHECK_ARG(d.batch > 0 && d.height > 0 && d.width > 0 && d.cin > 0 && d.cout > 0, "conv plan: bad sizes");
    TD_CHECK_ARG(d.x && d.y && d.w, "conv plan: null tensor pointer");
    TD_CHECK_ARG((d.x_dtype == TD_F32 || d.x_dtype == TD_BF16) && (d.y_dtype == TD_F32 || d.y_dtype == TD_BF16),
                 "conv plan: bad dtype");
    TD_CHECK_ARG(d.x_nchw || (d.ldx >= d.x_coff + d.cin), "conv plan: ldx too small");
    TD_CHECK_ARG(d.y_nchw || (d.ldy >= d.y_coff + d.cout), "conv plan: ldy too small");
    td_conv_plan* p = new td_conv_plan();
    memset(p, 0, sizeof(*p));
    p->d = d;
    p->engine = engine;
    int st = TD_OK;
    if (engine == TD_CONV_TC) {
        st = tc_plan_init(p);
    } else if (engine == TD_CONV_DIRECT) {
        const bool first = d.cin <= 8 && d.cout % 8 == 0 && d.x_dtype == TD_F32 && !d.y_nchw;
        const int V = d.x_dtype == TD_BF16 ? 8 : 4;
        const int L = d.cin / V;
        const bool last = d.cin > 8 && (d.cout == 1 || d.cout == 4) && d.y_dtype == TD_F32 && !d.x_nchw &&
                          d.cin % V == 0 && (L & (L - 1)) == 0 && d.ldx % V == 0 && d.x_coff % V == 0;
        if (!first && !last) {
            set_error("conv plan: shape not supported by the direct engine (cin=%d cout=%d)", d.cin, d.cout);
            st = TD_ERR_UNSUPPORTED;
        }
        if (first && (d.ldy % 8 != 0 || d.y_coff % 8 != 0)) {
            set_error("conv plan: direct first conv needs ldy, y_coff multiples of 8");
            st = TD_ERR_UNSUPPORTED;
        }
        if ((size_t)d.cout * 9 * d.cin * sizeof(float) > 48 * 1024) {
            set_error("conv plan: direct engine weights exceed 48 KB shared memory");
            st = TD_ERR_UNSUPPORTED;
        }
    } else if (engine == TD_CONV_SIMT) {
        if (d.stats) { set_error("conv plan: SIMT engine does not emit statistics"); st = TD_ERR_UNSUPPORTED; }
    } else {
        set_error("conv plan: unknown engine %d", engine);
        st = TD_ERR_ARG;
    }
    if (st != TD_OK) {
        delete p;
        return st;
    }
    *plan = p;
    return TD_OK;
}

extern "C" int td_conv3x3_pool_fused(const td_conv_plan* plan) {
    return plan && plan->engine == TD_CONV_TC && plan->halo && plan->h_pool ? 1 : 0;
}

extern "C" int td_conv3x3_stats_rows(const td_conv_plan* plan) {
    if (!plan || !plan->d.stats || plan->engine != TD_CONV_TC) return 0;
    if (plan->halo) return plan->h_grid;                              // one row per (persistent) CTA
    return plan->tiles_w * plan->tiles_h * plan->tiles_n;              // per-tap kernel: one row per M tile
}

extern "C" int64_t td_conv3x3_splitk_workspace(const td_conv3x3_desc* d) {
    if (!d || d->cin % 64 != 0 || d->cout % 64 != 0) return 0;
    const int64_t pixels = (int64_t)d->batch * d->height * d->width;
    const int64_t min_ctas = ceil_div(pixels, 128) * ceil_div(d->cout, 256);
    if (2 * min_ctas > kNumSMs) return 0;                               // enough tiles: never split
    const int64_t splits = std::min<int64_t>(8, kNumSMs / min_ctas);
    return splits * pixels * d->cout;
}

extern "C" int td_conv3x3_run(const td_conv_plan* plan, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(plan, "td_conv3x3_run: null plan");
    cudaStream_t s = (cudaStream_t)stream;
    switch (plan->engine) {
        case TD_CONV_TC: return tc_plan_run(plan, s);
        case TD_CONV_DIRECT: return run_direct(plan, s);
        default: return run_simt(plan, s);
    }
}

extern "C" void td_conv3x3_plan_destroy(td_conv_plan* plan) { delete plan; }

extern "C" double td_conv3x3_flops(const td_conv_plan* plan) {
    if (!plan) return 0.0;
    const td_conv3x3_desc& d = plan->d;
    return 2.0 * d.batch * d.height * d.width * (double)d.cout * 9.0 * d.cin;
}
