// Kernels of the reference's conv / attention VAE (vae_laion.py), fp32, NHWC:
//   * stride-2 4x4 convolution (vae_laion.py:98-131) and its transpose ConvTranspose2d(4, stride 2, padding 1) (:138-165) as
//     gather-form implicit GEMMs (FFMA; the layers are 0.05-0.3 GFLOP per image -- the attention below is where the work is);
//   * spectral normalisation (torch.nn.utils.spectral_norm: sigma = u^T W v, optional power iteration);
//   * SelfAttention (:50-65) flash-style: softmax(Q K^T) V over N = H*W up to 16384 keys with d_qk = C/8 in {4, 8, 16}
//     and d_v = C, online softmax over key tiles -- the (HW)^2 matrix the reference materialises (1 GiB per image at
//     128 x 128) never exists -- fused with the gamma * out + x residual.
#include <algorithm>

#include "common.cuh"

namespace td {

// ---------------------------------------------------------------------------------------------
// Gather-form implicit GEMM.  M = output pixels (of one parity class for the transpose), N = Cout, K = taps * Cin.
//   MODE 0  Conv2d(4, stride 2, padding 1):   k = (ky*4 + kx)*Cin + ci,  ih = 2*oh - 1 + ky
//   MODE 1  ConvTranspose2d(4, 2, 1), class (ah, aw) = (oh & 1, ow & 1) = blockIdx.z:
//           k = (j*2 + i)*Cin + ci,  ky = 1 - ah + 2j,  ih = (oh + ah)/2 - j     (two taps per dimension reach a pixel)
// Weights are packed [class][Cout][K] (already divided by sigma).  64 x 64 tile, BK = 16, 256 threads, 4 x 4 per thread.
// ---------------------------------------------------------------------------------------------
struct VlConv {
    const float* x; const float* w; const float* bias; float* y;
    int B, Hin, Win, Cin, Hout, Wout, Cout;
    int x_nchw, y_nchw, act;
};

constexpr int VL_BM = 64, VL_BN = 64, VL_BK = 16, VL_THREADS = 256;

__device__ inline float vl_act(float v, int act) {
    if (act == TD_ACT_RELU) return fmaxf(v, 0.f);
    if (act == TD_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(VL_THREADS)
vl_conv_kernel(const VlConv p) {
    td::pdl_sync();
    __shared__ float As[VL_BK][VL_BM + 4];
    __shared__ float Bs[VL_BK][VL_BN + 4];
    const int cls = MODE == 1 ? blockIdx.z : 0;
    const int ah = cls >> 1, aw = cls & 1;
    const int Hm = MODE == 1 ? p.Hin : p.Hout, Wm = MODE == 1 ? p.Win : p.Wout;      // pixel grid of this GEMM
    const int64_t M = (int64_t)p.B * Hm * Wm;
    const int taps = MODE == 1 ? 4 : 16;
    const int K = taps * p.Cin;
    const float* __restrict__ wt = p.w + (int64_t)cls * p.Cout * K;
    const int64_t m0 = (int64_t)blockIdx.x * VL_BM;
    const int n0 = blockIdx.y * VL_BN;
    const int tid = threadIdx.x;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int64_t pm = m0 + lrow;
    const bool pvalid = pm < M;
    int pb = 0, pr = 0, ps = 0;
    if (pvalid) {
        ps = (int)(pm % Wm);
        const int64_t r = pm / Wm;
        pr = (int)(r % Hm);
        pb = (int)(r / Hm);
    }
    const int wn = n0 + lrow;
    const bool wvalid = wn < p.Cout;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += VL_BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float av = 0.f, bv = 0.f;
            if (k < K) {
                const int tap = k / p.Cin, c = k - tap * p.Cin;
                int ih, iw;
                if (MODE == 0) { ih = 2 * pr - 1 + (tap >> 2); iw = 2 * ps - 1 + (tap & 3); }
                else { ih = pr + ah - (tap >> 1); iw = ps + aw - (tap & 1); }
                if (pvalid && ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win)
                    av = p.x_nchw ? p.x[(((int64_t)pb * p.Cin + c) * p.Hin + ih) * p.Win + iw]
                                  : p.x[(((int64_t)pb * p.Hin + ih) * p.Win + iw) * p.Cin + c];
                if (wvalid) bv = wt[(int64_t)wn * K + k];
            }
            As[lk + q][lrow] = av;
            Bs[lk + q][lrow] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VL_BK; ++k) {
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int s_ = (int)(m % Wm);
        const int64_t r = m / Wm;
        const int r_ = (int)(r % Hm);
        const int b_ = (int)(r / Hm);
        const int oh = MODE == 1 ? 2 * r_ + ah : r_, ow = MODE == 1 ? 2 * s_ + aw : s_;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            const float v = vl_act(acc[i][j] + (p.bias ? p.bias[n] : 0.f), p.act);
            if (p.y_nchw) p.y[(((int64_t)b_ * p.Cout + n) * p.Hout + oh) * p.Wout + ow] = v;
            else p.y[(((int64_t)b_ * p.Hout + oh) * p.Wout + ow) * p.Cout + n] = v;
        }
    }
}

// weight packing (x 1/sigma).  Conv2d weight (Cout, Cin, 4, 4) -> [Cout][(ky*4+kx)*Cin + ci];
// ConvTranspose2d weight (Cin, Cout, 4, 4) -> [class][Cout][(j*2+i)*Cin + ci] with ky = 1 - ah + 2j, kx = 1 - aw + 2i.
__global__ void __launch_bounds__(256)
vl_pack_kernel(const float* __restrict__ w, const float* __restrict__ sigma, float* __restrict__ out, int cout, int cin,
               int transposed) {
    td::pdl_sync();
    const float inv = sigma ? 1.f / sigma[0] : 1.f;
    const int64_t total = (int64_t)cout * cin * 16;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (!transposed) {
            const int ci = (int)(e % cin);
            const int64_t r = e / cin;
            const int tap = (int)(r % 16), co = (int)(r / 16);
            out[e] = w[((int64_t)co * cin + ci) * 16 + tap] * inv;
        } else {
            const int K = 4 * cin;
            const int k = (int)(e % K);
            const int64_t r = e / K;
            const int co = (int)(r % cout), cls = (int)(r / cout);
            const int ci = k % cin, t = k / cin;
            const int ky = 1 - (cls >> 1) + 2 * (t >> 1), kx = 1 - (cls & 1) + 2 * (t & 1);
            out[e] = w[(((int64_t)ci * cout + co) * 4 + ky) * 4 + kx] * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// torch.nn.utils.spectral_norm: W_mat = weight with `dim` first, flattened to [rows][cols].
//   power iteration (training forward):  v = normalize(W^T u),  u = normalize(W v)   (eps 1e-12)
//   sigma = u^T W v
// One CTA (the matrices are at most 256 x 2048); element (r, c) of W_mat: dim 0 -> w[r*cols + c];
// dim 1 (ConvTranspose2d, weight (Cin, Cout, kh, kw)) -> w[((c / khw)*rows + r)*khw + c % khw].
// ---------------------------------------------------------------------------------------------
__device__ inline float vl_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(1024)
spectral_sigma_kernel(const float* __restrict__ w, int rows, int cols, int dim1, int khw, float* __restrict__ u,
                      float* __restrict__ v, int power_iterations, float eps, float* __restrict__ sigma_out,
                      float* __restrict__ wv) {
    td::pdl_sync();
    __shared__ float red[32];
    auto W = [&](int r, int c) -> float {
        return dim1 ? w[((int64_t)(c / khw) * rows + r) * khw + (c % khw)] : w[(int64_t)r * cols + c];
    };
    for (int it = 0; it < power_iterations; ++it) {
        // v = normalize(W^T u)
        float nrm = 0.f;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s = fmaf(W(r, c), u[r], s);
            v[c] = s;
            nrm = fmaf(s, s, nrm);
        }
        nrm = fmaxf(sqrtf(vl_block_sum(nrm, red)), eps);
        for (int c = threadIdx.x; c < cols; c += blockDim.x) v[c] = v[c] / nrm;
        __syncthreads();
        // u = normalize(W v)
        float nu = 0.f;
        for (int r = threadIdx.x >> 5; r < rows; r += blockDim.x >> 5) {          // one warp per row
            float s = 0.f;
            for (int c = threadIdx.x & 31; c < cols; c += 32) s = fmaf(W(r, c), v[c], s);
            s = warp_sum(s);
            if ((threadIdx.x & 31) == 0) { wv[r] = s; nu = fmaf(s, s, nu); }
        }
        nu = fmaxf(sqrtf(vl_block_sum(nu, red)), eps);
        for (int r = threadIdx.x; r < rows; r += blockDim.x) u[r] = wv[r] / nu;
        __syncthreads();
    }
    float part = 0.f;
    for (int r = threadIdx.x >> 5; r < rows; r += blockDim.x >> 5) {
        float s = 0.f;
        for (int c = threadIdx.x & 31; c < cols; c += 32) s = fmaf(W(r, c), v[c], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) part = fmaf(u[r], s, part);
    }
    const float sig = vl_block_sum(part, red);
    if (threadIdx.x == 0) sigma_out[0] = sig;
}

__global__ void vl_scale_inv_kernel(float* __restrict__ x, const float* __restrict__ sigma, int n) {
    td::pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] / sigma[0];
}

// ---------------------------------------------------------------------------------------------
// Flash-style self-attention (vae_laion.py:57-65):  out_i = sum_j softmax_j(q_i . k_j) v_j ;  y = gamma * out + x.
// qkv: [B][N][2*DQ + DV] rows (q | k | v) from ONE fused 1x1-conv GEMM; x / y: NHWC [B][N][DV].
// A CTA owns 128 queries; thread (q, part) keeps its query (DQ registers), the running max / sum and DV / TPQ output
// channels.  K / V tiles of KB keys are staged in shared memory; all threads of a warp read the SAME key / value element
// (they differ only in the query), so every shared-memory read is a broadcast.  Online softmax per tile.
// ---------------------------------------------------------------------------------------------
template <int DQ, int DV, int TPQ>
__global__ void __launch_bounds__(128 * TPQ)
vl_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                    float* __restrict__ y, int N) {
    td::pdl_sync();
    constexpr int KB = 32, LD = 2 * DQ + DV, CV = DV / TPQ;
    __shared__ __align__(16) float Ks[KB][DQ];
    __shared__ __align__(16) float Vs[KB][DV];
    const int b = blockIdx.y;
    const int qi = blockIdx.x * 128 + (threadIdx.x & 127);
    const int part = threadIdx.x >> 7;
    const float* base = qkv + (int64_t)b * N * LD;
    float q[DQ];
#pragma unroll
    for (int d = 0; d < DQ; ++d) q[d] = base[(int64_t)qi * LD + d];
    float acc[CV];
#pragma unroll
    for (int c = 0; c < CV; ++c) acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < N; j0 += KB) {
        __syncthreads();
        for (int e = threadIdx.x; e < KB * (DQ + DV) / 4; e += blockDim.x) {      // (k | v) of a key row are contiguous
            const int j = e / ((DQ + DV) / 4), r4 = e - j * ((DQ + DV) / 4);
            const float4 v4 = *reinterpret_cast<const float4*>(base + (int64_t)(j0 + j) * LD + DQ + r4 * 4);
            if (r4 * 4 < DQ) *reinterpret_cast<float4*>(&Ks[j][r4 * 4]) = v4;
            else *reinterpret_cast<float4*>(&Vs[j][r4 * 4 - DQ]) = v4;
        }
        __syncthreads();
        float s[KB];
        float tmax = m;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            float t = 0.f;
#pragma unroll
            for (int d = 0; d < DQ; ++d) t = fmaf(q[d], Ks[j][d], t);
            s[j] = t;
            tmax = fmaxf(tmax, t);
        }
        const float alpha = __expf(m - tmax);            // 0 on the first tile (m = -inf)
        m = tmax;
        l *= alpha;
#pragma unroll
        for (int c = 0; c < CV; ++c) acc[c] *= alpha;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            const float pj = __expf(s[j] - m);
            l += pj;
#pragma unroll
            for (int c = 0; c < CV; c += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][part * CV + c]);
                acc[c] = fmaf(pj, v4.x, acc[c]);
                acc[c + 1] = fmaf(pj, v4.y, acc[c + 1]);
                acc[c + 2] = fmaf(pj, v4.z, acc[c + 2]);
                acc[c + 3] = fmaf(pj, v4.w, acc[c + 3]);
            }
        }
    }
    const float g = gamma[0], inv = 1.f / l;
    const int64_t o = ((int64_t)b * N + qi) * DV + part * CV;
#pragma unroll
    for (int c = 0; c < CV; c += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + o + c);
        float4 r;
        r.x = fmaf(g, acc[c] * inv, xv.x);
        r.y = fmaf(g, acc[c + 1] * inv, xv.y);
        r.z = fmaf(g, acc[c + 2] * inv, xv.z);
        r.w = fmaf(g, acc[c + 3] * inv, xv.w);
        *reinterpret_cast<float4*>(y + o + c) = r;
    }
}

}  // namespace td

using namespace td;

static int vl_conv_launch(int mode, const VlConv& p, cudaStream_t s) {
    const int Hm = mode == 1 ? p.Hin : p.Hout, Wm = mode == 1 ? p.Win : p.Wout;
    const int64_t M = (int64_t)p.B * Hm * Wm;
    const dim3 grid((unsigned)ceil_div(M, VL_BM), (unsigned)ceil_div(p.Cout, VL_BN), mode == 1 ? 4 : 1);
    if (mode == 0) td::launch(vl_conv_kernel<0>, td::LaunchCfg(grid, VL_THREADS, 0, s), p);
    else td::launch(vl_conv_kernel<1>, td::LaunchCfg(grid, VL_THREADS, 0, s), p);
    return launch_status(mode == 0 ? "conv4x4s2" : "convT4x4s2");
}

extern "C" int td_conv4x4s2_fwd(const float* x, const float* w_packed, const float* bias, float* y, int batch, int hin, int win,
                                int cin, int cout, int x_nchw, int act, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && w_packed && y && batch > 0 && hin > 0 && win > 0 && hin % 2 == 0 && win % 2 == 0 && cin > 0 && cout > 0,
                 "td_conv4x4s2_fwd: bad args");
    VlConv p{x, w_packed, bias, y, batch, hin, win, cin, hin / 2, win / 2, cout, x_nchw, 0, act};
    return vl_conv_launch(0, p, (cudaStream_t)stream);
}

extern "C" int td_convT4x4s2_fwd(const float* x, const float* w_packed, const float* bias, float* y, int batch, int hin, int win,
                                 int cin, int cout, int y_nchw, int act, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && w_packed && y && batch > 0 && hin > 0 && win > 0 && cin > 0 && cout > 0, "td_convT4x4s2_fwd: bad args");
    VlConv p{x, w_packed, bias, y, batch, hin, win, cin, 2 * hin, 2 * win, cout, 0, y_nchw, act};
    return vl_conv_launch(1, p, (cudaStream_t)stream);
}

extern "C" int td_pack_conv4x4_weight(const float* w, const float* sigma_dev, float* out, int cout, int cin, int transposed,
                                      void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(w && out && cout > 0 && cin > 0, "td_pack_conv4x4_weight: bad args");
    const int64_t n = (int64_t)cout * cin * 16;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8));
    td::launch(vl_pack_kernel, td::LaunchCfg(grid, 256, 0, (cudaStream_t)stream), w, sigma_dev, out, cout, cin, transposed);
    return launch_status("pack_conv4x4_weight");
}

extern "C" int td_spectral_sigma(const float* w, int rows, int cols, int dim1, int khw, float* u, float* v,
                                 int power_iterations, float eps, float* sigma_out, float* scratch_rows, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(w && u && v && sigma_out && scratch_rows && rows > 0 && cols > 0 && khw > 0 && power_iterations >= 0,
                 "td_spectral_sigma: bad args");
    TD_CHECK_ARG(!dim1 || cols % khw == 0, "td_spectral_sigma: cols must be a multiple of kh*kw for dim = 1");
    td::launch(spectral_sigma_kernel, td::LaunchCfg(1, 1024, 0, (cudaStream_t)stream), w, rows, cols, dim1, khw, u, v,
               power_iterations, eps, sigma_out, scratch_rows);
    return launch_status("spectral_sigma");
}

extern "C" int td_scale_by_inv_sigma(float* x, const float* sigma_dev, int n, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && sigma_dev && n > 0, "td_scale_by_inv_sigma: bad args");
    td::launch(vl_scale_inv_kernel, td::LaunchCfg((n + 127) / 128, 128, 0, (cudaStream_t)stream), x, sigma_dev, n);
    return launch_status("scale_by_inv_sigma");
}

extern "C" int td_self_attention_fwd(const float* qkv, const float* x, const float* gamma, float* y, int batch, int n,
                                     int dq, int dv, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(qkv && x && gamma && y && batch > 0 && n > 0 && n % 128 == 0, "td_self_attention_fwd: n must be a positive multiple of 128");
    TD_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)x | (uintptr_t)y) & 15) == 0, "td_self_attention_fwd: tensors must be 16-byte aligned");
    const dim3 grid((unsigned)(n / 128), (unsigned)batch);
    cudaStream_t s = (cudaStream_t)stream;
    if (dq == 4 && dv == 32) td::launch(vl_attention_kernel<4, 32, 1>, td::LaunchCfg(grid, 128, 0, s), qkv, x, gamma, y, n);
    else if (dq == 8 && dv == 64) td::launch(vl_attention_kernel<8, 64, 1>, td::LaunchCfg(grid, 128, 0, s), qkv, x, gamma, y, n);
    else if (dq == 16 && dv == 128) td::launch(vl_attention_kernel<16, 128, 2>, td::LaunchCfg(grid, 256, 0, s), qkv, x, gamma, y, n);
    else TD_CHECK_ARG(false, "td_self_attention_fwd: (d_qk, d_v) = (%d, %d) not instantiated (C/8, C with C in {32, 64, 128})", dq, dv);
    return launch_status("self_attention");
}
