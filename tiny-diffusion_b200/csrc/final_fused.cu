// Network tail in one kernel (eval path): bilinear resize (align_corners=True) of the last decoder
// activation followed by final_conv (3x3, C -> 1 channel, + bias), diffusion.py:157-160.
//
// Both steps are linear, so they commute with the per-tap channel contraction:
//   eps[p] = b + sum_t sum_c R(x)[p + o_t, c] w[t, c] = b + sum_t R(e_t)[p + o_t],   e_t[s] = sum_c x[s, c] w[t, c]
// The C-channel activation is read ONCE at its own resolution and contracted to nine scalars per pixel
// (e_t, in shared memory); the resize and the 3x3 stencil then run on scalars.  The resized C-channel
// tensor is never materialised.  HBM-bound: B*Hi*Wi*C*sizeof(T) bytes in, B*Ho*Wo*4 out.
#include "common.cuh"

namespace td {

struct FinalArgs {
    const void* x;
    int ldx, x_coff, B, Hi, Wi, C, Ho, Wo, BH;      // BH = output rows per CTA
    const float* w;      // [1][9][C] fp32
    const float* bias;   // [1] or NULL
    float* y;            // [B][1][Ho][Wo]
    int max_src_rows;    // source rows any band can touch
};

__device__ inline void bil1(int dst, int in, int out, int& i0, int& i1, float& l0, float& l1) {
    // same fp32 op order as ATen's area_pixel_compute_source_index (align_corners=True)
    const float scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float src = scale * (float)dst;
    i0 = min((int)floorf(src), in - 1);
    l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
}

// Sum of each of the nine per-tap partials over the L lanes of a pixel.  L == 8: values 0..7 go through a halving
// butterfly (4 + 2 + 1 shuffles: every step each lane keeps one half of its live values and sends the other), value 8
// through a plain one (3 shuffles); lane i ends up with the total of tap i in acc[0], every lane with tap 8 in acc[8].
template <int L>
__device__ inline void reduce_taps(float (&acc)[9], int lane) {
    if constexpr (L == 8) {
#pragma unroll
        for (int s = 4; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const float send = up ? acc[i] : acc[i + s];
                const float keep = up ? acc[i + s] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) acc[8] += __shfl_xor_sync(0xffffffffu, acc[8], o);
    } else {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
    }
}

template <typename T, int L>
__global__ void __launch_bounds__(256, 2)
final_resize_conv_kernel(const FinalArgs a) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    extern __shared__ float sm[];
    const int bands = (a.Ho + a.BH - 1) / a.BH;
    const int b = blockIdx.x / bands, band = blockIdx.x - b * bands;
    const int r0 = band * a.BH, r1 = min(r0 + a.BH, a.Ho);          // output rows [r0, r1)
    const int q0 = max(r0 - 1, 0), q1 = min(r1 + 1, a.Ho);           // resized rows the stencil reads
    int s0, s1, tmp;
    float f0, f1;
    bil1(q0, a.Hi, a.Ho, s0, tmp, f0, f1);
    bil1(q1 - 1, a.Hi, a.Ho, tmp, s1, f0, f1);
    const int src_rows = s1 - s0 + 1;
    float* e = sm;                                        // [9][max_src_rows][Wi]
    float* dq = sm + 9 * a.max_src_rows * a.Wi;           // [9][BH + 2][Wo]
    const int e_plane = a.max_src_rows * a.Wi, d_plane = (a.BH + 2) * a.Wo;

    // ---- phase 1: e_t[s] = sum_c x[s, c] w[t, c] --------------------------------------------------------
    const int lane = threadIdx.x % L, slot = threadIdx.x / L, slots = blockDim.x / L;
    float w[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < V; k += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(a.w + t * a.C + lane * V + k));
            w[t][k] = q.x; w[t][k + 1] = q.y; w[t][k + 2] = q.z; w[t][k + 3] = q.w;
        }
    const T* xb = reinterpret_cast<const T*>(a.x) + ((int64_t)b * a.Hi + s0) * a.Wi * a.ldx + a.x_coff + lane * V;
    const int npix = src_rows * a.Wi;
    constexpr int P = 4;                                             // pixels per thread and pass: P independent loads in flight
    for (int base = 0; base < npix; base += slots * P) {             // uniform trip count: shuffles stay converged
        Vec<T> raw[P];
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const int pix = base + u * slots + slot;
            if (pix < npix) raw[u] = Vec<T>::load(xb + (int64_t)pix * a.ldx);
        }
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const int pix = base + u * slots + slot;
            float acc[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[t] = 0.f;
            if (pix < npix) {
                float f[V];
                raw[u].unpack(f);
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[t] = fmaf(f[k], w[t][k], acc[t]);
            }
            reduce_taps<L>(acc, lane);
            if (pix < npix) {
                if constexpr (L == 8) {
                    e[lane * e_plane + pix] = acc[0];          // each step kept the upper half where the lane bit was set: tap == lane
                    if (lane == 0) e[8 * e_plane + pix] = acc[8];
                } else if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) e[t * e_plane + pix] = acc[t];
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 2: resized scalars dq_t[q] = R(e_t)[q] for rows q0 .. q1-1 ----------------------------------
    const int nq = (q1 - q0) * a.Wo;
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        const int qr = i / a.Wo, qc = i - qr * a.Wo;
        int h0, h1, c0, c1;
        float lh0, lh1, lw0, lw1;
        bil1(q0 + qr, a.Hi, a.Ho, h0, h1, lh0, lh1);
        bil1(qc, a.Wi, a.Wo, c0, c1, lw0, lw1);
        const int o00 = (h0 - s0) * a.Wi + c0, o01 = (h0 - s0) * a.Wi + c1, o10 = (h1 - s0) * a.Wi + c0, o11 = (h1 - s0) * a.Wi + c1;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float* et = e + t * e_plane;
            dq[t * d_plane + i] = lh0 * (lw0 * et[o00] + lw1 * et[o01]) + lh1 * (lw0 * et[o10] + lw1 * et[o11]);
        }
    }
    __syncthreads();
    // ---- phase 3: 3x3 stencil over the scalars, zero padding at the image border --------------------------
    const float bias = a.bias ? a.bias[0] : 0.f;
    const int nout = (r1 - r0) * a.Wo;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        const int rr = i / a.Wo, cc = i - rr * a.Wo;
        const int r = r0 + rr;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hh = r + t / 3 - 1, ww = cc + t % 3 - 1;
            if (hh >= 0 && hh < a.Ho && ww >= 0 && ww < a.Wo) acc += dq[t * d_plane + (hh - q0) * a.Wo + ww];
        }
        a.y[((int64_t)b * a.Ho + r) * a.Wo + cc] = acc + bias;
    }
}

}  // namespace td

using namespace td;

extern "C" int td_final_resize_conv(const void* x, int dtype, int ldx, int x_coff, int batch, int hi, int wi, int c,
                                    const float* w_ohwi, const float* bias, int ho, int wo, float* y_nchw, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && w_ohwi && y_nchw && batch > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "td_final_resize_conv: bad args");
    const int V = dtype == TD_BF16 ? 8 : 4;
    TD_CHECK_ARG(dtype == TD_BF16 || dtype == TD_F32, "td_final_resize_conv: unknown dtype %d", dtype);
    const int L = c / V;
    TD_CHECK_ARG(c % V == 0 && L >= 1 && L <= 32 && (L & (L - 1)) == 0 && ldx % V == 0 && x_coff % V == 0,
                 "td_final_resize_conv: channels (%d) must be V * 2^k <= 32 * V", c);
    FinalArgs a;
    a.x = x; a.ldx = ldx; a.x_coff = x_coff; a.B = batch; a.Hi = hi; a.Wi = wi; a.C = c; a.Ho = ho; a.Wo = wo;
    a.w = w_ohwi; a.bias = bias; a.y = y_nchw;
    // rows per CTA: at least one CTA per SM at the given batch (taller bands re-read fewer halo rows), at least 4 rows
    int bh = ho;
    while (bh > 4 && (int64_t)batch * ((ho + bh - 1) / bh) < kNumSMs) bh = (bh + 1) / 2;
    a.BH = bh;
    // source rows a band of bh+2 resized rows can touch: ceil((bh+1) * scale) + 2
    const double scale = ho > 1 ? (double)(hi - 1) / (double)(ho - 1) : 0.0;
    int msr = (int)((bh + 1) * scale) + 3;
    if (msr > hi) msr = hi;
    a.max_src_rows = msr;
    const size_t smem = ((size_t)9 * msr * wi + (size_t)9 * (bh + 2) * wo) * sizeof(float);
    TD_CHECK_ARG(smem <= 200 * 1024, "td_final_resize_conv: feature map too large for the shared-memory plan (%zu bytes)", smem);
    const int grid = batch * ((ho + bh - 1) / bh);
    cudaStream_t s = (cudaStream_t)stream;
#define TD_FINAL(T, LL)                                                                                                       \
    do {                                                                                                                      \
        static size_t cfg = 0;                                                                                                \
        if (smem > cfg) {                                                                                                     \
            TD_CUDA(cudaFuncSetAttribute(final_resize_conv_kernel<T, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cfg = smem;                                                                                                       \
        }                                                                                                                     \
        td::launch(final_resize_conv_kernel<T, LL>, td::LaunchCfg(grid, 256, smem, s), a);                                                           \
    } while (0)
#define TD_FINAL_L(T)                                       \
    switch (L) {                                            \
        case 1: TD_FINAL(T, 1); break;                      \
        case 2: TD_FINAL(T, 2); break;                      \
        case 4: TD_FINAL(T, 4); break;                      \
        case 8: TD_FINAL(T, 8); break;                      \
        case 16: TD_FINAL(T, 16); break;                    \
        default: TD_FINAL(T, 32); break;                    \
    }
    if (dtype == TD_BF16) { TD_FINAL_L(__nv_bfloat16) } else { TD_FINAL_L(float) }
#undef TD_FINAL_L
#undef TD_FINAL
    return launch_status("final_resize_conv");
}
