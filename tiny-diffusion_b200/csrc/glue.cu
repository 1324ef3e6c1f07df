// NHWC glue kernels around the convolutions (all HBM/L2-bound, 16-byte vector accesses):
// max-pool, decoder-input assembly (bilinear up-sample + skip/embedding add + resize + concat),
// bilinear resize, the conditioning head and weight packing.
#include <algorithm>

#include "common.cuh"

namespace td {

constexpr int kThreads = 256;

struct Bil { int i0, i1; float l0, l1; };
// align_corners=True source index, same fp32 op order as ATen's area_pixel_compute_* helpers.
__device__ inline Bil bil(int dst, int in, int out) {
    Bil b;
    const float scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float src = scale * (float)dst;
    b.i0 = min((int)floorf(src), in - 1);
    b.l1 = fminf(fmaxf(src - (float)b.i0, 0.f), 1.f);
    b.l0 = 1.f - b.l1;
    b.i1 = b.i0 + ((b.i0 < in - 1) ? 1 : 0);
    return b;
}

// Thread layout shared by the NHWC glue kernels: one CTA per output row (b, h); threadIdx.x walks the
// 16-byte channel vectors of a pixel, threadIdx.y the pixels of the row -- no integer division on the
// per-element path (these kernels were instruction-bound on 64-bit div/mod and bilinear setup).
static inline dim3 row_block(int cv) {
    const int bx = cv < 256 ? cv : 256;
    int by = 256 / bx;
    if (by < 1) by = 1;
    return dim3((unsigned)bx, (unsigned)by, 1);
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2, ceil_mode)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    const int cv = C / V;
    const int b = blockIdx.x / Ho, ho = blockIdx.x - b * Ho;
    const T* xb = x + (int64_t)b * H * W * C;
    T* yrow = y + ((int64_t)b * Ho + ho) * Wo * C;
    for (int cvi = threadIdx.x; cvi < cv; cvi += blockDim.x) {
        const int c = cvi * V;
        for (int wo = threadIdx.y; wo < Wo; wo += blockDim.y) {
            float m[V];
#pragma unroll
            for (int k = 0; k < V; ++k) m[k] = -INFINITY;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int h = 2 * ho + dy;
                if (h >= H) continue;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int w = 2 * wo + dx;
                    if (w >= W) continue;
                    float f[V];
                    Vec<T>::load(xb + ((int64_t)h * W + w) * C + c).unpack(f);
#pragma unroll
                    for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], f[k]);
                }
            }
            Vec<T>::pack(m).store(yrow + (int64_t)wo * C + c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// out = [ up2x(low) | resize(skip + temb) ]
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRowW = 64;      // widest output row the column tables are sized for

// column interpolation table of one output row, built once per CTA
__device__ inline void build_col_table(Bil* tab, int in, int out) {
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < out; i += blockDim.x * blockDim.y) tab[i] = bil(i, in, out);
}

// One thread produces one 16-byte channel vector of EVERY pixel of an output row, walking the row left to right.
// The two source columns it interpolates between stay in registers, already combined vertically, so a source
// vector is loaded and converted once instead of once per output tap (the per-pixel 4-tap version of this
// kernel was issue-bound: ~300 instructions per 16 bytes of output, ncu r01h).
//   out[w] = l0(w) * v[i0(w)] + l1(w) * v[i1(w)] (+ add),     v[i] = lh0 * r0[i] + lh1 * r1[i]
template <typename T>
__device__ inline void resample_row(const T* __restrict__ r0, const T* __restrict__ r1, int cstride, float lh0, float lh1,
                                    const Bil* __restrict__ col, int w_begin, int w_end, const float* add, T* __restrict__ dst,
                                    int dstride) {
    constexpr int V = Vec<T>::N;
    float va[V], vb[V];
    int ia = -1, ib = -1;
    auto fetch = [&](int i, float* v) {
        float a[V], b[V];
        Vec<T>::load(r0 + (int64_t)i * cstride).unpack(a);
        Vec<T>::load(r1 + (int64_t)i * cstride).unpack(b);
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] = lh0 * a[k] + lh1 * b[k];
    };
    for (int w = w_begin; w < w_end; ++w) {
        const Bil bw = col[w];
        if (bw.i0 != ia) {
            if (bw.i0 == ib) {
#pragma unroll
                for (int k = 0; k < V; ++k) va[k] = vb[k];
            } else {
                fetch(bw.i0, va);
            }
            ia = bw.i0;
        }
        if (bw.i1 != ib) {
            if (bw.i1 == ia) {
#pragma unroll
                for (int k = 0; k < V; ++k) vb[k] = va[k];
            } else {
                fetch(bw.i1, vb);
            }
            ib = bw.i1;
        }
        float r[V];
#pragma unroll
        for (int k = 0; k < V; ++k) r[k] = bw.l0 * va[k] + bw.l1 * vb[k];
        if (add) {
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] += add[k];
        }
        Vec<T>::pack(r).store(dst + (int64_t)w * dstride);
    }
}

// grid: (row groups, channel-vector groups, 2 halves); block: (vectors, rows).  blockIdx.z = 0 writes the
// up-sampled `low` channels, 1 the resized skip (+ embedding) channels; the threads of a warp share the column
// table of their half, so the register-cache branches above are warp-uniform.
template <typename T>
__global__ void __launch_bounds__(256)
upcat_kernel(const T* __restrict__ low, const T* __restrict__ skip, const float* __restrict__ temb, int ld_temb,
             int temb_off, T* __restrict__ out, int B, int Ho, int Wo, int Cu, int Hs, int Ws, int Cs, int nseg) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    __shared__ Bil col[kMaxRowW];
    const int Ct = Cu + Cs;
    const bool skip_half = blockIdx.z == 1;
    const int Ch = skip_half ? Cs : Cu;                      // channels of this half
    const int Hin = skip_half ? Hs : Ho / 2, Win = skip_half ? Ws : Wo / 2;
    const T* src = skip_half ? skip : low;
    const bool same = skip_half && Hs == Ho && Ws == Wo;
    if (!same) build_col_table(col, Win, Wo);
    __syncthreads();
    const int cvi = blockIdx.y * blockDim.x + threadIdx.x;
    const int seg = blockIdx.x % nseg;                       // segment of the output row walked by this thread
    const int row = (blockIdx.x / nseg) * blockDim.y + threadIdx.y;   // (b, ho) flattened
    if (cvi * V >= Ch || row >= B * Ho) return;
    const int sw = (Wo + nseg - 1) / nseg;
    const int w_begin = seg * sw, w_end = min(Wo, w_begin + sw);
    const int c = cvi * V;
    const int b = row / Ho, ho = row - b * Ho;
    T* orow = out + (int64_t)row * Wo * Ct + (skip_half ? Cu : 0) + c;
    float te[V];
    if (skip_half) {
        // the embedding is constant over space and the bilinear weights sum to one: resize(skip + t) == resize(skip) + t
#pragma unroll
        for (int k = 0; k < V; ++k) te[k] = temb[(int64_t)b * ld_temb + temb_off + c + k];
    }
    if (same) {
        const T* r0 = src + ((int64_t)row * Ws) * Cs + c;
        for (int w = w_begin; w < w_end; ++w) {
            float r[V];
            Vec<T>::load(r0 + (int64_t)w * Cs).unpack(r);
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] += te[k];
            Vec<T>::pack(r).store(orow + (int64_t)w * Ct);
        }
        return;
    }
    const Bil bh = bil(ho, Hin, Ho);
    const T* r0 = src + (((int64_t)b * Hin + bh.i0) * Win) * Ch + c;
    const T* r1 = src + (((int64_t)b * Hin + bh.i1) * Win) * Ch + c;
    resample_row<T>(r0, r1, Ch, bh.l0, bh.l1, col, w_begin, w_end, skip_half ? te : nullptr, orow, Ct);
}

template <typename T>
__global__ void __launch_bounds__(256)
resize_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Hi, int Wi, int Ho, int Wo, int C, int nseg) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    __shared__ Bil col[kMaxRowW];
    build_col_table(col, Wi, Wo);
    __syncthreads();
    const int cvi = blockIdx.y * blockDim.x + threadIdx.x;
    const int seg = blockIdx.x % nseg;
    const int row = (blockIdx.x / nseg) * blockDim.y + threadIdx.y;
    if (cvi * V >= C || row >= B * Ho) return;
    const int sw = (Wo + nseg - 1) / nseg;
    const int w_begin = seg * sw, w_end = min(Wo, w_begin + sw);
    const int c = cvi * V;
    const int b = row / Ho, ho = row - b * Ho;
    const Bil bh = bil(ho, Hi, Ho);
    const T* r0 = x + (((int64_t)b * Hi + bh.i0) * Wi) * C + c;
    const T* r1 = x + (((int64_t)b * Hi + bh.i1) * Wi) * C + c;
    resample_row<T>(r0, r1, C, bh.l0, bh.l1, col, w_begin, w_end, nullptr, y + (int64_t)row * Wo * C + c, C);
}

// block (vectors of one half, rows) for the row-walking kernels: up to 32 channel vectors wide, 256 threads
static inline dim3 walk_block(int cv) {
    int bx = 1;
    while (bx < cv && bx < 32) bx <<= 1;
    return dim3((unsigned)bx, (unsigned)(256 / bx), 1);
}
// row segments per thread walk: enough threads in flight (~1500 per SM) to cover the load latency
static inline int walk_segments(int64_t threads, int width) {
    int64_t n = ceil_div((int64_t)kNumSMs * 1536, std::max<int64_t>(threads, 1));
    n = std::max<int64_t>(1, std::min<int64_t>(n, std::max(1, width / 4)));
    const int sw = (int)ceil_div(width, n);
    return (int)ceil_div(width, sw);
}

// OIHW fp32 -> OHWI (fp32 or bf16)
template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_weight_kernel(const float* __restrict__ oihw, T* __restrict__ ohwi, int cout, int cin) {
    td::pdl_sync();
    const int64_t total = (int64_t)cout * 9 * cin;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total; i += gridDim.x * blockDim.x) {
        int c = (int)(i % cin);
        uint32_t r = i / cin;
        int tap = (int)(r % 9);
        int o = (int)(r / 9);
        ohwi[i] = from_f32<T>(oihw[((int64_t)o * cin + c) * 9 + tap]);
    }
}

// All bf16 conv layers of a model re-packed in ONE launch after an optimizer step: OIHW fp32 -> OHWI bf16 (forward
// operand) and, optionally, [Cin][3][3][Cout] with flipped taps (data-gradient operand).  One CTA transposes a
// 32 (cout) x 32 (cin) x 9 (taps) tile through shared memory: global reads are 288 contiguous floats per output
// channel, both writes are 64-byte runs; the 26 per-layer launches this replaces cost ~200 us per train step.
struct PackEntry { const float* src; __nv_bfloat16* fwd; __nv_bfloat16* dgrad; int cout, cin, tile_begin, pad; };
constexpr int PK_T = 32;
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackEntry* __restrict__ tab, int n_entries) {
    td::pdl_sync();
    __shared__ float tile[PK_T][PK_T * 9 + 1];
    int e = 0;
    while (e + 1 < n_entries && (int)blockIdx.x >= tab[e + 1].tile_begin) ++e;
    const PackEntry L = tab[e];
    const int t = blockIdx.x - L.tile_begin;
    const int cb = L.cin / PK_T;
    const int o0 = (t / cb) * PK_T, c0 = (t % cb) * PK_T;
    // 32 rows of 288 contiguous floats (16-byte aligned: c0 is a multiple of 32): 9 float4 loads per thread, all in flight
    // before the first shared-memory store (the loop was a chain of 36 dependent-latency scalar loads)
    constexpr int ROW4 = PK_T * 9 / 4;                           // float4 per row
    float4 v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int i = threadIdx.x + k * 256;                     // 9 * 256 == 32 * ROW4
        const int o = i / ROW4, r4 = i - o * ROW4;
        v[k] = __ldg(reinterpret_cast<const float4*>(L.src + ((int64_t)(o0 + o) * L.cin + c0) * 9) + r4);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int i = threadIdx.x + k * 256;
        const int o = i / ROW4, r = (i - o * ROW4) * 4;
        tile[o][r] = v[k].x; tile[o][r + 1] = v[k].y; tile[o][r + 2] = v[k].z; tile[o][r + 3] = v[k].w;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (int q = grp; q < PK_T * 9; q += 8) {               // forward layout: (o, tap) rows of 32 consecutive cin
        const int o = q / 9, tap = q - o * 9;
        L.fwd[((int64_t)(o0 + o) * 9 + tap) * L.cin + c0 + lane] = __float2bfloat16_rn(tile[o][lane * 9 + tap]);
    }
    if (L.dgrad) {
        for (int q = grp; q < PK_T * 9; q += 8) {           // dgrad layout: (cin, flipped tap) rows of 32 consecutive cout
            const int c = q / 9, tap = q - c * 9;
            L.dgrad[((int64_t)(c0 + c) * 9 + (8 - tap)) * L.cout + o0 + lane] = __float2bfloat16_rn(tile[lane][c * 9 + tap]);
        }
    }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, float eps, float* __restrict__ scale,
                               float* __restrict__ shift, int c) {
    td::pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] + ((conv_bias ? conv_bias[i] : 0.f) - mean[i]) * s;
}

static inline int grid_for(int64_t items) {
    int64_t blocks = ceil_div(items, kThreads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

}  // namespace td

using namespace td;

extern "C" int td_maxpool2_fwd(const void* x, void* y, int dtype, int batch, int h, int w, int c, int ceil_mode,
                               void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && y && batch > 0 && h > 0 && w > 0 && c > 0, "td_maxpool2_fwd: bad args");
    const int ho = ceil_mode ? (h + 1) / 2 : h / 2, wo = ceil_mode ? (w + 1) / 2 : w / 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TD_BF16) {
        TD_CHECK_ARG(c % 8 == 0, "td_maxpool2_fwd: channels must be a multiple of 8 for bf16");
        td::launch(maxpool2_kernel<__nv_bfloat16>, td::LaunchCfg(batch * ho, row_block(c / 8), 0, s), 
            (const __nv_bfloat16*)x, (__nv_bfloat16*)y, batch, h, w, c, ho, wo);
    } else if (dtype == TD_F32) {
        TD_CHECK_ARG(c % 4 == 0, "td_maxpool2_fwd: channels must be a multiple of 4 for fp32");
        td::launch(maxpool2_kernel<float>, td::LaunchCfg(batch * ho, row_block(c / 4), 0, s), 
            (const float*)x, (float*)y, batch, h, w, c, ho, wo);
    } else {
        TD_CHECK_ARG(false, "td_maxpool2_fwd: unknown dtype %d", dtype);
    }
    return launch_status("maxpool2");
}

extern "C" int td_upcat_fwd(const void* low, const void* skip, const float* temb, int ld_temb, int temb_off,
                            void* out, int dtype, int batch, int ho, int wo, int cu, int hs, int ws, int cs,
                            void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(low && skip && temb && out, "td_upcat_fwd: null pointer");
    TD_CHECK_ARG(batch > 0 && ho > 0 && wo > 0 && ho % 2 == 0 && wo % 2 == 0 && wo <= kMaxRowW, "td_upcat_fwd: bad output size");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TD_BF16) {
        TD_CHECK_ARG(cu % 8 == 0 && cs % 8 == 0, "td_upcat_fwd: channel counts must be multiples of 8");
        const int cvh = std::max(cu, cs) / 8;
        const dim3 blk = walk_block(cvh);
        const int nseg = walk_segments((int64_t)batch * ho * cvh * 2, wo);
        const dim3 grd((unsigned)(ceil_div((int64_t)batch * ho, blk.y) * nseg), (unsigned)ceil_div(cvh, blk.x), 2);
        td::launch(upcat_kernel<__nv_bfloat16>, td::LaunchCfg(grd, blk, 0, s), 
            (const __nv_bfloat16*)low, (const __nv_bfloat16*)skip, temb, ld_temb, temb_off, (__nv_bfloat16*)out,
            batch, ho, wo, cu, hs, ws, cs, nseg);
    } else if (dtype == TD_F32) {
        TD_CHECK_ARG(cu % 4 == 0 && cs % 4 == 0, "td_upcat_fwd: channel counts must be multiples of 4");
        const int cvh = std::max(cu, cs) / 4;
        const dim3 blk = walk_block(cvh);
        const int nseg = walk_segments((int64_t)batch * ho * cvh * 2, wo);
        const dim3 grd((unsigned)(ceil_div((int64_t)batch * ho, blk.y) * nseg), (unsigned)ceil_div(cvh, blk.x), 2);
        td::launch(upcat_kernel<float>, td::LaunchCfg(grd, blk, 0, s), 
            (const float*)low, (const float*)skip, temb, ld_temb, temb_off, (float*)out, batch, ho, wo, cu, hs, ws, cs, nseg);
    } else {
        TD_CHECK_ARG(false, "td_upcat_fwd: unknown dtype %d", dtype);
    }
    return launch_status("upcat");
}

extern "C" int td_resize_bilinear_fwd(const void* x, void* y, int dtype, int batch, int hi, int wi, int ho, int wo,
                                      int c, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && y && batch > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && wo <= kMaxRowW && c > 0, "td_resize_bilinear_fwd: bad args");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TD_BF16) {
        TD_CHECK_ARG(c % 8 == 0, "td_resize_bilinear_fwd: channels must be a multiple of 8 for bf16");
        const dim3 blk = walk_block(c / 8);
        const int nseg = walk_segments((int64_t)batch * ho * (c / 8), wo);
        const dim3 grd((unsigned)(ceil_div((int64_t)batch * ho, blk.y) * nseg), (unsigned)ceil_div(c / 8, blk.x), 1);
        td::launch(resize_kernel<__nv_bfloat16>, td::LaunchCfg(grd, blk, 0, s), 
            (const __nv_bfloat16*)x, (__nv_bfloat16*)y, batch, hi, wi, ho, wo, c, nseg);
    } else if (dtype == TD_F32) {
        TD_CHECK_ARG(c % 4 == 0, "td_resize_bilinear_fwd: channels must be a multiple of 4 for fp32");
        const dim3 blk = walk_block(c / 4);
        const int nseg = walk_segments((int64_t)batch * ho * (c / 4), wo);
        const dim3 grd((unsigned)(ceil_div((int64_t)batch * ho, blk.y) * nseg), (unsigned)ceil_div(c / 4, blk.x), 1);
        td::launch(resize_kernel<float>, td::LaunchCfg(grd, blk, 0, s), (const float*)x, (float*)y, batch, hi, wi, ho, wo, c, nseg);
    } else {
        TD_CHECK_ARG(false, "td_resize_bilinear_fwd: unknown dtype %d", dtype);
    }
    return launch_status("resize_bilinear");
}

extern "C" int td_pack_conv_weight(const float* oihw, void* ohwi, int out_dtype, int cout, int cin, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(oihw && ohwi && cout > 0 && cin > 0, "td_pack_conv_weight: bad args");
    const int64_t n = (int64_t)cout * cin * 9;
    if (out_dtype == TD_BF16)
        td::launch(pack_weight_kernel<__nv_bfloat16>, td::LaunchCfg(grid_for(n), kThreads, 0, (cudaStream_t)stream), oihw, (__nv_bfloat16*)ohwi, cout, cin);
    else if (out_dtype == TD_F32)
        td::launch(pack_weight_kernel<float>, td::LaunchCfg(grid_for(n), kThreads, 0, (cudaStream_t)stream), oihw, (float*)ohwi, cout, cin);
    else
        TD_CHECK_ARG(false, "td_pack_conv_weight: unknown dtype %d", out_dtype);
    return launch_status("pack_conv_weight");
}

extern "C" int td_pack_conv_weights_multi(const void* table, int n_entries, int total_tiles, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(table && n_entries > 0 && total_tiles > 0, "td_pack_conv_weights_multi: bad args");
    td::launch(pack_weights_multi_kernel, td::LaunchCfg(total_tiles, 256, 0, (cudaStream_t)stream), reinterpret_cast<const PackEntry*>(table), n_entries);
    return launch_status("pack_conv_weights_multi");
}

extern "C" int td_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                          const float* conv_bias, float eps, float* scale, float* shift, int c, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(gamma && beta && mean && var && scale && shift && c > 0, "td_bn_fold: bad args");
    td::launch(bn_fold_kernel, td::LaunchCfg((c + 127) / 128, 128, 0, (cudaStream_t)stream), gamma, beta, mean, var, conv_bias, eps, scale, shift, c);
    return launch_status("bn_fold");
}
