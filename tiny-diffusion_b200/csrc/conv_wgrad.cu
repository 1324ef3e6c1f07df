// Weight gradient of the 3x3 convolutions (backward of diffusion.py:28-98):
//   dW[o, c, ky, kx] = sum_{b,h,w} dY[b,h,w,o] * X[b, h+ky-1, w+kx-1, c]
// as an implicit GEMM with M = Cout, N = 9*Cin, K = B*H*W pixels, split along K across CTAs.
// Partials go to a workspace [splits][Cout][9*Cin]; a second kernel sums them in fixed order
// (deterministic) and writes PyTorch's OIHW layout.
//   TD_CONV_SIMT : fp32 FFMA, any shape / dtype / NCHW boundary tensors (parity path, first/last conv)
//   TD_CONV_TC   : tcgen05, conv_wgrad_tc.cu
#include "conv_plan.h"

namespace td {

constexpr int WG_BM = 64, WG_BN = 64, WG_BK = 16, WG_THREADS = 256;

template <typename Tx, typename Tdy>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_simt_kernel(const td_wgrad_desc d, int pixels_per_split) {
    td::pdl_sync();
    __shared__ float As[WG_BK][WG_BM + 4];   // dY  [k = pixel][i = cout]
    __shared__ float Bs[WG_BK][WG_BN + 4];   // X   [k = pixel][j = tap*cin + c]
    const Tx* __restrict__ x = reinterpret_cast<const Tx*>(d.x);
    const Tdy* __restrict__ dy = reinterpret_cast<const Tdy*>(d.dy);
    const int N = 9 * d.cin;
    const int64_t P = (int64_t)d.batch * d.height * d.width;
    const int n0 = blockIdx.x * WG_BN, m0 = blockIdx.y * WG_BM;
    const int64_t pbeg = (int64_t)blockIdx.z * pixels_per_split;
    const int64_t pend = min(P, pbeg + pixels_per_split);
    const int tid = threadIdx.x;
    const int li = tid & 63, lk = tid >> 6;          // loader: column li, pixel rows lk, lk+4, lk+8, lk+12
    const int ty = tid >> 4, tx = tid & 15;

    // the B column this thread loads is fixed: tap / channel decomposition once
    const int jn = n0 + li;
    const bool jvalid = jn < N;
    const int tap = jvalid ? jn / d.cin : 0, cc = jvalid ? jn - tap * d.cin : 0;
    const int ddy = tap / 3 - 1, ddx = tap % 3 - 1;
    const int io = m0 + li;
    const bool ivalid = io < d.cout;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t p0 = pbeg; p0 < pend; p0 += WG_BK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int k = lk + l * 4;
            const int64_t p = p0 + k;
            float av = 0.f, bv = 0.f;
            if (p < pend) {
                const int w_ = (int)(p % d.width);
                const int64_t r = p / d.width;
                const int h_ = (int)(r % d.height);
                const int b_ = (int)(r / d.height);
                if (ivalid) {
                    av = d.dy_nchw ? to_f32(dy[(((int64_t)b_ * d.cout + io) * d.height + h_) * d.width + w_])
                                   : to_f32(dy[p * d.lddy + d.dy_coff + io]);
                }
                const int hh = h_ + ddy, ww = w_ + ddx;
                if (jvalid && hh >= 0 && hh < d.height && ww >= 0 && ww < d.width) {
                    bv = d.x_nchw ? to_f32(x[(((int64_t)b_ * d.cin + cc) * d.height + hh) * d.width + ww])
                                  : to_f32(x[(((int64_t)b_ * d.height + hh) * d.width + ww) * d.ldx + d.x_coff + cc]);
                }
            }
            As[k][li] = av;
            Bs[k][li] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WG_BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* ws = d.workspace + (int64_t)blockIdx.z * d.cout * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = m0 + ty * 4 + i;
        if (o >= d.cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) ws[(int64_t)o * N + n] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Network-boundary layers: one side of the product has 1..4 channels (NCHW fp32: the network input
// of initial_conv, or d(eps) of final_conv), the other is a wide NHWC tensor.  Pure bandwidth: the
// wide tensor is read once with 16-byte vectors, the narrow one comes from L1/L2 as scalars.
//   final_conv  (narrow = dY, channel s = o):  dW[s][tap][c] = sum_p X[p, c]  * dY[p - off(tap), s]
//   initial_conv(narrow = X,  channel s = i):  dW[c][tap][s] = sum_p dY[p, c] * X[p + off(tap), s]
// grid = (CTAs over pixels, narrow channels); CTA partial -> ws[blockIdx.x][cout][9][cin].
// ---------------------------------------------------------------------------------------------
// A thread (channel vector c0, slot) walks whole image rows left to right with the 3 x 3 window of the narrow channel in registers:
// three new scalars and one 16-byte vector per pixel feed 9 * V FMAs, no per-tap bounds checks or index arithmetic (the per-pixel
// version spent ~150 instructions on 72 FMAs).
template <typename T, bool NARROW_IS_DY>
__global__ void __launch_bounds__(256)
wgrad_narrow_kernel(const T* __restrict__ wide, int ldw, int w_coff, int cw, const float* __restrict__ narrow, int cn,
                    int B, int H, int W, int cin, int cout, float* __restrict__ ws) {
    td::pdl_sync();
    constexpr int V = Vec<T>::N;
    constexpr int SGN = NARROW_IS_DY ? -1 : 1;
    constexpr bool narrow_is_dy = NARROW_IS_DY;
    extern __shared__ float red[];                  // [rows][cw]
    const int lanesC = cw / V;
    const int rows = 256 / lanesC;
    const int lane = threadIdx.x % lanesC, row = threadIdx.x / lanesC;
    const int c0 = lane * V;
    const int s = blockIdx.y;
    float acc[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[t][k] = 0.f;
    if (row < rows) {
        const int64_t R = (int64_t)B * H;           // image rows
        for (int64_t it = (int64_t)blockIdx.x * rows + row; it < R; it += (int64_t)gridDim.x * rows) {
            const int h_ = (int)(it % H);
            const int b_ = (int)(it / H);
            const float* nb = narrow + ((int64_t)b_ * cn + s) * H * W;
            const T* wp = wide + it * W * ldw + w_coff + c0;
            const float* nr[3];                     // window rows r = 0, 1, 2 <-> image rows h_ - 1, h_, h_ + 1
            bool rv[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int hh = h_ + r - 1;
                rv[r] = hh >= 0 && hh < H;
                nr[r] = nb + (rv[r] ? hh : 0) * W;
            }
            float win[3][3];                        // window columns w_ - 1, w_, w_ + 1
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                win[r][0] = 0.f;
                win[r][1] = rv[r] ? __ldg(nr[r]) : 0.f;
                win[r][2] = (rv[r] && W > 1) ? __ldg(nr[r] + 1) : 0.f;
            }
            Vec<T> cur = Vec<T>::load(wp);
            for (int w_ = 0; w_ < W; ++w_) {
                float nxt[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) nxt[r] = (rv[r] && w_ + 2 < W) ? __ldg(nr[r] + w_ + 2) : 0.f;
                Vec<T> nv = cur;
                if (w_ + 1 < W) nv = Vec<T>::load(wp + (int64_t)(w_ + 1) * ldw);
                float f[V];
                cur.unpack(f);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float x = win[1 + SGN * (t / 3 - 1)][1 + SGN * (t % 3 - 1)];      // narrow at (h_, w_) + SGN * off(tap)
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[t][k] = fmaf(f[k], x, acc[t][k]);
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) { win[r][0] = win[r][1]; win[r][1] = win[r][2]; win[r][2] = nxt[r]; }
                cur = nv;
            }
        }
    }
    // cross-row reduction, one tap at a time through a [rows][cw] staging buffer (fixed order)
    float* out = ws + (int64_t)blockIdx.x * cout * 9 * cin;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        if (row < rows) {
#pragma unroll
            for (int k = 0; k < V; ++k) red[(size_t)row * cw + c0 + k] = acc[t][k];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < cw; c += 256) {
            float tsum = 0.f;
            for (int r = 0; r < rows; ++r) tsum += red[(size_t)r * cw + c];
            if (narrow_is_dy) out[((int64_t)s * 9 + t) * cin + c] = tsum;          // o = s, cin = cw
            else out[((int64_t)c * 9 + t) * cin + s] = tsum;                       // o = c, cin = cn
        }
        __syncthreads();
    }
}

static bool narrow_ok(const td_wgrad_desc& d, int* grid_x) {
    const bool fin = d.dy_nchw && d.dy_dtype == TD_F32 && d.cout <= 4 && !d.x_nchw;       // final_conv
    const bool ini = d.x_nchw && d.x_dtype == TD_F32 && d.cin <= 4 && !d.dy_nchw;         // initial_conv
    if (!fin && !ini) return false;
    const int cw = fin ? d.cin : d.cout;
    const int wdt = fin ? d.x_dtype : d.dy_dtype;
    const int V = wdt == TD_BF16 ? 8 : 4;
    if (cw % V != 0) return false;
    const int lanesC = cw / V;
    if (lanesC > 256 || (lanesC & (lanesC - 1)) != 0) return false;
    const int rows = 256 / lanesC;
    if ((size_t)rows * cw * sizeof(float) > 48 * 1024) return false;
    const int64_t R = (int64_t)d.batch * d.height;              // one image row per slot and pass
    *grid_x = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, (int64_t)rows), kNumSMs));
    return true;
}

// dw[o][c][tap] = sum_s ws[s][o][tap*cin + c]     (OHWI partials -> OIHW gradient)
// grid (cin chunks of 32, cout), 288 threads = 4 split groups x 9 taps x 8 channel quads: every thread sums one float4
// (4 input channels of one tap) over a contiguous quarter of the splits with all of its loads in flight at once (the
// partials were just written and sit in the L2: the pass is a latency chain, not bandwidth), the four groups are combined in
// fixed order through shared memory -- deterministic -- and the [c][tap] run is written contiguously (OHWI -> OIHW transpose).
constexpr int WR_C = 32;
constexpr int WR_G = 4;
__global__ void __launch_bounds__(9 * WR_C)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int cout, int cin, float* __restrict__ dw) {
    td::pdl_sync();
    __shared__ float4 part[WR_G][9][WR_C / 4];
    __shared__ float tile[9][WR_C + 1];
    const int o = blockIdx.y;
    const int c0 = blockIdx.x * WR_C;
    const int nc = min(WR_C, cin - c0);
    const int64_t per = (int64_t)cout * 9 * cin;
    const int q = threadIdx.x % (WR_C / 4), t = (threadIdx.x / (WR_C / 4)) % 9, g = threadIdx.x / (9 * WR_C / 4);
    const int per_g = (splits + WR_G - 1) / WR_G;
    const int z0 = g * per_g, z1 = min(splits, z0 + per_g);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q * 4 < nc) {
        const float* src = ws + ((int64_t)o * 9 + t) * cin + c0 + q * 4;
        for (int z = z0; z < z1; z += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (z + k < z1) ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)(z + k) * per)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        }
    }
    part[g][t][q] = acc;
    __syncthreads();
    if (g == 0 && q * 4 < nc) {
        float4 s4 = part[0][t][q];
#pragma unroll
        for (int k = 1; k < WR_G; ++k) { const float4 v = part[k][t][q]; s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w; }
        tile[t][q * 4] = s4.x; tile[t][q * 4 + 1] = s4.y; tile[t][q * 4 + 2] = s4.z; tile[t][q * 4 + 3] = s4.w;
    }
    __syncthreads();
    float* dst = dw + ((int64_t)o * cin + c0) * 9;
    for (int e = threadIdx.x; e < nc * 9; e += 9 * WR_C) dst[e] = tile[e % 9][e / 9];
}

// scalar variant for channel counts that are not multiples of 4 (the 1- and 4-channel boundary convolutions)
__global__ void __launch_bounds__(9 * WR_C)
wgrad_reduce_scalar_kernel(const float* __restrict__ ws, int splits, int cout, int cin, float* __restrict__ dw) {
    td::pdl_sync();
    __shared__ float tile[9][WR_C + 1];
    const int o = blockIdx.y;
    const int c0 = blockIdx.x * WR_C;
    const int nc = min(WR_C, cin - c0);
    const int64_t per = (int64_t)cout * 9 * cin;
    const int t = threadIdx.x / WR_C, c = threadIdx.x % WR_C;
    if (c < nc) {
        const float* src = ws + ((int64_t)o * 9 + t) * cin + c0 + c;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {
            const float v0 = __ldg(src + (int64_t)z * per), v1 = __ldg(src + (int64_t)(z + 1) * per);
            const float v2 = __ldg(src + (int64_t)(z + 2) * per), v3 = __ldg(src + (int64_t)(z + 3) * per);
            a0 += v0; a1 += v1; a2 += v2; a3 += v3;
        }
        for (; z < splits; ++z) a0 += src[(int64_t)z * per];
        tile[t][c] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    float* dst = dw + ((int64_t)o * cin + c0) * 9;
    for (int e = threadIdx.x; e < nc * 9; e += 9 * WR_C) dst[e] = tile[e % 9][e / 9];
}

// OIHW fp32 -> [Cin][3][3][Cout] with the taps flipped: the weight operand of the data gradient,
// dX = conv3x3(dY, Wd),  Wd[i][ky][kx][o] = W[o][i][2-ky][2-kx]
template <typename T>
__global__ void __launch_bounds__(256)
pack_weight_dgrad_kernel(const float* __restrict__ oihw, T* __restrict__ out, int cout, int cin) {
    td::pdl_sync();
    const int64_t total = (int64_t)cout * 9 * cin;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(e % cout);
        const int64_t r = e / cout;
        const int tap = (int)(r % 9);
        const int i = (int)(r / 9);
        out[e] = from_f32<T>(oihw[((int64_t)o * cin + i) * 9 + (8 - tap)]);
    }
}

static int simt_splits(const td_wgrad_desc& d) {
    int gx = 0;
    if (narrow_ok(d, &gx)) return gx;
    const int64_t P = (int64_t)d.batch * d.height * d.width;
    const int64_t base = ceil_div(9 * d.cin, WG_BN) * ceil_div(d.cout, WG_BM);
    int64_t s = ceil_div(2 * kNumSMs, base);
    s = std::min<int64_t>(s, std::max<int64_t>(1, P / 256));
    return (int)std::max<int64_t>(1, std::min<int64_t>(s, 512));
}

}  // namespace td

using namespace td;

extern "C" int64_t td_conv3x3_wgrad_workspace(const td_wgrad_desc* d, int engine) {
    if (!d) return 0;
    const int64_t per = (int64_t)d->cout * 9 * d->cin;
    if (engine == TD_CONV_TC) return (int64_t)wgrad_tc_splits(*d) * per;
    return (int64_t)simt_splits(*d) * per;
}

extern "C" int td_conv3x3_wgrad_plan_create(td_wgrad_plan** plan, const td_wgrad_desc* desc, int engine) {
    TD_CHECK_ARG(plan && desc, "wgrad plan: null pointer");
    const td_wgrad_desc& d = *desc;
    TD_CHECK_ARG(d.batch > 0 && d.height > 0 && d.width > 0 && d.cin > 0 && d.cout > 0, "wgrad plan: bad sizes");
    TD_CHECK_ARG(d.x && d.dy && d.dw && d.workspace, "wgrad plan: null tensor pointer");
    td_wgrad_plan* p = new td_wgrad_plan();
    memset(p, 0, sizeof(*p));
    p->d = d;
    p->engine = engine;
    int st = TD_OK;
    if (engine == TD_CONV_TC) {
        st = wgrad_tc_plan_init(p);
    } else if (engine == TD_CONV_SIMT) {
        p->splits = simt_splits(d);
    } else {
        set_error("wgrad plan: unknown engine %d", engine);
        st = TD_ERR_ARG;
    }
    if (st != TD_OK) { delete p; return st; }
    *plan = p;
    return TD_OK;
}

extern "C" int td_conv3x3_wgrad_run(const td_wgrad_plan* p, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(p, "td_conv3x3_wgrad_run: null plan");
    cudaStream_t s = (cudaStream_t)stream;
    const td_wgrad_desc& d = p->d;
    if (p->engine == TD_CONV_TC) {
        int st = wgrad_tc_plan_run(p, s);
        if (st != TD_OK) return st;
    } else if (int gx = 0; narrow_ok(d, &gx)) {
        const bool fin = d.dy_nchw != 0;
        const int cw = fin ? d.cin : d.cout, cn = fin ? d.cout : d.cin;
        const int wdt = fin ? d.x_dtype : d.dy_dtype;
        const int V = wdt == TD_BF16 ? 8 : 4;
        const size_t smem = (size_t)(256 / (cw / V)) * cw * sizeof(float);
        const void* wide = fin ? d.x : d.dy;
        const float* narrow = (const float*)(fin ? d.dy : d.x);
        const int ldw = fin ? d.ldx : d.lddy, wcoff = fin ? d.x_coff : d.dy_coff;
        dim3 grid((unsigned)gx, (unsigned)cn);
#define TD_NARROW(T, FIN) td::launch(wgrad_narrow_kernel<T, FIN>, td::LaunchCfg(grid, 256, smem, s), (const T*)wide, ldw, wcoff, cw, narrow, \
                                     cn, d.batch, d.height, d.width, d.cin, d.cout, d.workspace)
        if (wdt == TD_BF16) { if (fin) TD_NARROW(__nv_bfloat16, true); else TD_NARROW(__nv_bfloat16, false); }
        else { if (fin) TD_NARROW(float, true); else TD_NARROW(float, false); }
#undef TD_NARROW
        int st = launch_status("wgrad_narrow");
        if (st != TD_OK) return st;
    } else {
        const int64_t P = (int64_t)d.batch * d.height * d.width;
        const int per = (int)(ceil_div(ceil_div(P, p->splits), WG_BK) * WG_BK);
        dim3 grid((unsigned)ceil_div(9 * d.cin, WG_BN), (unsigned)ceil_div(d.cout, WG_BM), (unsigned)p->splits);
        if (d.x_dtype == TD_F32 && d.dy_dtype == TD_F32) td::launch(wgrad_simt_kernel<float, float>, td::LaunchCfg(grid, WG_THREADS, 0, s), d, per);
        else if (d.x_dtype == TD_BF16 && d.dy_dtype == TD_BF16) td::launch(wgrad_simt_kernel<__nv_bfloat16, __nv_bfloat16>, td::LaunchCfg(grid, WG_THREADS, 0, s), d, per);
        else if (d.x_dtype == TD_BF16 && d.dy_dtype == TD_F32) td::launch(wgrad_simt_kernel<__nv_bfloat16, float>, td::LaunchCfg(grid, WG_THREADS, 0, s), d, per);
        else td::launch(wgrad_simt_kernel<float, __nv_bfloat16>, td::LaunchCfg(grid, WG_THREADS, 0, s), d, per);
        int st = launch_status("wgrad_simt");
        if (st != TD_OK) return st;
    }
    dim3 rgrid((unsigned)ceil_div(d.cin, WR_C), (unsigned)d.cout);
    if (d.cin % 4 == 0 && (((uintptr_t)d.workspace) & 15) == 0)
        td::launch(wgrad_reduce_kernel, td::LaunchCfg(rgrid, 9 * WR_C, 0, s), d.workspace, p->splits, d.cout, d.cin, d.dw);
    else
        td::launch(wgrad_reduce_scalar_kernel, td::LaunchCfg(rgrid, 9 * WR_C, 0, s), d.workspace, p->splits, d.cout, d.cin, d.dw);
    return launch_status("wgrad_reduce");
}

extern "C" void td_conv3x3_wgrad_plan_destroy(td_wgrad_plan* p) { delete p; }

extern "C" int td_pack_conv_weight_dgrad(const float* oihw, void* out, int out_dtype, int cout, int cin, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(oihw && out && cout > 0 && cin > 0, "td_pack_conv_weight_dgrad: bad args");
    const int64_t n = (int64_t)cout * cin * 9;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), kNumSMs * 16));
    if (out_dtype == TD_BF16)
        td::launch(pack_weight_dgrad_kernel<__nv_bfloat16>, td::LaunchCfg(grid, 256, 0, (cudaStream_t)stream), oihw, (__nv_bfloat16*)out, cout, cin);
    else if (out_dtype == TD_F32)
        td::launch(pack_weight_dgrad_kernel<float>, td::LaunchCfg(grid, 256, 0, (cudaStream_t)stream), oihw, (float*)out, cout, cin);
    else
        TD_CHECK_ARG(false, "td_pack_conv_weight_dgrad: unknown dtype %d", out_dtype);
    return launch_status("pack_conv_weight_dgrad");
}
