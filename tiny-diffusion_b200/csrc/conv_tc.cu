// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, c} X[pixel + tap_offset, c] * W[cout, tap, c]
//
//   * M tile  = 128 TMEM lanes = one spatial box (bw x bh x bn output pixels, <= 128) of the NHWC
//               activation tensor; N tile = BLOCK_N output channels; K loop = 9 taps x Cin/64.
//   * A operand: one 4-D TMA box load per (tap, 64-channel chunk) at coordinates shifted by the
//               tap offset; the zero padding of the convolution is TMA's out-of-bounds zero fill.
//               The box lands as 128-byte rows with SWIZZLE_128B = the canonical K-major UMMA layout.
//   * B operand: 2-D TMA box (64 k x BLOCK_N couts) of the OHWI bf16 weight matrix, same layout.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BLOCK_N, K=16) issued by one elected thread,
//               fp32 accumulator in TMEM; smem slots recycled through tcgen05.commit -> mbarrier.
//   * Epilogue: 4 warps read their TMEM lane quadrant with tcgen05.ld, apply the folded
//               BatchNorm/bias affine (+ReLU) in fp32 and store NHWC bf16/fp32 rows.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (warp 2 also owns TMEM alloc).
#include <mutex>

#include "conv_plan.h"
#include "sm100.cuh"

namespace td {

using namespace sm100;

struct TcParams {
    int B, H, W, cin, x_coff;
    int cout, ldy, y_coff, y_dtype;
    void* y;
    const float* scale;
    const float* shift;
    int relu;
    int bw, bh, bn, tiles_w, tiles_h;
    int stages;
    uint32_t a_bytes;
    float* stats;             // train-mode BatchNorm partials [m tiles][2][cout] (+ zero K row), or NULL
    int split_iters;          // K-loop iterations (tap x 64-channel chunk) per split; == 9*cin/64 without split-K
    float* ws;                // split-K partials [gridDim.z][B*H*W][cout] fp32 (NULL without split-K)
};

constexpr int TC_THREADS = 192;

constexpr int TC_A_STAGE = 128 * 128;   // 128 rows x 128 bytes (64 bf16)

template <int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_STAGE = BLOCK_N * 128;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);     // SWIZZLE_128B needs 1024-byte alignment
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * TC_A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * B_STAGE);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    // [BLOCK_N] epilogue affine, staged once per CTA (16-byte aligned for float4 reads)
    float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 1) + 15) & ~(uintptr_t)15);
    float* s_shift = s_scale + BLOCK_N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int mt = blockIdx.x;
    const int tw = mt % p.tiles_w;
    const int th = (mt / p.tiles_w) % p.tiles_h;
    const int tn = mt / (p.tiles_w * p.tiles_h);
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
    const int nt = blockIdx.y;
    const int kchunks = p.cin >> 6;
    const int it_begin = blockIdx.z * p.split_iters;
    const int iters = min(9 * kchunks - it_begin, p.split_iters);     // this CTA's share of the K loop

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<BLOCK_N>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    td::pdl_sync();              // everything above is on-chip setup; global memory is touched only below
    if (warp >= 2) {
        for (int c = threadIdx.x - 64; c < BLOCK_N; c += TC_THREADS - 64) {
            s_scale[c] = p.scale ? __ldg(p.scale + nt * BLOCK_N + c) : 1.f;
            s_shift[c] = p.shift ? __ldg(p.shift + nt * BLOCK_N + c) : 0.f;
        }
    }
    __syncthreads();

    if (warp == 0) {
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], p.a_bytes + (uint32_t)B_STAGE);
                const int git = it_begin + it;
                const int tap = git / kchunks, cc = git - tap * kchunks;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                tma_load_4d(smem_a + (size_t)s * TC_A_STAGE, &tmap_x, &full_bar[s], p.x_coff + cc * 64, w0 + dx, h0 + dy, n0);
                tma_load_2d(smem_b + (size_t)s * B_STAGE, &tmap_w, &full_bar[s], tap * p.cin + cc * 64, nt * BLOCK_N);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + (size_t)s * TC_A_STAGE);
                const uint32_t b_addr = smem_u32(smem_b + (size_t)s * B_STAGE);
#pragma unroll
                for (int k = 0; k < 4; ++k) {       // 4 x (K = 16 bf16 = 32 bytes) per 128-byte swizzle row
                    const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                    umma_bf16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs retire
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ---- epilogue: TMEM -> registers -> global ------------------------------------------------
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int rows_valid = p.bw * p.bh * p.bn;
        const int rw = row % p.bw;
        const int rh = (row / p.bw) % p.bh;
        const int rn = row / (p.bw * p.bh);
        const int w_ = w0 + rw, h_ = h0 + rh, n_ = n0 + rn;
        const bool valid = row < rows_valid && w_ < p.W && h_ < p.H && n_ < p.B;
        const int64_t pix = ((int64_t)n_ * p.H + h_) * p.W + w_;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        // Output staging (as in conv_halo.cu): a thread owns one position, so direct stores would touch 32 cache lines per
        // warp instruction.  Every 32 x 32 chunk is transposed through a padded tile -- in the pipeline stages, idle by now --
        // and leaves as whole 128-byte (fp32) / 64-byte (bf16) row segments.
        float* tile = reinterpret_cast<float*>(smem_a) + 8 * BLOCK_N + q * (32 * 36);
        const bool out_bf16 = !p.ws && p.y_dtype == TD_BF16;
        float* const out_f32 = p.ws ? p.ws + (int64_t)blockIdx.z * ((int64_t)p.B * p.H * p.W) * p.cout : reinterpret_cast<float*>(p.y);
        const int64_t ld_out = p.ws ? p.cout : p.ldy;
        const int64_t my_off = valid ? pix * ld_out + (p.ws ? 0 : p.y_coff) + nt * BLOCK_N : (int64_t)-1;
        const int rpi = out_bf16 ? 8 : 4, lpr = 32 / rpi;          // rows per store instruction, lanes per row
        int64_t row_off[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rw_ = it * rpi + lane / lpr;
            const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)(my_off & 0xffffffffll), rw_ & 31);
            const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)my_off >> 32), rw_ & 31);
            row_off[it] = (int64_t)(((uint64_t)hi << 32) | lo);
        }
        auto store_chunk = [&](const float* v, int c0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(tile + lane * 36 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            __syncwarp();
            if (out_bf16) {
                const int col = (lane & 3) * 8;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rw_ = it * 8 + (lane >> 2);
                    const float4 a = *reinterpret_cast<const float4*>(tile + rw_ * 36 + col);
                    const float4 b = *reinterpret_cast<const float4*>(tile + rw_ * 36 + col + 4);
                    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    if (row_off[it] >= 0) Vec<__nv_bfloat16>::pack(f).store(reinterpret_cast<__nv_bfloat16*>(p.y) + row_off[it] + c0 + col);
                }
            } else {
                const int col = (lane & 7) * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rw_ = it * 4 + (lane >> 3);
                    const float4 a = *reinterpret_cast<const float4*>(tile + rw_ * 36 + col);
                    if (row_off[it] >= 0) *reinterpret_cast<float4*>(out_f32 + row_off[it] + c0 + col) = a;
                }
            }
            __syncwarp();
        };
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (p.ws) {          // split-K: raw fp32 partial, the affine / activation run in the reduce kernel
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                store_chunk(v, c0);
                continue;
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + j);
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j);
                v[j] = fmaf(__uint_as_float(r[j]), sc.x, sh.x);
                v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
                v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z);
                v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            store_chunk(v, c0);
            if (p.stats) {
                // per-channel sum / sum of squares over this warp's 32 rows, from the fp32 accumulators;
                // the pipeline stages are idle by now, so their shared memory stages the 4 warps' partials
                float sq[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = valid ? v[j] : 0.f;
                    sq[j] = v[j] * v[j];
                }
                const float cs = warp_column_sums(v, lane);
                const float cq = warp_column_sums(sq, lane);
                float* st = reinterpret_cast<float*>(smem_a);
                st[(q * 2 + 0) * BLOCK_N + c0 + lane] = cs;
                st[(q * 2 + 1) * BLOCK_N + c0 + lane] = cq;
            }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
            const float* st = reinterpret_cast<const float*>(smem_a);
            const int tid = threadIdx.x - 64;
            float* row = p.stats + (int64_t)blockIdx.x * 2 * p.cout + nt * BLOCK_N;
            for (int c = tid; c < 2 * BLOCK_N; c += 128) {
                const int which = c / BLOCK_N, cc = c - which * BLOCK_N;
                const float t = (st[(0 * 2 + which) * BLOCK_N + cc] + st[(1 * 2 + which) * BLOCK_N + cc]) +
                                (st[(2 * 2 + which) * BLOCK_N + cc] + st[(3 * 2 + which) * BLOCK_N + cc]);
                row[which * p.cout + cc] = t;
            }
            if (blockIdx.x == 0)           // the finalize kernel reads a "shift" row K after the partial rows: zero here
                for (int c = tid; c < BLOCK_N; c += 128) p.stats[(int64_t)gridDim.x * 2 * p.cout + nt * BLOCK_N + c] = 0.f;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<BLOCK_N>(tmem_base);
    }
}

// split-K second pass: y = act(scale * sum_s ws[s] + shift), fixed summation order
template <typename Tout>
__global__ void __launch_bounds__(256)
conv_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t pixels, int cout, const float* __restrict__ scale,
                          const float* __restrict__ shift, int relu, Tout* __restrict__ y, int ldy, int y_coff) {
    td::pdl_sync();
    const int cq = cout >> 2;
    const int64_t total = pixels * cq;
    const int64_t slice = pixels * cout;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = e / cq;
        const int c = (int)(e - pix * cq) * 4;
        float4 a = *reinterpret_cast<const float4*>(ws + pix * cout + c);
        for (int z = 1; z < splits; ++z) {
            const float4 b = *reinterpret_cast<const float4*>(ws + z * slice + pix * cout + c);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = fmaf(v[k], scale ? scale[c + k] : 1.f, shift ? shift[c + k] : 0.f);
            if (relu) v[k] = fmaxf(v[k], 0.f);
        }
        Tout* dst = y + pix * ldy + y_coff + c;
        if constexpr (sizeof(Tout) == 2) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dst) = u;
        } else {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
EncodeTiledFn tc_get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

static int pick_block_n(int cout) {
    if (const char* e = getenv("TD_TC_BLOCK_N")) {
        int v = atoi(e);
        if ((v == 64 || v == 128 || v == 256) && cout % v == 0) return v;
    }
    if (cout % 256 == 0) return 256;
    if (cout % 128 == 0) return 128;
    return 64;
}

int tc_plan_init(td_conv_plan* p) {
    const td_conv3x3_desc& d = p->d;
    if (d.x_dtype != TD_BF16) { set_error("tc conv: input must be bf16"); return TD_ERR_UNSUPPORTED; }
    if (d.cin % 64 != 0 || d.cout % 64 != 0) {
        set_error("tc conv: cin (%d) and cout (%d) must be multiples of 64", d.cin, d.cout);
        return TD_ERR_UNSUPPORTED;
    }
    if (d.ldx % 8 != 0 || d.x_coff % 8 != 0 || d.ldy % 8 != 0 || d.y_coff % 8 != 0) {
        set_error("tc conv: channel strides/offsets must be multiples of 8");
        return TD_ERR_UNSUPPORTED;
    }
    if (d.x_nchw || d.y_nchw) { set_error("tc conv: NCHW tensors not supported"); return TD_ERR_UNSUPPORTED; }
    if (((uintptr_t)d.x & 15) || ((uintptr_t)d.w & 15) || ((uintptr_t)d.y & 15)) {
        set_error("tc conv: tensors must be 16-byte aligned");
        return TD_ERR_ARG;
    }
    {   // layers whose feature map tiles well with a halo: one activation box per chunk instead of nine
        int st = TD_OK;
        if (halo_plan_init(p, &st)) return st;
    }
    EncodeTiledFn encode = tc_get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled not available from the driver"); return TD_ERR_DRIVER; }

    // spatial box: maximise useful rows per 128-row MMA tile
    const int bw = d.width <= 32 ? d.width : 32;
    const int tiles_w = (int)ceil_div(d.width, bw);
    double best = -1.0;
    int bbh = 1, bbn = 1;
    for (int bh = 1; bh <= d.height && bw * bh <= 128; ++bh) {
        int bn = 128 / (bw * bh);
        if (bn > d.batch) bn = d.batch;
        if (bn > 256) bn = 256;
        if (bn < 1) continue;
        const double tiles = (double)tiles_w * ceil_div(d.height, bh) * ceil_div(d.batch, bn);
        const double eff = (double)d.batch * d.height * d.width / (tiles * 128.0);
        if (eff > best + 1e-9 || (eff > best - 1e-9 && bh > bbh)) { best = eff; bbh = bh; bbn = bn; }
    }
    p->bw = bw; p->bh = bbh; p->bn = bbn;
    p->tiles_w = tiles_w;
    p->tiles_h = (int)ceil_div(d.height, bbh);
    p->tiles_n = (int)ceil_div(d.batch, bbn);
    // N tile and split-K together.  The per-tap kernel is bound by L2 -> SM operand delivery (~45 B/clk/SM):
    // one 64-channel stage moves 16 KB of activations + BLOCK_N x 128 B of weights, so a stage costs ~ (16 + BLOCK_N / 8) KB
    // / 45 B/clk, and a layer costs waves x stages-per-CTA x that (+ a second pass when the K loop is split).  On the 8x8 and
    // 4x4 maps the 256-wide tile leaves most SMs idle (256 -> 256 @ 8x8: 64 CTAs), a narrower tile fills them.
    const int64_t m_tiles = (int64_t)p->tiles_w * p->tiles_h * p->tiles_n;
    const int iters = 9 * (d.cin / 64);
    const bool may_split = d.splitk_ws && !d.stats;
    auto split_for = [&](int64_t ctas) {
        // Too few tiles to fill the 148 SMs: split the K loop (taps x channel chunks) across CTAs.  Each split writes an fp32
        // partial tile; a second kernel sums them in fixed order and applies the epilogue.  Short K loops lose more to the
        // second pass than they gain.
        if (!may_split || 2 * ctas > kNumSMs || iters < 64) return 1;
        int sk = (int)(kNumSMs / ctas);
        if (sk > 8) sk = 8;
        while (sk > 1 && iters / sk < 8) --sk;
        return sk;
    };
    int best_bn = pick_block_n(d.cout), best_sk = split_for(m_tiles * (d.cout / best_bn));
    if (!getenv("TD_TC_BLOCK_N")) {
        double best_cost = 0.0;
        bool have = false;
        for (int bn : {256, 128, 64}) {
            if (d.cout % bn) continue;
            const int64_t ctas = m_tiles * (d.cout / bn);
            const int sk = split_for(ctas);
            const double stage = (16.0 + bn / 8.0) * 1024.0 / 45.0;
            const double waves = (double)ceil_div(ctas * sk, (int64_t)kNumSMs);
            double cost = waves * (double)ceil_div(iters, sk) * std::max(stage, 4.0 * (bn == 256 ? 128.0 : 64.0));
            if (sk > 1) cost += 12000.0;                     // the reduce pass (~6 us)
            if (!have || cost < best_cost * 0.97) { have = true; best_cost = cost; best_bn = bn; best_sk = sk; }
        }
    }
    p->block_n = best_bn;
    p->n_tiles = d.cout / p->block_n;
    const int stage_bytes = TC_A_STAGE + p->block_n * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    // Enough tiles for two waves of CTAs: keep two CTAs resident per SM (<= ~100 KB of stages each, TMEM
    // 2 x BLOCK_N <= 512 columns) so that one CTA's epilogue overlaps the other's main loop.  Measured on
    // B200 (tools/tc_sweep.py): 28x28 128->128 70 -> 45 us, 16x16 512->128 60 -> 37 us.
    const int64_t ctas = m_tiles * p->n_tiles;
    if (ctas > kNumSMs) {
        int s2 = (100 * 1024) / stage_bytes;
        if (s2 >= 2) stages = s2 > 4 ? 4 : s2;
    }
    p->split_k = best_sk;
    if (const char* e = getenv("TD_TC_SPLIT_K")) { int v = atoi(e); if (v >= 1 && v <= split_for(ctas)) p->split_k = v; }
    if (const char* e = getenv("TD_TC_STAGES")) { int v = atoi(e); if (v >= 2 && v <= stages) stages = v; }
    p->stages = stages;
    p->smem_bytes = stages * stage_bytes + (2 * stages + 1) * 8 + 32 + 2 * p->block_n * 4 + 1024;

    {   // activations: (C, W, H, N)
        cuuint64_t gdim[4] = {(cuuint64_t)d.ldx, (cuuint64_t)d.width, (cuuint64_t)d.height, (cuuint64_t)d.batch};
        cuuint64_t gstr[3] = {(cuuint64_t)d.ldx * 2, (cuuint64_t)d.width * d.ldx * 2,
                              (cuuint64_t)d.height * d.width * d.ldx * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p->bw, (cuuint32_t)p->bh, (cuuint32_t)p->bn};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&p->tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return TD_ERR_DRIVER; }
    }
    {   // weights: (K = 9*Cin, Cout)
        cuuint64_t gdim[2] = {(cuuint64_t)9 * d.cin, (cuuint64_t)d.cout};
        cuuint64_t gstr[1] = {(cuuint64_t)9 * d.cin * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p->block_n};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&p->tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.w), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return TD_ERR_DRIVER; }
    }
    return TD_OK;
}

template <int BLOCK_N>
static int launch_tc(const td_conv_plan* p, const TcParams& prm, cudaStream_t s) {
    static int configured_smem = 0;
    if (p->smem_bytes > configured_smem) {
        TD_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
        configured_smem = p->smem_bytes;
    }
    dim3 grid((unsigned)(p->tiles_w * p->tiles_h * p->tiles_n), (unsigned)p->n_tiles, (unsigned)p->split_k);
    td::launch(conv3x3_tc_kernel<BLOCK_N>, td::LaunchCfg(grid, TC_THREADS, p->smem_bytes, s), p->tmap_x, p->tmap_w, prm);
    int st = launch_status("conv3x3_tc");
    if (st != TD_OK || p->split_k == 1) return st;
    const td_conv3x3_desc& d = p->d;
    const int64_t pixels = (int64_t)d.batch * d.height * d.width;
    const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(pixels * (d.cout / 4), 256), kNumSMs * 8));
    if (d.y_dtype == TD_BF16)
        td::launch(conv_splitk_reduce_kernel<__nv_bfloat16>, td::LaunchCfg(rgrid, 256, 0, s), d.splitk_ws, p->split_k, pixels, d.cout, d.scale, d.shift,
                                                                       d.relu, (__nv_bfloat16*)d.y, d.ldy, d.y_coff);
    else
        td::launch(conv_splitk_reduce_kernel<float>, td::LaunchCfg(rgrid, 256, 0, s), d.splitk_ws, p->split_k, pixels, d.cout, d.scale, d.shift,
                                                               d.relu, (float*)d.y, d.ldy, d.y_coff);
    return launch_status("conv_splitk_reduce");
}

int tc_plan_run(const td_conv_plan* p, cudaStream_t s) {
    if (p->halo) return halo_plan_run(p, s);
    const td_conv3x3_desc& d = p->d;
    TcParams prm;
    prm.B = d.batch; prm.H = d.height; prm.W = d.width; prm.cin = d.cin; prm.x_coff = d.x_coff;
    prm.cout = d.cout; prm.ldy = d.ldy; prm.y_coff = d.y_coff; prm.y_dtype = d.y_dtype;
    prm.y = d.y; prm.scale = d.scale; prm.shift = d.shift; prm.relu = d.relu;
    prm.bw = p->bw; prm.bh = p->bh; prm.bn = p->bn; prm.tiles_w = p->tiles_w; prm.tiles_h = p->tiles_h;
    prm.stages = p->stages;
    prm.a_bytes = (uint32_t)(64 * p->bw * p->bh * p->bn * 2);
    const int iters = 9 * (d.cin / 64);
    prm.split_iters = (int)ceil_div(iters, p->split_k);
    prm.ws = p->split_k > 1 ? d.splitk_ws : nullptr;
    prm.stats = d.stats;
    switch (p->block_n) {
        case 64: return launch_tc<64>(p, prm, s);
        case 128: return launch_tc<128>(p, prm, s);
        case 256: return launch_tc<256>(p, prm, s);
    }
    set_error("tc conv: bad block_n %d", p->block_n);
    return TD_ERR_ARG;
}

}  // namespace td
