// Weight gradient of the 3x3 convolutions on the 5th-generation tensor cores (sm_100a).
//
//   dW[o, tap, c] = sum_pixels dY[pixel, o] * X[pixel + tap_offset, c]
//
// is a GEMM whose reduction dimension is the pixel index, so BOTH operands are "MN-major" in
// shared memory (channels contiguous, one 128-byte row of 64 bf16 channels per pixel):
//   * a K step is one spatial box of pixels (bw x bh x bn, a multiple of 16 rows); the box of the
//     shifted operand X is loaded at tap-shifted coordinates, TMA out-of-bounds zero fill is
//     the convolution padding;
//   * 64-channel groups land as separate TMA boxes [rows][128 B] with SWIZZLE_128B; the UMMA
//     shared-memory descriptor walks them with LBO = box bytes (next 64 channels) and
//     SBO = 1024 B (next 8 pixels); every tcgen05.mma consumes 16 pixels (2048 B further on);
//   * one CTA owns (128 M-side channels) x (BLOCK_N N-side channels) x (one tap) x (a slice of the
//     pixel boxes); the fp32 accumulator lives in TMEM and is written as a split-K partial.
// The "M side" is dY (rows = Cout) unless Cout < 128 <= Cin, in which case the roles are swapped so
// that no half-empty 128-row MMA is issued; the shift always applies to the X side.
// TAPS = 3 ("halo") variant: one CTA owns the three horizontal taps of one kernel row.  The pixel box is laid out
// flat with a zero pad column (PW = W + 1 positions per image row, TMA out-of-bounds fill), so the shifted operand of
// tap dx is the SAME shared-memory box read one 128-byte row earlier / later: an MN-major SWIZZLE_128B descriptor may
// start at any row (tools/probe_umma_layouts.cu).  Both operands are loaded once per three taps instead of once per
// tap (the per-tap kernel moved 128 B/clk/SM through the L2 -> SM path at full MMA rate and was bound by it).  Rows
// past the box and the row in front of it are zero guards (never written by TMA), so the last partial K step and
// the +-1 shifts multiply zeros.  Three accumulators (3 x BLOCK_N TMEM columns).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (warp 2 owns the TMEM allocation).
#include <mutex>

#include "conv_plan.h"
#include "sm100.cuh"

namespace td {

using namespace sm100;

struct WgParams {
    int cin, cout;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;   // pixel boxes
    int rows;                                     // bw*bh*bn (multiple of 16)
    int m_boxes, n_boxes;                         // 64-channel TMA boxes per operand tile
    int m_coff, n_coff;                           // channel offsets of the tensors inside their rows
    int m_tiles, n_tiles;
    int x_on_m;                                   // 1: M side is X (shifted), 0: M side is dY
    int stages, splits, boxes_per_split;
    int slot_bytes;                               // TAPS = 3: bytes of one 64-channel box slot (1 KB front guard + box + zero tail)
    int ksteps;                                   // 16-pixel MMA steps per box
    float* ws;
    int epi16;                                    // 1 (TD_WG_EPI16=1): direct 8-byte stores from the 16x256b register layout; default: transpose tile
};

constexpr int WG_TC_THREADS = 192;

template <int BLOCK_N, int TAPS>
__global__ void __launch_bounds__(WG_TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_m, const __grid_constant__ CUtensorMap tmap_n, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int TMEM_COLS = TAPS == 1 ? BLOCK_N : (BLOCK_N == 128 ? 512 : 256);
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const uint32_t box_bytes = (uint32_t)p.rows * 128u;                          // bytes one TMA box delivers
    const uint32_t slot = TAPS == 1 ? box_bytes : (uint32_t)p.slot_bytes;        // distance between 64-channel boxes
    const uint32_t lead = TAPS == 1 ? 0u : 1024u;                                // zero guard in front of each box
    const uint32_t a_stage = slot * (uint32_t)p.m_boxes;
    const uint32_t b_stage = slot * (uint32_t)p.n_boxes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * a_stage;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // blockIdx.x -> (m tile, n tile, tap [TAPS = 1] or kernel row [TAPS = 3]); blockIdx.y -> split
    int bx = blockIdx.x;
    constexpr int TAP_GROUPS = 9 / TAPS;
    const int tap = (bx % TAP_GROUPS) * TAPS; bx /= TAP_GROUPS;       // first tap of this CTA
    const int nt = bx % p.n_tiles;
    const int mt = bx / p.n_tiles;
    const int split = blockIdx.y;
    const int dyy = tap / 3 - 1, dxx = TAPS == 1 ? tap % 3 - 1 : 0;
    const int total_boxes = p.tiles_w * p.tiles_h * p.tiles_n;
    const int box_beg = split * p.boxes_per_split;
    const int box_end = min(total_boxes, box_beg + p.boxes_per_split);
    const int iters = max(box_end - box_beg, 0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_m);
        prefetch_tmap(&tmap_n);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr);
    if (TAPS != 1) {
        // guards (the row in front of every box, the rows behind it) must read as zeros: clear the operand ring once
        uint4* z = reinterpret_cast<uint4*>(smem);
        const int n16 = (int)(((size_t)p.stages * (a_stage + b_stage)) >> 4);
        for (int i = threadIdx.x; i < n16; i += WG_TC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    td::pdl_sync();              // everything above is on-chip setup; global memory is touched only below

    if (warp == 0) {
        if (elect_one()) {
            const int wbase = TAPS == 1 ? 0 : -1;          // flat layout: the box starts at the zero pad column
            const int msx = wbase + (p.x_on_m ? dxx : 0), msy = p.x_on_m ? dyy : 0;
            const int nsx = wbase + (p.x_on_m ? 0 : dxx), nsy = p.x_on_m ? 0 : dyy;
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], box_bytes * (uint32_t)(p.m_boxes + p.n_boxes));
                int b = box_beg + it;
                const int tw = b % p.tiles_w; b /= p.tiles_w;
                const int th = b % p.tiles_h;
                const int tn = b / p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                for (int g = 0; g < p.m_boxes; ++g)
                    tma_load_4d(smem_a + (size_t)s * a_stage + (size_t)g * slot + lead, &tmap_m, &full_bar[s],
                                p.m_coff + (mt * 2 + g) * 64, w0 + msx, h0 + msy, n0);
                for (int g = 0; g < p.n_boxes; ++g)
                    tma_load_4d(smem_b + (size_t)s * b_stage + (size_t)g * slot + lead, &tmap_n, &full_bar[s],
                                p.n_coff + (nt * (BLOCK_N / 64) + g) * 64, w0 + nsx, h0 + nsy, n0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);     // both operands MN-major
            const int ksteps = TAPS == 1 ? (p.rows >> 4) : p.ksteps;
            // descriptor = constant fields + (address >> 4): one 64-bit add per MMA keeps this single-thread issue loop
            // under the 64 cycles an MMA takes (building the descriptor from scratch every time made it issue-bound)
            const uint64_t desc0 = make_smem_desc_sw128(0, slot, 1024);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a16 = (smem_u32(smem_a + (size_t)s * a_stage) + lead) >> 4;
                const uint32_t b16 = (smem_u32(smem_b + (size_t)s * b_stage) + lead) >> 4;
#pragma unroll
                for (int t = 0; t < TAPS; ++t) {
                    // tap dx = t - 1 reads the shifted (X) operand one pixel row (128 B = 8 descriptor units) earlier / later
                    const int shift = TAPS == 1 ? 0 : (t - 1) * 8;
                    uint64_t da = desc0 + (uint64_t)(a16 + (p.x_on_m ? shift : 0));
                    uint64_t db = desc0 + (uint64_t)(b16 + (p.x_on_m ? 0 : shift));
                    uint32_t acc = it > 0 ? 1u : 0u;
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16(tmem_base + t * BLOCK_N, da, db, idesc, acc);
                        da += 128;                       // 16 pixel rows = 2048 bytes
                        db += 128;
                        acc = 1u;
                    }
                }
                umma_commit(&empty_bar[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                   // accumulator row = M-side channel within the tile
        const int64_t per = (int64_t)p.cout * 9 * p.cin;
        float* ws = p.ws + (int64_t)split * per;
        const int mch = mt * 128 + m;
        const int m_limit = p.x_on_m ? p.cin : p.cout;
        if (iters > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
        }
        if (!p.x_on_m && iters > 0 && p.epi16) {
            // row = cout, columns = cin, straight from the accumulators: tcgen05.ld.16x256b hands thread l the registers
            // [16h + 4j + 2u + w] = (row 16h + 8u + l/4, column 8j + 2(l%4) + w) (tools/probe_tmem_ld_shapes.cu), so one 8-byte store
            // per (row, j) and thread makes every warp instruction 8 rows x 32 contiguous bytes = 8 whole sectors.  No transpose
            // through shared memory, and the next chunk's load is in flight during the stores (the epilogue of this kernel is not
            // overlapped with anything: it is the drain of every launch).
            uint32_t r[2][32];
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
            tmem_ld_16x256_pair(t0, r[0]);
#pragma unroll 1
            for (int cc = 0; cc < TAPS * BLOCK_N; cc += 64) {
                tmem_ld_wait();
                tmem_ld_16x256_pair(t0 + (uint32_t)(cc + 32), r[1]);          // TAPS * BLOCK_N is a multiple of 64
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) {
                        tmem_ld_wait();
                        if (cc + 64 < TAPS * BLOCK_N) tmem_ld_16x256_pair(t0 + (uint32_t)(cc + 64), r[0]);
                    }
                    const int c = cc + 32 * half;
                    const int t_ = c / BLOCK_N, c0 = c - t_ * BLOCK_N;
                    const int nch0 = nt * BLOCK_N + c0;
                    float* wt = ws + (int64_t)(tap + t_) * p.cin + nch0 + 2 * (lane & 3);
#pragma unroll
                    for (int hu = 0; hu < 4; ++hu) {
                        const int mrow = mt * 128 + q * 32 + 8 * hu + (lane >> 2);
                        if (mrow >= m_limit) continue;
                        float* wr = wt + (int64_t)mrow * 9 * p.cin;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (nch0 + 8 * j + 2 * (lane & 3) < p.cin)
                                *reinterpret_cast<float2*>(wr + 8 * j) = make_float2(__uint_as_float(r[half][16 * (hu >> 1) + 4 * j + 2 * (hu & 1)]),
                                                                                    __uint_as_float(r[half][16 * (hu >> 1) + 4 * j + 2 * (hu & 1) + 1]));
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int cc = 0; cc < TAPS * BLOCK_N; cc += 32) {
                const int t_ = cc / BLOCK_N, c0 = cc - t_ * BLOCK_N;
                const int tap_ = tap + t_;
                uint32_t r[32];
                if (iters > 0) {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, r);
                    tmem_ld_wait();
                } else {
    #pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                const int nch0 = nt * BLOCK_N + c0;
                if (p.x_on_m && mch < m_limit) {
                    // row = cin, columns = cout: lanes are consecutive cin -> coalesced per column
    #pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nch0 + j < p.cout) ws[((int64_t)(nch0 + j) * 9 + tap_) * p.cin + mch] = __uint_as_float(r[j]);
                }
                if (!p.x_on_m) {
                    // row = cout, columns = cin.  A thread owns one row, so direct stores would touch 32 cache lines per warp
                    // instruction; the 32 x 32 chunk is transposed through a padded tile in the (idle) operand ring and leaves
                    // as 128-byte row segments, four rows per instruction.
                    float* tile = reinterpret_cast<float*>(smem) + q * (32 * 36);
    #pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(tile + lane * 36 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    __syncwarp();
                    const int col = (lane & 7) * 4;
    #pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rw_ = it * 4 + (lane >> 3);
                        const int mrow = mt * 128 + q * 32 + rw_;
                        const float4 a = *reinterpret_cast<const float4*>(tile + rw_ * 36 + col);
                        if (mrow < m_limit && nch0 + col < p.cin)
                            *reinterpret_cast<float4*>(ws + ((int64_t)mrow * 9 + tap_) * p.cin + nch0 + col) = a;
                    }
                    __syncwarp();
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode_fn();   // conv_tc.cu

struct WgGeom {
    int bw, bh, bn, rows, tiles_w, tiles_h, tiles_n;
    int x_on_m, block_n, m_tiles, n_tiles, m_boxes, n_boxes, stages, splits, boxes_per_split, smem_bytes;
    int taps, slot_bytes, ksteps;       // taps = 3: halo variant (flat boxes with a pad column, one CTA per kernel row)
};

// Halo variant: box = (64 ch, PW = W + 1, bh, bn) positions, P = PW*bh*bn <= 128 K rows per box.
static bool wg_geometry_halo(const td_wgrad_desc& d, WgGeom& g) {
    if (const char* e = getenv("TD_WG_HALO")) if (atoi(e) == 0) return false;
    if (d.cin % 64 != 0 || d.cout % 64 != 0 || d.width > 127) return false;
    const int PW = d.width + 1;
    double best = -1.0;
    int bbh = 0, bbn = 0;
    auto consider = [&](int bh, int bn) {
        const int P = PW * bh * bn;
        if (P > 128 || bh > 256 || bn > 256) return;
        const int krows = (P + 15) / 16 * 16;
        const double tiles = (double)ceil_div(d.height, bh) * ceil_div(d.batch, bn);
        const double eff = (double)d.batch * d.height * d.width / (tiles * krows);
        const double score = eff + 1e-3 * P / 128.0;            // ties: the larger box (more MMAs per pipeline stage)
        if (score > best) { best = score; bbh = bh; bbn = bn; }
    };
    for (int bh = 1; bh <= d.height; ++bh) consider(bh, 1);
    for (int bn = 2; bn <= d.batch; ++bn) consider(d.height, bn);       // whole images
    if (best < 0.6) return false;
    g.taps = 3;
    g.bw = PW; g.bh = bbh; g.bn = bbn; g.rows = PW * bbh * bbn;
    g.ksteps = (g.rows + 15) / 16;
    g.tiles_w = 1; g.tiles_h = (int)ceil_div(d.height, bbh); g.tiles_n = (int)ceil_div(d.batch, bbn);
    g.x_on_m = (d.cout < 128 && d.cin >= 128) ? 1 : 0;
    const int cm = g.x_on_m ? d.cin : d.cout, cn = g.x_on_m ? d.cout : d.cin;
    g.m_tiles = (int)ceil_div(cm, 128);
    g.m_boxes = 2;
    g.block_n = cn % 128 == 0 ? 128 : 64;       // three accumulators: 3 x BLOCK_N <= 512 TMEM columns
    g.n_boxes = g.block_n / 64;
    g.n_tiles = cn / g.block_n;
    // slot: 1 KB zero guard in front (row -1 of the dx = -1 tap), the box, zero rows up to ksteps*16 + 1 (dx = +1)
    g.slot_bytes = 1024 + ((g.ksteps * 16 + 1) * 128 + 1023) / 1024 * 1024;
    const int stage = (g.m_boxes + g.n_boxes) * g.slot_bytes;
    g.stages = std::min(4, (216 * 1024) / stage);
    if (g.stages < 2) return false;
    if (const char* e = getenv("TD_WG_STAGES")) { int v = atoi(e); if (v >= 2 && v <= g.stages) g.stages = v; }
    g.smem_bytes = g.stages * stage + (2 * g.stages + 1) * 8 + 16 + 1024;
    // one CTA per SM (the operand ring fills the shared memory): a single wave of <= 148 CTAs
    const int total_boxes = g.tiles_h * g.tiles_n;
    const int base = g.m_tiles * g.n_tiles * 3;
    int splits = std::max(1, sm_budget() / base);
    if (const char* e = getenv("TD_WG_TARGET_CTAS")) { int v = atoi(e); if (v > 0) splits = std::max(1, v / base); }
    splits = std::max(1, std::min(splits, std::max(1, total_boxes / 2)));
    g.boxes_per_split = (int)ceil_div(total_boxes, splits);
    g.splits = (int)ceil_div(total_boxes, g.boxes_per_split);
    return true;
}

static bool wg_geometry(const td_wgrad_desc& d, WgGeom& g) {
    if (wg_geometry_halo(d, g)) return true;
    g.taps = 1; g.slot_bytes = 0; g.ksteps = 0;
    if (d.cin % 64 != 0 || d.cout % 64 != 0) return false;
    // pixel box: full-width rows, rows*... a multiple of 16 and <= 128; maximise useful pixels, then size
    const int bw = d.width <= 32 ? d.width : 32;
    const int tiles_w = (int)ceil_div(d.width, bw);
    double best = -1.0;
    int bbh = 0, bbn = 0;
    for (int bh = 1; bh <= d.height; ++bh)
        for (int bn = 1; bn <= d.batch && bw * bh * bn <= 128; ++bn) {
            const int rows = bw * bh * bn;
            if (rows % 16 != 0) continue;
            const double tiles = (double)tiles_w * ceil_div(d.height, bh) * ceil_div(d.batch, bn);
            const double eff = (double)d.batch * d.height * d.width / (tiles * rows);
            const double score = eff + 1e-3 * rows / 128.0;
            if (score > best) { best = score; bbh = bh; bbn = bn; }
        }
    if (best < 0) return false;
    g.bw = bw; g.bh = bbh; g.bn = bbn; g.rows = bw * bbh * bbn;
    g.tiles_w = tiles_w; g.tiles_h = (int)ceil_div(d.height, bbh); g.tiles_n = (int)ceil_div(d.batch, bbn);
    g.x_on_m = (d.cout < 128 && d.cin >= 128) ? 1 : 0;
    const int cm = g.x_on_m ? d.cin : d.cout, cn = g.x_on_m ? d.cout : d.cin;
    g.m_tiles = (int)ceil_div(cm, 128);
    g.m_boxes = 2;
    const int box_bytes = g.rows * 128;
    int bn_ch = cn % 256 == 0 ? 256 : (cn % 128 == 0 ? 128 : 64);
    if (const char* e = getenv("TD_WG_BLOCK_N")) { int v = atoi(e); if ((v == 64 || v == 128 || v == 256) && cn % v == 0) bn_ch = v; }
    // keep at least 3 pipeline stages in ~200 KB
    while (bn_ch > 64 && (2 + bn_ch / 64) * box_bytes * 3 > 200 * 1024) bn_ch >>= 1;
    g.block_n = bn_ch;
    g.n_boxes = bn_ch / 64;
    g.n_tiles = cn / bn_ch;
    const int stage = (g.m_boxes + g.n_boxes) * box_bytes;
    g.stages = std::max(2, std::min(8, (200 * 1024) / stage));
    // 64-channel N tiles: two stages (96 KB) keep two CTAs per SM resident (tools/wg_sweep.py: 28x28 64->128
    // 48 -> 33 us, 32x32 256->64 110 -> 89 us)
    if (g.block_n == 64 && 2 * stage <= 100 * 1024) g.stages = 2;
    if (const char* e = getenv("TD_WG_STAGES")) { int v = atoi(e); if (v >= 2 && v <= g.stages) g.stages = v; }
    g.smem_bytes = g.stages * stage + (2 * g.stages + 1) * 8 + 16 + 1024;
    const int total_boxes = g.tiles_w * g.tiles_h * g.tiles_n;
    const int base = g.m_tiles * g.n_tiles * 9;
    int target = 2 * kNumSMs;
    if (const char* e = getenv("TD_WG_TARGET_CTAS")) { int v = atoi(e); if (v > 0) target = v; }
    int splits = (int)ceil_div(target, base);
    splits = std::max(1, std::min(splits, std::max(1, total_boxes / 4)));
    g.boxes_per_split = (int)ceil_div(total_boxes, splits);
    g.splits = (int)ceil_div(total_boxes, g.boxes_per_split);
    return true;
}

int wgrad_tc_splits(const td_wgrad_desc& d) {
    WgGeom g;
    return wg_geometry(d, g) ? g.splits : 1;
}

static int encode_act(EncodeTiledFn encode, CUtensorMap* map, const void* ptr, int ld, int W, int H, int B, int bw, int bh,
                      int bn) {
    cuuint64_t gdim[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r); return TD_ERR_DRIVER; }
    return TD_OK;
}

int wgrad_tc_plan_init(td_wgrad_plan* p) {
    const td_wgrad_desc& d = p->d;
    if (d.x_dtype != TD_BF16 || d.dy_dtype != TD_BF16 || d.x_nchw || d.dy_nchw) {
        set_error("tc wgrad: operands must be NHWC bf16");
        return TD_ERR_UNSUPPORTED;
    }
    if (d.ldx % 8 != 0 || d.x_coff % 8 != 0 || d.lddy % 8 != 0 || d.dy_coff % 8 != 0 || ((uintptr_t)d.x & 15) ||
        ((uintptr_t)d.dy & 15) || ((uintptr_t)d.workspace & 15)) {
        set_error("tc wgrad: strides/offsets must be multiples of 8 channels and tensors 16-byte aligned");
        return TD_ERR_UNSUPPORTED;
    }
    WgGeom g;
    if (!wg_geometry(d, g)) {
        set_error("tc wgrad: shape not supported (cin=%d cout=%d %dx%d)", d.cin, d.cout, d.height, d.width);
        return TD_ERR_UNSUPPORTED;
    }
    EncodeTiledFn encode = tc_get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled not available from the driver"); return TD_ERR_DRIVER; }
    p->bw = g.bw; p->bh = g.bh; p->bn = g.bn; p->rows = g.rows;
    p->tiles_w = g.tiles_w; p->tiles_h = g.tiles_h; p->tiles_n = g.tiles_n;
    p->x_on_m = g.x_on_m; p->block_n = g.block_n; p->m_tiles = g.m_tiles; p->n_tiles = g.n_tiles;
    p->m_boxes = g.m_boxes; p->n_boxes = g.n_boxes; p->stages = g.stages; p->splits = g.splits;
    p->boxes_per_split = g.boxes_per_split; p->smem_bytes = g.smem_bytes;
    p->taps = g.taps; p->slot_bytes = g.slot_bytes; p->ksteps = g.ksteps;
    // the tensor map's channel extent is the live channel range, so the second 64-channel box of a
    // 64-channel M operand is entirely out of bounds (zero rows in the MMA)
    const void* mp = g.x_on_m ? d.x : d.dy;
    const void* np = g.x_on_m ? d.dy : d.x;
    const int mld = g.x_on_m ? d.ldx : d.lddy, nld = g.x_on_m ? d.lddy : d.ldx;
    int st = encode_act(encode, &p->tmap_m, mp, mld, d.width, d.height, d.batch, g.bw, g.bh, g.bn);
    if (st != TD_OK) return st;
    return encode_act(encode, &p->tmap_n, np, nld, d.width, d.height, d.batch, g.bw, g.bh, g.bn);
}

template <int BLOCK_N, int TAPS>
static int launch_wg(const td_wgrad_plan* p, const WgParams& prm, cudaStream_t s) {
    static int configured_smem = 0;
    if (p->smem_bytes > configured_smem) {
        TD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
        configured_smem = p->smem_bytes;
    }
    dim3 grid((unsigned)(p->m_tiles * p->n_tiles * (9 / TAPS)), (unsigned)p->splits);
    td::launch(wgrad_tc_kernel<BLOCK_N, TAPS>, td::LaunchCfg(grid, WG_TC_THREADS, p->smem_bytes, s), p->tmap_m, p->tmap_n, prm);
    return launch_status("wgrad_tc");
}

int wgrad_tc_plan_run(const td_wgrad_plan* p, cudaStream_t s) {
    const td_wgrad_desc& d = p->d;
    WgParams prm;
    prm.cin = d.cin; prm.cout = d.cout;
    prm.bw = p->bw; prm.bh = p->bh; prm.bn = p->bn;
    prm.tiles_w = p->tiles_w; prm.tiles_h = p->tiles_h; prm.tiles_n = p->tiles_n;
    prm.rows = p->rows; prm.m_boxes = p->m_boxes; prm.n_boxes = p->n_boxes;
    prm.m_coff = p->x_on_m ? d.x_coff : d.dy_coff;
    prm.n_coff = p->x_on_m ? d.dy_coff : d.x_coff;
    prm.m_tiles = p->m_tiles; prm.n_tiles = p->n_tiles; prm.x_on_m = p->x_on_m;
    prm.stages = p->stages; prm.splits = p->splits; prm.boxes_per_split = p->boxes_per_split;
    prm.ws = d.workspace;
    {
        static int epi16 = -1;
        if (epi16 < 0) { const char* e = getenv("TD_WG_EPI16"); epi16 = (e && atoi(e) == 1) ? 1 : 0; }     // opt-in: measured neutral
        prm.epi16 = epi16;
    }
    prm.slot_bytes = p->slot_bytes; prm.ksteps = p->ksteps;
    if (p->taps == 3) return p->block_n == 128 ? launch_wg<128, 3>(p, prm, s) : launch_wg<64, 3>(p, prm, s);
    switch (p->block_n) {
        case 64: return launch_wg<64, 1>(p, prm, s);
        case 128: return launch_wg<128, 1>(p, prm, s);
        case 256: return launch_wg<256, 1>(p, prm, s);
    }
    set_error("tc wgrad: bad block_n %d", p->block_n);
    return TD_ERR_ARG;
}

}  // namespace td
