// 3x3 convolution on the tcgen05 tensor cores, halo variant: the activation tile is loaded ONCE per
// 64-channel chunk together with its one-pixel halo, and the nine taps of the stencil read it through
// shifted shared-memory descriptors.
//
// Why: the per-tap kernel (conv_tc.cu) moves 9 activation boxes + 9 weight boxes per chunk through the
// L2 -> SM path and is bound by it (measured ~45 B/clk/SM against 128 B/clk needed at full MMA rate).
// Here an M subtile is 128 consecutive positions of the *padded* image held in shared memory
//      position(h, w) = (h - h0 + 1) * PW + (w - w_box0),        PW = pixels per row of the box,
// so the A operand of tap (dy, dx) is the same box read from row offset dy*PW + dx: a K-major SWIZZLE_128B
// UMMA descriptor may start at any 128-byte row of a TMA-written tile (the swizzle is a function of the
// absolute shared-memory address; tools/probe_umma_rowoff.cu checks this on the device).  Two layouts:
//   G = 1 ("flat")  box columns -1 .. PW-2 with PW >= W + 1: the zero column that TMA fills right of the
//                   image doubles as the left halo of the next row; positions with w >= W are computed
//                   and dropped.  One box per chunk serves all 9 taps.
//   G = 3 ("dx")    W % 8 == 0, PW = W, no dead positions: one box per horizontal shift dx (columns
//                   dx .. dx+W-1, zero filled by TMA), each serving the three dy taps.  (opt-in: 3x the A bytes)
//   strip           W % 8 == 0: a subtile is a strip of 8 columns x 16 rows read out of ONE box of 10 columns x 18
//                   rows (PW = 10).  Each 8-row group of the A operand is one image row of the strip, so the groups
//                   are PW*128 bytes apart: the descriptor's stride byte offset is 1280 instead of 1024
//                   (tools/probe_umma_layouts.cu).  No dead positions and one box per chunk.
// Two subtiles (each with its own box) share every weight stage, accumulators are double buffered in
// TMEM, and the kernel is persistent: each CTA walks a contiguous range of (n-tile, subtile) units so
// that the epilogue of one pair overlaps the main loop of the next.
// Warp roles: 0 = TMA producer, 1-2 = MMA issuers (one per subtile), 3..6 = epilogue of subtile 0, 7..10 = epilogue of
// subtile 1 (with one group of four warps the epilogue was as long as the main loop on the K = 576 layers).
// Epilogue: TMEM -> registers -> one padded 32 x 32 shared-memory tile per warp -> folded BatchNorm affine / ReLU ->
// whole row segments to global memory (a thread owns one position, so direct stores would hit 32 cache lines per warp
// instruction); the TMEM load of the next chunk is in flight during the read-back of the current one.
#include <algorithm>

#include "conv_plan.h"
#include "sm100.cuh"

namespace td {

using namespace sm100;

struct HaloParams {
    int B, H, W, cin, x_coff;
    int cout, ldy, y_coff, y_dtype;
    void* y;
    const float* scale;
    const float* shift;
    int relu;
    float* stats;            // train-mode BatchNorm partials [n_sub][2][cout] (+ zero row), or NULL
    int G;                   // 1 or 3 (see above)
    int PW, bh, BN, RH;
    int bw;                  // image columns per subtile (W for flat / dx, 8 for strip)
    int ew;                  // positions per subtile row in the M index (PW for flat, W for dx, 8 for strip)
    int a_sbo;               // bytes between the 8-row groups of the A operand (1024, or PW*128 for strip)
    int tiles_w, tiles_h, n_sub, units;
    int NA, NB;
    uint32_t a_box_bytes, a_slot_bytes;
    int dbg;                 // TD_TC_HALO_DBG=1: per-CTA wait-cycle counters into g_halo_dbg
    void* pool_y;            // eval mode: MaxPool2d(2) of the output written beside it (windows never straddle a subtile), or NULL
    int pool_ceil, Hp, Wp;   // pooled map size
};

// [CTA][8]: 0 total, 1 producer wait A slot, 2 producer wait B stage, 3 MMA wait A, 4 MMA wait B, 5 MMA wait
// accumulator, 6 epilogue wait accumulator, 7 epilogue busy (clock64 cycles; tuning aid, read by td_conv3x3_debug_counters)
__device__ unsigned long long g_halo_dbg[kNumSMs * 8];
#define HALO_T0() const long long _t0 = p.dbg ? clock64() : 0
#define HALO_ACC(var) do { if (p.dbg) var += clock64() - _t0; } while (0)

constexpr int HALO_THREADS = 352;   // warp 0 TMA producer, warps 1-2 MMA issuers (one per subtile), warps 3-6 / 7-10 epilogue groups
constexpr int HALO_S = 2;           // subtiles sharing one weight stage

struct UnitWalk {             // the same walk is replayed by the producer, the MMA issuers and the epilogue
    int u, u_hi, n_sub;
    __device__ bool next(int& nt, int& s0, int& cnt) {
        if (u >= u_hi) return false;
        nt = u / n_sub;
        s0 = u - nt * n_sub;
        cnt = min(HALO_S, min(u_hi - u, n_sub - s0));
        u += cnt;
        return true;
    }
};

struct Ring {                 // position in a ring of n mbarrier-guarded slots: index + phase parity, no divisions
    uint32_t idx, phase, n;
    __device__ void advance() { if (++idx == n) { idx = 0; phase ^= 1u; } }
    __device__ void advance(uint32_t k) { idx += k; if (idx >= n) { idx -= n; phase ^= 1u; } }
};

// EPI: epilogue variant.  0: fp32 transpose tile, affine after the transpose (fp32 outputs); 1: bf16 output without affine / ReLU /
// pooling (train mode): rounded in the 32x32b layout, bf16 tile; 2: bf16 output with the folded affine (eval): 16x256b TMEM loads, a
// thread owns 4 rows x 8 columns, affine + ReLU + rounding in registers, bf16 tile; 3: variant 1 WITH train-mode statistics in the
// 16x256b layout: a thread pre-sums its 4 rows per column, a halving butterfly over the 8 row groups finishes the warp's column sums in
// 14 shuffles (the 32x32b layout needs 62 per chunk, and shuffles share the pipe with shared memory).
template <int N_TILE, int EPI>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_STAGE = N_TILE * 128;
    constexpr int ACC_COLS = HALO_S * N_TILE;              // TMEM columns of one accumulator buffer
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + (size_t)p.NA * p.a_slot_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.NB * B_STAGE);
    uint64_t* a_empty = a_full + p.NA;
    uint64_t* b_full = a_empty + p.NA;
    uint64_t* b_empty = b_full + p.NB;
    uint64_t* acc_full = b_empty + p.NB;                   // [2 buffers][2 subtiles]
    uint64_t* acc_empty = acc_full + 4;                    // [2 buffers]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* s_affine = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 1) + 15) & ~(uintptr_t)15);  // [2 groups][2][2][N_TILE]
    float* s_stats = s_affine + 8 * N_TILE;                // [2 groups][4 warps][2][N_TILE]
    float* s_tile = s_stats + 16 * N_TILE;                 // [8 epilogue warps][32 rows][36]: output staging (see the epilogue)
    double* s_acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_tile + 8 * 32 * 36) + 7) & ~(uintptr_t)7);   // [2 groups][2][N_TILE]: per-CTA statistics

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u_lo = (int)((int64_t)blockIdx.x * p.units / gridDim.x);
    const int u_hi = (int)((int64_t)(blockIdx.x + 1) * p.units / gridDim.x);
    const int kchunks = p.cin >> 6;
    const int tpg = 9 / p.G;                               // taps per box group

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < p.NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], HALO_S); }
        for (int s = 0; s < 4; ++s) mbar_init(&acc_full[s], 1);
        for (int s = 0; s < 2; ++s) mbar_init(&acc_empty[s], 8);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<2 * ACC_COLS>(tmem_ptr);
    // The tail of every A slot (>= 1 KB past the box) is never written by TMA: zero it once.  In the flat layout the position one
    // past the box is the right-hand zero pad of the last row.
    for (int s = 0; s < p.NA; ++s) {
        uint32_t* tail = reinterpret_cast<uint32_t*>(smem_a + (size_t)s * p.a_slot_bytes + p.a_box_bytes);
        for (int i = threadIdx.x; i < (int)((p.a_slot_bytes - p.a_box_bytes) >> 2); i += HALO_THREADS) tail[i] = 0u;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    td::pdl_sync();              // everything above is on-chip setup; global memory is touched only below

    if (warp == 0) {
        if (elect_one()) {
            UnitWalk wk{u_lo, u_hi, p.n_sub};
            int nt, s0, cnt;
            Ring ra{0, 0, (uint32_t)p.NA}, rb{0, 0, (uint32_t)p.NB};
            long long w0 = 0, w1 = 0;
            const long long t_start = clock64();
            while (wk.next(nt, s0, cnt)) {
                for (int cc = 0; cc < kchunks; ++cc) {
                    for (int g = 0; g < p.G; ++g) {
                        for (int j = 0; j < cnt; ++j) {
                            const int sub = s0 + j;
                            const int tw = sub % p.tiles_w, th = (sub / p.tiles_w) % p.tiles_h, tn = sub / (p.tiles_w * p.tiles_h);
                            { HALO_T0(); mbar_wait(&a_empty[ra.idx], ra.phase ^ 1u); HALO_ACC(w0); }
                            if (p.dbg & 2) {                  // timing experiment: no A traffic
                                mbar_arrive(&a_full[ra.idx]);
                            } else {
                                mbar_arrive_expect_tx(&a_full[ra.idx], p.a_box_bytes);
                                tma_load_4d(smem_a + (size_t)ra.idx * p.a_slot_bytes, &tmap_x, &a_full[ra.idx], p.x_coff + cc * 64,
                                            tw * p.bw + (p.G == 3 ? g - 1 : -1), th * p.bh - 1, tn * p.BN);
                            }
                            ra.advance();
                        }
                        for (int ti = 0; ti < tpg; ++ti) {
                            const int tap = p.G == 3 ? ti * 3 + g : ti;
                            { HALO_T0(); mbar_wait(&b_empty[rb.idx], rb.phase ^ 1u); HALO_ACC(w1); }
                            if (p.dbg & 4) {                  // timing experiment: no B traffic
                                mbar_arrive(&b_full[rb.idx]);
                            } else {
                                mbar_arrive_expect_tx(&b_full[rb.idx], (uint32_t)B_STAGE);
                                tma_load_2d(smem_b + (size_t)rb.idx * B_STAGE, &tmap_w, &b_full[rb.idx], tap * p.cin + cc * 64, nt * N_TILE);
                            }
                            rb.advance();
                        }
                    }
                }
            }
            if (p.dbg & 1) {
                g_halo_dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
                g_halo_dbg[blockIdx.x * 8 + 1] = w0;
                g_halo_dbg[blockIdx.x * 8 + 2] = w1;
            }
        }
    } else if (warp <= HALO_S) {
        // ---- MMA issuers: warp 1 drives subtile 0, warp 2 subtile 1.  Two independent issue streams keep the tensor pipe
        // busy while one of them sits in an mbarrier wait (a wait is ordered behind the thread's own queued MMAs and costs
        // ~150 idle cycles per stage with a single issuer: tools/probe_umma_pipe.cu).
        if (elect_one()) {
            const int j = warp - 1;
            constexpr uint32_t idesc = make_idesc_bf16(128, N_TILE, 0, 0);
            const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
            const uint64_t desc0a = make_smem_desc_sw128(0, 16, (uint32_t)p.a_sbo);
            const uint32_t a_base = smem_u32(smem_a) >> 4, b_base = smem_u32(smem_b) >> 4;
            const uint32_t a_slot16 = p.a_slot_bytes >> 4;
            UnitWalk wk{u_lo, u_hi, p.n_sub};
            int nt, s0, cnt;
            Ring ra{0, 0, (uint32_t)p.NA}, rb{0, 0, (uint32_t)p.NB};
            uint32_t gi = 0;
            long long w0 = 0, w1 = 0, w2 = 0;
            while (wk.next(nt, s0, cnt)) {
                const uint32_t ab = gi & 1u;
                const bool active = j < cnt;
                { HALO_T0(); mbar_wait(&acc_empty[ab], ((gi >> 1) & 1u) ^ 1u); HALO_ACC(w2); }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ab * ACC_COLS + j * N_TILE;
                uint32_t accumulate = 0;
                for (int cc = 0; cc < kchunks; ++cc) {
                    for (int g = 0; g < p.G; ++g) {
                        uint32_t a16 = 0, my_slot = 0;
                        if (active) {
                            Ring mine = ra;
                            mine.advance((uint32_t)j);
                            my_slot = mine.idx;
                            { HALO_T0(); mbar_wait(&a_full[my_slot], mine.phase); HALO_ACC(w0); }
                            a16 = a_base + my_slot * a_slot16;
                        }
                        int dyi = 0, dxi = 0;
                        for (int ti = 0; ti < tpg; ++ti) {
                            const uint32_t rowoff = (uint32_t)(dyi * p.PW + (p.G == 3 ? 0 : dxi));      // 128-byte rows = 8 descriptor units
                            if (p.G == 3 || ++dxi == 3) { dxi = 0; ++dyi; }
                            const uint64_t da = desc0a + (uint64_t)(a16 + rowoff * 8u);
                            const uint64_t db = desc0 + (uint64_t)(b_base + rb.idx * (uint32_t)(B_STAGE >> 4));
                            { HALO_T0(); mbar_wait(&b_full[rb.idx], rb.phase); HALO_ACC(w1); }
                            if (active) {
                                tc_fence_after();
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accumulate | (uint32_t)k);
                                accumulate = 1;
                                umma_commit(&b_empty[rb.idx]);
                            } else {
                                mbar_arrive(&b_empty[rb.idx]);      // keeps the stage's two-arrival count uniform
                            }
                            rb.advance();
                        }
                        if (active) umma_commit(&a_empty[my_slot]);
                        ra.advance((uint32_t)cnt);
                    }
                }
                if (active) umma_commit(&acc_full[ab * 2 + j]);
                else mbar_arrive(&acc_full[ab * 2 + j]);
                ++gi;
            }
            if ((p.dbg & 1) && j == 0) {
                g_halo_dbg[blockIdx.x * 8 + 3] = w0;
                g_halo_dbg[blockIdx.x * 8 + 4] = w1;
                g_halo_dbg[blockIdx.x * 8 + 5] = w2;
            }
        }
    } else {
        // ---- epilogue: TMEM -> registers -> global, overlapped with the next pair's main loop ---------
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;                 // position inside the subtile
        const int img_rows = p.RH * p.PW;             // flat layout with whole images: positions per image (others: >= 128)
        const int n_rel = p.BN > 1 ? r / img_rows : 0;
        const int rem = r - n_rel * img_rows;
        const int h_rel = rem / p.ew;
        const int w_rel = rem - h_rel * p.ew;
        const int grp = (warp - 3) >> 2;             // epilogue group = the subtile of the pair it drains
        const int tid = threadIdx.x - 96 - grp * 128;
        const int bar_id = 1 + grp;
        // Fused MaxPool2d(2): the group's four 32 x 32 staging tiles are one 128-position tile; thread `tid` owns pooled pixel
        // tid >> 2 of the subtile and 8 of the chunk's 32 channels.  (pp_r[k] < 0: window element outside the image / subtile.)
        int pp_r[4] = {-1, -1, -1, -1}, pp_n = 0, pp_h = 0, pp_w = 0;
        bool pp_on = false;
        if (p.pool_y) {
            const int pw_cnt = (p.bw + 1) >> 1, ph_cnt = (p.bh + 1) >> 1;
            const int pp = tid >> 2, per_img = pw_cnt * ph_cnt;
            pp_n = pp / per_img;
            const int rem_p = pp - pp_n * per_img;
            pp_h = rem_p / pw_cnt;
            pp_w = rem_p - pp_h * pw_cnt;
            pp_on = pp_n < p.BN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int hh = 2 * pp_h + (k >> 1), ww = 2 * pp_w + (k & 1);
                if (pp_on && hh < p.bh && ww < p.bw) {
                    const int pos = pp_n * img_rows + hh * p.ew + ww;                // position inside the subtile = TMEM lane
                    pp_r[k] = ((((pos >> 5) + 1) & 3) << 5) + (pos & 31);           // its row of the group tile (warp slot = (quadrant + 1) & 3)
                }
            }
        }
        float* const aff_g = s_affine + grp * 4 * N_TILE;
        float* const stats_g = s_stats + grp * 8 * N_TILE;
        UnitWalk wk{u_lo, u_hi, p.n_sub};
        int nt, s0, cnt;
        uint32_t gi = 0;
        long long w0 = 0;
        const long long t_start = clock64();
        // Train-mode statistics: ONE partial row per CTA (row = blockIdx.x), so the consumer's finalize prologue sums <= 148
        // rows.  Both epilogue groups add their subtiles' column sums into per-group double accumulators (a (which, channel)
        // entry is owned by one thread of the group); when the walk moves to the next n-tile, and at its end, the 256
        // epilogue threads merge the two groups and write the row's slice of that n-tile.  N-tiles the CTA never visits are
        // written as zeros at the end, so every row is fully defined.
        int cur_nt = -1;
        uint32_t visited = 0;
        const int etid = threadIdx.x - 96;                     // 0..255 over both epilogue groups
        double* const acc_g = s_acc + grp * 2 * N_TILE;
        if (p.stats) {
            for (int c = tid; c < 2 * N_TILE; c += 128) acc_g[c] = 0.0;
            asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        auto flush_stats = [&](int nt_f) {
            asm volatile("bar.sync 3, 256;" ::: "memory");
            float* row = p.stats + (int64_t)blockIdx.x * 2 * p.cout + nt_f * N_TILE;
            for (int c = etid; c < 2 * N_TILE; c += 256) {
                const int which = c / N_TILE, cc = c - which * N_TILE;
                row[which * p.cout + cc] = (float)(s_acc[c] + s_acc[2 * N_TILE + c]);
                s_acc[c] = 0.0;
                s_acc[2 * N_TILE + c] = 0.0;
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
        };
        while (wk.next(nt, s0, cnt)) {
            if (p.stats && nt != cur_nt) {
                if (cur_nt >= 0) flush_stats(cur_nt);
                cur_nt = nt;
                visited |= 1u << nt;
            }
            const uint32_t ab = gi & 1u;
            float* sc_s = aff_g + ab * 2 * N_TILE;
            float* sh_s = sc_s + N_TILE;
            for (int c = tid; c < N_TILE; c += 128) {
                sc_s[c] = p.scale ? __ldg(p.scale + nt * N_TILE + c) : 1.f;
                sh_s[c] = p.shift ? __ldg(p.shift + nt * N_TILE + c) : 0.f;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            {
                const int j = grp;
                { HALO_T0(); mbar_wait(&acc_full[ab * 2 + j], (gi >> 1) & 1u); HALO_ACC(w0); }
                if (j < cnt) {
                tc_fence_after();
                const int sub = s0 + j;
                const int tw = sub % p.tiles_w, th = (sub / p.tiles_w) % p.tiles_h, tn = sub / (p.tiles_w * p.tiles_h);
                const int w_ = tw * p.bw + w_rel, h_ = th * p.bh + h_rel, n_ = tn * p.BN + n_rel;
                const bool valid = n_rel < p.BN && h_rel < p.bh && w_rel < p.bw && w_ < p.W && h_ < p.H && n_ < p.B;
                const int64_t pix = ((int64_t)n_ * p.H + h_) * p.W + w_;
                // Output staging.  A thread owns one position (= one TMEM lane), so direct stores touch 32 different cache
                // lines per warp instruction (16 bytes each) and the epilogue becomes LSU-bound: 4096 line accesses per fp32
                // subtile against 4608 MMA cycles at K = 576.  Each 32 x 32 chunk is transposed through a padded shared-memory
                // tile instead and leaves as whole 128-byte (fp32) / 64-byte (bf16) row segments, 4 / 8 rows per instruction.
                const uint32_t tile_s = smem_u32(s_tile + (warp - 3) * (32 * 36));
                const uint32_t sc_u = smem_u32(sc_s), sh_u = smem_u32(sh_s);
                const int64_t my_off = valid ? pix * p.ldy + p.y_coff + nt * N_TILE : (int64_t)-1;
                const bool out_bf16 = p.y_dtype == TD_BF16;
                const int rpi = out_bf16 ? 8 : 4;                      // rows per store instruction
                const int lpr = 32 / rpi;                              // lanes per row
                int64_t row_off[8];                                    // element offset of the row this lane stores in pass `it`
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = it * rpi + lane / lpr;
                    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)(my_off & 0xffffffffll), row & 31);
                    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)my_off >> 32), row & 31);
                    row_off[it] = (int64_t)(((uint64_t)hi << 32) | lo);
                }
                // Software pipeline over the 32-column chunks: the TMEM load of chunk k+1 is in flight while chunk k is read
                // back from the tile and stored (LDTM was the top stall of the epilogue); one register set suffices because
                // the read-back works from shared memory.
                uint32_t rr[32];
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * ACC_COLS + j * N_TILE;
                // bf16 outputs WITHOUT an affine (train mode: the raw conv output; BatchNorm follows as its own kernel): rounding
                // happens in the thread-per-position layout and the tile holds finished bf16 rows -- 64 bytes per position, 16-byte
                // piece g of row l at piece g ^ ((l >> 1) & 3): conflict-free for the stores (8 rows per quarter warp) and for the
                // read-back (2 rows x 4 pieces) -- half the shared-memory traffic of the fp32 tile.  With the folded eval-mode affine
                // the same layout needs 16 broadcast loads of scale / shift per chunk and thread and measured SLOWER (sampler 424.6
                // -> 434.2 us per reverse step): eval keeps the fp32 tile and applies the affine after the transpose.
                constexpr bool raw_bf16 = EPI == 1, epi16 = EPI == 2 || EPI == 3, stats16 = EPI == 3;      // (chosen by the host: halo_plan_run)
                auto stage = [&](int c0) {            // registers -> tile (+ train-mode statistics of this chunk)
                    if (stats16) {
                        // column sums / sums of squares of this warp's 32 positions (invalid positions count as zero)
                        const uint32_t vm = __ballot_sync(0xffffffffu, valid);
                        float V[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) V[i] = 0.f;
#pragma unroll
                        for (int hu = 0; hu < 4; ++hu) {
                            const bool rv = (vm >> (8 * hu + (lane >> 2))) & 1u;
#pragma unroll
                            for (int jw = 0; jw < 8; ++jw) {
                                const float v = rv ? __uint_as_float(rr[16 * (hu >> 1) + 4 * (jw >> 1) + 2 * (hu & 1) + (jw & 1)]) : 0.f;
                                V[jw] += v;
                                V[8 + jw] = fmaf(v, v, V[8 + jw]);
                            }
                        }
                        // halving butterfly over the row groups (lane bits 4, 3, 2): every step a lane keeps one half of its live values
                        // and sends the other; lane l ends with kind = bit 4 (0 sum, 1 squares) of columns 8(2 b3 + b2) + 2(l%4) + {0, 1}
#pragma unroll
                        for (int st = 0; st < 3; ++st) {
                            const int n = 8 >> st, xo = 16 >> st;
                            const bool hi = (lane & xo) != 0;
#pragma unroll
                            for (int i = 0; i < n; ++i) {
                                const float send = hi ? V[i] : V[n + i];
                                const float keep = hi ? V[n + i] : V[i];
                                V[i] = keep + __shfl_xor_sync(0xffffffffu, send, xo);
                            }
                        }
                        const int kind = (lane >> 4) & 1, col = 8 * (2 * ((lane >> 3) & 1) + ((lane >> 2) & 1)) + 2 * (lane & 3);
                        stats_g[(q * 2 + kind) * N_TILE + c0 + col] = V[0];
                        stats_g[(q * 2 + kind) * N_TILE + c0 + col + 1] = V[1];
                    }
                    if (epi16) {
                        // rr[16h + 4j + 2u + w] = (row 16h + 8u + lane/4, column 8j + 2(lane%4) + w): eight columns per thread, so scale /
                        // shift are 16 registers (eight 8-byte shared-memory reads); the rounded pair (w = 0, 1) is one 4-byte store into
                        // the bf16 tile of the raw path (piece j of row r at piece j ^ ((r >> 1) & 3): 8 rows x 4 words per store
                        // instruction hit 32 distinct banks)
                        float2 sc2[4], sh2[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            sc2[jj] = stats16 ? make_float2(1.f, 1.f) : lds64(sc_u + (uint32_t)(c0 + 8 * jj + 2 * (lane & 3)) * 4u);
                            sh2[jj] = stats16 ? make_float2(0.f, 0.f) : lds64(sh_u + (uint32_t)(c0 + 8 * jj + 2 * (lane & 3)) * 4u);
                        }
#pragma unroll
                        for (int hu = 0; hu < 4; ++hu) {
                            const uint32_t row = (uint32_t)(8 * hu + (lane >> 2));        // 16h + 8u = 8 * (2h + u)
                            const uint32_t row_s = tile_s + row * 64u + (uint32_t)(lane & 3) * 4u, sw = (row >> 1) & 3u;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float a = __uint_as_float(rr[16 * (hu >> 1) + 4 * jj + 2 * (hu & 1) + 0]);
                                float b = __uint_as_float(rr[16 * (hu >> 1) + 4 * jj + 2 * (hu & 1) + 1]);
                                if (!stats16) {
                                    a = fmaf(a, sc2[jj].x, sh2[jj].x);
                                    b = fmaf(b, sc2[jj].y, sh2[jj].y);
                                    if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                                }
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                                sts32(row_s + (((uint32_t)jj ^ sw) << 4), *reinterpret_cast<const uint32_t*>(&h2));
                            }
                        }
                        __syncwarp();
                        return;
                    }
                    if (raw_bf16) {
                        if (p.stats) {
                            float v[32], sq[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                v[i] = valid ? __uint_as_float(rr[i]) : 0.f;
                                sq[i] = v[i] * v[i];
                            }
                            const float cs = warp_column_sums(v, lane);
                            const float cq = warp_column_sums(sq, lane);
                            stats_g[(q * 2 + 0) * N_TILE + c0 + lane] = cs;
                            stats_g[(q * 2 + 1) * N_TILE + c0 + lane] = cq;
                        }
                        const uint32_t row_s = tile_s + (uint32_t)lane * 64u, sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint32_t w[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(rr[8 * g + 2 * k]), __uint_as_float(rr[8 * g + 2 * k + 1]));
                                w[k] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            sts128(row_s + (((uint32_t)g ^ sw) << 4), w[0], w[1], w[2], w[3]);
                        }
                        __syncwarp();
                        return;
                    }
                    // raw accumulators go through the tile; the folded BatchNorm affine and the ReLU are applied after the
                    // transpose, where a lane owns a fixed group of 8 (bf16) / 4 (fp32) channels for all rows: 4 loads of
                    // scale / shift per chunk instead of 16 per thread (they were ~25 % of the epilogue's stall samples)
#pragma unroll
                    for (int i = 0; i < 32; i += 4) sts128(tile_s + (uint32_t)(lane * 36 + i) * 4u, rr[i], rr[i + 1], rr[i + 2], rr[i + 3]);
                    if (p.stats) {
                        // statistics of the stored values (train mode stores the raw conv output: scale / shift NULL, no ReLU)
                        float v[32], sq[32];
                        const bool affine = p.scale || p.shift;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float t = __uint_as_float(rr[i]);
                            if (affine) t = fmaf(t, sc_s[c0 + i], sh_s[c0 + i]);
                            if (p.relu) t = fmaxf(t, 0.f);
                            v[i] = valid ? t : 0.f;
                            sq[i] = v[i] * v[i];
                        }
                        const float cs = warp_column_sums(v, lane);
                        const float cq = warp_column_sums(sq, lane);
                        stats_g[(q * 2 + 0) * N_TILE + c0 + lane] = cs;
                        stats_g[(q * 2 + 1) * N_TILE + c0 + lane] = cq;
                    }
                    __syncwarp();
                };
                // pooled pixel of this thread in this subtile (the window elements' validity against the image edge)
                const int p_h0 = th * p.bh, p_w0 = tw * p.bw;
                const int p_n = tn * p.BN + pp_n, p_ph = (p_h0 >> 1) + pp_h, p_pw = (p_w0 >> 1) + pp_w;
                const bool p_store = pp_on && j < cnt && p_n < p.B && p_ph < p.Hp && p_pw < p.Wp;
                const uint32_t gtile_s = smem_u32(s_tile + grp * (4 * 32 * 36));
                if (epi16) tmem_ld_16x256_pair(t_addr, rr);
                else tmem_ld_32x32(t_addr, rr);
                tmem_ld_wait();
                stage(0);
#pragma unroll 1
                for (int c0 = 0; c0 < N_TILE; c0 += 32) {
                    const bool more = c0 + 32 < N_TILE;
                    if (more) {
                        if (epi16) tmem_ld_16x256_pair(t_addr + (uint32_t)(c0 + 32), rr);
                        else tmem_ld_32x32(t_addr + (uint32_t)(c0 + 32), rr);
                    }
                    if (raw_bf16 || epi16) {
                        // finished bf16 rows: lane = (row it * 8 + lane / 4, 16-byte piece lane % 4) -> 64-byte row segments
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int row = it * 8 + (lane >> 2);
                            const float4 v = lds128(tile_s + (uint32_t)row * 64u + ((((uint32_t)lane & 3u) ^ ((uint32_t)(row >> 1) & 3u)) << 4));
                            if (row_off[it] >= 0)
                                *reinterpret_cast<float4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + row_off[it] + c0 + (lane & 3) * 8) = v;
                        }
                    } else if (out_bf16) {
                        const int col = (lane & 3) * 8;
                        float sc[8], sh[8];
                        *reinterpret_cast<float4*>(sc) = lds128(sc_u + (uint32_t)(c0 + col) * 4u);
                        *reinterpret_cast<float4*>(sc + 4) = lds128(sc_u + (uint32_t)(c0 + col + 4) * 4u);
                        *reinterpret_cast<float4*>(sh) = lds128(sh_u + (uint32_t)(c0 + col) * 4u);
                        *reinterpret_cast<float4*>(sh + 4) = lds128(sh_u + (uint32_t)(c0 + col + 4) * 4u);
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int row = it * 8 + (lane >> 2);
                            const float4 a = lds128(tile_s + (uint32_t)(row * 36 + col) * 4u);
                            const float4 b = lds128(tile_s + (uint32_t)(row * 36 + col + 4) * 4u);
                            float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                f[k] = fmaf(f[k], sc[k], sh[k]);
                                if (p.relu) f[k] = fmaxf(f[k], 0.f);
                            }
                            const Vec<__nv_bfloat16> packed = Vec<__nv_bfloat16>::pack(f);
                            if (row_off[it] >= 0) packed.store(reinterpret_cast<__nv_bfloat16*>(p.y) + row_off[it] + c0 + col);
                            // fused max-pool: the finished bf16 values replace the row's first 64 bytes in place (the four lanes of a
                            // row read its 128 fp32 bytes in the two loads above, before this store of the same warp instruction)
                            if (p.pool_y) sts128(tile_s + (uint32_t)(row * 36) * 4u + (uint32_t)(lane & 3) * 16u, packed.v.x, packed.v.y, packed.v.z, packed.v.w);
                        }
                    } else {
                        const int col = (lane & 7) * 4;
                        const float4 sc = lds128(sc_u + (uint32_t)(c0 + col) * 4u);
                        const float4 sh = lds128(sh_u + (uint32_t)(c0 + col) * 4u);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = it * 4 + (lane >> 3);
                            float4 a = lds128(tile_s + (uint32_t)(row * 36 + col) * 4u);
                            a.x = fmaf(a.x, sc.x, sh.x); a.y = fmaf(a.y, sc.y, sh.y); a.z = fmaf(a.z, sc.z, sh.z); a.w = fmaf(a.w, sc.w, sh.w);
                            if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                            if (row_off[it] >= 0) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + row_off[it] + c0 + col) = a;
                        }
                    }
                    if (p.pool_y) {
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // the four warps' rows hold their bf16 results
                        if (p_store) {
                            const uint32_t NEG_INF2 = 0xFF80FF80u;                          // two bf16 -inf
                            uint32_t m[4] = {NEG_INF2, NEG_INF2, NEG_INF2, NEG_INF2};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int hh = p_h0 + 2 * pp_h + (e >> 1), ww = p_w0 + 2 * pp_w + (e & 1);
                                if (pp_r[e] < 0 || hh >= p.H || ww >= p.W) continue;
                                const uint32_t pr = (uint32_t)pp_r[e];      // warp slot pr >> 5, row pr & 31 of that warp's tile
                                const float4 a = epi16 ? lds128(gtile_s + (pr >> 5) * (32u * 36u * 4u) + (pr & 31u) * 64u +
                                                                ((((uint32_t)tid & 3u) ^ ((pr >> 1) & 3u)) << 4))
                                                       : lds128(gtile_s + (uint32_t)(pp_r[e] * 36) * 4u + (uint32_t)(tid & 3) * 16u);
                                const uint32_t v[4] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&m[k]),
                                                                     *reinterpret_cast<const __nv_bfloat162*>(&v[k]));
                                    m[k] = *reinterpret_cast<const uint32_t*>(&r);
                                }
                            }
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.pool_y) +
                                                      (((int64_t)p_n * p.Hp + p_ph) * p.Wp + p_pw) * p.cout + nt * N_TILE + c0 + (tid & 3) * 8) =
                                make_uint4(m[0], m[1], m[2], m[3]);
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // pooled reads done: the tile may be restaged
                    } else {
                        __syncwarp();
                    }
                    if (more) {
                        tmem_ld_wait();
                        stage(c0 + 32);
                    }
                }
                if (p.stats) {
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    for (int c = tid; c < 2 * N_TILE; c += 128) {
                        const int which = c / N_TILE, cc = c - which * N_TILE;
                        const float t = (stats_g[(0 * 2 + which) * N_TILE + cc] + stats_g[(1 * 2 + which) * N_TILE + cc]) +
                                        (stats_g[(2 * 2 + which) * N_TILE + cc] + stats_g[(3 * 2 + which) * N_TILE + cc]);
                        acc_g[c] += (double)t;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
            ++gi;
        }
        if (p.stats) {
            if (cur_nt >= 0) flush_stats(cur_nt);
            const int n_tiles = p.cout / N_TILE;
            for (int t2 = 0; t2 < n_tiles; ++t2) {
                if (visited & (1u << t2)) continue;
                float* row = p.stats + (int64_t)blockIdx.x * 2 * p.cout + t2 * N_TILE;
                for (int c = etid; c < 2 * N_TILE; c += 256) {
                    const int which = c / N_TILE, cc = c - which * N_TILE;
                    row[which * p.cout + cc] = 0.f;
                }
            }
            if (blockIdx.x == 0)        // the finalize reads a "shift" row K after the partial rows: zero for the conv epilogue
                for (int c = etid; c < p.cout; c += 256) p.stats[(int64_t)gridDim.x * 2 * p.cout + c] = 0.f;
        }
        if ((p.dbg & 1) && tid == 0 && grp == 0) {
            g_halo_dbg[blockIdx.x * 8 + 6] = w0;
            g_halo_dbg[blockIdx.x * 8 + 7] = clock64() - t_start - w0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: geometry, tensor maps, launch
// ---------------------------------------------------------------------------------------------
struct HaloGeom {
    int G, PW, bh, BN;
    double eff;        // useful positions / computed positions
    int strip;         // 1: 8-column strips (PW = 10, 16 rows per subtile)
};

// Best subtile (<= 128 positions) of an H x W feature map for each layout; eff = 0 when not applicable.
static HaloGeom best_flat(int B, int H, int W, bool even_bh) {
    HaloGeom g{1, W + 1, 0, 1, 0.0, 0};
    const int PW = W + 1;
    for (int bh = 1; bh <= H; ++bh) {               // bh rows of one image
        if ((bh - 1) * PW + W - 1 >= 128) break;
        if (even_bh && (bh & 1) && bh < H) continue;        // fused 2x2 max-pool: windows must not straddle two subtiles
        const double eff = (double)H * W / ((double)ceil_div(H, bh) * 128.0);
        if (eff > g.eff + 1e-9) { g.eff = eff; g.bh = bh; g.BN = 1; }
    }
    const int img = (H + 2) * PW;                   // BN whole images
    for (int bn = 2; bn <= B; ++bn) {
        if ((bn - 1) * img + (H - 1) * PW + W - 1 >= 128 || bn * (H + 2) > 256 || bn > 256) break;
        const double eff = (double)B * H * W / ((double)ceil_div(B, bn) * 128.0);
        if (eff > g.eff + 1e-9) { g.eff = eff; g.bh = H; g.BN = bn; }
    }
    return g;
}
static HaloGeom best_dx(int H, int W) {
    HaloGeom g{3, W, 0, 1, 0.0, 0};
    if (W % 8 != 0 || W > 128) return g;
    for (int bh = 1; bh <= H && bh * W <= 128; ++bh) {
        const double eff = (double)H * W / ((double)ceil_div(H, bh) * 128.0);
        if (eff > g.eff + 1e-9) { g.eff = eff; g.bh = bh; }
    }
    return g;
}

static HaloGeom best_strip(int H, int W) {
    HaloGeom g{1, 10, 16, 1, 0.0, 1};
    if (W % 8 != 0) return g;
    g.eff = (double)H / ((double)ceil_div(H, 16) * 16.0);
    return g;
}

bool halo_plan_init(td_conv_plan* p, int* status) {
    const td_conv3x3_desc& d = p->d;
    *status = TD_OK;
    if (const char* e = getenv("TD_TC_HALO")) if (atoi(e) == 0) return false;
    if (d.width > 255 || d.height > 253) return false;
    const int n_tile = d.cout % 128 == 0 ? 128 : 64;
    // Relative cost per useful output: MMA cycles ~ 1/eff; L2 -> SM bytes ~ (A boxes + shared B) at ~45 B/clk.
    // eval-mode layers followed by MaxPool2d(2): the epilogue writes the pooled map too when every window lies inside a subtile
    const bool want_pool = d.pool_y && !d.stats && d.y_dtype == TD_BF16 && getenv("TD_TC_HALO_POOL") == nullptr;
    HaloGeom cand[3] = {best_flat(d.batch, d.height, d.width, want_pool), best_strip(d.height, d.width), best_dx(d.height, d.width)};
    int best = -1;
    double best_cost = 0;
    for (int i = 0; i < 2; ++i) {      // the dx layout moves 3x the activation bytes and measures slower than the per-tap kernel: opt-in only
        const HaloGeom& g = cand[i];
        if (g.eff < 0.7) continue;
        const double mma_clk = 9.0 * 4.0 * n_tile / 2.0;                                   // per subtile and chunk
        const double a_bytes = (double)g.G * g.PW * (g.bh + 2) * g.BN * 128.0;
        const double b_bytes = 9.0 * n_tile * 128.0 / HALO_S;
        const double cost = std::max(mma_clk, (a_bytes + b_bytes) / 45.0) / g.eff;
        if (best < 0 || cost < best_cost) { best = i; best_cost = cost; }
    }
    if (const char* e = getenv("TD_TC_HALO_MODE")) {       // 1: flat, 2: strip, 3: dx (tests / sweeps)
        const int want = atoi(e);
        for (int i = 0; i < 3; ++i) {
            const int kind = cand[i].G == 3 ? 3 : (cand[i].strip ? 2 : 1);
            if (kind == want && cand[i].eff > 0) best = i;
        }
    }
    if (best < 0) return false;
    const HaloGeom g = cand[best];
    EncodeTiledFn encode = tc_get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled not available from the driver"); *status = TD_ERR_DRIVER; return true; }

    p->halo = 1;
    p->h_groups = g.G; p->h_pw = g.PW; p->h_bh = g.bh; p->h_bn = g.BN; p->h_rh = g.bh + 2;
    p->h_strip = g.strip;
    p->tiles_w = g.strip ? d.width / 8 : 1;
    p->tiles_h = (int)ceil_div(d.height, g.bh);
    p->tiles_n = (int)ceil_div(d.batch, g.BN);
    p->h_nsub = p->tiles_w * p->tiles_h * p->tiles_n;
    p->block_n = n_tile;
    p->n_tiles = d.cout / n_tile;
    p->h_units = p->h_nsub * p->n_tiles;
    {
        const int bw = g.strip ? 8 : d.width;
        p->h_pool = want_pool && g.G == 1 && (g.bh % 2 == 0 || p->tiles_h == 1) && (bw % 2 == 0 || p->tiles_w == 1) &&
                    ((bw + 1) / 2) * ((g.bh + 1) / 2) * g.BN <= 32;
    }
    p->h_grid = std::min(p->h_units, sm_budget());
    {   // The walk's duration is ceil(units / grid) units whatever the grid: take the SMALLEST grid with the same number of rounds
        // (896 / 512 / 256 units at batch 128 -> 128 CTAs of exactly 7 / 4 / 2 units instead of 148 CTAs of which a handful run the
        // extra one).  Every CTA then finishes together and the idle SMs' share of the power budget goes into the clock.
        const char* e = getenv("TD_TC_HALO_BALANCE");
        if (!e || atoi(e) != 0) {
            const int rounds = (int)ceil_div(p->h_units, p->h_grid);
            p->h_grid = (int)ceil_div(p->h_units, rounds);
        }
    }
    p->split_k = 1;
    const int box_bytes = g.PW * p->h_rh * g.BN * 128;
    p->h_slot_bytes = (box_bytes + 1023) / 1024 * 1024 + 1024;
    p->h_na = g.G == 3 ? 6 : 4;        // dx layout: a box group lasts only three taps, keep three groups in flight
    if (const char* e = getenv("TD_TC_HALO_NA")) { int v = atoi(e); if (v >= 2 && v <= 8) p->h_na = v; }
    const int b_stage = n_tile * 128;
    const int fixed = p->h_na * p->h_slot_bytes + (2 * p->h_na + 6) * 8 + 64 + (8 + 16) * n_tile * 4 + 8 * 32 * 36 * 4 + 1024 +
                      4 * n_tile * 8 + 16;
    int nb = (226 * 1024 - fixed) / (b_stage + 16);
    if (nb > 8) nb = 8;
    if (const char* e = getenv("TD_TC_HALO_NB")) { int v = atoi(e); if (v >= 2 && v <= nb) nb = v; }
    if (nb < 3) { p->halo = 0; return false; }
    p->h_nb = nb;
    p->stages = nb;
    p->smem_bytes = fixed + nb * (b_stage + 16);

    {   // activations: (C, W, H, N), box (64, PW, bh + 2, BN); out-of-bounds elements are zero filled
        cuuint64_t gdim[4] = {(cuuint64_t)d.ldx, (cuuint64_t)d.width, (cuuint64_t)d.height, (cuuint64_t)d.batch};
        cuuint64_t gstr[3] = {(cuuint64_t)d.ldx * 2, (cuuint64_t)d.width * d.ldx * 2,
                              (cuuint64_t)d.height * d.width * d.ldx * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)g.PW, (cuuint32_t)p->h_rh, (cuuint32_t)g.BN};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&p->tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo activations) failed: %d", (int)r); *status = TD_ERR_DRIVER; return true; }
    }
    {   // weights: (K = 9*Cin, Cout)
        cuuint64_t gdim[2] = {(cuuint64_t)9 * d.cin, (cuuint64_t)d.cout};
        cuuint64_t gstr[1] = {(cuuint64_t)9 * d.cin * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)n_tile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&p->tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.w), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo weights) failed: %d", (int)r); *status = TD_ERR_DRIVER; return true; }
    }
    return true;
}

template <int N_TILE, int EPI>
static int launch_halo(const td_conv_plan* p, const HaloParams& prm, cudaStream_t s) {
    static int configured_smem = 0;
    if (p->smem_bytes > configured_smem) {
        TD_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<N_TILE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes));
        configured_smem = p->smem_bytes;
    }
    const int grid = p->h_grid;
    td::launch(conv3x3_halo_kernel<N_TILE, EPI>, td::LaunchCfg(grid, HALO_THREADS, p->smem_bytes, s), p->tmap_x, p->tmap_w, prm);
    return launch_status("conv3x3_halo");
}

int halo_plan_run(const td_conv_plan* p, cudaStream_t s) {
    const td_conv3x3_desc& d = p->d;
    HaloParams prm;
    prm.B = d.batch; prm.H = d.height; prm.W = d.width; prm.cin = d.cin; prm.x_coff = d.x_coff;
    prm.cout = d.cout; prm.ldy = d.ldy; prm.y_coff = d.y_coff; prm.y_dtype = d.y_dtype;
    prm.y = d.y; prm.scale = d.scale; prm.shift = d.shift; prm.relu = d.relu; prm.stats = d.stats;
    prm.G = p->h_groups; prm.PW = p->h_pw; prm.bh = p->h_bh; prm.BN = p->h_bn; prm.RH = p->h_rh;
    prm.tiles_w = p->tiles_w; prm.tiles_h = p->tiles_h; prm.n_sub = p->h_nsub; prm.units = p->h_units;
    prm.bw = p->h_strip ? 8 : d.width;
    prm.ew = p->h_strip ? 8 : p->h_pw;
    prm.a_sbo = p->h_strip ? p->h_pw * 128 : 1024;
    prm.NA = p->h_na; prm.NB = p->h_nb;
    prm.a_box_bytes = (uint32_t)(p->h_pw * p->h_rh * p->h_bn * 128);
    prm.a_slot_bytes = (uint32_t)p->h_slot_bytes;
    {
        const char* e = getenv("TD_TC_HALO_DBG");
        prm.dbg = e ? atoi(e) : 0;
    }
    prm.pool_y = p->h_pool ? d.pool_y : nullptr;
    prm.pool_ceil = d.pool_ceil;
    prm.Hp = d.pool_ceil ? (d.height + 1) / 2 : d.height / 2;
    prm.Wp = d.pool_ceil ? (d.width + 1) / 2 : d.width / 2;
    static int raw_ok = -1, epi16_ok = -1;
    if (raw_ok < 0) { const char* e = getenv("TD_TC_HALO_RAW"); raw_ok = (e && atoi(e) == 0) ? 0 : 1; }
    if (epi16_ok < 0) { const char* e = getenv("TD_TC_HALO_EPI16"); epi16_ok = (e && atoi(e) == 0) ? 0 : 1; }
    const bool plain = !d.scale && !d.shift && !d.relu && !prm.pool_y;
    static int stats16_ok = -1;
    if (stats16_ok < 0) { const char* e = getenv("TD_TC_HALO_STATS16"); stats16_ok = (e && atoi(e) == 0) ? 0 : 1; }
    if (stats16_ok && raw_ok && d.y_dtype == TD_BF16 && plain && d.stats)
        return p->block_n == 128 ? launch_halo<128, 3>(p, prm, s) : launch_halo<64, 3>(p, prm, s);
    if (raw_ok && d.y_dtype == TD_BF16 && plain) return p->block_n == 128 ? launch_halo<128, 1>(p, prm, s) : launch_halo<64, 1>(p, prm, s);
    if (epi16_ok && d.y_dtype == TD_BF16 && !plain && !d.stats)
        return p->block_n == 128 ? launch_halo<128, 2>(p, prm, s) : launch_halo<64, 2>(p, prm, s);
    return p->block_n == 128 ? launch_halo<128, 0>(p, prm, s) : launch_halo<64, 0>(p, prm, s);
}

}  // namespace td

// Tuning aid: copies the wait-cycle counters of the last halo launch run with TD_TC_HALO_DBG=1 (synchronises).
extern "C" int td_conv3x3_debug_counters(unsigned long long* host_out, int n) {
    if (!host_out || n <= 0 || n > td::kNumSMs * 8) { td::set_error("td_conv3x3_debug_counters: bad arguments"); return TD_ERR_ARG; }
    TD_CUDA(cudaDeviceSynchronize());
    TD_CUDA(cudaMemcpyFromSymbol(host_out, td::g_halo_dbg, (size_t)n * sizeof(unsigned long long)));
    return TD_OK;
}
