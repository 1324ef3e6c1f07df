// Internal: the opaque plan object behind td_conv3x3_plan_create / td_conv3x3_run.
#pragma once
#include <cuda.h>

#include "common.cuh"

struct td_conv_plan {
    td_conv3x3_desc d;
    int engine;
    // --- tcgen05 engine state (conv_tc.cu) ---
    CUtensorMap tmap_x;     // activations: 4-D (C, W, H, N) box (64, bw, bh, bn), SWIZZLE_128B
    CUtensorMap tmap_w;     // weights:     2-D (9*Cin, Cout) box (64, block_n), SWIZZLE_128B
    int bw, bh, bn;         // spatial box of one M tile (bw*bh*bn <= 128 output pixels)
    int tiles_w, tiles_h, tiles_n;
    int block_n;            // UMMA N
    int n_tiles;            // cout / block_n
    int stages;
    int smem_bytes;
    int split_k;
    // --- halo variant of the tcgen05 engine (conv_halo.cu): one shared-memory box serves several taps ---
    int halo;               // 0: conv_tc.cu kernel, 1: conv_halo.cu kernel
    int h_groups;           // 1: one box (w0 = -1) serves the 9 taps; 3: one box per dx
    int h_pw, h_bh, h_bn, h_rh;   // box = (64 channels, pw, rh = bh + 2, bn images); pw = smem pixels per image row
    int h_units, h_nsub, h_na, h_nb, h_slot_bytes;
    int h_grid;             // CTAs of the persistent halo kernel = min(h_units, sm_budget() at plan creation)
    int h_strip;            // 1: 8-column strip subtiles (A descriptor group stride = h_pw * 128 bytes)
    int h_pool;             // 1: the epilogue also writes d.pool_y (2x2 max-pool windows never straddle a subtile)
};

struct td_wgrad_plan {
    td_wgrad_desc d;
    int engine;
    int splits;
    // --- tcgen05 engine state (conv_wgrad_tc.cu) ---
    CUtensorMap tmap_m, tmap_n;
    int bw, bh, bn, rows, tiles_w, tiles_h, tiles_n;
    int x_on_m, block_n, m_tiles, n_tiles, m_boxes, n_boxes, stages, boxes_per_split, smem_bytes;
    int taps, slot_bytes, ksteps;       // taps = 3: halo variant of the tcgen05 kernel (conv_wgrad_tc.cu)
};

namespace td {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode_fn();                                 // conv_tc.cu
int wgrad_tc_splits(const td_wgrad_desc& d);                      // conv_wgrad_tc.cu
int wgrad_tc_plan_init(td_wgrad_plan* p);                         // conv_wgrad_tc.cu
int wgrad_tc_plan_run(const td_wgrad_plan* p, cudaStream_t s);    // conv_wgrad_tc.cu
int tc_plan_init(td_conv_plan* p);                       // conv_tc.cu
int tc_plan_run(const td_conv_plan* p, cudaStream_t s);   // conv_tc.cu
bool halo_plan_init(td_conv_plan* p, int* status);       // conv_halo.cu; false: geometry not eligible
int halo_plan_run(const td_conv_plan* p, cudaStream_t s); // conv_halo.cu
}
