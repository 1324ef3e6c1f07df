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
};

namespace td {
int tc_plan_init(td_conv_plan* p);                       // conv_tc.cu
int tc_plan_run(const td_conv_plan* p, cudaStream_t s);   // conv_tc.cu
}
