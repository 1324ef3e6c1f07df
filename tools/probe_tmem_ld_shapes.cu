// Hardware probe (B200): register layout of tcgen05.ld.16x256b.x4 (16 TMEM lanes x 32 fp32 columns per instruction) against the
// 32x32b layout the kernels use (thread i = lane i, 32 consecutive columns).  TMEM is filled with value(lane, col) = lane * 1000 + col
// through tcgen05.st.32x32b; each thread then prints what the 16x256b load hands it.  The halo convolution's eval epilogue could round
// and pack in registers with this shape (a thread owns 2 rows x 8 columns per chunk: 16 scale / shift registers instead of 64).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc tools/probe_tmem_ld_shapes.cu -o tools/build/probe_tmem_ld_shapes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sm100.cuh"

using namespace td::sm100;

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<32>(&tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    const uint32_t row = warp * 32 + lane;                       // TMEM lane of this thread in the 32x32b shape
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = row * 1000u + (uint32_t)c;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // two 16x256b.x4 loads cover the warp's 32 lanes: lane offsets 0 and 16
    uint32_t r[2][16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t a = tmem_base + ((uint32_t)(warp * 32 + h * 16) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[h][0]), "=r"(r[h][1]), "=r"(r[h][2]), "=r"(r[h][3]), "=r"(r[h][4]), "=r"(r[h][5]), "=r"(r[h][6]), "=r"(r[h][7]),
              "=r"(r[h][8]), "=r"(r[h][9]), "=r"(r[h][10]), "=r"(r[h][11]), "=r"(r[h][12]), "=r"(r[h][13]), "=r"(r[h][14]), "=r"(r[h][15])
            : "r"(a)
            : "memory");
    }
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 16; ++i) out[(threadIdx.x * 2 + h) * 16 + i] = r[h][i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tmem_base); }
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 128 * 32 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // check the conjectured layout: register 4*j + 2*u + w of half hh holds (lane = warp*32 + hh*16 + 8*u + l/4, col = 8*j + 2*(l%4) + w)
    int bad = 0;
    for (int t = 0; t < 128; ++t)
        for (int hh = 0; hh < 2; ++hh)
            for (int i = 0; i < 16; ++i) {
                const int l = t & 31, wq = t >> 5, j = i >> 2, u = (i >> 1) & 1, w = i & 1;
                const uint32_t want = (uint32_t)(wq * 32 + hh * 16 + 8 * u + l / 4) * 1000u + (uint32_t)(8 * j + 2 * (l % 4) + w);
                if (h[(t * 2 + hh) * 16 + i] != want) ++bad;
            }
    printf("conjecture reg[4j+2u+w] = (lane base + 8u + l/4, col 8j + 2(l%%4) + w): %d mismatches of %d\n", bad, 128 * 32);
    for (int t : {0, 1, 5, 33}) {
        printf("thread %3d:", t);
        for (int hh = 0; hh < 2; ++hh)
            for (int i = 0; i < 16; ++i) printf(" %u", h[(t * 2 + hh) * 16 + i]);
        printf("\n");
    }
    return 0;
}
