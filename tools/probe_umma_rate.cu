// Hardware probe (B200): issue rate of tcgen05.mma kind::f16 (bf16 in, fp32 accumulate) from shared-memory
// operands, as a function of N and of the CTA group (1 CTA, M=128  vs  CTA pair, M=256).  No global traffic:
// operands are generated in shared memory, the same K=64 stage is multiplied over and over.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc \
//          tools/probe_umma_rate.cu -o tools/build/probe_umma_rate
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "sm100.cuh"

using namespace td::sm100;

// PAIR = 0: cta_group::1, M = 128, B tile = N rows.   PAIR = 1: cta_group::2, M = 256, each CTA holds N/2 rows of B.
template <int N, int PAIR, int STAGES, int MN = 0, int ROWOFF = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_ROWS = PAIR ? N / 2 : N;
    constexpr int B_BYTES = B_ROWS * 128;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* done = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5;
    // pseudo-random bf16 operands in [-1, 1)
    uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
    const int words = STAGES * (A_BYTES + B_BYTES) / 4;
    uint32_t st = 0x9E3779B9u * (threadIdx.x + 1 + blockIdx.x * 131u);
    for (int i = threadIdx.x; i < words; i += 128) {
        st = st * 1664525u + 1013904223u;
        const uint32_t lo = 0x3F00u | ((st >> 9) & 0x807Fu), hi = 0x3F00u | ((st >> 17) & 0x807Fu);
        w32[i] = lo | (hi << 16);
    }
    if (threadIdx.x == 0) {
        mbar_init(done, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    uint32_t rank = 0;
    if (PAIR) {
        rank = cluster_ctarank();
        cluster_sync_all();
    }
    if (warp == 0) {
        if (PAIR) tmem_alloc_pair<N>(tmem_ptr);
        else tmem_alloc<N>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, MN, MN);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // MN-major (weight-gradient form): a stage is [64 K rows][128 B] per 64-channel box, boxes 8 KB apart
                const uint64_t da = MN ? make_smem_desc_sw128(a_addr + ROWOFF * 128 + (k & 1) * 2048, 8192, 1024)
                                       : make_smem_desc_sw128(a_addr + ROWOFF * 128 + k * 32, 16, 1024);
                const uint64_t db = MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024) : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                if (PAIR) umma_bf16_pair(tmem_base, da, db, idesc, (it | k) ? 1u : 0u);
                else umma_bf16(tmem_base, da, db, idesc, (it | k) ? 1u : 0u);
            }
        }
        if (PAIR) umma_commit_pair(done);
        else umma_commit(done);
    }
    mbar_wait(done, 0);
    if (threadIdx.x == 0 && rank == 0) {
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair<N>(tmem_base);
        else tmem_dealloc<N>(tmem_base);
    }
}

template <int N, int PAIR, int STAGES, int MN = 0, int ROWOFF = 0>
static void run(const char* name, int ctas, int iters) {
    long long* d;
    cudaMalloc(&d, ctas * sizeof(long long));
    cudaMemset(d, 0, ctas * sizeof(long long));
    const int smem = STAGES * (128 * 128 + (PAIR ? N / 2 : N) * 128) + 64 + 1024;
    auto kern = rate_kernel<N, PAIR, STAGES, MN, ROWOFF>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t le = cudaLaunchKernelEx(&cfg, kern, d, iters);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (le != cudaSuccess || e != cudaSuccess) {
            printf("%s: CUDA error %s / %s\n", name, cudaGetErrorString(le), cudaGetErrorString(e));
            exit(1);
        }
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum = 0;
    int cnt = 0;
    for (int i = 0; i < ctas; ++i) if (h[i] > 0) { sum += h[i]; ++cnt; }
    const double cyc = sum / cnt / (iters * 4.0);
    const double flop_per_mma = 2.0 * (PAIR ? 256 : 128) * N * 16;
    const double sms = ctas;      // every CTA occupies one SM
    const double tflops = flop_per_mma * iters * 4.0 * (PAIR ? ctas / 2 : ctas) / (ms * 1e-3) / 1e12;
    printf("%-34s ctas %3d: %7.1f cyc per MMA instr = %6.1f FLOP/clk/SM ; kernel %.3f ms -> %7.1f TFLOP/s\n", name, ctas, cyc,
           flop_per_mma / cyc / (PAIR ? 2 : 1), ms, tflops);
    (void)sms;
    cudaFree(d);
}

int main() {
    const int iters = 4096;
    for (int ctas : {2, 148}) {
        run<64, 0, 4>("1-CTA M=128 N=64", ctas, iters);
        run<128, 0, 4>("1-CTA M=128 N=128", ctas, iters);
        run<256, 0, 4>("1-CTA M=128 N=256", ctas, iters);
        run<128, 0, 1>("1-CTA M=128 N=128 (one stage)", ctas, iters);
        run<128, 0, 4, 1>("1-CTA M=128 N=128 MN-major A and B", ctas, iters);
        run<64, 0, 4, 1>("1-CTA M=128 N=64  MN-major A and B", ctas, iters);
        run<128, 0, 4, 1, 1>("MN-major, A starts at row 1", ctas, iters);
        run<128, 0, 4, 1, 3>("MN-major, A starts at row 3", ctas, iters);
        run<128, 0, 4, 0, 1>("K-major, A starts at row 1", ctas, iters);
        run<128, 0, 4, 0, 5>("K-major, A starts at row 5", ctas, iters);
        run<64, 1, 4>("CTA pair M=256 N=64", ctas, iters);
        run<128, 1, 4>("CTA pair M=256 N=128", ctas, iters);
        run<256, 1, 4>("CTA pair M=256 N=256", ctas, iters);
    }
    return 0;
}
