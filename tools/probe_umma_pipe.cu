// Hardware probe (B200): what slows a tcgen05.mma stream below its 64-cycle issue rate (M=128, N=128, K=16)?
// Adds, one at a time, the elements of a real pipelined kernel to the bare issue loop of probe_umma_rate.cu:
//   mode 0  bare loop                      mode 1  + tcgen05.commit to a (never waited) mbarrier after every 8 MMAs
//   mode 2  + full/empty ring handshake with a producer thread (no data movement, ring depth 6)
//   mode 3  mode 2 + four warps streaming the accumulator out with tcgen05.ld (epilogue traffic)
//   mode 4  mode 0 + the epilogue warps only
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc \
//          tools/probe_umma_pipe.cu -o tools/build/probe_umma_pipe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "sm100.cuh"

using namespace td::sm100;

constexpr int N = 128, STAGES = 6, A_BYTES = 128 * 128, B_BYTES = N * 128;

__global__ void __launch_bounds__(192, 1) pipe_kernel(long long* out, int taps, int mode, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    uint64_t* dummy = done + 1;
    volatile int* stop = reinterpret_cast<volatile int*>(dummy + 1);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(const_cast<int*>(stop) + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
    uint32_t st = 0x9E3779B9u * (threadIdx.x + 1 + blockIdx.x * 131u);
    for (int i = threadIdx.x; i < STAGES * (A_BYTES + B_BYTES) / 4; i += 192) {
        st = st * 1664525u + 1013904223u;
        w32[i] = (0x3F00u | ((st >> 9) & 0x807Fu)) | ((0x3F00u | ((st >> 17) & 0x807Fu)) << 16);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_init(dummy, 1);
        *stop = 0;
        fence_barrier_init();
    }
    fence_proxy_async();
    if (warp == 2) tmem_alloc<256>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool ring = mode == 2 || mode == 3 || mode == 6 || mode == 9;   // 6: ring without tcgen05.fence::after_thread_sync
    volatile int* ready = stop + 2;      // mode 9: a watcher warp waits on the full barriers and publishes a plain counter
    if (threadIdx.x == 0) *ready = 0;
    if (mode == 8 && threadIdx.x == 0) mbar_arrive(dummy);      // 8: bare + wait on an already completed barrier per tap
    if (warp == 0) {
        if (ring && elect_one()) {
            for (int t = 0; t < taps; ++t) {
                const int s = t % STAGES;
                mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
                mbar_arrive(&full[s]);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
            if (mode == 5) {
                // thread-side cost of each step, starting from an idle tensor pipe
                long long c[6];
                const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
                mbar_arrive(&full[0]);
                c[0] = clock64();
                mbar_wait(&full[0], 0);
                tc_fence_after();
                c[1] = clock64();
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_bf16(tmem_base + j * N, da, db, idesc, k ? 1u : 0u);
                    }
                }
                c[2] = clock64();
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + A_BYTES + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(b_addr + B_BYTES + k * 32, 16, 1024);
                        umma_bf16(tmem_base + j * N, da, db, idesc, 1u);
                    }
                }
                c[3] = clock64();
                umma_commit(done);
                c[4] = clock64();
                mbar_wait(done, 0);
                c[5] = clock64();
                for (int i = 0; i < 5; ++i) out[blockIdx.x * 8 + i] = c[i + 1] - c[i];
                *stop = 1;
            } else {
            const long long t0 = clock64();
            for (int t = 0; t < taps; ++t) {
                const int s = t % STAGES;
                if (mode == 9) {
                    while (*ready <= t) { }
                    tc_fence_after();
                } else if (ring) {
                    mbar_wait(&full[s], (t / STAGES) & 1);
                    if (mode != 6) tc_fence_after();
                }
                if (mode == 7) tc_fence_after();                  // 7: bare + fence per tap
                if (mode == 8) mbar_wait(dummy, 0);
                const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_bf16(tmem_base + j * N, da, db, idesc, (t | k) ? 1u : 0u);
                    }
                }
                if (ring) umma_commit(&empty[s]);
                else if (mode == 1) umma_commit(dummy);
            }
            umma_commit(done);
            mbar_wait(done, 0);
            out[blockIdx.x] = clock64() - t0;
            *stop = 1;
            }
        }
    } else if (mode == 9) {
        if (warp == 2 && elect_one()) {
            for (int t = 0; t < taps; ++t) {
                mbar_wait(&full[t % STAGES], (t / STAGES) & 1);
                *ready = t + 1;
            }
        }
    } else if (mode >= 3 && mode != 9) {
        const int q = warp & 3;
        float acc = 0.f;
        while (!*stop) {
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                tmem_ld_wait();
                for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
            }
        }
        if (acc == 123.456f) sink[threadIdx.x] = acc;
        (void)lane;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

int main() {
    const int taps = 2048, ctas = 148;
    long long* d;
    float* sink;
    cudaMalloc(&d, ctas * 8 * sizeof(long long));
    cudaMalloc(&sink, 1024);
    const int smem = STAGES * (A_BYTES + B_BYTES) + 256 + 1024;
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"bare issue loop", "+ commit every 8 MMAs", "+ full/empty ring with producer", "ring + tcgen05.ld epilogue warps",
                           "bare + tcgen05.ld epilogue warps", "", "ring, no fence::after_thread_sync", "bare + fence per tap",
                           "bare + wait on completed mbarrier", "ring, waits relayed by a watcher warp"};
    for (int mode = 0; mode < 10; ++mode) {
        if (mode == 5) continue;
        for (int rep = 0; rep < 2; ++rep) {
            pipe_kernel<<<ctas, 192, smem>>>(d, taps, mode, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> h(ctas);
        cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum = 0;
        for (int i = 0; i < ctas; ++i) sum += h[i];
        printf("mode %d %-36s: %6.1f cycles per MMA (M=128 N=128 K=16; 64 = full rate)\n", mode, names[mode], sum / ctas / (taps * 8.0));
    }
    pipe_kernel<<<ctas, 192, smem>>>(d, taps, 5, sink);
    cudaDeviceSynchronize();
    pipe_kernel<<<ctas, 192, smem>>>(d, taps, 5, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode 5 failed\n"); return 1; }
    std::vector<long long> h(ctas * 8);
    cudaMemcpy(h.data(), d, ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("mode 5 thread-side cycles (CTA 0 / CTA 77): wait on completed mbarrier + fence %lld / %lld ; issue 8 MMAs (idle pipe) %lld / %lld ; "
           "issue 8 more %lld / %lld ; commit %lld / %lld ; wait for completion %lld / %lld\n",
           h[0], h[77 * 8], h[1], h[77 * 8 + 1], h[2], h[77 * 8 + 2], h[3], h[77 * 8 + 3], h[4], h[77 * 8 + 4]);
    return 0;
}
