// Hardware probe (B200): may a K-major SWIZZLE_128B UMMA descriptor start at an arbitrary 128-byte row of a
// TMA-written tile (not only at a 1024-byte swizzle-atom boundary)?  The halo convolution kernel relies on it
// to read the nine taps of a 3x3 stencil out of ONE shared-memory box by shifting the descriptor start.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc \
//          tools/probe_umma_rowoff.cu -o tools/build/probe_umma_rowoff -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "sm100.cuh"

using namespace td::sm100;

constexpr int ROWS = 192, N = 64, K = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out, int rowoff,
             int base_off_mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + ROWS * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_b + N * 128);
    uint64_t* done = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<64>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, ROWS * 128 + N * 128);
        tma_load_2d(smem_a, &tmap_a, bar, 0, 0);
        tma_load_2d(smem_b, &tmap_b, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t a_addr = smem_u32(smem_a) + rowoff * 128;
        const uint32_t b_addr = smem_u32(smem_b);
        for (int k = 0; k < 4; ++k) {
            uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            if (base_off_mode) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, int rows) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)K, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    std::vector<__nv_bfloat16> ha(ROWS * K), hb(N * K);
    std::vector<float> fa(ROWS * K), fb(N * K);
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { fa[i] = (float)(rand() % 17 - 8); ha[i] = __float2bfloat16(fa[i]); }
    for (int i = 0; i < N * K; ++i) { fb[i] = (float)(rand() % 9 - 4); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma = make_map(enc, da, ROWS), mb = make_map(enc, db, N);
    const int smem = ROWS * 128 + N * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int offs[] = {0, 8, 1, 2, 3, 5, 7, 9, 29, 33, 34, 63};
    std::vector<float> hout(128 * N);
    for (int mode = 0; mode < 2; ++mode)
        for (int off : offs) {
            cudaMemset(dout, 0, 128 * N * 4);
            probe_kernel<<<1, 128, smem>>>(ma, mb, dout, off, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d rowoff %d: CUDA error %s\n", mode, off, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    float ref = 0;
                    for (int k = 0; k < K; ++k) ref += fa[(off + m) * K + k] * fb[n * K + k];
                    const double err = fabs(ref - hout[m * N + n]);
                    if (err > 1e-3) ++bad;
                    if (err > maxerr) maxerr = err;
                }
            printf("base_offset_mode %d rowoff %2d: %s (mismatches %d / %d, max err %.1f)\n", mode, off, bad ? "WRONG" : "exact", bad,
                   128 * N, maxerr);
        }
    return 0;
}
