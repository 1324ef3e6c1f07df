// Hardware probe (B200), companion of probe_umma_rowoff.cu.  Two more descriptor forms the halo kernels want:
//  (1) K-major SWIZZLE_128B A operand whose 8-row groups are NOT 1024 bytes apart: stride byte offset = pitch * 128 with
//      pitch = 10 rows, start at an arbitrary row.  (M rows = 16 image rows x 8 pixels read out of a 10-pixel-pitch box.)
//  (2) MN-major SWIZZLE_128B A operand (rows = K = pixels, 64-channel boxes, LBO = box bytes, SBO = 1024) starting at
//      an arbitrary row: the weight-gradient kernel reading tap-shifted pixels out of one halo box.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc \
//          tools/probe_umma_layouts.cu -o tools/build/probe_umma_layouts -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "sm100.cuh"

using namespace td::sm100;

constexpr int ROWS = 192, N = 64;

// mode 1: K-major, SBO = pitch*128.   mode 2: MN-major A (128 channels = 2 boxes) and MN-major B (64 channels), K = 64 rows.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out, int rowoff,
             int pitch, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int a_boxes = mode == 2 ? 2 : 1;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a_boxes * ROWS * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_b + 64 * 128);
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
        mbar_arrive_expect_tx(bar, a_boxes * ROWS * 128 + 64 * 128);
        for (int g = 0; g < a_boxes; ++g) tma_load_2d(smem_a + g * ROWS * 128, &tmap_a, bar, g * 64, 0);
        tma_load_2d(smem_b, &tmap_b, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a) + rowoff * 128;
        const uint32_t b_addr = smem_u32(smem_b);
        if (mode == 1) {
            constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base, make_smem_desc_sw128(a_addr + k * 32, 16, pitch * 128), make_smem_desc_sw128(b_addr + k * 32, 16, 1024),
                          idesc, k > 0 ? 1u : 0u);
        } else {
            constexpr uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
            for (int k = 0; k < 4; ++k)      // 16 K rows (pixels) per MMA
                umma_bf16(tmem_base, make_smem_desc_sw128(a_addr + k * 2048, ROWS * 128, 1024), make_smem_desc_sw128(b_addr + k * 2048, 64 * 128, 1024),
                          idesc, k > 0 ? 1u : 0u);
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

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, int cols, int rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
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
    srand(2);
    // ---- mode 1 data: A [ROWS][64] K-major, B [64 n][64 k]
    std::vector<float> fa(ROWS * 64), fb(64 * 64), fa2(ROWS * 128), fb2(64 * 64);
    std::vector<__nv_bfloat16> ha(fa.size()), hb(fb.size()), ha2(fa2.size()), hb2(fb2.size());
    for (size_t i = 0; i < fa.size(); ++i) { fa[i] = (float)(rand() % 17 - 8); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < fb.size(); ++i) { fb[i] = (float)(rand() % 9 - 4); hb[i] = __float2bfloat16(fb[i]); }
    for (size_t i = 0; i < fa2.size(); ++i) { fa2[i] = (float)(rand() % 17 - 8); ha2[i] = __float2bfloat16(fa2[i]); }
    for (size_t i = 0; i < fb2.size(); ++i) { fb2[i] = (float)(rand() % 9 - 4); hb2[i] = __float2bfloat16(fb2[i]); }
    __nv_bfloat16 *da, *db, *da2, *db2;
    float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&da2, ha2.size() * 2); cudaMalloc(&db2, hb2.size() * 2);
    cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(da2, ha2.data(), ha2.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db2, hb2.data(), hb2.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma = make_map(enc, da, 64, ROWS, ROWS), mb = make_map(enc, db, 64, 64, 64);
    CUtensorMap ma2 = make_map(enc, da2, 128, ROWS, ROWS), mb2 = make_map(enc, db2, 64, 64, 64);
    const int smem = 2 * ROWS * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hout(128 * N);
    const int offs[] = {0, 1, 3, 8, 11, 21};
    for (int pitch : {8, 9, 10})
        for (int off : offs) {
            if (off + 15 * pitch + 8 > ROWS) continue;
            cudaMemset(dout, 0, 128 * N * 4);
            probe_kernel<<<1, 128, smem>>>(ma, mb, dout, off, pitch, 1);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode 1 pitch %d off %d: CUDA error\n", pitch, off); return 1; }
            cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    float ref = 0;
                    const int row = off + (m / 8) * pitch + m % 8;
                    for (int k = 0; k < 64; ++k) ref += fa[row * 64 + k] * fb[n * 64 + k];
                    if (fabs(ref - hout[m * N + n]) > 1e-3) ++bad;
                }
            printf("K-major   SBO = %2d rows, start row %2d: %s (%d mismatches)\n", pitch, off, bad ? "WRONG" : "exact", bad);
        }
    for (int off : {0, 1, 3, 8, 11, 34, 127}) {
        cudaMemset(dout, 0, 128 * N * 4);
        probe_kernel<<<1, 128, smem>>>(ma2, mb2, dout, off, 8, 2);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode 2 off %d: CUDA error\n", off); return 1; }
        cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += fa2[(off + k) * 128 + m] * fb2[k * 64 + n];
                if (fabs(ref - hout[m * N + n]) > 1e-3) ++bad;
            }
        printf("MN-major  A start row %3d (K = pixels): %s (%d mismatches)\n", off, bad ? "WRONG" : "exact", bad);
    }
    return 0;
}
