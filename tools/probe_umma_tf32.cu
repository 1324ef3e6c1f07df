// Hardware probe (B200): tcgen05.mma kind::tf32 on fp32 operands straight out of TMA-written shared memory -- the building
// block of a tensor-core GEMM for the dense denoisers (latent MLP / DiT), whose activations and weights are fp32
// (DESIGN.md 4.7).  Questions: (1) instruction-descriptor format code for tf32 (A/B format field = 2), K = 8 per instruction,
// K-major SWIZZLE_128B operands with 32 fp32 per 128-byte row; (2) are MN-major fp32 operands (the two backward products of
// a Linear layer reduce over the batch dimension) accepted for kind::tf32?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiny-diffusion_b200/csrc \
//          tools/probe_umma_tf32.cu -o tools/build/probe_umma_tf32 -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "sm100.cuh"

using namespace td::sm100;

constexpr int M = 128, N = 64, K = 64;       // K = 64 fp32 = two 128-byte swizzle rows per operand row

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (kind::tf32: 2 = tf32)  [15] A major  [16] B major  N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// mode 1: A [M][K] and B [N][K] K-major: boxes of 32 fp32 (128 B) x rows, two boxes along K per operand.
// mode 2: A [K][M] and B [K][N] MN-major: boxes of 32 fp32 along M / N x K rows.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;                       // 32 KB
    uint8_t* smem_b = smem + 32768;               // 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_b + 16384);
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
        mbar_arrive_expect_tx(bar, M * K * 4 + N * K * 4);
        if (mode == 1) {
            for (int g = 0; g < 2; ++g) tma_load_2d(smem_a + g * M * 128, &tmap_a, bar, g * 32, 0);     // box (32 k, 128 m)
            for (int g = 0; g < 2; ++g) tma_load_2d(smem_b + g * N * 128, &tmap_b, bar, g * 32, 0);     // box (32 k, 64 n)
        } else {
            for (int g = 0; g < 4; ++g) tma_load_2d(smem_a + g * K * 128, &tmap_a, bar, g * 32, 0);     // box (32 m, 64 k)
            for (int g = 0; g < 2; ++g) tma_load_2d(smem_b + g * K * 128, &tmap_b, bar, g * 32, 0);     // box (32 n, 64 k)
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
        if (mode == 1) {
            constexpr uint32_t idesc = make_idesc_tf32(M, N, 0, 0);
            for (int k = 0; k < K / 8; ++k) {                     // 8 fp32 = 32 bytes per instruction, 4 per swizzle row
                const uint32_t ao = (k / 4) * M * 128 + (k % 4) * 32, bo = (k / 4) * N * 128 + (k % 4) * 32;
                umma_tf32(tmem_base, make_smem_desc_sw128(a_addr + ao, 16, 1024), make_smem_desc_sw128(b_addr + bo, 16, 1024), idesc,
                          k > 0 ? 1u : 0u);
            }
        } else {
            constexpr uint32_t idesc = make_idesc_tf32(M, N, 1, 1);
            for (int k = 0; k < K / 8; ++k)                       // 8 K rows = one 1024-byte group per instruction
                umma_tf32(tmem_base, make_smem_desc_sw128(a_addr + k * 1024, K * 128, 1024), make_smem_desc_sw128(b_addr + k * 1024, K * 128, 1024),
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
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    srand(4);
    // values with at most 10 significant mantissa bits are exact in tf32: small integers / 8
    std::vector<float> a(M * K), b(N * K), at(K * M), bt(K * N), ar(M * K), br(N * K);
    for (auto& v : a) v = (float)(rand() % 33 - 16) / 8.f;
    for (auto& v : b) v = (float)(rand() % 17 - 8) / 8.f;
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) at[k * M + m] = a[m * K + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) bt[k * N + n] = b[n * K + k];
    // a second data set with full fp32 mantissas: the result must match a tf32-rounded reference to ~1e-3
    for (auto& v : ar) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : br) v = (float)rand() / RAND_MAX - 0.5f;
    float *da, *db, *dat, *dbt, *dar, *dbr, *dout;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dat, at.size() * 4); cudaMalloc(&dbt, bt.size() * 4);
    cudaMalloc(&dar, ar.size() * 4); cudaMalloc(&dbr, br.size() * 4); cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dat, at.data(), at.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dbt, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dar, ar.data(), ar.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dbr, br.data(), br.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 32768 + 16384 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(M * N);
    auto run = [&](const char* what, CUtensorMap ma, CUtensorMap mb, int mode, const std::vector<float>& fa, const std::vector<float>& fb,
                   double tol) {
        cudaMemset(dout, 0, M * N * 4);
        probe_kernel<<<1, 128, smem>>>(ma, mb, dout, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", what, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        double worst = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)fa[m * K + k] * fb[n * K + k];
                const double d = fabs(ref - out[m * N + n]);
                worst = fmax(worst, d);
                if (d > tol) ++bad;
            }
        printf("%-44s %s (%d mismatches, max abs err %.3g)\n", what, bad ? "WRONG" : "ok", bad, worst);
    };
    run("tf32 K-major, tf32-exact data", make_map(enc, da, K, M, M), make_map(enc, db, K, N, N), 1, a, b, 1e-4);
    run("tf32 K-major, full-mantissa data (tol 2e-2)", make_map(enc, dar, K, M, M), make_map(enc, dbr, K, N, N), 1, ar, br, 2e-2);
    run("tf32 MN-major A and B, tf32-exact data", make_map(enc, dat, M, K, K), make_map(enc, dbt, N, K, K), 2, a, b, 1e-4);
    return 0;
}
