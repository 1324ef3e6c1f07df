// Tensor-core path of td_gemm_f32 for large batches: tcgen05.mma kind::tf32 on the fp32 operands as they are.
//
// nn.Linear forward y = x W^T + b has both operands K-major (x [M][K], W [N][K]), which is the only form kind::tf32 accepts
// for shared-memory operands (tools/probe_umma_tf32.cu: MN-major fp32 operands give wrong results).  TMA brings boxes of
// 32 fp32 (one 128-byte swizzle row) x 128 rows of x and x BLOCK_N rows of W; one elected thread issues four K = 8 MMAs per
// box; the fp32 accumulator (128 x BLOCK_N) lives in TMEM.  The epilogue -- bias, pre-activation copy, activation,
// residual, embedding-row gather, accumulate: the contract of td_gemm_args -- runs after a transpose through shared memory
// (a thread owns an accumulator row, so direct stores would touch 32 cache lines per warp instruction; see conv_halo.cu):
// a lane then owns four consecutive columns of four rows per pass and every global access is a 16-byte vector.
//
// Opt-in per call (td_gemm_args.allow_tf32; the dense engines set it at batch >= 2048; TD_GEMM_TF32=0 disables it): at the reference batch sizes the FFMA kernels of linear.cu are
// latency-bound anyway and keep the fp32 parity tolerances; at 4096-65536 samples per step they reach 5-7 TFLOP/s.
// TF32 keeps 10 mantissa bits of each operand: results agree with fp32 to ~1e-3 (the tolerance of the bf16 conv engine is 1e-2).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (warp 2 owns the TMEM allocation).  Two CTAs per SM.
#include <stdlib.h>

#include "conv_plan.h"
#include "sm100.cuh"

namespace td {

using namespace sm100;

__device__ inline float act_apply_tc(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return fmaxf(v, 0.f);
        case TD_ACT_SILU: return v / (1.0f + expf(-v));
        case TD_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        case TD_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (kind::tf32: 2)  [17,23) N >> 3  [24,29) M >> 4; both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int GT_THREADS = 192;
constexpr int GT_STAGES = 3;
constexpr int GT_A_STAGE = 128 * 128;        // 128 rows x 32 fp32

template <int BLOCK_N>
__global__ void __launch_bounds__(GT_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const td_gemm_args g,
                 const int splits, const int ipz) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_STAGE = BLOCK_N * 128;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + GT_STAGES * GT_A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + GT_STAGES * B_STAGE);
    uint64_t* empty_bar = full_bar + GT_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GT_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * BLOCK_N;
    const int iters_all = (g.K + 31) / 32;                // TMA zero-fills the K tail
    // split-K (blockIdx.z): this CTA reduces K steps [it0, it1) and stores the RAW accumulators to splitk_ws[z][M][N]; the second
    // pass (gemm_splitk_reduce_kernel) sums the slices in fixed order and applies the epilogue.  The batch-reducing weight
    // gradients of the dense denoisers are 2..64 output tiles over 2048 K steps: without the split 16-64 CTAs do all the work.
    const int it0 = splits > 1 ? (int)blockIdx.z * ipz : 0;
    const int it1 = splits > 1 ? min(iters_all, it0 + ipz) : iters_all;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<BLOCK_N>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    td::pdl_sync();              // everything above is on-chip setup; global memory is touched only below

    if (warp == 0) {
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(GT_A_STAGE + B_STAGE));
                tma_load_2d(smem_a + (size_t)s * GT_A_STAGE, &tmap_a, &full_bar[s], it * 32, m0);
                tma_load_2d(smem_b + (size_t)s * B_STAGE, &tmap_b, &full_bar[s], it * 32, n0);
                if (++s == GT_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_tf32(128, BLOCK_N);
            int s = 0;
            uint32_t ph = 0;
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + (size_t)s * GT_A_STAGE), 16, 1024);
                const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + (size_t)s * B_STAGE), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > it0 || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[s]);
                if (++s == GT_STAGES) { s = 0; ph ^= 1u; }
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;                              // TMEM lane quadrant of this warp
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        // the operand ring is idle now: its first bytes stage the 32 x 32 transposes (one tile per epilogue warp)
        const uint32_t tile_s = smem_u32(smem_a) + (uint32_t)q * (32 * 36 * 4);
        const int col = (lane & 7) * 4;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            if (n0 + c0 >= g.N) break;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4) sts128(tile_s + (uint32_t)(lane * 36 + i) * 4u, r[i], r[i + 1], r[i + 2], r[i + 3]);
            __syncwarp();
            const int gj = n0 + c0 + col;
            if (splits > 1) {
                if (gj < g.N) {
                    float* wsz = g.splitk_ws + (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rw = it * 4 + (lane >> 3);
                        const int gi = m0 + q * 32 + rw;
                        if (gi < g.M) *reinterpret_cast<float4*>(wsz + (int64_t)gi * g.N + gj) = lds128(tile_s + (uint32_t)(rw * 36 + col) * 4u);
                    }
                }
                __syncwarp();
                continue;
            }
            if (gj < g.N) {                                  // N is a multiple of 4 (checked on the host)
                float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) bias = __ldg(reinterpret_cast<const float4*>(g.bias + gj));
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rw = it * 4 + (lane >> 3);
                    const int gi = m0 + q * 32 + rw;
                    if (gi >= g.M) continue;
                    float4 v = lds128(tile_s + (uint32_t)(rw * 36 + col) * 4u);
                    v.x = v.x * g.alpha + bias.x; v.y = v.y * g.alpha + bias.y; v.z = v.z * g.alpha + bias.z; v.w = v.w * g.alpha + bias.w;
                    if (g.pre_out) *reinterpret_cast<float4*>(g.pre_out + (int64_t)gi * g.ld_pre + gj) = v;
                    v.x = act_apply_tc(v.x, g.act); v.y = act_apply_tc(v.y, g.act); v.z = act_apply_tc(v.z, g.act); v.w = act_apply_tc(v.w, g.act);
                    if (g.residual) {
                        const float4 t = *reinterpret_cast<const float4*>(g.residual + (int64_t)gi * g.ldr + gj);
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    if (g.gather_idx && g.gather_table) {
                        const float4 t = *reinterpret_cast<const float4*>(g.gather_table + g.gather_idx[gi] * g.ld_table + gj);
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    float4* c = reinterpret_cast<float4*>(g.C + (int64_t)gi * g.ldc + gj);
                    if (g.accumulate) {
                        const float4 t = *c;
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    *c = v;
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<BLOCK_N>(tmem_base);
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// true: the tensor-core path ran (or failed: *status != TD_OK); false: not eligible, the caller falls back to the FFMA kernels
// K slices of the tcgen05 GEMM: only when the output grid leaves most SMs idle and K is long (the weight gradients at large batch)
int gemm_tf32_splits(int M, int N, int K) {
    const int bn = N > 64 ? 128 : 64;
    const int64_t tiles = ceil_div(M, 128) * ceil_div(N, bn);
    const int iters = (K + 31) / 32;
    if (2 * tiles > kNumSMs || K < 8192) return 1;
    int splits = (int)std::min<int64_t>(std::min<int64_t>(kNumSMs / tiles, iters / 32), 64);
    if (splits < 2) return 1;
    const int ipz = (iters + splits - 1) / splits;
    return (iters + ipz - 1) / ipz;
}

bool gemm_tf32_eligible(const td_gemm_args* a) {
    static const bool on = []() { const char* e = getenv("TD_GEMM_TF32"); return !(e && atoi(e) == 0); }();
    if (!on || !a->allow_tf32) return false;            // opt-in per call: the dense engines set it at batch >= 2048
    if (a->splitk_ws && !aligned16(a->splitk_ws)) return false;
    if (a->M < 32) return false;
    if (a->a_cs != 1 || a->b_rs != 1) return false;                                  // both operands K-major
    if (a->a_rs % 4 || a->b_cs % 4 || a->N % 4 || a->ldc % 4) return false;          // 16-byte rows for TMA and the epilogue
    if (!aligned16(a->A) || !aligned16(a->B) || !aligned16(a->C)) return false;
    if (a->bias && !aligned16(a->bias)) return false;
    if (a->pre_out && (a->ld_pre % 4 || !aligned16(a->pre_out))) return false;
    if (a->residual && (a->ldr % 4 || !aligned16(a->residual))) return false;
    if (a->gather_table && (a->ld_table % 4 || !aligned16(a->gather_table))) return false;
    if (a->K < 32 || a->N < 32) return false;
    return tc_get_encode_fn() != nullptr;
}

bool gemm_tf32_try(const td_gemm_args* a, cudaStream_t s, int* status, int* nz_out) {
    *status = TD_OK;
    *nz_out = 1;
    if (!gemm_tf32_eligible(a)) return false;
    const int splits = a->splitk_ws ? gemm_tf32_splits(a->M, a->N, a->K) : 1;
    const int ipz = (((a->K + 31) / 32) + splits - 1) / splits;
    *nz_out = splits;
    EncodeTiledFn encode = tc_get_encode_fn();
    const int bn = a->N > 64 ? 128 : 64;
    CUtensorMap ma, mb;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M};
        cuuint64_t gstr[1] = {(cuuint64_t)a->a_rs * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
        cuuint64_t gstr[1] = {(cuuint64_t)a->b_cs * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)bn};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a->B), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    const int smem = GT_STAGES * (GT_A_STAGE + bn * 128) + (2 * GT_STAGES + 1) * 8 + 16 + 1024;
    const dim3 grid((unsigned)ceil_div(a->N, bn), (unsigned)ceil_div(a->M, 128), (unsigned)splits);
    if (bn == 128) {
        static bool cfg = false;
        if (!cfg) { cudaFuncSetAttribute(gemm_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cfg = true; }
        td::launch(gemm_tf32_kernel<128>, td::LaunchCfg(grid, GT_THREADS, smem, s), ma, mb, *a, splits, ipz);
    } else {
        static bool cfg = false;
        if (!cfg) { cudaFuncSetAttribute(gemm_tf32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cfg = true; }
        td::launch(gemm_tf32_kernel<64>, td::LaunchCfg(grid, GT_THREADS, smem, s), ma, mb, *a, splits, ipz);
    }
    *status = launch_status("gemm_tf32");
    return true;
}

}  // namespace td

extern "C" int td_gemm_f32_path(const td_gemm_args* a) { return (a && td::gemm_tf32_eligible(a)) ? 1 : 0; }
