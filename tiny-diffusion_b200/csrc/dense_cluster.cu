// Eval-mode forward of the dense denoisers (latent MLP latent_diffusion.py:107-128, length-1-sequence DiT
// diffusion_transformer.py:81-109) at the reference batch: ONE kernel, thread-block clusters, no grid barrier.
//
// In eval mode every op of these models is independent per sample (Linear, LayerNorm, eval BatchNorm1d = per-feature affine,
// residual adds, embedding gathers).  So the batch is cut into slices of 8 rows and a cluster of 8 CTAs takes one slice
// through the WHOLE tape:
//   * the slice's activations live in shared memory, replicated in every CTA of the cluster (an "arena" whose buffer offsets
//     the host assigns by liveness); a Linear's output tile is pushed into all eight arenas with st.async (distributed shared
//     memory), every 16-byte piece completing bytes on an mbarrier of the RECEIVING CTA: a CTA goes on as soon as the R x N x 4
//     bytes of the op have landed in its own arena.  No cluster barrier between ops (barrier.cluster compiles to MEMBAR.ALL.GPU
//     + UCGABAR + CCTL.IVALL and drains the weight copies in flight; the grid-barrier tape of dense_fused.cu paid four
//     dependent L2 round trips per op: fence, arrival, observation, operand load);
//   * CTA c of the cluster owns the features [c*fcp, (c+1)*fcp) of every Linear; its weight slice streams from the L2
//     through a ring of stages in chunks of fcp rows x kc columns: TMA boxes of 32 floats x fcp rows (SWIZZLE_128B: eight
//     consecutive rows read the same k without a bank conflict, no padding), mbarrier completion.  (One cp.async per 16 bytes
//     and one bulk copy per row were both measured: 10-17 % of the kernel's issue slots went into requesting weights.)  The
//     stream does not depend on activations, so it runs ahead ACROSS ops: the L2 -> shared-memory pipe never drains between
//     layers;
//   * row-wise ops (LayerNorm, add, time features, input load) are computed redundantly by every CTA on its own copy: no
//     communication at all.
// FFMA tile: a thread owns 4 features x 8 rows over a slice of the chunk's k range (one 16-byte weight read feeds 32 FMAs,
// one broadcast activation read 16); the k-slices are summed in fixed order through shared memory, then bias / folded
// BatchNorm1d / activation / residual / gathered embedding row in the epilogue, as in the tape kernel.
#include <algorithm>
#include <stdlib.h>

#include "conv_plan.h"
#include "sm100.cuh"

namespace td {

struct ClusterOp {               // mirrored by tinydiff/dense.py (ctypes)
    int kind;                    // 0 linear, 1 layernorm, 2 add, 3 time features, 4 load (global -> arena)
    int N, K;
    int act, accumulate, tmode, bn_relu;
    int out_global;              // linear: the result goes to `gout` (global memory), not to the arena
    int x_off, ldx;              // arena offsets / row strides in floats
    int out_off, ldo;
    int res_off, ldr;            // res_off < 0: no residual
    int fcp;                     // linear: features per CTA (power of two, 4..128)
    int kc;                      // linear: k extent of a weight chunk (multiple of 32: TMA boxes of 32 floats x fcp rows); 0 = weights read straight from global memory
    const float* w;              // linear: [N][K] row-major; layernorm: gamma
    const float* bias;           // linear: [N] or NULL; layernorm: beta
    const void* tmap;            // linear with kc > 0: device copy of the weight's tensor map (td_dense_cluster_weight_map), 64-byte aligned
    long long pad_;
    const float* gx; long long gldx;      // load: global source
    float* gout; long long gldo;
    const long long* gidx; const float* gtab; long long ldt;
    const float* bn_mean; const float* bn_var; const float* bn_gamma; const float* bn_beta;
    float bn_eps, ln_eps;
    const long long* t; const int* t_dev;
};
static_assert(sizeof(ClusterOp) % 16 == 0, "ClusterOp is copied as 16-byte vectors");

constexpr int CK_THREADS = 256, CK_CL = 8;
constexpr int CK_RMAX = 9;                                // rows per cluster: 8, or 9 when that saves a pass (128 rows on 15 clusters)
constexpr int CK_STAGES = 3, CK_STAGE_FLOATS = 8704;      // 3 x 34 KB
constexpr int CK_ARENA_FLOATS = 18432;                    // 72 KB
constexpr int CK_RED_FLOATS = 256 * 4 * CK_RMAX + 256;    // [k-slices][rows][fcp] (+ one bank-shift pad per slice) whatever fcp is
constexpr int CK_MAX_OPS = 56;
constexpr int CK_SMEM = 1024 + CK_MAX_OPS * (int)sizeof(ClusterOp) + (CK_ARENA_FLOATS + CK_STAGES * CK_STAGE_FLOATS + CK_RED_FLOATS) * 4;   // 1 KB: alignment of the base
static_assert(CK_SMEM + 64 <= 227 * 1024 && (CK_STAGE_FLOATS * 4) % 1024 == 0, "dense_cluster_kernel: shared memory plan");

__device__ inline float ck_act(float v, int act) {
    switch (act) {
        case TD_ACT_RELU: return fmaxf(v, 0.f);
        case TD_ACT_SILU: return v / (1.f + expf(-v));
        case TD_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case TD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

__device__ inline void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ inline uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ inline uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ inline uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ inline uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// asynchronous stores into a peer's (or the own) shared memory; the bytes complete on the RECEIVER's mbarrier `bar`
__device__ inline void st_async_f32(uint32_t addr, float v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(bar) : "memory");
}
__device__ inline void st_async_v4(uint32_t addr, float4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(bar)
                 : "memory");
}
// wait for a phase of an mbarrier whose bytes are completed by peers of the cluster
__device__ inline void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && ++spins > TD_SPIN_LIMIT) { printf("tinydiff: dense_cluster activation barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    } while (!ok);
}
__device__ __noinline__ float4 ck_act4(float4 v, int act) {
    v.x = ck_act(v.x, act); v.y = ck_act(v.y, act); v.z = ck_act(v.z, act); v.w = ck_act(v.w, act);
    return v;
}

// Position in this CTA's weight stream: (pass over a row slice, op, chunk).  The producer side and the consumer side walk the
// same sequence; ops that do not stream (non-linear, kc == 0, or no feature of the op belongs to this CTA) are skipped.
struct ChunkCursor {
    int pass, oi, ch;
};
__device__ inline bool ck_streams(const ClusterOp& op, int rank) { return op.kind == 0 && op.kc > 0 && rank * op.fcp < op.N; }
__device__ inline bool ck_settle(ChunkCursor& c, const ClusterOp* ops, int n_ops, int n_pass, int rank) {
    while (c.pass < n_pass) {
        while (c.oi < n_ops) {
            const ClusterOp& op = ops[c.oi];
            if (ck_streams(op, rank) && c.ch * op.kc < op.K) return true;
            ++c.oi;
            c.ch = 0;
        }
        ++c.pass;
        c.oi = 0;
        c.ch = 0;
    }
    return false;
}

// request chunk `c` of this CTA's weight stream into `stage`: boxes of 32 floats x fcp rows (box b at stage + b * fcp * 32 floats),
// warp b issues box b.  Rows past N and columns past K are zero-filled by TMA.
__device__ inline void ck_issue(const ChunkCursor& c, const ClusterOp* ops, int rank, float* stage, uint64_t* bar) {
    const ClusterOp& op = ops[c.oi];
    const int k0 = c.ch * op.kc;
    const int nb = (min(op.kc, op.K - k0) + 31) >> 5;
    if (threadIdx.x == 0) sm100::mbar_arrive_expect_tx(bar, (uint32_t)(nb * op.fcp) * 128u);
    if ((threadIdx.x & 31) == 0)
        for (int b = (int)(threadIdx.x >> 5); b < nb; b += CK_THREADS / 32)
            sm100::tma_load_2d(stage + b * op.fcp * 32, reinterpret_cast<const CUtensorMap*>(op.tmap), bar, k0 + 32 * b, rank * op.fcp);
}

// tuning aid (TD_DENSE_CLUSTER_DBG=1): per CTA {globaltimer at start, at end, clock64 cycles in: weight-chunk waits, FFMA
// chunks, k-slice reduction + epilogue + pushes, cluster barriers, row-wise ops}
__device__ unsigned long long g_ck_dbg[kNumSMs * 16];
__device__ inline unsigned long long ck_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define CK_T(var) do { if (dbg) { const long long _n = clock64(); var += _n - t_last; t_last = _n; } } while (0)

template <int R>
__global__ void __launch_bounds__(CK_THREADS, 1)
dense_cluster_kernel(const ClusterOp* __restrict__ g_ops, int n_ops, int M, int dbg) {
    extern __shared__ __align__(16) unsigned char ck_smem[];
    __shared__ __align__(8) uint64_t s_full[CK_STAGES];      // weight stage landed (bulk-copy bytes)
    __shared__ __align__(8) uint64_t s_act[2];               // a Linear's output landed in THIS CTA's arena (peers' st.async bytes)
    unsigned char* const base = ck_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(ck_smem) & 1023u)) & 1023u);   // swizzle = f(address)
    float* const ring = reinterpret_cast<float*>(base);             // stages are 1024-byte aligned (34 KB each)
    float* const arena = ring + CK_STAGES * CK_STAGE_FLOATS;
    float* const red = arena + CK_ARENA_FLOATS;
    ClusterOp* const ops = reinterpret_cast<ClusterOp*>(red + CK_RED_FLOATS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank();
    const int n_clusters = (int)cluster_count_x(), cid = (int)cluster_id_x();
    const int n_slices = (M + R - 1) / R;
    const int n_pass = cid < n_slices ? (n_slices - cid + n_clusters - 1) / n_clusters : 0;

    {   // the tape is host-written and constant: copied before the dependency wait
        const int4* src = reinterpret_cast<const int4*>(g_ops);
        int4* dst = reinterpret_cast<int4*>(ops);
        for (int i = tid; i < n_ops * (int)(sizeof(ClusterOp) / 16); i += CK_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int i = 0; i < CK_STAGES; ++i) sm100::mbar_init(&s_full[i], 1);
        sm100::mbar_init(&s_act[0], 1);
        sm100::mbar_init(&s_act[1], 1);
        sm100::fence_barrier_init();
    }
    td::pdl_sync();
    __syncthreads();
    const unsigned long long g_start = dbg ? ck_gtime() : 0ull;
    long long t_last = dbg ? clock64() : 0, c_wait = 0, c_fma = 0, c_epi = 0, c_bar = 0, c_row = 0, c_e1 = 0, c_e2 = 0, c_pre = 0;

    // weight stream prologue: CK_STAGES - 1 chunks in flight
    ChunkCursor pc{0, 0, 0};
    uint32_t p_stage = 0, c_stage = 0, c_phase = 0;
#pragma unroll 1
    for (int s = 0; s < CK_STAGES - 1; ++s) {
        if (ck_settle(pc, ops, n_ops, n_pass, rank)) {
            ck_issue(pc, ops, rank, ring + p_stage * CK_STAGE_FLOATS, &s_full[p_stage]);
            ++pc.ch;
        }
        p_stage = p_stage + 1 == CK_STAGES ? 0 : p_stage + 1;
    }
    // the peers' mbarriers are initialised and their shared memory may be written
    cluster_sync_all();
    // gamma / beta of the NEXT LayerNorm (<= 256 features) are requested while the Linear in front of it runs
    constexpr int LN_CACHE = 8;
    float ln_g[LN_CACHE], ln_b[LN_CACHE];
    int ln_for = -1;                                       // op index ln_g / ln_b were requested for
#pragma unroll
    for (int i = 0; i < LN_CACHE; ++i) { ln_g[i] = 0.f; ln_b[i] = 0.f; }
    uint32_t act_n = 0;                                    // Linears pushed so far: barrier s_act[act_n & 1], parity (act_n >> 1) & 1
    const uint32_t act0 = (uint32_t)__cvta_generic_to_shared(&s_act[0]);

#pragma unroll 1
    for (int pass = 0; pass < n_pass; ++pass) {
        const int row0 = (cid + pass * n_clusters) * R;
        if (pass > 0) cluster_sync_all();                  // every CTA of the cluster has left the previous slice
        CK_T(c_bar);
#pragma unroll 1
        for (int oi = 0; oi < n_ops; ++oi) {
            const ClusterOp& op = ops[oi];
            if (op.kind == 0) {
                const int fcp = op.fcp, nfg = fcp >> 2, lg_nfg = 31 - __clz(nfg), nks = CK_THREADS >> lg_nfg;      // fcp is a power of two
                const int kstride = R * fcp + (nfg >= 4 ? nfg : 0);    // floats between the k-slices of `red` (bank shift: conflict-free stores)
                const int fbase = rank * fcp;
                const bool mine = fbase < op.N;
                // epilogue items of this thread: rows er0 and er0 + 256 / nfg, features fbase + 4*eg .. + 3 (256 % nfg == 0: the same
                // features for both).  Their per-feature operands are REQUESTED now (raw values, no dependent use: the loads fly
                // beside the weight chunks), combined after the reduction.
                const int er0 = tid >> lg_nfg, eg = tid & (nfg - 1), er_step = nks;
                const int en = fbase + 4 * eg;
                float e_bias[4] = {0.f, 0.f, 0.f, 0.f};
                float e_var[4] = {1.f, 1.f, 1.f, 1.f}, e_mean[4] = {0.f, 0.f, 0.f, 0.f}, e_gamma[4] = {1.f, 1.f, 1.f, 1.f}, e_beta[4] = {0.f, 0.f, 0.f, 0.f};
                long long e_gi0 = 0;
                if (mine && er0 < R) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = min(en + j, op.N - 1);
                        if (op.bias) e_bias[j] = __ldg(op.bias + n);
                        if (op.bn_mean) {
                            e_var[j] = __ldg(op.bn_var + n);
                            e_mean[j] = __ldg(op.bn_mean + n);
                            e_gamma[j] = __ldg(op.bn_gamma + n);
                            e_beta[j] = __ldg(op.bn_beta + n);
                        }
                    }
                    if (op.gidx) e_gi0 = __ldg(op.gidx + min(row0 + er0, M - 1));
                }
                if (oi + 1 < n_ops && ops[oi + 1].kind == 1 && ops[oi + 1].N <= 32 * LN_CACHE) {
                    const ClusterOp& ln = ops[oi + 1];
#pragma unroll
                    for (int i = 0; i < LN_CACHE; ++i) {
                        const int n = min(lane + 32 * i, ln.N - 1);
                        ln_g[i] = __ldg(ln.w + n);
                        ln_b[i] = __ldg(ln.bias + n);
                    }
                    ln_for = oi + 1;
                }
                CK_T(c_pre);
                if (mine) {
                    if (op.kc > 0) {
                        const int fg = eg, ks = er0;
                        float acc[4][R];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
                        const float* xbase = arena + op.x_off;
                        const int ldx = op.ldx, K = op.K, kc = op.kc;
#pragma unroll 1
                        for (int k0 = 0; k0 < K; k0 += kc) {
                            sm100::mbar_wait(&s_full[c_stage], c_phase);
                            __syncthreads();                 // every thread is done with the stage refilled below
                            CK_T(c_wait);
                            if (ck_settle(pc, ops, n_ops, n_pass, rank)) {
                                ck_issue(pc, ops, rank, ring + p_stage * CK_STAGE_FLOATS, &s_full[p_stage]);
                                ++pc.ch;
                            }
                            p_stage = p_stage + 1 == CK_STAGES ? 0 : p_stage + 1;
                            const float* wst = ring + c_stage * CK_STAGE_FLOATS;
                            const int kq = min(kc, K - k0) >> 2;
                            const float* xk = xbase + k0;
                            // software pipeline: the weights of the NEXT k-piece and the next activation row are in flight during the
                            // 16 FMAs of the current row (all warps run this loop in step: with load-then-compute phases the FMA pipe
                            // idled while the shared-memory pipe served 13 loads per thread, and vice versa)
                            auto w_at = [&](int q, int j) -> float4 {   // 16-byte piece c of row rr sits at piece c ^ (rr & 7) of its 128-byte line
                                const int rr = (q >> 3) * fcp + fg + nfg * j;
                                return *reinterpret_cast<const float4*>(wst + rr * 32 + (((q & 7) ^ (rr & 7)) << 2));
                            };
                            int q = ks;
                            if (q < kq) {
                                float4 wv[4], xc;
#pragma unroll
                                for (int j = 0; j < 4; ++j) wv[j] = w_at(q, j);
                                xc = *reinterpret_cast<const float4*>(xk + 4 * q);
#pragma unroll 1
                                while (true) {
                                    const int qn = q + nks;
                                    const bool more = qn < kq;
                                    const int ql = more ? qn : q;                 // (a finished thread re-reads its last piece: no branch)
                                    float4 wn[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) wn[j] = w_at(ql, j);
#pragma unroll
                                    for (int r = 0; r < R; ++r) {
                                        const float4 xn = r + 1 < R ? *reinterpret_cast<const float4*>(xk + (r + 1) * ldx + 4 * q)
                                                                    : *reinterpret_cast<const float4*>(xk + 4 * ql);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xc.x, wv[j].x, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xc.y, wv[j].y, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xc.z, wv[j].z, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xc.w, wv[j].w, acc[j][r]);
                                        xc = xn;
                                    }
                                    if (!more) break;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) wv[j] = wn[j];
                                    q = qn;
                                }
                            }
                            if (++c_stage == CK_STAGES) { c_stage = 0; c_phase ^= 1u; }
                            CK_T(c_fma);
                        }
                        // partial sums of the k-slices: red[ks][row][feature]
                        float* rk = red + ks * kstride + fg;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int r = 0; r < R; ++r) rk[r * fcp + nfg * j] = acc[j][r];
                    } else {
                        // weights straight from global memory (K not a multiple of 4, or an unaligned slice): thread = (row, feature)
                        for (int e = tid; e < R * fcp; e += CK_THREADS) {
                            const int r = e / fcp, fr = e - r * fcp;
                            const int n = min(fbase + fr, op.N - 1);
                            const float* wr = op.w + (long long)n * op.K;
                            const float* xr = arena + op.x_off + r * op.ldx;
                            float v = 0.f;
                            for (int k = 0; k < op.K; ++k) v = fmaf(xr[k], __ldg(wr + k), v);
                            red[r * fcp + fr] = v;
                        }
                    }
                }
                uint64_t* const my_act = &s_act[act_n & 1u];
                const uint32_t act_parity = (act_n >> 1) & 1u;
                // bytes the PEERS will complete here; the CTA's own tile is a plain store ordered by the CTA barriers below
                if (!op.out_global && tid == 0)
                    sm100::mbar_arrive_expect_tx(my_act, (uint32_t)(R * (op.N - (mine ? min(fcp, op.N - fbase) : 0))) * 4u);
                __syncthreads();
                CK_T(c_e1);
                if (mine) {
                    const int nsl = op.kc > 0 ? nks : 1;
                    const uint32_t bar_local = act0 + 8u * (act_n & 1u);
#pragma unroll 1
                    for (int er = er0; er < R; er += er_step) {
                        const int em = row0 + er;
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float* rp = red + er * fcp + 4 * eg;
#pragma unroll 8
                        for (int q = 0; q < nsl; ++q) {
                            const float4 p = *reinterpret_cast<const float4*>(rp + q * kstride);
                            s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
                        }
                        float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float t = v[j] + e_bias[j];
                            if (op.bn_mean) {
                                t = (t - e_mean[j]) * (1.f / sqrtf(e_var[j] + op.bn_eps)) * e_gamma[j] + e_beta[j];
                                if (op.bn_relu) t = fmaxf(t, 0.f);
                            }
                            v[j] = t;
                        }
                        if (op.act != TD_ACT_NONE) {
                            const float4 a = ck_act4(make_float4(v[0], v[1], v[2], v[3]), op.act);
                            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                        }
                        if (op.res_off >= 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] += arena[op.res_off + er * op.ldr + min(en + j, op.N - 1)];
                        }
                        if (op.gidx) {
                            const long long gi = er == er0 ? e_gi0 : __ldg(op.gidx + min(em, M - 1));
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] += __ldg(op.gtab + gi * op.ldt + min(en + j, op.N - 1));
                        }
                        if (er == er0) CK_T(c_e2);
                        if (op.out_global) {
                            if (em < M)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (en + j < op.N) op.gout[(long long)em * op.gldo + en + j] = v[j];
                        } else {
                            const int o = op.out_off + er * op.ldo + en;
                            const uint32_t local = (uint32_t)__cvta_generic_to_shared(arena + o);
                            const bool vec = en + 3 < op.N && (o & 3) == 0;
                            if (vec) {
                                *reinterpret_cast<float4*>(arena + o) = make_float4(v[0], v[1], v[2], v[3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (en + j < op.N) arena[o + j] = v[j];
                            }
#pragma unroll
                            for (int d = 1; d < CK_CL; ++d) {
                                const uint32_t peer = (uint32_t)((rank + d) & (CK_CL - 1));
                                const uint32_t dst = map_to_rank(local, peer), dbar = map_to_rank(bar_local, peer);
                                if (vec) {
                                    st_async_v4(dst, make_float4(v[0], v[1], v[2], v[3]), dbar);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (en + j < op.N) st_async_f32(dst + 4u * j, v[j], dbar);
                                }
                            }
                        }
                    }
                }
                CK_T(c_epi);
                if (!op.out_global) {
                    sm100::mbar_wait(my_act, act_parity);          // R x N x 4 bytes have landed in this CTA's arena (complete_tx by the peers' st.async)
                    ++act_n;
                    // a CTA that owns no feature of this op is not held back by its own pushes: realign the cluster (rare, tiny layers)
                    if ((CK_CL - 1) * fcp >= op.N) cluster_sync_all();
                }
                __syncthreads();                                  // `red` and the row-wise scratch are free for the next op
                CK_T(c_bar);
                continue;
            }
            if (op.kind == 1) {
                // LayerNorm on the CTA's own copy: a warp per row, the row in registers, two-pass (mean, then centred variance) like
                // ATen with the lane-strided summation order of the other kernels
                if (op.N <= 32 * LN_CACHE) {
                    if (ln_for != oi) {
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i) {
                            const int n = min(lane + 32 * i, op.N - 1);
                            ln_g[i] = __ldg(op.w + n);
                            ln_b[i] = __ldg(op.bias + n);
                        }
                    }
#pragma unroll 1
                    for (int r = warp; r < R; r += CK_THREADS / 32) {
                        const float* xr = arena + op.x_off + r * op.ldx;
                        float* orow = arena + op.out_off + r * op.ldo;
                        float xv[LN_CACHE];
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i) xv[i] = lane + 32 * i < op.N ? xr[lane + 32 * i] : 0.f;
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i) if (lane + 32 * i < op.N) s += xv[i];
                        const float mean = warp_sum(s) / (float)op.N;
                        float q = 0.f;
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i)
                            if (lane + 32 * i < op.N) { const float d = xv[i] - mean; q = fmaf(d, d, q); }
                        const float rstd = 1.f / sqrtf(warp_sum(q) / (float)op.N + op.ln_eps);
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i)
                            if (lane + 32 * i < op.N) orow[lane + 32 * i] = (xv[i] - mean) * rstd * ln_g[i] + ln_b[i];
                    }
                } else {
#pragma unroll 1
                    for (int r = warp; r < R; r += CK_THREADS / 32) {
                        const float* xr = arena + op.x_off + r * op.ldx;
                        float* orow = arena + op.out_off + r * op.ldo;
                        float s = 0.f;
                        for (int n = lane; n < op.N; n += 32) s += xr[n];
                        const float mean = warp_sum(s) / (float)op.N;
                        float q = 0.f;
                        for (int n = lane; n < op.N; n += 32) { const float d = xr[n] - mean; q = fmaf(d, d, q); }
                        const float rstd = 1.f / sqrtf(warp_sum(q) / (float)op.N + op.ln_eps);
                        for (int n = lane; n < op.N; n += 32) orow[n] = (xr[n] - mean) * rstd * __ldg(op.w + n) + __ldg(op.bias + n);
                    }
                }
            } else if (op.kind == 2) {
                for (int e = tid; e < R * op.N; e += CK_THREADS) {
                    const int r = e / op.N, n = e - r * op.N;
                    const float v = arena[op.x_off + r * op.ldx + n];
                    float* o = arena + op.out_off + r * op.ldo + n;
                    *o = op.accumulate ? *o + v : v;
                }
            } else if (op.kind == 3) {
                const int width = op.tmode == 2 ? op.N : 1;
                for (int e = tid; e < R * width; e += CK_THREADS) {
                    const int r = e / width, j = e - r * width;
                    const int m = min(row0 + r, M - 1);
                    float tv = op.t ? (float)op.t[m] : (float)op.t_dev[0];
                    if (op.tmode == 1) tv = tv / 1000.0f;
                    float v = tv;
                    if (op.tmode == 2) {
                        const int half = op.N / 2;
                        v = 0.f;
                        if (j < 2 * half) {
                            const int jj = (j < half) ? j : j - half;
                            const float arg = tv * expf(-logf(10000.0f) * (float)jj / (float)(half - 1));
                            v = (j < half) ? sinf(arg) : cosf(arg);
                        }
                    }
                    arena[op.out_off + r * op.ldo + j] = v;
                }
            } else {
                for (int e = tid; e < R * op.N; e += CK_THREADS) {
                    const int r = e / op.N, n = e - r * op.N;
                    const int m = row0 + r;
                    arena[op.out_off + r * op.ldo + n] = m < M ? __ldcg(op.gx + (long long)m * op.gldx + n) : 0.f;
                }
            }
            __syncthreads();
            CK_T(c_row);
        }
    }
    cluster_sync_all();          // no CTA leaves while a peer may still write into its shared memory
    if (dbg && tid == 0 && blockIdx.x < kNumSMs) {
        unsigned long long* d = g_ck_dbg + blockIdx.x * 16;
        d[0] = g_start; d[1] = ck_gtime();
        d[2] = c_wait; d[3] = c_fma; d[4] = c_epi; d[5] = c_bar; d[6] = c_row; d[7] = c_pre; d[8] = c_e1; d[9] = c_e2;
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Two-group variant (the default): the CTA's eight compute warps are TWO independent groups of four, each taking half of the
// cluster's rows through the tape with its own named barrier, its own activation mbarriers and its own k-slice scratch; a ninth
// warp is the weight producer (TMA boxes into the shared ring, full / empty mbarriers).  A Linear is a chain of dependent
// latencies (weight wait -> FMA tile -> partial stores -> barrier -> k-slice sum -> pushes -> wait for the slowest peer); with one
// chain per SM the FP32 pipe ran at 30 %.  Two chains interleave on the same SM -- and share every weight stage, so the L2 -> SM
// traffic does not grow.
// ------------------------------------------------------------------------------------------------------------------------------
// Reverse-step update folded into the tape's last Linear (the one that writes eps): x <- c1[t] (x - c2[t] eps) + c3[t] z with
// the roundings, the noise-table rows and the Philox indexing of psample_kernel (elementwise.cu), and t_dev[0] = t - 1 by the last
// CTA of the grid to finish.  x == NULL: plain forward.
struct ClusterStep {
    float* x;                    // [M][N] the sampler state (the tape's input buffer)
    const float* coef;           // [T] float4 {c1, c2, c3, -}
    const float* z;              // noise table [T][M * N] or NULL
    long long z_step_stride;
    const unsigned long long* seed;   // {seed, subsequence base} or NULL
    int* t_dev;
    unsigned int* ticket;
    int num_timesteps;
};

constexpr int CK2_GT = 128;                                // threads per group
constexpr int CK2_GMAX = 3;                                // groups per CTA: 2 or 3
constexpr int CK2_RED_FLOATS = 2 * (128 * 4 * 5 + 128);    // all groups: [k-slices][rows of the group][fcp] + bank-shift pad (2 x 5 rows >= 3 x 3 rows)
constexpr int CK2_SMEM = 1024 + CK_MAX_OPS * (int)sizeof(ClusterOp) + (CK_ARENA_FLOATS + CK_STAGES * CK_STAGE_FLOATS + CK2_RED_FLOATS) * 4;

__device__ inline void ck2_issue(const ChunkCursor& c, const ClusterOp* ops, int rank, float* stage, uint64_t* bar) {   // one thread
    const ClusterOp& op = ops[c.oi];
    const int k0 = c.ch * op.kc;
    const int nb = (min(op.kc, op.K - k0) + 31) >> 5;
    sm100::mbar_arrive_expect_tx(bar, (uint32_t)(nb * op.fcp) * 128u);
    for (int b = 0; b < nb; ++b)
        sm100::tma_load_2d(stage + b * op.fcp * 32, reinterpret_cast<const CUtensorMap*>(op.tmap), bar, k0 + 32 * b, rank * op.fcp);
}

template <int R, int G>
__global__ void __launch_bounds__(G * CK2_GT + 32, 1)
dense_cluster2_kernel(const ClusterOp* __restrict__ g_ops, int n_ops, int M, int dbg, const ClusterStep step) {
    constexpr int RG = (R + G - 1) / G;                      // rows of a group's register tile (the last group may own fewer)
    constexpr int CK2_THREADS = G * CK2_GT + 32;             // + producer warp
    constexpr int RED_G = 128 * (4 * RG + 1);                // floats of one group's k-slice scratch
    static_assert(G * RED_G <= CK2_RED_FLOATS && G <= CK2_GMAX, "k-slice scratch");
    extern __shared__ __align__(16) unsigned char ck_smem[];
    __shared__ __align__(8) uint64_t s_full[CK_STAGES];      // weight stage landed (TMA bytes)
    __shared__ __align__(8) uint64_t s_empty[CK_STAGES];     // both groups are done with the stage
    __shared__ __align__(8) uint64_t s_act[CK2_GMAX][2];     // [group][parity]: a Linear's output rows landed in THIS CTA's arena
    unsigned char* const base = ck_smem + ((1024u - ((uint32_t)__cvta_generic_to_shared(ck_smem) & 1023u)) & 1023u);
    float* const ring = reinterpret_cast<float*>(base);
    float* const arena = ring + CK_STAGES * CK_STAGE_FLOATS;
    float* const red_all = arena + CK_ARENA_FLOATS;
    ClusterOp* const ops = reinterpret_cast<ClusterOp*>(red_all + CK2_RED_FLOATS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank();
    const int n_clusters = (int)cluster_count_x(), cid = (int)cluster_id_x();
    const int n_slices = (M + R - 1) / R;
    const int n_pass = cid < n_slices ? (n_slices - cid + n_clusters - 1) / n_clusters : 0;

    {
        const int4* src = reinterpret_cast<const int4*>(g_ops);
        int4* dst = reinterpret_cast<int4*>(ops);
        for (int i = tid; i < n_ops * (int)(sizeof(ClusterOp) / 16); i += CK2_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int i = 0; i < CK_STAGES; ++i) { sm100::mbar_init(&s_full[i], 1); sm100::mbar_init(&s_empty[i], G); }
        for (int i = 0; i < 2 * G; ++i) sm100::mbar_init(&s_act[i >> 1][i & 1], 1);
        sm100::fence_barrier_init();
    }
    td::pdl_sync();
    __syncthreads();
    cluster_sync_all();          // the peers' mbarriers are initialised and their shared memory may be written

    if (warp == G * CK2_GT / 32) {
        // ---- weight producer: walks this CTA's chunk stream, at most CK_STAGES chunks ahead of the slower group
        ChunkCursor pc{0, 0, 0};
        uint32_t stage = 0, phase = 0;
        int cur_pass = 0;
        while (ck_settle(pc, ops, n_ops, n_pass, rank)) {
            while (cur_pass < pc.pass) { cluster_sync_all(); ++cur_pass; }      // the groups meet the cluster at every slice boundary
            if (lane == 0) {
                sm100::mbar_wait(&s_empty[stage], phase ^ 1u);
                ck2_issue(pc, ops, rank, ring + stage * CK_STAGE_FLOATS, &s_full[stage]);
            }
            __syncwarp();
            ++pc.ch;
            if (++stage == CK_STAGES) { stage = 0; phase ^= 1u; }
        }
        while (cur_pass < n_pass - 1) { cluster_sync_all(); ++cur_pass; }
    } else {
        const int g = warp >> 2, gtid = tid & (CK2_GT - 1), gwarp = warp & 3;
        const int rbase = g * RG, rcnt = min(RG, R - g * RG);                 // this group's rows of the slice
        float* const red = red_all + g * RED_G;
        const uint32_t bar_id = 1u + (uint32_t)g;
        auto gbar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
        const unsigned long long g_start = dbg ? ck_gtime() : 0ull;
        long long t_last = dbg ? clock64() : 0, c_wait = 0, c_fma = 0, c_epi = 0, c_bar = 0, c_row = 0, c_e1 = 0, c_e2 = 0, c_pre = 0;
        constexpr int LN_CACHE = 8;
        float ln_g[LN_CACHE], ln_b[LN_CACHE];
        int ln_for = -1;
#pragma unroll
        for (int i = 0; i < LN_CACHE; ++i) { ln_g[i] = 0.f; ln_b[i] = 0.f; }
        uint32_t act_n = 0, c_stage = 0, c_phase = 0;
        const uint32_t act0 = (uint32_t)__cvta_generic_to_shared(&s_act[g][0]);

#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
            const int row0 = (cid + pass * n_clusters) * R + rbase;          // first global row of this group
            if (pass > 0) cluster_sync_all();
            CK_T(c_bar);
#pragma unroll 1
            for (int oi = 0; oi < n_ops; ++oi) {
                const ClusterOp& op = ops[oi];
                if (op.kind == 0) {
                    const int fcp = op.fcp, nfg = fcp >> 2, lg_nfg = 31 - __clz(nfg), nks = CK2_GT >> lg_nfg;
                    const int kstride = RG * fcp + (nfg >= 4 ? nfg : 0);
                    const int fbase = rank * fcp;
                    const bool mine = fbase < op.N;
                    const int er0 = gtid >> lg_nfg, eg = gtid & (nfg - 1), er_step = nks;
                    const int en = fbase + 4 * eg;
                    float e_bias[4] = {0.f, 0.f, 0.f, 0.f};
                    float e_var[4] = {1.f, 1.f, 1.f, 1.f}, e_mean[4] = {0.f, 0.f, 0.f, 0.f}, e_gamma[4] = {1.f, 1.f, 1.f, 1.f}, e_beta[4] = {0.f, 0.f, 0.f, 0.f};
                    long long e_gi0 = 0;
                    if (mine && er0 < rcnt) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int n = min(en + j, op.N - 1);
                            if (op.bias) e_bias[j] = __ldg(op.bias + n);
                            if (op.bn_mean) {
                                e_var[j] = __ldg(op.bn_var + n);
                                e_mean[j] = __ldg(op.bn_mean + n);
                                e_gamma[j] = __ldg(op.bn_gamma + n);
                                e_beta[j] = __ldg(op.bn_beta + n);
                            }
                        }
                        if (op.gidx) e_gi0 = __ldg(op.gidx + min(row0 + er0, M - 1));
                    }
                    if (oi + 1 < n_ops && ops[oi + 1].kind == 1 && ops[oi + 1].N <= 32 * LN_CACHE) {
                        const ClusterOp& ln = ops[oi + 1];
#pragma unroll
                        for (int i = 0; i < LN_CACHE; ++i) {
                            const int n = min(lane + 32 * i, ln.N - 1);
                            ln_g[i] = __ldg(ln.w + n);
                            ln_b[i] = __ldg(ln.bias + n);
                        }
                        ln_for = oi + 1;
                    }
                    CK_T(c_pre);
                    if (mine) {
                        if (op.kc > 0) {
                            const int fg = eg, ks = er0;
                            float acc[4][RG];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int r = 0; r < RG; ++r) acc[j][r] = 0.f;
                            const int ldx = op.ldx, K = op.K, kc = op.kc;
                            // the tile's rows (a short group repeats its last row: the copy is never stored)
                            const float* xrow[RG];
#pragma unroll
                            for (int r = 0; r < RG; ++r) xrow[r] = arena + op.x_off + (rbase + min(r, rcnt - 1)) * ldx;
#pragma unroll 1
                            for (int k0 = 0; k0 < K; k0 += kc) {
                                sm100::mbar_wait(&s_full[c_stage], c_phase);
                                CK_T(c_wait);
                                const float* wst = ring + c_stage * CK_STAGE_FLOATS;
                                const int kq = min(kc, K - k0) >> 2;
                                auto w_at = [&](int q, int j) -> float4 {   // 16-byte piece c of row rr sits at piece c ^ (rr & 7) of its 128-byte line
                                    const int rr = (q >> 3) * fcp + fg + nfg * j;
                                    return *reinterpret_cast<const float4*>(wst + rr * 32 + (((q & 7) ^ (rr & 7)) << 2));
                                };
#pragma unroll 2
                                for (int q = ks; q < kq; q += nks) {
                                    float4 wv[4], xv[RG];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) wv[j] = w_at(q, j);
#pragma unroll
                                    for (int r = 0; r < RG; ++r) xv[r] = *reinterpret_cast<const float4*>(xrow[r] + k0 + 4 * q);
#pragma unroll
                                    for (int r = 0; r < RG; ++r) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xv[r].x, wv[j].x, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xv[r].y, wv[j].y, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xv[r].z, wv[j].z, acc[j][r]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) acc[j][r] = fmaf(xv[r].w, wv[j].w, acc[j][r]);
                                    }
                                }
                                gbar();                                          // the whole group is done with the stage
                                if (gtid == 0) sm100::mbar_arrive(&s_empty[c_stage]);
                                if (++c_stage == CK_STAGES) { c_stage = 0; c_phase ^= 1u; }
                                CK_T(c_fma);
                            }
                            float* rk = red + ks * kstride + fg;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int r = 0; r < RG; ++r) rk[r * fcp + nfg * j] = acc[j][r];
                        } else {
                            for (int e = gtid; e < rcnt * fcp; e += CK2_GT) {
                                const int r = e / fcp, fr = e - r * fcp;
                                const int n = min(fbase + fr, op.N - 1);
                                const float* wr = op.w + (long long)n * op.K;
                                const float* xr = arena + op.x_off + (rbase + r) * op.ldx;
                                float v = 0.f;
                                for (int k = 0; k < op.K; ++k) v = fmaf(xr[k], __ldg(wr + k), v);
                                red[r * fcp + fr] = v;
                            }
                        }
                    }
                    uint64_t* const my_act = &s_act[g][act_n & 1u];
                    const uint32_t act_parity = (act_n >> 1) & 1u;
                    if (!op.out_global && gtid == 0)
                        sm100::mbar_arrive_expect_tx(my_act, (uint32_t)(rcnt * (op.N - (mine ? min(fcp, op.N - fbase) : 0))) * 4u);
                    gbar();
                    CK_T(c_e1);
                    if (mine) {
                        const int nsl = op.kc > 0 ? nks : 1;
                        const uint32_t bar_local = act0 + 8u * (act_n & 1u);
#pragma unroll 1
                        for (int er = er0; er < rcnt; er += er_step) {
                            const int em = row0 + er;
                            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                            const float* rp = red + er * fcp + 4 * eg;
#pragma unroll 8
                            for (int q = 0; q < nsl; ++q) {
                                const float4 p = *reinterpret_cast<const float4*>(rp + q * kstride);
                                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
                            }
                            float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float t = v[j] + e_bias[j];
                                if (op.bn_mean) {
                                    t = (t - e_mean[j]) * (1.f / sqrtf(e_var[j] + op.bn_eps)) * e_gamma[j] + e_beta[j];
                                    if (op.bn_relu) t = fmaxf(t, 0.f);
                                }
                                v[j] = t;
                            }
                            if (op.act != TD_ACT_NONE) {
                                const float4 a = ck_act4(make_float4(v[0], v[1], v[2], v[3]), op.act);
                                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                            }
                            if (op.res_off >= 0) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] += arena[op.res_off + (rbase + er) * op.ldr + min(en + j, op.N - 1)];
                            }
                            if (op.gidx) {
                                const long long gi = er == er0 ? e_gi0 : __ldg(op.gidx + min(em, M - 1));
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] += __ldg(op.gtab + gi * op.ldt + min(en + j, op.N - 1));
                            }
                            if (er == er0) CK_T(c_e2);
                            if (op.out_global) {
                                if (em < M) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (en + j < op.N) op.gout[(long long)em * op.gldo + en + j] = v[j];
                                    if (step.x) {
                                        // the reverse-step update of these (<= 4) elements; N % 4 == 0 (checked on the host), so
                                        // they are one 16-byte piece of the flat state: piece index i of psample_kernel
                                        const int t = step.t_dev[0];
                                        if (t >= 0 && t < step.num_timesteps && en + 3 < op.N) {
                                            const float4 c = reinterpret_cast<const float4*>(step.coef)[t];
                                            const long long i4 = ((long long)em * op.N + en) >> 2;
                                            float zz[4] = {0.f, 0.f, 0.f, 0.f};
                                            if (t > 0) {
                                                if (step.z) {
                                                    const float4 zv = reinterpret_cast<const float4*>(step.z + (long long)t * step.z_step_stride)[i4];
                                                    zz[0] = zv.x; zz[1] = zv.y; zz[2] = zv.z; zz[3] = zv.w;
                                                } else if (step.seed) {
                                                    Philox rng(step.seed[0]);
                                                    rng.normal4((uint64_t)i4, step.seed[1] + (uint64_t)t, zz);
                                                }
                                            }
                                            float4* xp = reinterpret_cast<float4*>(step.x) + i4;
                                            const float4 xv = *xp;
                                            float4 o4;
                                            o4.x = __fadd_rn(__fmul_rn(c.x, __fsub_rn(xv.x, __fmul_rn(c.y, v[0]))), __fmul_rn(c.z, zz[0]));
                                            o4.y = __fadd_rn(__fmul_rn(c.x, __fsub_rn(xv.y, __fmul_rn(c.y, v[1]))), __fmul_rn(c.z, zz[1]));
                                            o4.z = __fadd_rn(__fmul_rn(c.x, __fsub_rn(xv.z, __fmul_rn(c.y, v[2]))), __fmul_rn(c.z, zz[2]));
                                            o4.w = __fadd_rn(__fmul_rn(c.x, __fsub_rn(xv.w, __fmul_rn(c.y, v[3]))), __fmul_rn(c.z, zz[3]));
                                            *xp = o4;
                                        }
                                    }
                                }
                            } else {
                                const int o = op.out_off + (rbase + er) * op.ldo + en;
                                const uint32_t local = (uint32_t)__cvta_generic_to_shared(arena + o);
                                const bool vec = en + 3 < op.N && (o & 3) == 0;
                                if (vec) {
                                    *reinterpret_cast<float4*>(arena + o) = make_float4(v[0], v[1], v[2], v[3]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (en + j < op.N) arena[o + j] = v[j];
                                }
#pragma unroll
                                for (int d = 1; d < CK_CL; ++d) {
                                    const uint32_t peer = (uint32_t)((rank + d) & (CK_CL - 1));
                                    const uint32_t dst = map_to_rank(local, peer), dbar = map_to_rank(bar_local, peer);
                                    if (vec) {
                                        st_async_v4(dst, make_float4(v[0], v[1], v[2], v[3]), dbar);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            if (en + j < op.N) st_async_f32(dst + 4u * j, v[j], dbar);
                                    }
                                }
                            }
                        }
                    }
                    CK_T(c_epi);
                    if (!op.out_global) {
                        sm100::mbar_wait(my_act, act_parity);          // this group's rows of the op have landed (complete_tx by the peers' st.async)
                        ++act_n;
                    }
                    gbar();                                            // `red`, and the own tile stored above, are settled for the group
                    CK_T(c_bar);
                    continue;
                }
                if (op.kind == 1) {
                    if (op.N <= 32 * LN_CACHE) {
                        if (ln_for != oi) {
#pragma unroll
                            for (int i = 0; i < LN_CACHE; ++i) {
                                const int n = min(lane + 32 * i, op.N - 1);
                                ln_g[i] = __ldg(op.w + n);
                                ln_b[i] = __ldg(op.bias + n);
                            }
                        }
#pragma unroll 1
                        for (int r = gwarp; r < rcnt; r += CK2_GT / 32) {
                            const float* xr = arena + op.x_off + (rbase + r) * op.ldx;
                            float* orow = arena + op.out_off + (rbase + r) * op.ldo;
                            float xv[LN_CACHE];
#pragma unroll
                            for (int i = 0; i < LN_CACHE; ++i) xv[i] = lane + 32 * i < op.N ? xr[lane + 32 * i] : 0.f;
                            float s = 0.f;
#pragma unroll
                            for (int i = 0; i < LN_CACHE; ++i) if (lane + 32 * i < op.N) s += xv[i];
                            const float mean = warp_sum(s) / (float)op.N;
                            float q = 0.f;
#pragma unroll
                            for (int i = 0; i < LN_CACHE; ++i)
                                if (lane + 32 * i < op.N) { const float d = xv[i] - mean; q = fmaf(d, d, q); }
                            const float rstd = 1.f / sqrtf(warp_sum(q) / (float)op.N + op.ln_eps);
#pragma unroll
                            for (int i = 0; i < LN_CACHE; ++i)
                                if (lane + 32 * i < op.N) orow[lane + 32 * i] = (xv[i] - mean) * rstd * ln_g[i] + ln_b[i];
                        }
                    } else {
#pragma unroll 1
                        for (int r = gwarp; r < rcnt; r += CK2_GT / 32) {
                            const float* xr = arena + op.x_off + (rbase + r) * op.ldx;
                            float* orow = arena + op.out_off + (rbase + r) * op.ldo;
                            float s = 0.f;
                            for (int n = lane; n < op.N; n += 32) s += xr[n];
                            const float mean = warp_sum(s) / (float)op.N;
                            float q = 0.f;
                            for (int n = lane; n < op.N; n += 32) { const float d = xr[n] - mean; q = fmaf(d, d, q); }
                            const float rstd = 1.f / sqrtf(warp_sum(q) / (float)op.N + op.ln_eps);
                            for (int n = lane; n < op.N; n += 32) orow[n] = (xr[n] - mean) * rstd * __ldg(op.w + n) + __ldg(op.bias + n);
                        }
                    }
                } else if (op.kind == 2) {
                    for (int e = gtid; e < rcnt * op.N; e += CK2_GT) {
                        const int r = e / op.N, n = e - r * op.N;
                        const float v = arena[op.x_off + (rbase + r) * op.ldx + n];
                        float* o = arena + op.out_off + (rbase + r) * op.ldo + n;
                        *o = op.accumulate ? *o + v : v;
                    }
                } else if (op.kind == 3) {
                    const int width = op.tmode == 2 ? op.N : 1;
                    for (int e = gtid; e < rcnt * width; e += CK2_GT) {
                        const int r = e / width, j = e - r * width;
                        const int m = min(row0 + r, M - 1);
                        float tv = op.t ? (float)op.t[m] : (float)op.t_dev[0];
                        if (op.tmode == 1) tv = tv / 1000.0f;
                        float v = tv;
                        if (op.tmode == 2) {
                            const int half = op.N / 2;
                            v = 0.f;
                            if (j < 2 * half) {
                                const int jj = (j < half) ? j : j - half;
                                const float arg = tv * expf(-logf(10000.0f) * (float)jj / (float)(half - 1));
                                v = (j < half) ? sinf(arg) : cosf(arg);
                            }
                        }
                        arena[op.out_off + (rbase + r) * op.ldo + j] = v;
                    }
                } else {
                    for (int e = gtid; e < rcnt * op.N; e += CK2_GT) {
                        const int r = e / op.N, n = e - r * op.N;
                        const int m = row0 + r;
                        arena[op.out_off + (rbase + r) * op.ldo + n] = m < M ? __ldcg(op.gx + (long long)m * op.gldx + n) : 0.f;
                    }
                }
                gbar();
                CK_T(c_row);
            }
        }
        if (dbg && gtid == 0 && g == 0 && blockIdx.x < kNumSMs) {
            unsigned long long* d = g_ck_dbg + blockIdx.x * 16;
            d[0] = g_start; d[1] = ck_gtime();
            d[2] = c_wait; d[3] = c_fma; d[4] = c_epi; d[5] = c_bar; d[6] = c_row; d[7] = c_pre; d[8] = c_e1; d[9] = c_e2;
        }
    }
    cluster_sync_all();          // no CTA leaves while a peer may still write into its shared memory
    if (step.x && tid == 0) {    // every read of t_dev by this CTA is behind it: the last CTA of the grid counts the step down
        const int t = step.t_dev[0];
        __threadfence();
        if (atomicAdd(step.ticket, 1u) == gridDim.x - 1) {
            *step.ticket = 0u;
            step.t_dev[0] = t - 1;
        }
    }
}

}  // namespace td

using namespace td;

extern "C" int td_dense_cluster_op_bytes(void) { return (int)sizeof(ClusterOp); }

template <int R>
static void ck_fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int clusters, cudaStream_t s) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(clusters * CK_CL);
    cfg.blockDim = dim3(CK_THREADS);
    cfg.dynamicSmemBytes = CK_SMEM;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CK_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
}

// Clusters of 8 CTAs (one CTA per SM, ~223 KB of shared memory each) the device keeps resident at once: 15 on a B200 (measured;
// the GPCs do not all hold two such clusters).  Queried once; 15 when no device answers (tape construction on a CPU-only host).
static int ck_max_active_clusters() {
    static int cached = 0;
    if (cached > 0) return cached;
    int n = 0;
    if (cudaFuncSetAttribute(dense_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, CK_SMEM) == cudaSuccess &&
        cudaFuncSetAttribute(dense_cluster_kernel<8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0) == cudaSuccess) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr[2];
        ck_fill_cfg<8>(cfg, attr, 16, nullptr);
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, dense_cluster_kernel<8>, &cfg) != cudaSuccess) n = 0;
    }
    (void)cudaGetLastError();
    cached = n > 0 ? std::min(n, kNumSMs / CK_CL) : 15;
    return cached;
}

// rows per cluster for a batch: 8, or 9 when that saves a pass over the weights (128 rows: 16 slices of 8 on 15 clusters = 2 passes)
static int ck_rows(int batch) {
    const int maxc = ck_max_active_clusters();
    const int p8 = (int)ceil_div(ceil_div(batch, 8), maxc), p9 = (int)ceil_div(ceil_div(batch, 9), maxc);
    return p9 < p8 ? 9 : 8;
}

extern "C" int td_dense_cluster_limits(int batch, int* rows, int* cluster, int* arena_floats, int* stage_floats, int* max_ops,
                                       int* max_clusters) {
    if (rows) *rows = ck_rows(batch > 0 ? batch : 1);
    if (max_clusters) *max_clusters = ck_max_active_clusters();
    if (cluster) *cluster = CK_CL;
    if (arena_floats) *arena_floats = CK_ARENA_FLOATS;
    if (stage_floats) *stage_floats = CK_STAGE_FLOATS;
    if (max_ops) *max_ops = CK_MAX_OPS;
    return TD_OK;
}

// Tensor map of one Linear's weight [n][k] (fp32, row-major) for the cluster kernel: boxes of 32 floats x fcp rows, SWIZZLE_128B.
// Written to HOST memory (128 bytes); the caller copies it to a 64-byte aligned device buffer and points ClusterOp::tmap at it.
extern "C" int td_dense_cluster_weight_map(const float* w, int n, int k, int fcp, void* map_out_host) {
    TD_CHECK_ARG(w && map_out_host && n > 0 && k > 0 && k % 4 == 0 && fcp >= 4 && fcp <= 256 && (((uintptr_t)w) & 15) == 0,
                 "td_dense_cluster_weight_map: bad arguments");
    EncodeTiledFn encode = tc_get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled not available from the driver"); return TD_ERR_DRIVER; }
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)k * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)fcp};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dense weight %d x %d, box 32 x %d) failed: %d", n, k, fcp, (int)r); return TD_ERR_DRIVER; }
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    memcpy(map_out_host, &m, sizeof(m));
    return TD_OK;
}

// 2 (default) / 3: independent row groups + producer warp per CTA (dense_cluster2_kernel); 1: one group (dense_cluster_kernel)
static int ck_groups() {
    static int g = 0;
    if (!g) { const char* e = getenv("TD_DENSE_CLUSTER_GROUPS"); const int v = e ? atoi(e) : 2; g = (v == 1 || v == 3) ? v : 2; }
    return g;
}

template <int R>
static int ck_launch(const void* ops, int n_ops, int batch, int max_clusters, cudaStream_t s, const ClusterStep& step) {
    static bool configured = false;
    if (!configured) {
        TD_CUDA(cudaFuncSetAttribute(dense_cluster_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, CK_SMEM));
        TD_CUDA(cudaFuncSetAttribute(dense_cluster2_kernel<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CK2_SMEM));
        TD_CUDA(cudaFuncSetAttribute(dense_cluster2_kernel<R, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CK2_SMEM));
        configured = true;
    }
    int clusters = std::min((int)ceil_div(batch, R), ck_max_active_clusters());
    if (max_clusters > 0) clusters = std::min(clusters, max_clusters);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    ck_fill_cfg<R>(cfg, attr, clusters, s);
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TD_DENSE_CLUSTER_DBG"); dbg = e ? atoi(e) : 0; }
    if (dbg & 2) {
        fprintf(stderr, "tinydiff: dense_cluster: batch %d, %d rows per cluster, %d clusters of %d (max active %d)\n", batch, R, clusters,
                CK_CL, ck_max_active_clusters());
        dbg &= ~2;
    }
    count_launch();
    if (ck_groups() >= 2) {
        cfg.blockDim = dim3(ck_groups() * CK2_GT + 32);
        cfg.dynamicSmemBytes = CK2_SMEM;
        if (ck_groups() == 3)
            (void)cudaLaunchKernelEx(&cfg, dense_cluster2_kernel<R, 3>, reinterpret_cast<const ClusterOp*>(ops), n_ops, batch, dbg, step);
        else
            (void)cudaLaunchKernelEx(&cfg, dense_cluster2_kernel<R, 2>, reinterpret_cast<const ClusterOp*>(ops), n_ops, batch, dbg, step);
    } else {
        if (step.x) { set_error("td_dense_cluster_step: the one-group kernel (TD_DENSE_CLUSTER_GROUPS=1) has no fused reverse step"); return TD_ERR_UNSUPPORTED; }
        (void)cudaLaunchKernelEx(&cfg, dense_cluster_kernel<R>, reinterpret_cast<const ClusterOp*>(ops), n_ops, batch, dbg);
    }
    return launch_status("dense_cluster");
}

extern "C" int td_dense_cluster_run(const void* ops, int n_ops, int batch, int rows, int max_clusters, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(ops && n_ops > 0 && n_ops <= CK_MAX_OPS && batch > 0, "td_dense_cluster_run: bad args");
    TD_CHECK_ARG(rows == 8 || rows == 9, "td_dense_cluster_run: rows per cluster must be 8 or 9 (the tape's arena offsets were laid out for it)");
    const ClusterStep none{};
    return rows == 9 ? ck_launch<9>(ops, n_ops, batch, max_clusters, (cudaStream_t)stream, none)
                     : ck_launch<8>(ops, n_ops, batch, max_clusters, (cudaStream_t)stream, none);
}

// 1: td_dense_cluster_step is available (the multi-group kernel is the one that runs)
extern "C" int td_dense_cluster_step_fused(void) { return ck_groups() >= 2 ? 1 : 0; }

// One reverse step of the sampler in ONE launch: the tape (its last op must be the Linear that writes eps [batch][n] to global
// memory, n % 4 == 0) followed, in that Linear's epilogue, by x <- c1[t] (x - c2[t] eps) + c3[t] z and t_dev[0] -= 1 -- the
// arithmetic, noise-table layout and Philox indexing of td_psample_step_advance.  x: the sampler state the tape loads [batch][n].
extern "C" int td_dense_cluster_step(const void* ops, int n_ops, int batch, int rows, int max_clusters, float* x, const float* coef,
                                     const float* z, int64_t z_step_stride, const uint64_t* seed_ptr, int32_t* t_dev,
                                     unsigned int* ticket, int num_timesteps, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(ops && n_ops > 0 && n_ops <= CK_MAX_OPS && batch > 0, "td_dense_cluster_step: bad args");
    TD_CHECK_ARG(rows == 8 || rows == 9, "td_dense_cluster_step: rows per cluster must be 8 or 9");
    TD_CHECK_ARG(x && coef && t_dev && ticket && num_timesteps > 0, "td_dense_cluster_step: null pointer");
    TD_CHECK_ARG((((uintptr_t)x | (uintptr_t)z) & 15) == 0 && z_step_stride % 4 == 0, "td_dense_cluster_step: x and the noise rows must be 16-byte aligned");
    ClusterStep st{};
    st.x = x; st.coef = coef; st.z = z; st.z_step_stride = z_step_stride;
    st.seed = reinterpret_cast<const unsigned long long*>(seed_ptr);
    st.t_dev = t_dev; st.ticket = ticket; st.num_timesteps = num_timesteps;
    return rows == 9 ? ck_launch<9>(ops, n_ops, batch, max_clusters, (cudaStream_t)stream, st)
                     : ck_launch<8>(ops, n_ops, batch, max_clusters, (cudaStream_t)stream, st);
}

// Tuning aid: the per-CTA counters of the last launch run with TD_DENSE_CLUSTER_DBG=1 (synchronises).
extern "C" int td_dense_cluster_debug_counters(unsigned long long* host_out, int n) {
    if (!host_out || n <= 0 || n > kNumSMs * 16) { td::set_error("td_dense_cluster_debug_counters: bad arguments"); return TD_ERR_ARG; }
    TD_CUDA(cudaDeviceSynchronize());
    TD_CUDA(cudaMemcpyFromSymbol(host_out, td::g_ck_dbg, (size_t)n * sizeof(unsigned long long)));
    return TD_OK;
}
