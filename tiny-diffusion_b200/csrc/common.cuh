// Shared helpers for libtinydiff (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tinydiff.h"

namespace td {

void set_error(const char* fmt, ...);

#define TD_CHECK_ARG(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            td::set_error(__VA_ARGS__);              \
            return TD_ERR_ARG;                       \
        }                                            \
    } while (0)

#define TD_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            td::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

// Fails with TD_ERR_ARCH unless the current device is sm_100-class.  Cached per device.
int require_sm100();

#define TD_REQUIRE_ARCH()                 \
    do {                                  \
        int _a = td::require_sm100();     \
        if (_a != TD_OK) return _a;       \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return TD_OK;
}

constexpr int kNumSMs = 148;
// SMs the persistent / one-CTA-per-SM kernels size their grids for (td_set_sm_budget): fewer than kNumSMs when other work
// holds SMs for the duration of a step -- NCCL's all-reduce CTAs in the data-parallel train step -- so that a 148-CTA grid
// does not run as 140 + a second wave of 8.  Read when a PLAN is created and stored in it.
int sm_budget();

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute (unless TD_PDL=0) and runs
// pdl_sync() before it touches global memory: `griddepcontrol.wait` returns once the preceding kernel of the stream has
// completed and flushed, so a kernel's launch latency, block scheduling and on-chip prologue (barrier init, TMEM allocation,
// tensor-map prefetch) overlap the tail of its predecessor; `griddepcontrol.launch_dependents` then lets the successor do the
// same.  Because EVERY kernel waits, completion stays transitive along the stream (a kernel cannot finish before all earlier
// ones have).  Kernels that allocate TMEM call pdl_sync() after the allocation: a successor made resident on the same SM can
// then never take the columns first.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();      // api.cu: TD_PDL environment switch (default on), read once
void count_launch();     // api.cu: bumps the counter td_launch_count() reports

struct LaunchCfg {
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    LaunchCfg(dim3 g, dim3 b, size_t s = 0, cudaStream_t st = nullptr) : grid(g), block(b), smem(s), stream(st) {}
};

template <typename... P, typename... A>
inline void launch(void (*kernel)(P...), const LaunchCfg& c, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = c.grid;
    cfg.blockDim = c.block;
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    count_launch();
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);       // errors surface through launch_status()
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- dtype helpers -------------------------------------------------------------------------
template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float> {
    static constexpr int N = 4;
    float4 v;
    __device__ static Vec load(const float* p) { Vec r; r.v = *reinterpret_cast<const float4*>(p); return r; }
    __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ void unpack(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ static Vec pack(const float* f) { Vec r; r.v = make_float4(f[0], f[1], f[2], f[3]); return r; }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    uint4 v;
    __device__ static Vec load(const __nv_bfloat16* p) { Vec r; r.v = *reinterpret_cast<const uint4*>(p); return r; }
    __device__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
    __device__ void unpack(float* f) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ static Vec pack(const float* f) {
        Vec r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r.v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return r;
    }
};

__device__ inline float to_f32(float x) { return x; }
__device__ inline float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ inline T from_f32(float x);
template <> __device__ inline float from_f32<float>(float x) { return x; }
template <> __device__ inline __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ inline float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 counter-based RNG -> N(0,1) --------------------------------------------------
struct Philox {
    uint32_t key[2];
    __device__ Philox(uint64_t seed) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
    __device__ static inline void round(uint32_t* c, uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // 4 x uint32 for counter (idx, subsequence)
    __device__ inline void gen(uint64_t idx, uint64_t sub, uint32_t* out) const {
        uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)};
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // 4 standard normals (Box-Muller on two pairs)
    __device__ inline void normal4(uint64_t idx, uint64_t sub, float* z) const {
        uint32_t r[4];
        gen(idx, sub, r);
        const float k = 2.3283064365386963e-10f;   // 2^-32
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float u1 = ((float)r[2 * i] + 0.5f) * k;          // (0,1]
            float u2 = ((float)r[2 * i + 1] + 0.5f) * k;
            float rad = sqrtf(-2.0f * __logf(u1));
            float s, c;
            __sincosf(6.283185307179586f * u2, &s, &c);
            z[2 * i] = rad * c;
            z[2 * i + 1] = rad * s;
        }
    }
};

}  // namespace td
