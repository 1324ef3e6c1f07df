// Elementwise DDPM kernels: q_sample, MSE loss + gradient, p_sample step, fused Adam.
// All are HBM-bound streaming kernels: 128-bit accesses, grid sized from the element count,
// no shared-memory staging (no reuse to exploit).
#include "common.cuh"

namespace td {

constexpr int kEwThreads = 256;

// x_t = sqrt(abar[t]) * x0 + sqrt(1 - abar[t]) * noise          (diffusion.py:180-190)
// __fmul_rn/__fadd_rn keep the reference's separate roundings (no FMA contraction) so the
// result is bit-identical to the PyTorch expression.
__global__ void __launch_bounds__(kEwThreads)
qsample_kernel(const float* __restrict__ x0, float* __restrict__ noise, const int64_t* __restrict__ t,
               const float* __restrict__ abar, float* __restrict__ x_t, int64_t per_sample4,
               int num_timesteps, const uint64_t* __restrict__ seed_ptr) {
    td::pdl_sync();
    const int64_t b = blockIdx.y;
    int tt = (int)t[b];
    tt = min(max(tt, 0), num_timesteps - 1);
    const float ab = abar[tt];
    const float ca = sqrtf(ab);
    const float cb = sqrtf(__fsub_rn(1.0f, ab));
    const int64_t base = b * per_sample4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_sample4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = reinterpret_cast<const float4*>(x0)[base + i];
        float4 n;
        if (seed_ptr) {
            Philox rng(seed_ptr[0]);
            float z[4];
            rng.normal4((uint64_t)(base + i), seed_ptr[1], z);
            n = make_float4(z[0], z[1], z[2], z[3]);
            reinterpret_cast<float4*>(noise)[base + i] = n;
        } else {
            n = reinterpret_cast<const float4*>(noise)[base + i];
        }
        float4 o;
        o.x = __fadd_rn(__fmul_rn(ca, x.x), __fmul_rn(cb, n.x));
        o.y = __fadd_rn(__fmul_rn(ca, x.y), __fmul_rn(cb, n.y));
        o.z = __fadd_rn(__fmul_rn(ca, x.z), __fmul_rn(cb, n.z));
        o.w = __fadd_rn(__fmul_rn(ca, x.w), __fmul_rn(cb, n.w));
        reinterpret_cast<float4*>(x_t)[base + i] = o;
    }
}

// scalar tail-safe variant (per_sample not a multiple of 4, e.g. the 20-d latents)
__global__ void __launch_bounds__(kEwThreads)
qsample_scalar_kernel(const float* __restrict__ x0, float* __restrict__ noise, const int64_t* __restrict__ t,
                      const float* __restrict__ abar, float* __restrict__ x_t, int64_t batch,
                      int64_t per_sample, int num_timesteps, const uint64_t* __restrict__ seed_ptr) {
    td::pdl_sync();
    const int64_t total = batch * per_sample;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per_sample;
        int tt = (int)t[b];
        tt = min(max(tt, 0), num_timesteps - 1);
        const float ab = abar[tt];
        float n;
        if (seed_ptr) {
            Philox rng(seed_ptr[0]);
            float z[4];
            rng.normal4((uint64_t)(i >> 2), seed_ptr[1], z);
            n = z[i & 3];
            noise[i] = n;
        } else {
            n = noise[i];
        }
        x_t[i] = __fadd_rn(__fmul_rn(sqrtf(ab), x0[i]), __fmul_rn(sqrtf(__fsub_rn(1.0f, ab)), n));
    }
}

// loss = mean((pred-target)^2), grad = 2*(pred-target)/n          (diffusion.py:231)
__global__ void __launch_bounds__(kEwThreads)
mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ grad,
                float* __restrict__ loss_out, float* __restrict__ partials, unsigned int* __restrict__ counter,
                int64_t n, float inv_n) {
    td::pdl_sync();
    float acc = 0.f;
    const float g2 = 2.0f * inv_n;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = reinterpret_cast<const float4*>(pred)[i];
        float4 q = reinterpret_cast<const float4*>(target)[i];
        float4 d = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
        acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g2 * d.x, g2 * d.y, g2 * d.z, g2 * d.w);
    }
    if (blockIdx.x == 0) {
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float d = pred[i] - target[i];
            acc += d * d;
            if (grad) grad[i] = g2 * d;
        }
    }
    __shared__ float warp_part[kEwThreads / 32];
    __shared__ bool is_last;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kEwThreads / 32; ++i) s += warp_part[i];
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order final reduction by one warp (deterministic for a given grid)
        if (threadIdx.x < 32) {
            double s = 0.0;
            for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) s += (double)__ldcg(&partials[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (threadIdx.x == 0) {
                loss_out[0] = (float)(s * (double)inv_n);
                *counter = 0u;
            }
        }
    }
}

// t_dev[0] = t - 1 by the LAST block of the grid to get here (every block has read t by then); `ticket` is a zero-initialised
// counter the last block resets.  Folds the per-step td_counter_add launch of the reverse loop into the step kernel.
__device__ inline void step_counter_advance(int32_t* t_dev, int t, unsigned int* ticket) {
    if (!ticket) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            t_dev[0] = t - 1;
        }
    }
}

// x <- c1*(x - c2*eps) + c3*z        (diffusion.py:272-274), separate roundings as in PyTorch.
__global__ void __launch_bounds__(kEwThreads)
psample_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z_base,
               int64_t z_step_stride, const float* __restrict__ coef, int32_t* t_dev, int64_t n,
               int num_timesteps, const uint64_t* __restrict__ seed_ptr, unsigned int* ticket) {
    td::pdl_sync();
    const int t = t_dev[0];
    if (t < 0 || t >= num_timesteps) {                // a graph replayed past t = 0 (or started above T-1) is a no-op
        step_counter_advance(t_dev, t, ticket);
        return;
    }
    const float* __restrict__ z = z_base ? z_base + (int64_t)t * z_step_stride : nullptr;
    const float4 c = reinterpret_cast<const float4*>(coef)[t];
    const float c1 = c.x, c2 = c.y, c3 = c.z;
    const bool use_noise = (t > 0);
    // 16-byte accesses need every row of the noise table aligned too (odd latent sizes: stride % 4 != 0)
    const bool vec_ok = (((uintptr_t)x | (uintptr_t)eps | (uintptr_t)z) & 15) == 0;
    const int64_t n4 = vec_ok ? (n >> 2) : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 xv = reinterpret_cast<const float4*>(x)[i];
        float4 ev = reinterpret_cast<const float4*>(eps)[i];
        float zz[4] = {0.f, 0.f, 0.f, 0.f};
        if (use_noise) {
            if (z) {
                float4 zv = reinterpret_cast<const float4*>(z)[i];
                zz[0] = zv.x; zz[1] = zv.y; zz[2] = zv.z; zz[3] = zv.w;
            } else if (seed_ptr) {
                Philox rng(seed_ptr[0]);
                rng.normal4((uint64_t)i, seed_ptr[1] + (uint64_t)t, zz);
            }
        }
        float4 o;
        o.x = __fadd_rn(__fmul_rn(c1, __fsub_rn(xv.x, __fmul_rn(c2, ev.x))), __fmul_rn(c3, zz[0]));
        o.y = __fadd_rn(__fmul_rn(c1, __fsub_rn(xv.y, __fmul_rn(c2, ev.y))), __fmul_rn(c3, zz[1]));
        o.z = __fadd_rn(__fmul_rn(c1, __fsub_rn(xv.z, __fmul_rn(c2, ev.z))), __fmul_rn(c3, zz[2]));
        o.w = __fadd_rn(__fmul_rn(c1, __fsub_rn(xv.w, __fmul_rn(c2, ev.w))), __fmul_rn(c3, zz[3]));
        reinterpret_cast<float4*>(x)[i] = o;
    }
    {   // scalar tail (< 4 elements), or the whole tensor when the noise rows are not 16-byte aligned: grid-strided
        for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            float zz = 0.f;
            if (use_noise) {
                if (z) zz = z[i];
                else if (seed_ptr) {
                    Philox rng(seed_ptr[0]);
                    float q[4];
                    rng.normal4((uint64_t)(i >> 2), seed_ptr[1] + (uint64_t)t, q);
                    zz = q[i & 3];
                }
            }
            x[i] = __fadd_rn(__fmul_rn(c1, __fsub_rn(x[i], __fmul_rn(c2, eps[i]))), __fmul_rn(c3, zz));
        }
    }
    step_counter_advance(t_dev, t, ticket);
}

// Classifier-free-guidance reverse step (extension; the reference has no guidance).  The denoiser ran
// on a doubled batch: rows [0, n) with the labels, rows [n, 2n) with the null label, both halves holding the same x.
//   eps = eps_u + w*(eps_c - eps_u);   x <- c1*(x - c2*eps) + c3*z,   written to both halves (same noise for both).
// Separate roundings, so the result equals the torch composition bit for bit.  20 algorithmic bytes per element of n.
__global__ void __launch_bounds__(kEwThreads)
psample_cfg_kernel(float* __restrict__ x, const float* __restrict__ eps, int64_t n, float w,
                   const float* __restrict__ z_base, int64_t z_step_stride, const float* __restrict__ coef,
                   int32_t* t_dev, int num_timesteps, const uint64_t* __restrict__ seed_ptr, unsigned int* ticket) {
    td::pdl_sync();
    const int t = t_dev[0];
    if (t < 0 || t >= num_timesteps) {
        step_counter_advance(t_dev, t, ticket);
        return;
    }
    const float* __restrict__ z = z_base ? z_base + (int64_t)t * z_step_stride : nullptr;
    const float4 c = reinterpret_cast<const float4*>(coef)[t];
    const float c1 = c.x, c2 = c.y, c3 = c.z;
    const bool use_noise = (t > 0);
    const int64_t n4 = n >> 2;                         // n is a multiple of 4 (checked on the host)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        const float4 ec = reinterpret_cast<const float4*>(eps)[i];
        const float4 eu = reinterpret_cast<const float4*>(eps + n)[i];
        float zz[4] = {0.f, 0.f, 0.f, 0.f};
        if (use_noise) {
            if (z) {
                const float4 zv = reinterpret_cast<const float4*>(z)[i];
                zz[0] = zv.x; zz[1] = zv.y; zz[2] = zv.z; zz[3] = zv.w;
            } else if (seed_ptr) {
                Philox rng(seed_ptr[0]);
                rng.normal4((uint64_t)i, seed_ptr[1] + (uint64_t)t, zz);
            }
        }
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, cs[4] = {ec.x, ec.y, ec.z, ec.w}, us[4] = {eu.x, eu.y, eu.z, eu.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float e = __fadd_rn(us[k], __fmul_rn(w, __fsub_rn(cs[k], us[k])));
            o[k] = __fadd_rn(__fmul_rn(c1, __fsub_rn(xs[k], __fmul_rn(c2, e))), __fmul_rn(c3, zz[k]));
        }
        const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(x)[i] = ov;
        reinterpret_cast<float4*>(x + n)[i] = ov;
    }
    step_counter_advance(t_dev, t, ticket);
}

// Global gradient L2 norm + clip factor (torch.nn.utils.clip_grad_norm_, conditional_diffusion_laion.py:471) over the flat
// gradient buffer: per-block partial sums of squares, the last block to finish reduces them in fixed order (deterministic
// for a given grid) and writes
//   norm = pre_scale * sqrt(sum g^2)          (pre_scale = 1/world_size: the buffer holds the all-reduced SUM)
//   scale = pre_scale * min(1, max_norm / (norm + 1e-6))       -> td_adam_multi's grad_scale_dev
// max_norm <= 0 disables clipping (scale = pre_scale).  4 algorithmic bytes per gradient element.
__global__ void __launch_bounds__(kEwThreads)
grad_clip_kernel(const float* __restrict__ g, int64_t n, float pre_scale, float max_norm, float* __restrict__ partials,
                 unsigned int* __restrict__ counter, float* __restrict__ scale_out, float* __restrict__ norm_out) {
    td::pdl_sync();
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc += g[i] * g[i];
    __shared__ float warp_part[kEwThreads / 32];
    __shared__ bool is_last;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kEwThreads / 32; ++i) s += warp_part[i];
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double s = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) s += (double)__ldcg(&partials[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) {
            const float norm = pre_scale * (float)sqrt(s);
            float clip = 1.0f;
            if (max_norm > 0.f) clip = fminf(max_norm / (norm + 1e-6f), 1.0f);
            scale_out[0] = pre_scale * clip;
            if (norm_out) norm_out[0] = norm;
            *counter = 0u;
        }
    }
}

// t ~ randint(low, high) on the device (diffusion.py:220), Philox keyed by seed_ptr[0], subsequence seed_ptr[1] with the top
// bit set (so the stream is disjoint from the noise q_sample draws with the same pair).  Our own RNG stream, not torch's.
__global__ void __launch_bounds__(kEwThreads)
randint_kernel(int64_t* __restrict__ out, int64_t n, int low, int high, const uint64_t* __restrict__ seed_ptr) {
    td::pdl_sync();
    const Philox rng(seed_ptr[0]);
    const uint64_t sub = seed_ptr[1] | 0x8000000000000000ull;
    const uint32_t span = (uint32_t)(high - low);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n + 3) / 4; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        rng.gen((uint64_t)i, sub, r);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * i + k < n) out[4 * i + k] = (int64_t)low + (int64_t)__umulhi(r[k], span);      // floor(r * span / 2^32)
    }
}

__global__ void __launch_bounds__(kEwThreads) fill_f32_kernel(float* __restrict__ dst, int64_t n, float v) {
    td::pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = v;
}

__global__ void seed_advance_kernel(uint64_t* seed, uint64_t delta) {
    td::pdl_sync(); seed[1] += delta; }

__global__ void counter_add_kernel(int32_t* c, int32_t delta) {
    td::pdl_sync(); c[0] += delta; }

// Fused multi-tensor Adam (torch.optim.Adam defaults; diffusion.py:211,236).
__global__ void __launch_bounds__(kEwThreads)
adam_multi_kernel(float* const* __restrict__ p, const float* const* __restrict__ g, float* const* __restrict__ m,
                  float* const* __restrict__ v, const int64_t* __restrict__ numel,
                  const int32_t* __restrict__ chunk_tensor, const int64_t* __restrict__ chunk_offset,
                  int64_t chunk_elems, const int32_t* __restrict__ step_dev, float lr_host,
                  const float* __restrict__ lr_dev, double beta1_d, double beta2_d,
                  float eps, const float* __restrict__ grad_scale_dev, void* const* __restrict__ bf16_shadow) {
    td::pdl_sync();
    const int tid = chunk_tensor[blockIdx.x];
    const int64_t off = chunk_offset[blockIdx.x];
    const int64_t n = min(chunk_elems, numel[tid] - off);
    float* __restrict__ pp = p[tid] + off;
    const float* __restrict__ gp = g[tid] + off;
    float* __restrict__ mp = m[tid] + off;
    float* __restrict__ vp = v[tid] + off;
    __nv_bfloat16* sh = bf16_shadow ? reinterpret_cast<__nv_bfloat16*>(bf16_shadow[tid]) : nullptr;
    if (sh) sh += off;
    // Scalars as torch.optim.Adam forms them: bias corrections, step size and (1 - beta) in double from the double betas
    // (Python floats there), rounded to fp32 once.  In fp32, 1 - 0.999f is off by 1.3e-5 relative, which goes straight into
    // exp_avg_sq and the update.
    const double step = (double)step_dev[0];
    const float lr = lr_dev ? lr_dev[0] : lr_host;      // device scalar: schedulers change it under a captured graph
    const double bc1 = 1.0 - pow(beta1_d, step);
    const double bc2 = 1.0 - pow(beta2_d, step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float gs = grad_scale_dev ? grad_scale_dev[0] : 1.0f;
    const float beta2 = (float)beta2_d;
    const float omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d);

    auto upd = [&](float& pv, float gv, float& mv, float& vv) {
        gv *= gs;
        mv = mv + omb1 * (gv - mv);                  // exp_avg.lerp_(grad, 1-beta1)
        vv = vv * beta2 + omb2 * gv * gv;            // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = sqrtf(vv) / bc2_sqrt + eps;      // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        pv = pv - step_size * (mv / denom);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
    };
    // all four arrays share alignment only if the tensor base pointers are 16B aligned and
    // off % 4 == 0 (chunk_elems is a multiple of 4); torch allocations are 512B aligned.
    const bool vec_ok = ((((uintptr_t)pp | (uintptr_t)gp | (uintptr_t)mp | (uintptr_t)vp) & 15) == 0);
    const int64_t n4 = vec_ok ? (n >> 2) : 0;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(pp)[i];
        float4 gv = reinterpret_cast<const float4*>(gp)[i];
        float4 mv = reinterpret_cast<float4*>(mp)[i];
        float4 vv = reinterpret_cast<float4*>(vp)[i];
        upd(pv.x, gv.x, mv.x, vv.x);
        upd(pv.y, gv.y, mv.y, vv.y);
        upd(pv.z, gv.z, mv.z, vv.z);
        upd(pv.w, gv.w, mv.w, vv.w);
        reinterpret_cast<float4*>(pp)[i] = pv;
        reinterpret_cast<float4*>(mp)[i] = mv;
        reinterpret_cast<float4*>(vp)[i] = vv;
        if (sh) {
            __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
            if ((((uintptr_t)sh) & 7) == 0) {
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&a);
                u.y = *reinterpret_cast<uint32_t*>(&b);
                reinterpret_cast<uint2*>(sh)[i] = u;
            } else {
                sh[4 * i] = a.x; sh[4 * i + 1] = a.y; sh[4 * i + 2] = b.x; sh[4 * i + 3] = b.y;
            }
        }
    }
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
        float pv = pp[i], mv = mp[i], vv = vp[i];
        upd(pv, gp[i], mv, vv);
        pp[i] = pv; mp[i] = mv; vp[i] = vv;
        if (sh) sh[i] = __float2bfloat16_rn(pv);
    }
}

__global__ void __launch_bounds__(kEwThreads)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    td::pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

static inline int ew_grid(int64_t work_items) {
    int64_t blocks = ceil_div(work_items, kEwThreads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace td

using namespace td;

extern "C" int td_qsample(const float* x0, float* noise, const int64_t* t, const float* alphas_cumprod,
                          float* x_t, int64_t batch, int64_t per_sample, int num_timesteps,
                          const uint64_t* seed_ptr, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x0 && noise && t && alphas_cumprod && x_t, "td_qsample: null pointer");
    TD_CHECK_ARG(batch >= 0 && per_sample >= 0 && num_timesteps > 0, "td_qsample: bad sizes");
    if (batch == 0 || per_sample == 0) return TD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (per_sample % 4 == 0 && batch <= 65535) {
        int64_t ps4 = per_sample / 4;
        int gx = (int)std::min<int64_t>(ceil_div(ps4, kEwThreads), std::max<int64_t>(1, (kNumSMs * 16) / batch));
        dim3 grid(gx, (unsigned)batch);
        td::launch(qsample_kernel, td::LaunchCfg(grid, kEwThreads, 0, s), x0, noise, t, alphas_cumprod, x_t, ps4, num_timesteps, seed_ptr);
    } else {
        td::launch(qsample_scalar_kernel, td::LaunchCfg(ew_grid(batch * per_sample), kEwThreads, 0, s), 
            x0, noise, t, alphas_cumprod, x_t, batch, per_sample, num_timesteps, seed_ptr);
    }
    return launch_status("qsample");
}

extern "C" int64_t td_mse_num_partials(int64_t n) { return ew_grid(std::max<int64_t>(n / 4, 1)); }

extern "C" int td_mse_grad(const float* pred, const float* target, float* grad, float* loss_out, float* partials,
                           unsigned int* counter, int64_t n, float inv_n, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(pred && target && loss_out && partials && counter, "td_mse_grad: null pointer");
    TD_CHECK_ARG(n > 0, "td_mse_grad: n must be positive");
    td::launch(mse_grad_kernel, td::LaunchCfg((int)td_mse_num_partials(n), kEwThreads, 0, (cudaStream_t)stream), 
        pred, target, grad, loss_out, partials, counter, n, inv_n);
    return launch_status("mse_grad");
}

static int psample_step_launch(float* x, const float* eps, const float* z, int64_t z_step_stride, const float* coef, int32_t* t_dev,
                               int64_t n, int num_timesteps, const uint64_t* seed_ptr, unsigned int* ticket, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && eps && coef && t_dev, "td_psample_step: null pointer");
    TD_CHECK_ARG(n > 0 && num_timesteps > 0, "td_psample_step: n and num_timesteps must be positive");
    td::launch(psample_kernel, td::LaunchCfg(ew_grid(std::max<int64_t>(n / 4, 1)), kEwThreads, 0, (cudaStream_t)stream),
               x, eps, z, z_step_stride, coef, t_dev, n, num_timesteps, seed_ptr, ticket);
    return launch_status("psample_step");
}

extern "C" int td_psample_step(float* x, const float* eps, const float* z, int64_t z_step_stride, const float* coef,
                               const int32_t* t_dev, int64_t n, int num_timesteps, const uint64_t* seed_ptr, void* stream) {
    return psample_step_launch(x, eps, z, z_step_stride, coef, const_cast<int32_t*>(t_dev), n, num_timesteps, seed_ptr, nullptr, stream);
}

extern "C" int td_psample_step_advance(float* x, const float* eps, const float* z, int64_t z_step_stride, const float* coef,
                                       int32_t* t_dev, int64_t n, int num_timesteps, const uint64_t* seed_ptr, unsigned int* ticket,
                                       void* stream) {
    TD_CHECK_ARG(ticket, "td_psample_step_advance: null ticket");
    return psample_step_launch(x, eps, z, z_step_stride, coef, t_dev, n, num_timesteps, seed_ptr, ticket, stream);
}

static int psample_step_cfg_launch(float* x, const float* eps, int64_t n, float guidance, const float* z, int64_t z_step_stride,
                                   const float* coef, int32_t* t_dev, int num_timesteps, const uint64_t* seed_ptr, unsigned int* ticket,
                                   void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(x && eps && coef && t_dev, "td_psample_step_cfg: null pointer");
    TD_CHECK_ARG(n > 0 && n % 4 == 0 && num_timesteps > 0, "td_psample_step_cfg: n (elements of one half) must be a positive multiple of 4");
    TD_CHECK_ARG(!z || ((((uintptr_t)z) & 15) == 0 && z_step_stride % 4 == 0), "td_psample_step_cfg: noise table rows must be 16-byte aligned");
    td::launch(psample_cfg_kernel, td::LaunchCfg(ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)stream), x, eps, n, guidance, z,
               z_step_stride, coef, t_dev, num_timesteps, seed_ptr, ticket);
    return launch_status("psample_step_cfg");
}

extern "C" int td_psample_step_cfg(float* x, const float* eps, int64_t n, float guidance, const float* z,
                                   int64_t z_step_stride, const float* coef, const int32_t* t_dev, int num_timesteps,
                                   const uint64_t* seed_ptr, void* stream) {
    return psample_step_cfg_launch(x, eps, n, guidance, z, z_step_stride, coef, const_cast<int32_t*>(t_dev), num_timesteps, seed_ptr,
                                   nullptr, stream);
}

extern "C" int td_psample_step_cfg_advance(float* x, const float* eps, int64_t n, float guidance, const float* z,
                                           int64_t z_step_stride, const float* coef, int32_t* t_dev, int num_timesteps,
                                           const uint64_t* seed_ptr, unsigned int* ticket, void* stream) {
    TD_CHECK_ARG(ticket, "td_psample_step_cfg_advance: null ticket");
    return psample_step_cfg_launch(x, eps, n, guidance, z, z_step_stride, coef, t_dev, num_timesteps, seed_ptr, ticket, stream);
}

extern "C" int64_t td_grad_clip_num_partials(int64_t n) { return ew_grid(std::max<int64_t>(n / 4, 1)); }

extern "C" int td_grad_clip_scale(const float* g, int64_t n, float pre_scale, float max_norm, float* partials,
                                  unsigned int* counter, float* scale_out, float* norm_out, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(g && partials && counter && scale_out, "td_grad_clip_scale: null pointer");
    TD_CHECK_ARG(n > 0 && ((((uintptr_t)g) & 15) == 0), "td_grad_clip_scale: n must be positive and g 16-byte aligned");
    td::launch(grad_clip_kernel, td::LaunchCfg((int)td_grad_clip_num_partials(n), kEwThreads, 0, (cudaStream_t)stream),
               g, n, pre_scale, max_norm, partials, counter, scale_out, norm_out);
    return launch_status("grad_clip_scale");
}

extern "C" int td_randint(int64_t* out, int64_t n, int low, int high, const uint64_t* seed_ptr, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(out && seed_ptr, "td_randint: null pointer");
    TD_CHECK_ARG(n > 0 && high > low, "td_randint: need n > 0 and high > low");
    td::launch(randint_kernel, td::LaunchCfg(ew_grid((n + 3) / 4), kEwThreads, 0, (cudaStream_t)stream), out, n, low, high, seed_ptr);
    return launch_status("randint");
}

extern "C" int td_fill_f32(float* dst, int64_t n, float value, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(dst && n > 0, "td_fill_f32: bad arguments");
    td::launch(fill_f32_kernel, td::LaunchCfg(ew_grid(n), kEwThreads, 0, (cudaStream_t)stream), dst, n, value);
    return launch_status("fill_f32");
}

extern "C" int td_seed_advance(uint64_t* seed_ptr, int64_t delta, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(seed_ptr, "td_seed_advance: null pointer");
    td::launch(seed_advance_kernel, td::LaunchCfg(1, 1, 0, (cudaStream_t)stream), seed_ptr, (uint64_t)delta);
    return launch_status("seed_advance");
}

extern "C" int td_counter_add(int32_t* t_dev, int32_t delta, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(t_dev, "td_counter_add: null pointer");
    td::launch(counter_add_kernel, td::LaunchCfg(1, 1, 0, (cudaStream_t)stream), t_dev, delta);
    return launch_status("counter_add");
}

extern "C" int td_adam_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                             const int64_t* numel, const int32_t* chunk_tensor, const int64_t* chunk_offset,
                             int64_t num_chunks, int64_t chunk_elems, const int32_t* step_dev, float lr,
                             const float* lr_dev, double beta1, double beta2, float eps, const float* grad_scale_dev,
                             void* const* bf16_shadow, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(p && g && m && v && numel && chunk_tensor && chunk_offset && step_dev, "td_adam_multi: null pointer");
    TD_CHECK_ARG(chunk_elems > 0 && chunk_elems % 4 == 0, "td_adam_multi: chunk_elems must be a positive multiple of 4");
    if (num_chunks == 0) return TD_OK;
    td::launch(adam_multi_kernel, td::LaunchCfg((unsigned)num_chunks, kEwThreads, 0, (cudaStream_t)stream), 
        p, g, m, v, numel, chunk_tensor, chunk_offset, chunk_elems, step_dev, lr, lr_dev, beta1, beta2, eps,
        grad_scale_dev, bf16_shadow);
    return launch_status("adam_multi");
}

extern "C" int td_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
    TD_REQUIRE_ARCH();
    TD_CHECK_ARG(src && dst && n >= 0, "td_cast_f32_to_bf16: bad args");
    if (n == 0) return TD_OK;
    td::launch(cast_bf16_kernel, td::LaunchCfg(ew_grid(n), kEwThreads, 0, (cudaStream_t)stream), src, (__nv_bfloat16*)dst, n);
    return launch_status("cast_bf16");
}
