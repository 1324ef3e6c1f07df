/*
 * tinydiff.h -- C ABI of libtinydiff.so: the B200 (sm_100a) kernels behind the DDPM hot path
 * of david-wb/tiny-diffusion.
 *
 * The reference has no FFI / plugin interface of its own: its hot path is a
 * sequence of PyTorch library calls.  Each entry point below replaces the library call(s) at
 * the cited reference lines.  Conventions, identical for every function:
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*;
 *   - returns 0 on success, <0 for an argument / architecture error, >0 = cudaError_t;
 *     td_last_error_string() describes the last failure on the calling thread;
 *   - never allocates or frees caller memory, never synchronises the stream, and is safe to
 *     call during CUDA-graph capture on `stream` (the only host-side state is a per-plan
 *     TMA descriptor block created by the *_plan_create calls);
 *   - all tensors are dense, activations are NHWC ("channels last") unless stated otherwise;
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every
 *     compute entry point fails with TD_ERR_ARCH.
 */
#ifndef TINYDIFF_H_
#define TINYDIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD_OK 0
#define TD_ERR_ARG (-1)
#define TD_ERR_ARCH (-2)
#define TD_ERR_UNSUPPORTED (-3)
#define TD_ERR_DRIVER (-4)

/* activation element types */
#define TD_F32 0
#define TD_BF16 1

int td_version(void);
const char* td_last_error_string(void);
/* 0 if `device` is an sm_100-class GPU this library can run on, TD_ERR_ARCH otherwise. */
int td_device_check(int device);

/* ------------------------------------------------------------------------------------------
 * Elementwise DDPM kernels (HBM-bound)
 * ---------------------------------------------------------------------------------------- */

/* ForwardProcess.q_sample, diffusion.py:177-190 (2-D variant latent_diffusion.py:143-154):
 *   x_t[b,i] = sqrt(abar[t[b]]) * x0[b,i] + sqrt(1 - abar[t[b]]) * noise[b,i]
 * `noise` is caller-provided (injected) when `seed_ptr == NULL`; otherwise it is drawn here
 * (Philox4x32-10 + Box-Muller keyed by seed_ptr[0], offset seed_ptr[1]) and WRITTEN to `noise`.
 * 12 algorithmic bytes per element with injected noise. */
int td_qsample(const float* x0, float* noise, const int64_t* t, const float* alphas_cumprod,
               float* x_t, int64_t batch, int64_t per_sample, int num_timesteps,
               const uint64_t* seed_ptr, void* stream);

/* F.mse_loss + the seed of its backward, diffusion.py:231,235:
 *   loss_out[0] = sum((pred-target)^2) * inv_n ;  grad[i] = 2*(pred[i]-target[i])*inv_n
 * `grad` may be NULL (validation pass).  Deterministic: per-block partials are written to
 * `partials` (>= td_mse_num_partials(n) floats) and the last block to finish (tracked through
 * `counter`, one zero-initialised uint32 that the kernel resets) reduces them in fixed order. */
int64_t td_mse_num_partials(int64_t n);
int td_mse_grad(const float* pred, const float* target, float* grad, float* loss_out,
                float* partials, unsigned int* counter, int64_t n, float inv_n, void* stream);

/* One reverse step, diffusion.py:263-274:
 *   x <- c1[t]*(x - c2[t]*eps) + c3[t]*z      (z = 0 at t == 0)
 * coef is a device table [num_timesteps][4] = {c1, c2, c3, 0} built on the host with the
 * reference's exact fp32 op order.  t is read from t_dev[0] (device-side step counter, so the
 * whole loop is CUDA-graph capturable).  z: injected noise; step t reads z + t*z_step_stride
 * (stride 0: one buffer refilled by the caller each step; stride n: a [T, n] table);
 * if z == NULL and seed_ptr != NULL the noise is drawn in-kernel (Philox, subsequence = t).
 * 16 algorithmic bytes per element (12 at t == 0). */
int td_psample_step(float* x, const float* eps, const float* z, int64_t z_step_stride,
                    const float* coef, const int32_t* t_dev, int64_t n, const uint64_t* seed_ptr,
                    void* stream);
/* t_dev[0] += delta ; used between captured steps. */
int td_counter_add(int32_t* t_dev, int32_t delta, void* stream);

/* torch.optim.Adam(lr, betas=(b1,b2), eps) single fused multi-tensor step, diffusion.py:211,236.
 * Tensors are described by a device-resident table: for chunk c (one CTA's worth of work)
 * chunk_tensor[c] / chunk_offset[c] give the tensor id and element offset; per tensor the four
 * pointer arrays give p, g, m, v.  `step_dev[0]` is the 1-based step (read on the device so the
 * launch can be graph-captured); grad_scale multiplies g first (1/world_size or clip factor,
 * read from grad_scale_dev[0] if non-NULL).  If `bf16_shadow` is non-NULL, each updated
 * parameter is also written as bf16 at bf16_shadow[tensor][i] (packed operand copy for the
 * tcgen05 convolutions).  28 algorithmic bytes per parameter. */
int td_adam_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                  const int64_t* numel, const int32_t* chunk_tensor, const int64_t* chunk_offset,
                  int64_t num_chunks, int64_t chunk_elems, const int32_t* step_dev, float lr,
                  float beta1, float beta2, float eps, const float* grad_scale_dev,
                  void* const* bf16_shadow, void* stream);

/* ------------------------------------------------------------------------------------------
 * Conditioning head: time MLP (+class / text embedding) + the three 1x1 "time_proj" convs.
 * diffusion.py:21-25,105-107,111-113,130-132; conditional_diffusion.py:31,117-125;
 * conditional_diffusion_laion.py:223-232,239-243,307-310.  fp32 throughout (raw t up to 999).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int batch;
    int dim;                 /* time_dim D */
    int in_mode;             /* 0: raw t (Linear(1,D)); 1: t/1000 (DiT); 2: sinusoidal(D) */
    int proj_out;            /* rows of proj_w (sum of the three time_proj Cout); 0 = none */
    const int64_t* t;        /* [batch] or NULL */
    const int32_t* t_dev;    /* used for every sample when t == NULL (sampler step counter) */
    const float* w0;         /* [D, Din] */
    const float* b0;         /* [D] */
    const float* w2;         /* [D, D] */
    const float* b2;         /* [D] */
    const int64_t* y;        /* class labels [batch] or NULL */
    const float* class_table;/* [num_classes, D] or NULL */
    const float* text;       /* [batch, D] additive embedding or NULL */
    const float* proj_w;     /* [proj_out, D] */
    const float* proj_b;     /* [proj_out] */
    float* emb_out;          /* [batch, D]  (combined embedding; needed by backward) or NULL */
    float* h_out;            /* [batch, D]  pre-SiLU hidden (needed by backward) or NULL */
    float* proj_out_ptr;     /* [batch, proj_out] */
} td_embed_args;
int td_embed_head_fwd(const td_embed_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3x3 / stride 1 / pad 1 convolutions, NHWC.  Replaces nn.Conv2d(.,.,3,padding=1) (+ the
 * eval-mode BatchNorm2d + ReLU that follow it), diffusion.py:28-98.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int batch, height, width;
    int cin, cout;
    int x_dtype, y_dtype;    /* TD_F32 / TD_BF16 */
    const void* x;           /* [batch, height, width, ldx] ; channels [x_coff, x_coff+cin) */
    int ldx, x_coff;
    void* y;                 /* [batch, height, width, ldy] ; channels [y_coff, y_coff+cout) */
    int ldy, y_coff;
    const void* w;           /* packed weights, see td_conv3x3_weight_layout */
    const float* scale;      /* [cout] or NULL (=1)  : y = relu?(acc*scale + shift)           */
    const float* shift;      /* [cout] or NULL (=0)                                         */
    int relu;
    float* stats;            /* NULL, or [2*cout]: += per-channel sum / sum-of-squares of   */
                             /* (acc*scale+shift) -- train-mode BatchNorm batch statistics  */
    int x_nchw, y_nchw;      /* direct kernels only: tensor is NCHW fp32 (network input/output) */
} td_conv3x3_desc;

/* Engines.  Weight layout expected by each (w is [cout][3][3][cin] "OHWI" in every case):
 *   TD_CONV_SIMT : fp32 OHWI, any cin/cout; fp32 FFMA implicit GEMM (the fp32 parity path)
 *   TD_CONV_TC   : bf16 OHWI, cin % 64 == 0, cout % 64 == 0 ; TMA -> smem -> tcgen05.mma -> TMEM
 *   TD_CONV_DIRECT : fp32 OHWI, small cin (<=8) or small cout (<=8); bandwidth kernel        */
#define TD_CONV_SIMT 0
#define TD_CONV_TC 1
#define TD_CONV_DIRECT 2

typedef struct td_conv_plan td_conv_plan;
int td_conv3x3_plan_create(td_conv_plan** plan, const td_conv3x3_desc* desc, int engine);
int td_conv3x3_run(const td_conv_plan* plan, void* stream);
void td_conv3x3_plan_destroy(td_conv_plan* plan);
/* algorithmic FLOPs (2*MAC) of one run of the plan */
double td_conv3x3_flops(const td_conv_plan* plan);

/* ------------------------------------------------------------------------------------------
 * NHWC glue kernels (HBM-bound)
 * ---------------------------------------------------------------------------------------- */
/* nn.MaxPool2d(2, ceil_mode) diffusion.py:101 ; conditional_diffusion_laion.py:302 */
int td_maxpool2_fwd(const void* x, void* y, int dtype, int batch, int h, int w, int c,
                    int ceil_mode, void* stream);

/* Decoder input assembly, diffusion.py:135-140 (and :142-154): writes
 *   out[b,h,w, 0:cu]      = Upsample(x2, bilinear, align_corners=True)(low)[b,h,w,:]
 *   out[b,h,w, cu:cu+cs]  = interpolate(skip + temb[b,:], size=(ho,wo), bilinear, align_corners=True)
 * (plain add when the skip already has the output size: conditional_diffusion_laion.py:323-331)
 * low: [b, ho/2, wo/2, cu]; skip: [b, hs, ws, cs]; temb: fp32 [b, ld_temb] at column temb_off. */
int td_upcat_fwd(const void* low, const void* skip, const float* temb, int ld_temb, int temb_off,
                 void* out, int dtype, int batch, int ho, int wo, int cu, int hs, int ws, int cs,
                 void* stream);

/* F.interpolate(x, size=(ho,wo), mode="bilinear", align_corners=True), diffusion.py:157-159 */
int td_resize_bilinear_fwd(const void* x, void* y, int dtype, int batch, int hi, int wi, int ho,
                           int wo, int c, void* stream);

/* Eval-mode BatchNorm folded into the conv epilogue (diffusion.py:34 in eval mode):
 *   scale = gamma / sqrt(running_var + eps) ; shift = beta + (conv_bias - running_mean) * scale */
int td_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
               const float* conv_bias, float eps, float* scale, float* shift, int c, void* stream);

/* dtype conversion / weight packing helpers */
int td_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* OIHW fp32 (PyTorch conv weight) -> OHWI fp32 or bf16 */
int td_pack_conv_weight(const float* oihw, void* ohwi, int out_dtype, int cout, int cin,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYDIFF_H_ */
