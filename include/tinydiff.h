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
/* Programmatic dependent launch (every kernel's launch latency and on-chip prologue overlap the tail of the preceding kernel
 * of the stream; on by default, TD_PDL=0 in the environment turns it off).  Returns the previous setting. */
int td_set_pdl(int on);
/* SMs the one-CTA-per-SM kernels (persistent halo convolution, tcgen05 weight gradient) size their grids for; plans read it
 * when they are created.  Default 148.  The data-parallel train step sets 148 - (NCCL CTAs) so that a grid never runs as
 * 140 CTAs + a second wave of 8 while an all-reduce holds SMs.  Returns the previous value. */
int td_set_sm_budget(int sms);
/* Number of kernels this library has launched (or captured into a graph) in this process so far; bench.py reads the
 * difference over one eager step to report `gpu_launches`. */
int64_t td_launch_count(void);

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
 * A step counter outside [0, num_timesteps) makes the launch a no-op (a graph replayed past t = 0).
 * Noise rows that are not 16-byte aligned (z_step_stride % 4 != 0) take the scalar path.
 * 16 algorithmic bytes per element (12 at t == 0). */
int td_psample_step(float* x, const float* eps, const float* z, int64_t z_step_stride,
                    const float* coef, const int32_t* t_dev, int64_t n, int num_timesteps,
                    const uint64_t* seed_ptr, void* stream);
/* Classifier-free-guidance reverse step -- an EXTENSION: the reference has no guidance, BASELINE.json's
 * config 2 names it.  x and eps hold a doubled batch, rows [0, n) conditional and rows [n, 2n) null-label:
 *   e = eps_u + guidance*(eps_c - eps_u);  x <- c1[t]*(x - c2[t]*e) + c3[t]*z, stored to both halves.
 * n = elements of ONE half (multiple of 4); z / Philox indexing as td_psample_step over n elements. */
int td_psample_step_cfg(float* x, const float* eps, int64_t n, float guidance, const float* z, int64_t z_step_stride,
                        const float* coef, const int32_t* t_dev, int num_timesteps, const uint64_t* seed_ptr, void* stream);
/* The same steps with the counter update folded in: the last block of the grid to finish writes t_dev[0] = t - 1 (every block has
 * read t by then), so a reverse step needs no td_counter_add launch.  ticket: one zero-initialised uint32 (left at zero). */
int td_psample_step_advance(float* x, const float* eps, const float* z, int64_t z_step_stride, const float* coef, int32_t* t_dev,
                            int64_t n, int num_timesteps, const uint64_t* seed_ptr, unsigned int* ticket, void* stream);
int td_psample_step_cfg_advance(float* x, const float* eps, int64_t n, float guidance, const float* z, int64_t z_step_stride,
                                const float* coef, int32_t* t_dev, int num_timesteps, const uint64_t* seed_ptr, unsigned int* ticket,
                                void* stream);
/* t_dev[0] += delta ; used between captured steps. */
int td_counter_add(int32_t* t_dev, int32_t delta, void* stream);

/* torch.optim.Adam(lr, betas=(b1,b2), eps) single fused multi-tensor step, diffusion.py:211,236.
 * Tensors are described by a device-resident table: for chunk c (one CTA's worth of work)
 * chunk_tensor[c] / chunk_offset[c] give the tensor id and element offset; per tensor the four
 * pointer arrays give p, g, m, v.  `step_dev[0]` is the 1-based step (read on the device so the
 * launch can be graph-captured); the learning rate is lr_dev[0] when lr_dev is non-NULL (a device
 * scalar, so CosineAnnealingLR -- diffusion_transformer.py:176-177,288; conditional_diffusion_laion.py:
 * 434-438,473 -- can drive a captured step), else `lr`; grad_scale multiplies g first (1/world_size
 * and / or the clip factor of td_grad_clip_scale, read from grad_scale_dev[0] if non-NULL).  If `bf16_shadow` is non-NULL, each updated
 * parameter is also written as bf16 at bf16_shadow[tensor][i] (packed operand copy for the
 * tcgen05 convolutions).  The betas are doubles: bias corrections, step size and (1 - beta) are formed
 * in double and rounded to fp32 once, as torch does with its Python-float betas.  28 algorithmic bytes
 * per parameter. */
int td_adam_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                  const int64_t* numel, const int32_t* chunk_tensor, const int64_t* chunk_offset,
                  int64_t num_chunks, int64_t chunk_elems, const int32_t* step_dev, float lr,
                  const float* lr_dev, double beta1, double beta2, float eps, const float* grad_scale_dev,
                  void* const* bf16_shadow, void* stream);

/* torch.nn.utils.clip_grad_norm_(parameters, max_norm), conditional_diffusion_laion.py:471, over one flat
 * gradient buffer g[0, n) (padding must be zero):
 *   norm_out[0] = pre_scale * ||g||_2 ;  scale_out[0] = pre_scale * min(1, max_norm / (norm + 1e-6))
 * pre_scale = 1/world_size when g holds the all-reduced SUM of the ranks' gradients, else 1.  scale_out feeds
 * td_adam_multi's grad_scale_dev, so the clipped gradient is never materialised.  max_norm <= 0: no clipping.
 * Deterministic two-stage reduction like td_mse_grad (partials: td_grad_clip_num_partials(n) floats; counter: one
 * zero-initialised uint32 the kernel resets).  4 algorithmic bytes per element. */
int64_t td_grad_clip_num_partials(int64_t n);
int td_grad_clip_scale(const float* g, int64_t n, float pre_scale, float max_norm, float* partials,
                       unsigned int* counter, float* scale_out, float* norm_out, void* stream);

/* t = torch.randint(low, high, (n,)) of the train step (diffusion.py:220) drawn on the device so that the whole
 * step is graph-capturable: Philox4x32-10 keyed by seed_ptr[0], subsequence seed_ptr[1] | 2^63 (disjoint from the
 * noise td_qsample draws with the same pair).  td_seed_advance bumps seed_ptr[1] between steps. */
int td_randint(int64_t* out, int64_t n, int low, int high, const uint64_t* seed_ptr, void* stream);
int td_seed_advance(uint64_t* seed_ptr, int64_t delta, void* stream);
/* dst[0, n) = value (device scalars such as the learning rate td_adam_multi reads through lr_dev). */
int td_fill_f32(float* dst, int64_t n, float value, void* stream);

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
    const float* w0;         /* [D, Din]   Din = 1 (modes 0, 1) or D (mode 2) */
    const float* b0;         /* [D] */
    const float* w2;         /* [D, D] */
    const float* b2;         /* [D] */
    const int64_t* y;        /* class labels [batch] or NULL */
    const float* class_table;/* [num_classes, D] or NULL */
    const float* text;       /* [batch, D] additive embedding or NULL */
    const float* proj_w;     /* [proj_out, D] */
    const float* proj_b;     /* [proj_out] */
    float* saved;            /* td_embed_head_saved_floats() floats, consecutive row-major blocks:      */
                             /* feat [B, Din] | h_pre [B, D] | h = silu(h_pre) [B, D] | emb [B, D]      */
    float* proj_out_ptr;     /* [batch, proj_out] */
} td_embed_args;
/* gradients of the head's parameters given d_proj = dL/d(proj_out_ptr); `a` must be the forward's
 * argument block (same `saved`).  scratch: 2*batch*dim floats. */
typedef struct {
    const float* d_proj;     /* [batch, proj_out] */
    float* scratch;
    float* d_w0; float* d_b0; float* d_w2; float* d_b2;
    float* d_class_table;    /* [num_classes, D] or NULL */
    int num_classes;
    float* d_proj_w; float* d_proj_b;
} td_embed_grads;
int64_t td_embed_head_saved_floats(int batch, int dim, int in_mode);
int td_embed_head_fwd(const td_embed_args* a, void* stream);
int td_embed_head_bwd(const td_embed_args* a, const td_embed_grads* g, void* stream);

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
    float* stats;            /* NULL, or train-mode BatchNorm partial sums emitted by the tcgen05 epilogue:  */
                             /* [td_conv3x3_stats_rows(plan)][2][cout] { sum, sum of squares } of the fp32   */
                             /* outputs, followed by one zero row [cout] -- the layout td_bn_finalize reads  */
    int x_nchw, y_nchw;      /* direct kernels only: tensor is NCHW fp32 (network input/output) */
    float* splitk_ws;        /* NULL, or td_conv3x3_splitk_workspace() floats: lets the tcgen05 engine   */
                             /* split the K loop across CTAs when the layer has too few tiles            */
    void* pool_y;            /* NULL, or MaxPool2d(2, ceil_mode = pool_ceil) of y (diffusion.py:101), same dtype, dense   */
                             /* [batch, Hp, Wp, cout]: written by the epilogue of the tcgen05 halo kernel when the plan   */
                             /* accepts it (td_conv3x3_pool_fused); otherwise the caller runs td_maxpool2_fwd on y        */
    int pool_ceil;
} td_conv3x3_desc;

/* Engines.  Weight layout expected by each (w is [cout][3][3][cin] "OHWI" in every case):
 *   TD_CONV_SIMT : fp32 OHWI, any cin/cout; fp32 FFMA implicit GEMM (the fp32 parity path)
 *   TD_CONV_TC   : bf16 OHWI, cin % 64 == 0, cout % 64 == 0 ; TMA -> smem -> tcgen05.mma -> TMEM
 *   TD_CONV_DIRECT : fp32 OHWI, small cin (<=8) or small cout (<=8); bandwidth kernel        */
#define TD_CONV_SIMT 0
#define TD_CONV_TC 1
#define TD_CONV_DIRECT 2

typedef struct td_conv_plan td_conv_plan;
/* floats of split-K workspace the tcgen05 engine may use for this layer (0: no split) */
int64_t td_conv3x3_splitk_workspace(const td_conv3x3_desc* desc);
int td_conv3x3_plan_create(td_conv_plan** plan, const td_conv3x3_desc* desc, int engine);
int td_conv3x3_run(const td_conv_plan* plan, void* stream);
/* 1: the plan writes desc.pool_y itself (eval mode, bf16 output, halo kernel with pooling windows inside a subtile) */
int td_conv3x3_pool_fused(const td_conv_plan* plan);
/* number of partial rows the plan writes to desc.stats per run (0: statistics are not fused, use td_bn_stats) */
int td_conv3x3_stats_rows(const td_conv_plan* plan);
void td_conv3x3_plan_destroy(td_conv_plan* plan);
/* algorithmic FLOPs (2*MAC) of one run of the plan */
double td_conv3x3_flops(const td_conv_plan* plan);
/* tuning aid: per-CTA wait-cycle counters [148][8] of the last halo-kernel launch run with TD_TC_HALO_DBG=1 */
int td_conv3x3_debug_counters(unsigned long long* host_out, int n);

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

/* Network tail, eval path: F.interpolate(x, size=(ho,wo), bilinear, align_corners=True) followed by final_conv
 * (3x3, c -> 1 channel, + bias), diffusion.py:157-160, in one pass over x (the resized c-channel tensor is never
 * written).  x: NHWC [batch, hi, wi, ldx] channels [x_coff, x_coff + c); w_ohwi: fp32 [1][3][3][c]; y: fp32 NCHW. */
int td_final_resize_conv(const void* x, int dtype, int ldx, int x_coff, int batch, int hi, int wi, int c,
                         const float* w_ohwi, const float* bias, int ho, int wo, float* y_nchw, void* stream);

/* Eval-mode BatchNorm folded into the conv epilogue (diffusion.py:34 in eval mode):
 *   scale = gamma / sqrt(running_var + eps) ; shift = beta + (conv_bias - running_mean) * scale */
int td_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
               const float* conv_bias, float eps, float* scale, float* shift, int c, void* stream);

/* Every bf16 conv layer of a model re-packed in one launch (after an optimizer step).  table: device array of
 *   struct { const float* src_oihw; bf16* dst_ohwi; bf16* dst_dgrad_or_NULL; int cout, cin, tile_begin, pad; }
 * (40 bytes per entry, cout and cin multiples of 32, tile_begin = running sum of (cout/32)*(cin/32)). */
int td_pack_conv_weights_multi(const void* table, int n_entries, int total_tiles, void* stream);

/* dtype conversion / weight packing helpers */
int td_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* OIHW fp32 (PyTorch conv weight) -> OHWI fp32 or bf16 */
int td_pack_conv_weight(const float* oihw, void* ohwi, int out_dtype, int cout, int cin,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Training: data / weight gradients of the 3x3 convolutions (autograd of diffusion.py:28-98).
 * The data gradient is td_conv3x3_run on the flipped-transposed weights
 * (td_pack_conv_weight_dgrad); the weight gradient is the plan below.
 * ---------------------------------------------------------------------------------------- */
/* OIHW fp32 -> [cin][3][3][cout] (fp32 / bf16), taps flipped: dX = conv3x3(dY, Wd) */
int td_pack_conv_weight_dgrad(const float* oihw, void* out, int out_dtype, int cout, int cin, void* stream);

typedef struct {
    int batch, height, width;
    int cin, cout;
    int x_dtype, dy_dtype;
    const void* x;           /* conv input  [batch, height, width, ldx] channels [x_coff, x_coff+cin)   */
    int ldx, x_coff, x_nchw; /* x_nchw: fp32 NCHW network input (SIMT engine only)                      */
    const void* dy;          /* grad of conv output [batch, height, width, lddy] at dy_coff             */
    int lddy, dy_coff, dy_nchw;
    float* dw;               /* OIHW fp32 [cout][cin][3][3]: PyTorch's .grad layout (overwritten)       */
    float* workspace;        /* td_conv3x3_wgrad_workspace() floats of split-K partials                 */
} td_wgrad_desc;
typedef struct td_wgrad_plan td_wgrad_plan;
int64_t td_conv3x3_wgrad_workspace(const td_wgrad_desc* desc, int engine);
int td_conv3x3_wgrad_plan_create(td_wgrad_plan** plan, const td_wgrad_desc* desc, int engine); /* TD_CONV_SIMT | TD_CONV_TC */
int td_conv3x3_wgrad_run(const td_wgrad_plan* plan, void* stream);
void td_conv3x3_wgrad_plan_destroy(td_wgrad_plan* plan);

/* ------------------------------------------------------------------------------------------
 * Training: BatchNorm2d (+ReLU) in train mode and its backward, nn.BatchNorm2d/nn.ReLU at
 * diffusion.py:34-35.  NHWC, `pixels` = batch*height*width rows.  Reductions write one row of
 * [2][channels] partial sums per CTA (td_chan_reduce_rows rows) and are finalised in fixed order.
 * ---------------------------------------------------------------------------------------- */
int td_chan_reduce_rows(int dtype, int64_t pixels, int channels);
/* partials[r] = { sum (x-K), sum (x-K)^2 } of x (the raw fp32 conv output, conv bias NOT included, for
 * BatchNorm; any NHWC tensor for per-channel sums), followed by the row K[channels]: K = x[pixel 0]
 * when `shifted` (guards E[x^2]-mean^2 against cancellation), else 0.
 * partials holds td_chan_reduce_rows()*2*channels + channels floats. */
int td_bn_stats(const void* x, int dtype, int64_t ldx, int x_coff, int64_t pixels, int channels, float* partials,
                int shifted, void* stream);
/* batch mean / biased variance -> scale = gamma*invstd, shift = beta - mean*scale; running stats
 * momentum update with the unbiased variance and mean + conv_bias; num_batches_tracked += 1 */
int td_bn_finalize(const float* partials, int nrows, int channels, int64_t count, const float* gamma,
                   const float* beta, const float* conv_bias, float eps, float momentum, float* running_mean,
                   float* running_var, int64_t* num_batches_tracked, float* scale, float* shift, float* save_mean,
                   float* save_invstd, void* stream);
/* a[p, a_coff + c] = relu?(y[p, c] * scale[c] + shift[c]);  y: fp32 [pixels, channels]; a: `dtype` */
int td_bn_relu_apply(const float* y, const float* scale, const float* shift, void* a, int dtype, int64_t lda,
                     int a_coff, int64_t pixels, int channels, int relu, void* stream);
/* partials[r] = { sum g, sum g*(y-mean) },  g = da * [y*scale+shift > 0]   (da: `dtype`, y: fp32) */
int td_bn_relu_bwd_reduce(const void* da, int64_t ldda, int da_coff, const float* y, int dtype, const float* scale,
                          const float* shift, const float* save_mean, int64_t pixels, int channels, float* partials,
                          void* stream);
/* dgamma, dbeta and the per-channel coefficients coef[3][channels] of the apply pass */
int td_bn_bwd_finalize(const float* partials, int nrows, int channels, int64_t count, const float* scale,
                       const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* coef,
                       void* stream);
/* dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) */
int td_bn_relu_bwd_apply(const void* da, int64_t ldda, int da_coff, const float* y, int dtype, const float* scale,
                         const float* shift, const float* coef, void* dy, int64_t pixels, int channels, void* stream);

/* Fused BatchNorm passes (train mode): the finalize runs in the prologue of the streaming kernel that consumes it, so a
 * BatchNorm2d + ReLU is conv -> td_bn_apply_fused (forward) and td_bn_bwd_reduce -> td_bn_bwd_apply_fused (backward), with
 * no finalize launch in between.  partials: [nrows][2][channels] + K[channels] as produced by a convolution's `stats`
 * epilogue, td_bn_stats or td_bn_bwd_reduce.  channels % 32 == 0.
 *   td_bn_apply_fused      = td_bn_finalize + td_bn_relu_apply   (also writes scale / shift / save_mean / save_invstd and
 *                            updates the running statistics, diffusion.py:34-35)
 *   td_bn_bwd_reduce       = td_bn_relu_bwd_reduce with one partial row per pixel chunk (td_bn_bwd_reduce_rows() <= 148)
 *   td_bn_bwd_apply_fused  = td_bn_bwd_finalize + td_bn_relu_bwd_apply (writes dgamma / dbeta)
 * y (the raw conv output) is fp32, or bf16 (y_dtype = TD_BF16, with bf16 activations) on layers whose outputs carry no large
 * per-sample offsets -- every layer except the three that consume a decoder concat.
 * The three evaluate the pre-activation as scale * (y - mean) + beta (beta = the BatchNorm bias parameter): under the raw-t
 * time embedding y carries per-channel offsets of O(1e2..1e3), and y * scale + shift cancels in fp32. */
int td_bn_apply_fused(const void* y, int y_dtype, const float* partials, int nrows, int64_t count, const float* gamma,
                      const float* beta, const float* conv_bias, float eps, float momentum, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, float* scale, float* shift, float* save_mean,
                      float* save_invstd, void* a, int dtype, int64_t lda, int a_coff, int64_t pixels, int channels,
                      int relu, void* stream);
int td_bn_bwd_reduce_rows(int dtype, int64_t pixels, int channels);
int td_bn_bwd_reduce(const void* da, int64_t ldda, int da_coff, const void* y, int y_dtype, int dtype, const float* scale,
                     const float* beta, const float* save_mean, int64_t pixels, int channels, float* partials,
                     void* stream);
int td_bn_bwd_apply_fused(const void* da, int64_t ldda, int da_coff, const void* y, int y_dtype, int dtype, const float* partials,
                          int nrows, int64_t count, const float* scale, const float* beta, const float* save_mean,
                          const float* save_invstd, float* dgamma, float* dbeta, void* dy, int64_t pixels, int channels,
                          void* stream);

/* backward of td_maxpool2_fwd (first maximum in scan order takes the gradient, like ATen) */
int td_maxpool2_bwd(const void* x, const void* dy, void* dx, int dtype, int batch, int h, int w, int c, int ceil_mode,
                    int accumulate, void* stream);
/* backward of td_resize_bilinear_fwd: dx[b,hi,wi,:] from dy [b,ho,wo,ld_dy] at channel dy_coff (gather form) */
int td_resize_bilinear_bwd(const void* dy, int64_t ld_dy, int dy_coff, void* dx, int dtype, int batch, int hi, int wi,
                           int ho, int wo, int c, void* stream);
/* backward of td_upcat_fwd: dlow [b,ho/2,wo/2,cu], dskip [b,hs,ws,cs] (overwritten) and
 * dtemb[b, temb_off + c] = sum_{h,w} dout[b,h,w,cu+c] */
int td_upcat_bwd(const void* dout, void* dlow, void* dskip, float* dtemb, int ld_temb, int temb_off, int dtype,
                 int batch, int ho, int wo, int cu, int hs, int ws, int cs, void* stream);
/* out[c] = sum_r partials[r][which][c] */
int td_partial_sum(const float* partials, int nrows, int channels, int which, float* out, void* stream);
/* out[c] = sum_{b,hw} x[b,c,hw]  (final_conv bias gradient); workspace: 128*channels floats */
int td_nchw_chansum(const float* x, int batch, int channels, int hw, float* out, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32 dense layers: nn.Linear / nn.LayerNorm / nn.BatchNorm1d / nn.GELU / nn.SiLU of the
 * conditioning head (diffusion.py:21-25), the latent MLP denoiser (latent_diffusion.py:24-105) and
 * the DiT blocks (diffusion_transformer.py:16-109).
 * ---------------------------------------------------------------------------------------- */
#define TD_ACT_NONE 0
#define TD_ACT_RELU 1
#define TD_ACT_SILU 2
#define TD_ACT_GELU 3      /* exact (erf) GELU, nn.GELU() default */
#define TD_ACT_SIGMOID 4
/* C[i,j] = epilogue( alpha * sum_k A(i,k) * B(k,j) ),  A(i,k) = A[i*a_rs + k*a_cs], B(k,j) = B[k*b_rs + j*b_cs]
 * epilogue(v): v += bias[j]; pre_out[i,j] = v; v = act(v); v += residual[i,j]; v += gather_table[gather_idx[i], j];
 *              if (accumulate) v += C[i,j]
 * nn.Linear forward y = x W^T + b: A = x (a_rs=K, a_cs=1), B = W [N,K] (b_rs=1, b_cs=K). */
typedef struct {
    int M, N, K;
    const float* A; int64_t a_rs, a_cs;
    const float* B; int64_t b_rs, b_cs;
    float* C; int64_t ldc;
    float alpha;
    const float* bias;
    int act;
    float* pre_out; int64_t ld_pre;
    const float* residual; int64_t ldr;
    const int64_t* gather_idx; const float* gather_table; int64_t ld_table;
    int accumulate;
    float* splitk_ws;        /* NULL, or td_gemm_f32_workspace(M,N,K) floats: deterministic split-K */
    int allow_tf32;          /* 0: exact fp32 products (FFMA kernels); 1: the caller accepts 10-bit operand mantissas, the
                              * tcgen05 kind::tf32 kernel runs when the operands qualify (see td_gemm_f32_path) */
} td_gemm_args;
int64_t td_gemm_f32_workspace(int M, int N, int K);
int td_gemm_f32(const td_gemm_args* a, void* stream);
/* Which kernel family td_gemm_f32 would run for these arguments: 0 = fp32 FFMA (exact fp32 products), 1 = tcgen05 kind::tf32
 * (allow_tf32 set, both operands K-major, 16-byte aligned rows, M, N, K >= 32; ~1e-3 relative; TD_GEMM_TF32=0 in the
 * environment keeps everything on path 0). */
int td_gemm_f32_path(const td_gemm_args* a);
int td_colsum_f32(const float* x, int64_t ldx, float* out, int M, int N, int accumulate, void* stream);
int td_act_bwd_f32(const float* dy, const float* pre, float* dx, int64_t n, int act, void* stream);
/* dst[i,j] (+)= src[i,j] on strided fp32 matrices (residual-branch gradients) */
int td_add2d_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int accumulate,
                 void* stream);
/* dst[j, i] = src[i, j] on strided fp32 matrices (rows <= 2^21) */
int td_transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream);
/* nn.Dropout / attention-weight dropout at L=1 (diffusion_transformer.py:19,27,29): out = x*keep/(1-p), one
 * Bernoulli draw per `group` consecutive columns; Philox keyed by seed_ptr[0], subsequence seed_ptr[1].
 * Applying the same call to a gradient is the backward. */
int td_dropout_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int rows, int cols, int group, float p,
                   const uint64_t* seed_ptr, void* stream);
/* table_grad[c,:] = sum_{i: idx[i]==c} g[i,:]   (nn.Embedding backward, conditional_diffusion.py:31).  With a workspace
 * (>= 2*num_rows*D floats; more = more row chunks, up to 296) and M >= 4096 the rows are split over many CTAs (table-shaped
 * partial sums, fixed-order finalize); workspace NULL: one CTA per table row. */
int td_embedding_bwd(const float* g, int64_t ldg, const int64_t* idx, float* table_grad, int M, int D, int num_rows,
                     int accumulate, float* workspace, int64_t workspace_floats, void* stream);
/* out[b, :] = t (mode 0), t/1000 (mode 1) [width 1] or the sinusoidal embedding [width dim] (mode 2) */
int td_time_features(const int64_t* t, const int32_t* t_dev, float* out, int batch, int dim, int mode, void* stream);
int td_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M,
                     int D, float eps, void* stream);
/* workspace (optional, >= 4*D floats; M >= 4096): dgamma / dbeta from many-CTA partial sums + fixed-order finalize */
int td_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                     float* dx, float* dgamma, float* dbeta, int M, int D, float* workspace, int64_t workspace_floats,
                     void* stream);
int td_bn1d_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float* save_mean, float* save_rstd, float* y, int64_t ldy, int M, int N, float eps,
                float momentum, int training, int relu, void* stream);
int td_bn1d_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* y_out, int64_t ldy,
                const float* gamma, const float* save_mean, const float* save_rstd, float* dx, int64_t lddx,
                float* dgamma, float* dbeta, int M, int N, int relu, void* stream);

/* ------------------------------------------------------------------------------------------
 * The conv / attention VAE of the reference's vae_laion.py, fp32, NHWC activations.
 * ---------------------------------------------------------------------------------------- */

/* torch.nn.utils.spectral_norm (vae_laion.py:72-78,98-131,138-165): W_mat = the weight with `dim` first, flattened to
 * [rows][cols] (dim1 = 0: plain Conv2d weight; dim1 = 1: ConvTranspose2d weight (Cin, Cout, kh, kw), rows = Cout,
 * cols = Cin*khw).  Runs `power_iterations` updates v = normalize(W^T u), u = normalize(W v) in place (1 in a training-mode
 * forward, 0 in eval mode), then sigma_out[0] = u^T W v.  scratch_rows: `rows` floats. */
int td_spectral_sigma(const float* w, int rows, int cols, int dim1, int khw, float* u, float* v, int power_iterations,
                      float eps, float* sigma_out, float* scratch_rows, void* stream);
/* x[0, n) /= sigma_dev[0]  (folds 1/sigma of a spectral-norm 3x3 convolution into its BatchNorm scale) */
int td_scale_by_inv_sigma(float* x, const float* sigma_dev, int n, void* stream);
/* Operand packing of the 4x4 stride-2 layers, multiplied by 1 / sigma_dev[0] (sigma_dev may be NULL):
 * transposed = 0: Conv2d weight (Cout, Cin, 4, 4)          -> [Cout][(ky*4+kx)*Cin + ci]
 * transposed = 1: ConvTranspose2d weight (Cin, Cout, 4, 4) -> [4 output-parity classes][Cout][(j*2+i)*Cin + ci]. */
int td_pack_conv4x4_weight(const float* w, const float* sigma_dev, float* out, int cout, int cin, int transposed,
                           void* stream);
/* nn.Conv2d(cin, cout, 4, stride=2, padding=1) (+ bias, TD_ACT_*), vae_laion.py:98-131.  x: NHWC (or NCHW when x_nchw);
 * y: NHWC [batch][hin/2][win/2][cout]. */
int td_conv4x4s2_fwd(const float* x, const float* w_packed, const float* bias, float* y, int batch, int hin, int win, int cin,
                     int cout, int x_nchw, int act, void* stream);
/* nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1) (+ bias, TD_ACT_*), vae_laion.py:138-165.  x: NHWC; y: NHWC
 * [batch][2*hin][2*win][cout] (or NCHW when y_nchw). */
int td_convT4x4s2_fwd(const float* x, const float* w_packed, const float* bias, float* y, int batch, int hin, int win, int cin,
                      int cout, int y_nchw, int act, void* stream);
/* SelfAttention.forward (vae_laion.py:57-65), flash-style: y = gamma * softmax(Q K^T) V + x with the softmax over all n = H*W
 * keys (no scaling), never materialising the n x n matrix.  qkv: [batch][n][2*dq + dv] rows (q | k | v) of the fused 1x1
 * convolutions; x, y: NHWC [batch][n][dv].  (dq, dv) in {(4, 32), (8, 64), (16, 128)}; n % 128 == 0. */
int td_self_attention_fwd(const float* qkv, const float* x, const float* gamma, float* y, int batch, int n, int dq, int dv,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused eval-mode forward of the dense denoisers (latent_diffusion.py:107-128, diffusion_transformer.py:81-109) at the
 * reference batch sizes: one persistent kernel walks a device-resident tape of ops (Linear with bias / eval BatchNorm1d +
 * ReLU / activation / residual / embedding gather in the epilogue, LayerNorm, add, time features) with grid barriers
 * between dependent ops.  ops: `n_ops` records of td_dense_tape_op_bytes() bytes (layout: csrc/dense_fused.cu TapeOp);
 * barrier: one zero-initialised uint64 arrival counter (never reset; launches that share it must use the same max_ctas).
 * max_ctas: 0 = default (<= 128, all co-resident).
 * ---------------------------------------------------------------------------------------- */
int td_dense_tape_op_bytes(void);
int td_dense_tape_run(const void* ops, int n_ops, int batch, unsigned long long* barrier, int max_ctas, void* stream);

/* The same forward with thread-block clusters instead of a grid barrier (csrc/dense_cluster.cu): a cluster of 8 CTAs takes 8
 * rows of the batch through the whole tape, activations replicated in the CTAs' shared memory (st.shared::cluster pushes +
 * barrier.cluster), weight slices streamed from the L2 through a cp.async ring that runs ahead across ops.  ops: `n_ops`
 * (<= max_ops) records of td_dense_cluster_op_bytes() bytes (layout: ClusterOp; activation operands are float offsets into
 * the arena of `arena_floats` floats per CTA, assigned by the host for `rows` rows per cluster: td_dense_cluster_limits picks
 * 8, or 9 when that saves a pass -- 128 rows on the 15 clusters a B200 keeps resident).  max_clusters: 0 = default. */
int td_dense_cluster_op_bytes(void);
int td_dense_cluster_limits(int batch, int* rows, int* cluster, int* arena_floats, int* stage_floats, int* max_ops,
                            int* max_clusters);
int td_dense_cluster_run(const void* ops, int n_ops, int batch, int rows, int max_clusters, void* stream);
/* One reverse step of the sampler in ONE launch (multi-group kernel only: td_dense_cluster_step_fused() == 1): the tape -- whose
 * last op must be the Linear that writes eps [batch][n], n % 4 == 0 -- and, in that Linear's epilogue, the update of
 * td_psample_step_advance on the sampler state x [batch][n] (the buffer the tape loads): same roundings, noise-table rows and
 * Philox indexing; the last CTA of the grid to finish writes t_dev[0] = t - 1 (ticket: one zero-initialised uint32). */
int td_dense_cluster_step_fused(void);
int td_dense_cluster_step(const void* ops, int n_ops, int batch, int rows, int max_clusters, float* x, const float* coef,
                          const float* z, int64_t z_step_stride, const uint64_t* seed_ptr, int32_t* t_dev, unsigned int* ticket,
                          int num_timesteps, void* stream);
/* 128-byte tensor map (written to HOST memory) of one Linear's fp32 weight [n][k] for the kernel's weight stream: TMA boxes of
 * 32 floats x fcp rows, SWIZZLE_128B.  The caller keeps a 64-byte aligned device copy and points ClusterOp::tmap at it. */
int td_dense_cluster_weight_map(const float* w, int n, int k, int fcp, void* map_out_host);
/* tuning aid: 16 counters per CTA of the last launch made with TD_DENSE_CLUSTER_DBG=1 (globaltimer start / end, clock64 cycles
 * in weight waits, FFMA chunks, output pushes, cluster barriers, row-wise ops, epilogue-operand requests, partial stores + CTA
 * barrier, k-slice sum + epilogue math, 6 unused); synchronises the device */
int td_dense_cluster_debug_counters(unsigned long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* TINYDIFF_H_ */
