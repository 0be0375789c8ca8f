/*
 * dreamb200.h -- C-ABI of the B200-native DREAM hot path (libdreamb200.so).
 *
 * Every entry point replaces a torch/cuDNN/scipy library call made by the
 * reference on its belief-map path.  Citations are into NVlabs/DREAM:
 *
 *   dreamb200_conv2d_fwd        nn.Conv2d / nn.ConvTranspose2d (+bias +BN-fold +ReLU +residual)
 *                               dream/models.py:591-615,690-747 (vgg trunk/decoder/heads),
 *                               :22-32,:37-136 (resnet trunk + deconv head)
 *   dreamb200_im2col_first      first-layer patch gather feeding conv2d_fwd
 *                               (models.py:591-599 3x3 p1; torchvision resnet conv1 7x7 s2 p3, models.py:23)
 *   dreamb200_maxpool_nhwc      nn.MaxPool2d(2) models.py:589 ; resnet maxpool k3 s2 p1 models.py:27
 *   dreamb200_upsample2_nhwc    nn.Upsample(scale_factor=2) (nearest) models.py:691,703
 *   dreamb200_peaks             dream/image_proc.py:914-1018 peaks_from_belief_maps
 *                               + top-2 bookkeeping for dream/network.py:548-577
 *   dreamb200_softargmax        dream/spatial_softmax.py:24-95 SoftArgmaxPavlo.forward
 *   backward entry points       autograd of the same ops (network.py:328-338 loss.backward())
 *
 * Conventions: raw device pointers, plain ints, an explicit cudaStream_t (passed as
 * void*), int return (0 = ok, <0 = error; text via dreamb200_last_error()).  No
 * allocation, no torch types, no exceptions.  Activations are NHWC fp16 with the
 * channel count padded to a multiple of 64; weights are fp16 [tap][Cout_pad][Cin_pad].
 */
#ifndef DREAMB200_H_
#define DREAMB200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DREAMB200_MAX_TAPS 16

/* output modes of conv2d_fwd */
#define DREAMB200_OUT_NHWC_F16 0   /* fp16 NHWC through TMA store (strided view allowed) */
#define DREAMB200_OUT_NCHW_F32 1   /* fp32 NCHW, first cout_real channels (network head)  */

typedef struct dreamb200_conv_desc {
  /* input activation, NHWC fp16, dense */
  const void* x; int32_t B, H, W, Cin;      /* Cin % 64 == 0 */
  int32_t in_stride;                        /* spatial stride of the conv (1 or 2) */
  /* weights fp16 [taps][Cout_pad][Cin], bias fp32 [Cout_pad] (may be NULL) */
  const void* w; const float* bias;
  int32_t taps; int32_t Cout_pad;           /* Cout_pad % 64 == 0, or 16 for the NCHW_F32 head */
  int8_t tap_dy[DREAMB200_MAX_TAPS];        /* input offset of each tap, in input pixels   */
  int8_t tap_dx[DREAMB200_MAX_TAPS];
  /* output: logical size Ho x Wo; element strides let a deconv phase write an
     interleaved view of a larger tensor (y points at the phase's first pixel). */
  void* y; int32_t Ho, Wo;
  int64_t y_stride_w, y_stride_h, y_stride_b;   /* in elements of the output dtype */
  int32_t out_mode; int32_t cout_real;      /* cout_real only for NCHW_F32 */
  /* fused epilogue */
  const void* residual;                     /* fp16 NHWC [B,Ho,Wo,Cout_pad] dense, or NULL */
  int32_t relu;
  /* fp32 residual stream (ResNet identity path kept in fp32 across the 33 bottlenecks):
     residual_f32 is added like `residual`; y_f32, when set, receives an fp32 NHWC dense copy
     of the (post-ReLU) output next to the fp16 one. Both [B,Ho,Wo,Cout_pad] or NULL. */
  const float* residual_f32;
  float* y_f32;
  /* fused nn.MaxPool2d(2) (models.py:589): when set, the 2x2/s2 (floor) max pool of the output is written
     here, fp16 NHWC dense [B,Ho/2,Wo/2,Cout_pad]; `y` may then be NULL (un-pooled tensor not stored). */
  void* y_pool;
  /* when set: *absmax = max(*absmax, max |y|) over the fp16 outputs written by this launch (device float holding a
     non-negative value, zeroed by the caller) -- lets the backward pass pick the next layer's loss scale for free */
  float* absmax;
  /* backward-pass epilogue (data gradient of layer l feeding the ReLU of layer l-1, autograd of models.py:761-827):
     gate, fp16 NHWC [B,Ho,Wo,Cout_pad] or NULL: outputs where gate <= 0 are zeroed (the ReLU mask, taken from the
     saved forward activation); out_scale, device pointer to one float or NULL: every output is multiplied by it
     (power-of-two loss re-scaling chosen on the device).  Applied before `absmax`. */
  const void* gate;
  const float* out_scale;
  /* optional fp32 [Cout_pad], zeroed by the caller: colsum[c] += sum over all output pixels of the (gated, scaled)
     fp32 results -- the bias gradient of the layer this data gradient flows into (NHWC_F16 mode only) */
  float* colsum;
} dreamb200_conv_desc;

/* fraction of the 128 accumulator rows a conv with this output size keeps busy, for the free tile
   choice (even=0) or with even tile sides as fused pooling needs (even=1); lets callers decide on fusion */
double dreamb200_conv_tile_utilization(int Wo, int Ho, int in_stride, int even);

const char* dreamb200_last_error(void);
int dreamb200_version(void);
/* number of kernels this library has launched since load (bench "gpu_launches") */
int64_t dreamb200_launch_count(void);

int dreamb200_conv2d_fwd(const dreamb200_conv_desc* d, void* stream);
/* The sub-pixel phases of a stride-2 ConvTranspose2d (dream/models.py:618-686, 37-136) or of nn.Upsample(2) + 3x3 conv
   (models.py:691-709) are `n_phases` (<= 4) convolutions over the same input with their own weights, tap offsets and
   interleaved output view: descs[0..n_phases).  When the group is uniform (same shapes / tap count / epilogue, plain
   bias + ReLU) and Cout_pad % 128 == 0 it runs as ONE launch; otherwise exactly like n_phases dreamb200_conv2d_fwd
   calls.  Results are identical either way. */
int dreamb200_conv2d_fwd_phases(const dreamb200_conv_desc* descs, int n_phases, void* stream);

/* The network's first conv (3->64, 3x3 s1 p1, +bias +ReLU; models.py:591-599) fused with the input pack:
   x fp32 NCHW [B,3,H,W], w fp16 [1][64][64] (k=(r*3+s)*3+c, zero padded), bias fp32 [64] -> y fp16 NHWC [B,H,W,64] */
int dreamb200_first_conv3x3(const float* x, const void* w, const float* bias, void* y, int B, int H, int W,
                            void* stream);
/* the same layer fed with raw uint8 frames [B,H,W,3] (PIL / decoder layout): the dataset's
 * ToTensor + Normalize (dream/datasets.py:60-75,177-179) is applied while gathering, with the same fp32
 * operations ((u/255 - mean)/std), so the result is bit-identical to dreamb200_first_conv3x3 on the host-normalised
 * tensor.  mean3 / std3: host pointers to 3 floats (config key image_normalization). */
int dreamb200_first_conv3x3_u8(const void* x_u8, const float* mean3, const float* std3, const void* w,
                               const float* bias, void* y, int B, int H, int W, void* stream);
/* uint8 [B,H,W,3] -> fp32 NCHW [B,3,H,W] = Normalize(mean,std)(ToTensor(frame)), bit-exact
 * (dream/datasets.py:60-75,177-179); mean3 / std3 host pointers */
int dreamb200_normalize_u8(const void* x_u8_nhwc, float* y_nchw, int B, int H, int W, const float* mean3,
                           const float* std3, void* stream);
/* training targets (dream/image_proc.py:866-910 create_belief_map): pts fp32 [n_maps,2] (x,y) device pointer;
 * out fp32 [n_maps,h,w]; a (2*window_radius+1)^2 stamp table[(dx^2+dy^2)] (host pointer, 2*radius^2+1 floats =
 * float32(exp(-d2/(2 sigma^2))) computed by the caller in fp64) around the truncated integer centre, only when
 * the window lies strictly inside the frame (u-r>=0, u+r+1<w, ...); zeros elsewhere. */
int dreamb200_belief_targets(const float* pts, int n_maps, int h, int w, int window_radius, const float* table,
                             float* out, void* stream);
/* x fp32 NCHW [B,3,H,W] -> out fp16 NHWC [B,Ho,Wo,Kpad]; k=(r*S+s)*3+c, zero padded */
int dreamb200_im2col_first(const float* x, void* out, int B, int H, int W,
                           int R, int S, int stride, int pad, int Ho, int Wo, int Kpad, void* stream);
/* k x k / stride s / pad p max pool, NHWC fp16, floor mode */
int dreamb200_maxpool_nhwc(const void* x, void* y, int B, int H, int W, int C,
                           int k, int s, int p, int Ho, int Wo, void* stream);
int dreamb200_upsample2_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* y += x elementwise, fp16, n % 8 == 0 (hourglass skip connections, models.py:775-799) */
int dreamb200_add_f16(void* y, const void* x, long long n, void* stream);
/* fp16 NHWC [B,H,W,Cpad] -> fp32 NCHW [B,C,H,W] and back (debug / boundary) */
int dreamb200_nhwc_f16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int Cpad, int C, void* stream);
int dreamb200_nchw_f32_to_nhwc_f16(const float* x, void* y, int B, int H, int W, int C, int Cpad, void* stream);

/* peaks_from_belief_maps for n_maps = B*K maps of h x w fp32 (contiguous).
   gauss_w: 13 fp64 taps w[0..12] for |offset| 0..12 (host computes them like scipy).
   Three implementations, chosen by dreamb200_peaks_plan (same results bit for bit): mode 0 = one CTA per map, the
   whole map in shared memory (radius 12, maps up to ~160x160); mode 1 = one CTA per band of rows, the last band of a
   map to finish assembles the table (larger maps: 208x208, 400x400, 480x640); mode 2 = three generic kernels.
   scratch: `scratch_floats` floats as reported by dreamb200_peaks_plan for the same (n_maps, h, w, radius, cap), 8-byte
   aligned, contents irrelevant; NULL when 0.  peak table: capacity `cap` per map, rows
   (x:f64, y:f64, score:f32, pad) ; counts[n_maps] = true count (may exceed cap).
   summary[n_maps*4] doubles: best x, best y, best score, second score. */
int dreamb200_peaks(const float* maps, int n_maps, int h, int w, const double* gauss_w, int radius,
                    double offset, float* scratch, int cap, double* peak_xy, float* peak_score,
                    int32_t* peak_ij, int32_t* counts, double* summary, void* stream);
int dreamb200_peaks_plan(int n_maps, int h, int w, int radius, int cap, long long* scratch_floats, int* mode);
/* The smoothing stage of dreamb200_peaks alone: out[n_maps,h,w] = scipy.ndimage.gaussian_filter(map, sigma) bit for
   bit (dream/image_proc.py:935; fp64 accumulation in scipy's tap order, "reflect" borders, one rounding to fp32 per
   pass).  Same kernels as dreamb200_peaks; scratch (n_maps*h*w floats) unless dreamb200_peaks_plan reports mode 0. */
int dreamb200_gaussian_smooth(const float* maps, int n_maps, int h, int w, const double* gauss_w, int radius,
                              float* scratch, float* out, void* stream);

/* ---- backward (training) ------------------------------------------------------------------ */
/* dW[tap][co][ci] += sum_pixels dY[p][co] * X[p + tap][ci]; dy, x NHWC fp16 [B,H,W,Cout_pad|Cin_pad],
   dw fp32 [taps][Cout_pad][Cin_pad] (caller zeroes it); autograd of nn.Conv2d weights */
int dreamb200_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad, int Cin_pad,
                    int taps, const int8_t* tap_dy, const int8_t* tap_dx, void* stream);
/* weight gradient of a stride-2 ConvTranspose2d (k3 p1 op1 / k4 p1; output 2H x 2W):
   dW[tap][co][ci] += sum_pixels dY[2p + tap][co] * X[p][ci]; x [B,H,W,Cin_pad], dy [B,2H,2W,Cout_pad] */
int dreamb200_wgrad_deconv(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad, int Cin_pad,
                           int taps, const int8_t* tap_dy, const int8_t* tap_dx, void* stream);
/* weight gradient of a stride-2 convolution: dy [B,Ho,Wo,Cout_pad], x [B,Hx,Wx,Cin_pad]:
   dW[tap][co][ci] += sum_pixels dY[p][co] * X[2p + tap][ci] */
int dreamb200_wgrad_strided(const void* dy, const void* x, float* dw, int B, int Ho, int Wo, int Hx, int Wx,
                            int Cout_pad, int Cin_pad, int taps, const int8_t* tap_dy, const int8_t* tap_dx,
                            void* stream);
/* nn.BatchNorm2d in training mode (ResNet trunk / decoder, models.py:22-32,46-76), z fp16 [rows, C]:
   bn_stats: sum[c] = sum_rows z, sumsq[c] = sum_rows z^2 (C % 64 == 0);
   bn_apply: y = relu?(z*scale[c] + shift[c] (+ residual)) fp16;
   bn_bwd_reduce: sum_dy[c] = sum dy, sum_dyz[c] = sum dy*z;  bn_bwd_apply: dy = a[c]*dy + b[c]*z + c0[c] in place.
   The two reductions are DETERMINISTIC (two-stage, fixed summation order, no floating-point atomics -- what
   cudnn.deterministic gives the reference, dream/utilities.py:15-26): `partials` is scratch of
   dreamb200_bn_reduce_workspace() floats (contents irrelevant), `tickets` n_tickets uint32 that must be ZERO on
   entry and are zero again on exit (one persistent buffer serves every launch on a stream). */
int dreamb200_bn_reduce_workspace(long long rows, int C, long long* partial_floats, int* n_tickets);
int dreamb200_bn_stats_f16(const void* z, float* sum, float* sumsq, long long rows, int C, float* partials,
                           unsigned* tickets, void* stream);
int dreamb200_bn_apply_f16(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                           long long rows, int C, int relu, void* stream);
int dreamb200_bn_bwd_reduce_f16(const void* dy, const void* z, float* sum_dy, float* sum_dyz, long long rows, int C,
                                float* partials, unsigned* tickets, void* stream);
int dreamb200_bn_bwd_apply_f16(void* dy, const void* z, const float* a, const float* b, const float* c0,
                               long long rows, int C, void* stream);
/* autograd of MaxPool2d(3, stride 2, padding 1) (resnet stem): x [B,H,W,C], dy [B,(H-1)/2+1,(W-1)/2+1,C] */
int dreamb200_maxpool3_bwd_nhwc(const void* x, const void* dy, void* dx, int B, int H, int W, int C, void* stream);
/* dy = dy * (*scale) * (y > 0): autograd of nn.ReLU given its output, fused with the power-of-two
   re-scaling that keeps fp16 gradients in range; y and/or scale may be NULL */
int dreamb200_scale_mask_f16(void* dy, const void* y, const float* scale, long long n, void* stream);
/* scale_mask and the bias gradient in one pass: dy = dy*(*scale)*(y>0); db[c] += sum_rows dy (caller zeroes db) */
int dreamb200_scale_mask_bias_f16(void* dy, const void* y, const float* scale, float* db, long long rows, int C,
                                  void* stream);
/* weight gradient of the first 3x3 convolution (models.py:591-599) without a patch tensor: dy fp16 NHWC [B,H,W,64],
 * x fp32 NCHW [B,3,H,W] (W % 4 == 0), dw fp32 [64][64] zeroed by the caller:
 * dw[co][(r*3+s)*3+c] += sum_p dy[p][co] * x[p + (r-1, s-1)][c] */
int dreamb200_wgrad_first3x3(const void* dy, const float* x, float* dw, int B, int H, int W, void* stream);
/* one step of the fp16 backward pass' loss re-scaling, on the device (all pointers: one device float):
 * f = 2^floor(log2(target / max(*amax, 1e-30))) clamped to [2^-20, 2^20]; *cum *= f; *f_out = f; *inv_out = 1 / *cum */
int dreamb200_loss_scale_step(const float* amax, float* cum, float* f_out, float* inv_out, float target, void* stream);
/* *out = max(*out, max|x|) over n fp16 values (out: device float, zeroed by the caller) */
int dreamb200_absmax_f16(const void* x, long long n, float* out, void* stream);
/* autograd of nn.MaxPool2d(2) (models.py:589): x [B,H,W,C] forward input, dy [B,H/2,W/2,C].  relu_gate != 0: x is a
 * ReLU output and the gradient continues through that ReLU (windows whose maximum is 0 pass nothing). */
int dreamb200_maxpool2_bwd_nhwc(const void* x, const void* dy, void* dx, int B, int H, int W, int C, int relu_gate,
                                void* stream);
/* autograd of nn.Upsample(scale_factor=2): dy [B,2H,2W,C] -> dx [B,H,W,C] */
int dreamb200_upsample2_bwd_nhwc(const void* dy, void* dx, int B, int H, int W, int C, void* stream);
/* db[c] += sum_rows dy[row][c]; dy fp16 [rows,C], C % 64 == 0 (caller zeroes db) */
int dreamb200_bias_grad(const void* dy, float* db, long long rows, int C, void* stream);

int dreamb200_softargmax(const float* maps, const float* beta, float* out_xy,
                         int B, int K, int H, int W, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
