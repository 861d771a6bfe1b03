/* alignq_b200 -- C ABI of the B200-native AlignQ quantization hot path.
 *
 * The reference (tinganchen/AlignQ) is pure Python/PyTorch and has no FFI; this header is the
 * boundary a replacement shared library exports.  Every entry point names the reference
 * function it replaces (paths relative to the reference root; QA = cdf_alignment/<exp>/model/
 * quantization.py, QB = cdf_alignment_admm/resnet-56-cifar-10/model/quantization.py, QC =
 * cdf_alignment_admm/dann_office/model/quantization.py, OPT = <exp>/utils/optimizer.py,
 * ADMM = <exp>/utils/admm.py).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers into caller-owned, contiguous fp32 storage
 *     (NCHW activations viewed as [B, F] row-major; weights flat).  The library never allocates,
 *     never synchronises the host and never creates streams: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) and the call returns.
 *   - Return value: 0 on success, a positive cudaError_t if a launch failed, or a negative
 *     ALIGNQ_E* argument error.  Nothing throws.  alignq_error_string() describes any code.
 *   - NaN/Inf propagate through the arithmetic; the reference raises ValueError from
 *     torch.distributions' host-side validation instead (a device sync) -- documented deviation.
 *   - Thread-safe: no global mutable state; callable from autograd worker threads.
 *   - variant: 0 = QA (c = Phi), 1 = QB, 2 = QC (t = (2 Phi - 1) [* act_range]); QB and QC differ
 *     only in corr()'s eps, which is an explicit argument.
 */
#ifndef ALIGNQ_B200_H_
#define ALIGNQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALIGNQ_ABI_VERSION 1
#define ALIGNQ_CHUNK 4096          /* elements per multi-tensor work chunk */

typedef void* alignq_stream_t;     /* cudaStream_t */

enum { ALIGNQ_VARIANT_A = 0, ALIGNQ_VARIANT_B = 1, ALIGNQ_VARIANT_C = 2 };
enum { ALIGNQ_GRAM_FP32 = 0,       /* CUDA-core FFMA, fp32 parity (1e-5) */
       ALIGNQ_GRAM_TF32X3 = 1,     /* tcgen05 kind::tf32, 3-pass split, fp32 parity (1e-5) */
       ALIGNQ_GRAM_BF16 = 2 };     /* tcgen05 kind::f16 bf16 operands, 1e-2 */

int alignq_abi_version(void);
const char* alignq_error_string(int code);
/* Number of CUDA kernels this library has launched (or captured into a graph) since it was loaded. */
uint64_t alignq_launch_count(void);

/* ---- activation quantizer ------------------------------------------------------------------
 * activation_quantize_fn.forward (QA:91-103; QB:102-132 without the ADMM branch; QC:96-110):
 *   A: y = (round(Phi(x) n)/n * 2 - 1) * act_range          n = 2^a_bit - 1
 *   B/C: y = round(t n)/n, t = (2 Phi(x) - 1) * act_range
 * a_bit == 1 uses sign() (QA:22-23); a_bit == 32 with return_cdf != 0 is stage=='align' and
 * returns the un-rounded map (QA:100-101).  a_bit == 32 without return_cdf is the identity and
 * is the caller's job (ALIGNQ_EINVAL).  `codes` (nullable) receives round(.) as int16.
 * alignq_act_bwd is the whole autograd backward in one pass (uniform_quantize.backward QA:29-32
 * chained with erf'):  gx = gy * alignq_act_grad_scale(...) * exp(-x^2/2).                    */
int alignq_act_fwd(const float* x, float* y, int16_t* codes, int64_t numel, int a_bit, float act_range,
                   int variant, int return_cdf, alignq_stream_t stream);
int alignq_act_bwd(const float* x, const float* gy, float* gx, int64_t numel, int a_bit, float act_range,
                   int variant, int return_cdf, alignq_stream_t stream);
float alignq_act_grad_scale(int a_bit, float act_range, int variant, int return_cdf);

/* Stand-alone pieces of the reference surface (the hot path uses the fused kernels):
 * uniform_quantize(k).forward (QA:15-27): y = round(x n)/n, k == 1: sign(x).
 * cdf(m, s, src).forward (QA:45-50; QB:49-59): Normal(m, s).cdf mapped per variant, and
 * pdf = 2 exp(log_prob) (pdf_out nullable); m, s are DEVICE scalars.  alignq_cdf_bwd is d/dx of
 * both outputs with m, s held constant (g_cdf / g_pdf nullable).                                  */
int alignq_uniform_q_fwd(const float* x, float* y, int64_t numel, int k, alignq_stream_t stream);
int alignq_cdf_fwd(const float* x, const float* m, const float* s, int variant, int src_is_act,
                   float act_range, float* cdf_out, float* pdf_out, int64_t numel, alignq_stream_t stream);
int alignq_cdf_bwd(const float* x, const float* m, const float* s, int variant, int src_is_act,
                   float act_range, const float* g_cdf, const float* g_pdf, float* gx, int64_t numel,
                   alignq_stream_t stream);

/* ---- weight quantizer (multi-tensor) ---------------------------------------------------------
 * weight_quantize_fn.forward (QA:62-78; QB:71-85): per-TENSOR mean / unbiased std (QA:70), CDF
 * map, rounding, dequant; also the attributes weight_cdf / weight_pdf (QB:78) read by SGD.step.
 * Tensors are segments of one flat fp32 buffer: tensor t = flat[seg_off[t], seg_off[t+1]).
 * alignq_wq_plan (host helper) splits the segments into ALIGNQ_CHUNK-element chunks:
 *   chunk_seg[c] = segment of chunk c, seg_chunk0[t] = first chunk of segment t (nseg+1 entries);
 * it returns the chunk count (pass NULL outputs to size the tables).  seg_off / chunk_seg /
 * seg_chunk0 passed to the launch functions are DEVICE copies of those tables.
 * ws: nchunks*2 doubles of scratch.  stats: nseg*4 floats out = {mean, std, 1/std, numel}.
 * Backward (autograd through mean and std):
 *   gw_j = (1/s) [a_j - sum(a)/N - z_j sum(a z)/(N-1)],  a = 2 g phi(z).
 * Upstream gradients: either flat `g_wq` (same segmentation as `flat`) or `g_ptrs`, a HOST array of nseg
 * device pointers to per-tensor gradients, handed to the kernels by value (no table upload) (NULL entry = tensor not in this backward pass, left untouched);
 * accumulate != 0 adds into g_w.                                                                  */
int64_t alignq_wq_plan(const int64_t* seg_off_host, int nseg, int32_t* chunk_seg_host, int32_t* seg_chunk0_host);
int alignq_wq_forward(const float* flat, const int64_t* seg_off, const int32_t* chunk_seg,
                      const int32_t* seg_chunk0, int nseg, int64_t nchunks, int w_bit, int variant,
                      float* wq, float* w_cdf, float* w_pdf, int16_t* codes, float* stats, double* ws,
                      alignq_stream_t stream);
int alignq_wq_backward(const float* flat, const float* g_wq, const float* const* g_ptrs,
                       const int64_t* seg_off, const int32_t* chunk_seg, const int32_t* seg_chunk0,
                       int nseg, int64_t nchunks, int w_bit, const float* stats, float* g_w,
                       int accumulate, double* ws, alignq_stream_t stream);

/* ---- correlation / Gram ----------------------------------------------------------------------
 * corr(x, y) (QB:134-137 eps = 0; QC:158-161 eps = 1e-5): standardise every feature column over
 * the batch dim (unbiased std), G = Xs Ys^T / F.  x, y: [B, F] row-major (y may alias x).
 * ws: alignq_gram_ws_bytes(B, F) bytes of scratch.                                              */
size_t alignq_gram_ws_bytes(int B, int64_t F);
int alignq_corr_fwd(const float* x, const float* y, int B, int64_t F, float eps, float* G, void* ws,
                    size_t ws_bytes, int gram_mode, alignq_stream_t stream);
/* Autograd backward of corr(x, y) for an upstream dG [B, B] (the reference's corr is differentiable through
 * mean and std, QB:135-137): gXs = dG Ys / F, gYs = dG^T Xs / F, then the standardise backward of each operand
 * (the std term is masked at std == 0 as torch's std_backward does).  gx / gy are nullable;
 * when y aliases x there is one operand and its gradient (both terms) goes to gx.  fp32 FFMA kernels.
 * ws: at least 2 * roundup(B,4)^2 floats (alignq_gram_ws_bytes(B, F) always suffices).              */
int alignq_corr_bwd(const float* x, const float* y, const float* dG, int B, int64_t F, float eps, float* gx,
                    float* gy, void* ws, size_t ws_bytes, alignq_stream_t stream);

/* ---- fused activation quantizer + ADMM correlation term --------------------------------------
 * activation_quantize_fn.forward with method=='ours' (QB:102-132) / activation_quantize_fn2
 * (QC:126-156): one read of x produces y, corr(x), corr(t), D = corr(t) - corr(x) (QB:118-122),
 * trans_loss = ADMM.forward(D) (ADMM:24-33) and dLdD for the backward.  Z = alterD, U = gamma are
 * [dim, dim] with dim >= B; the [:B, :B] corner is used (ADMM:26-27).
 * Backward: gx = corr_bwd(x, -gloss dLdD) + (corr_bwd(t, gloss dLdD) + gy) * 2 ar phi(x);
 * gloss is a DEVICE scalar (the upstream gradient of trans_loss); gy may be NULL (no y-grad).  */
int alignq_act_admm_fwd(const float* x, int B, int64_t F, int a_bit, float act_range, float eps,
                        const float* Z, const float* U, int dim, float mu, float rho,
                        float* y, float* D, float* loss, float* dLdD, void* ws, size_t ws_bytes,
                        int gram_mode, alignq_stream_t stream);
int alignq_act_admm_bwd(const float* x, const float* gy, const float* dLdD, const float* gloss,
                        int B, int64_t F, int a_bit, float act_range, float eps, float* gx,
                        void* ws, size_t ws_bytes, int gram_mode, alignq_stream_t stream);

/* ---- data-parallel (feature-sharded) pieces of the same term ------------------------------------
 * With the batch sharded over P GPUs the Gram is [B_global, B_global] ACROSS samples: every rank receives
 * (all-to-all) all B_global rows of its slice of the feature columns and computes the UN-NORMALISED column sums
 *   sums[0] = sum_f Xs Xs^T,  sums[1] = sum_f Ts Ts^T      (Xs, Ts standardised over all B_global rows: exact)
 * (alignq_gram_sums_fwd); the sums of all ranks are added with ONE all-reduce (NCCL), then
 * alignq_gram_sums_to_d forms D = sums[1]/F - sums[0]/F with the single-device rounding (QB:118-122) for nmod
 * modules at once (inv_f: DEVICE array of 1/F per module), and alignq_admm_loss evaluates ADMM.forward.
 * Backward: alignq_act_admm_bwd on the slice with gy = NULL and *gloss / P, all-to-all back, and
 * alignq_act_bwd_add adds the straight-through term: gx = gadd + gy * 2 ar phi(x).                  */
int alignq_gram_sums_fwd(const float* x, int B, int64_t F, float act_range, float eps, float* sums, void* ws,
                         size_t ws_bytes, int gram_mode, alignq_stream_t stream);
int alignq_gram_sums_to_d(const float* sums, const float* inv_f, int B, int nmod, float* D, alignq_stream_t stream);
int alignq_act_bwd_add(const float* x, const float* gy, const float* gadd, float* gx, int64_t numel, int a_bit,
                       float act_range, int variant, int return_cdf, alignq_stream_t stream);

/* ---- ADMM loss and Z/U update (batched over nmod modules) ------------------------------------
 * ADMM.forward (ADMM:24-33): loss = mu mean|Z| + rho/2 sqrt(mean (D-Z)^2) + mean(U |D-Z|), and
 * dLdD.  Module m uses D + m*B*B, Z + m*dim*dim, U + m*dim*dim, loss[m],
 * dLdD + m*B*B.  loss, dLdD, dLdZ, dLdU ([dim,dim], zero outside the corner) are each nullable;
 * the gradients are scaled by the DEVICE scalar *gloss (gloss[m] if gloss_per_module; NULL = 1).
 * ADMM_OPT.step (OPT:97-124): V = pad(D) + U/rho; Z <- max(0, 1 - (mu/rho)/||V||_F) V;
 * U <- U + rho (pad(D) - Z).  No host sync (the reference branches on the host).                */
int alignq_admm_loss(const float* D, int B, const float* Z, const float* U, int dim, int nmod,
                     float mu, float rho, const float* gloss, int gloss_per_module,
                     float* loss, float* dLdD, float* dLdZ, float* dLdU, alignq_stream_t stream);
int alignq_admm_zu_update(float* Z, float* U, const float* D, int B, int dim, int nmod,
                          float mu, float rho, alignq_stream_t stream);

/* ---- multi-tensor SGD with the quantization-aware gradient surrogate --------------------------
 * SGD.step(idx, w_cdf, w_pdf, lam, lam2) (OPT:196-262): d_p = grad_scale g + wd p (in place on g;
 * grad_scale = 1 reproduces the reference, 1/world_size averages a summed data-parallel gradient);
 * buf = mom buf + (1-damp) d_p (first step: buf = d_p); d_p = nesterov ? d_p + mom buf : buf;
 * p -= lr d_p; g <- d_p, or for tensors with w_cdf != NULL:
 * g <- d_p * sigmoid'(((w_cdf+0.5)(2^bitW-1) mod 1) 2 lam2) lam * w_pdf (OPT:6-13, 232-249).
 * `tensors` is a DEVICE array; chunk tables as in alignq_wq_plan over numel.                    */
typedef struct alignq_sgd_tensor {
  float* p;
  float* g;
  float* buf;              /* NULL when momentum == 0 */
  const float* w_cdf;      /* NULL unless the tensor is in idx */
  const float* w_pdf;
  int64_t numel;
  float lr, momentum, dampening, weight_decay;
  int32_t nesterov;
  int32_t first_step;      /* 1: momentum buffer not yet initialised */
} alignq_sgd_tensor_t;
int alignq_sgd_step(const alignq_sgd_tensor_t* tensors, const int32_t* chunk_tensor,
                    const int32_t* tensor_chunk0, int ntensors, int64_t nchunks,
                    float lam, float lam2, int bitW, float grad_scale, alignq_stream_t stream);

/* ---- bf16 Gram on the tensor cores (cp.async- or TMA-fed tcgen05 pipeline, split-K) ------------------
 * G = X X^T (divided by F if divide_by_F) for X [B <= 256, F] bf16 row-major, F % 8 == 0, 16-byte
 * aligned: the dense contraction of corr() (QB:137) for operands already standardised and stored
 * in bf16 (the tensor-bound micro-shape of SURVEY.md 8d).  G: [B, B] fp32.
 * ws: alignq_gram_bf16_ws_bytes(B) bytes.                                                          */
size_t alignq_gram_bf16_ws_bytes(int B);
int alignq_gram_bf16(const void* x_bf16, int B, int64_t F, int divide_by_F, float* G, void* ws,
                     size_t ws_bytes, alignq_stream_t stream);

/* ---- fused BatchNorm2d -> activation quantizer -> (ReLU), NHWC ------------------------------------
 * The step either side of the quantizer in every model file, e.g.
 * `F.relu(self.act_q0(self.bn0(out)))` (cdf_alignment/resnet-20-cifar-10/model/resnet.py:72,121-123),
 * `self.relu(self.act_q0(self.bn1(x)))` (cdf_alignment/dense-cifar-10/model/densenet.py:32-34).
 * x, y, gy, gx: [rows = B*H*W, C] fp32 with C contiguous (channels_last), 16-byte aligned, C % 4 == 0,
 * C <= 1024.  training != 0: batch statistics (biased variance), running_mean/var updated with
 * `momentum` (unbiased variance) unless NULL, *num_batches_tracked += 1 unless NULL; training == 0:
 * running statistics.  save_mean /
 * save_invstd: [C] out (forward) / in (backward).  ws: alignq_bn_act_ws_doubles(C) doubles and
 * counter: TWO uint32 -- both must be ZERO before the first call and belong to ONE layer (one value of
 * C): the kernels keep state in them between calls (counter[0]: a ticket they re-arm; counter[1]: the
 * epoch of the single-launch backward, which says which of its two accumulator sets is clean).
 * `residual` (nullable, same layout): y = relu(q(bn(x)) + residual), the block tail `out += shortcut;
 * F.relu(out)` (resnet.py:77-78); g_residual (nullable) receives its gradient.
 * Backward: g_z = gy [y > 0 if relu] * 2 ar phi(z) (z = BN output), then the BatchNorm backward;
 * ggamma / gbeta ([C], nullable) receive the affine gradients.                                      */
size_t alignq_bn_act_ws_doubles(int C);
int alignq_bn_act_fwd(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float bn_eps, int training,
                      int a_bit, float act_range, int variant, int relu, const float* residual, float* y,
                      float* save_mean, float* save_invstd, double* ws, uint32_t* counter,
                      int64_t* num_batches_tracked, alignq_stream_t stream);
int alignq_bn_act_bwd(const float* x, const float* y, const float* gy, int64_t rows, int C,
                      const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                      int training, int a_bit, float act_range, int variant, int relu, float* gx,
                      float* g_residual, float* ggamma, float* gbeta, double* ws, uint32_t* counter,
                      alignq_stream_t stream);
/* Same, for an output with TWO consumers (the block input of `out = conv0(x) ... ; out += shortcut`, resnet.py:70-78):
 * the upstream gradient is gy + gy2 (gy2 nullable, same layout), added while it is read, so that autograd's separate
 * accumulate kernel never runs (model/fused.py:_GradFork).                                             */
int alignq_bn_act_bwd_sum(const float* x, const float* y, const float* gy, const float* gy2, int64_t rows, int C,
                          const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                          int training, int a_bit, float act_range, int variant, int relu, float* gx,
                          float* g_residual, float* ggamma, float* gbeta, double* ws, uint32_t* counter,
                          alignq_stream_t stream);
/* The apply pass of the backward alone, for a gy that came out of alignq_conv3x3_bwd_data_bnreduce (below): that
 * convolution's epilogue has already formed the two BatchNorm-backward sums and left their means in `ws`.   */
int alignq_bn_act_bwd_apply(const float* x, const float* y, const float* gy, int64_t rows, int C, const float* gamma,
                            const float* beta, const float* save_mean, const float* save_invstd, int a_bit,
                            float act_range, int variant, int relu, float* gx, float* g_residual, double* ws,
                            alignq_stream_t stream);


/* ---- 3x3 convolution of the quantized conv layers on the tensor cores (tcgen05) ------------------------------
 * `F.conv2d(input, weight_q, None, 1, 1)` of Conv2d_Q.forward (QA:116-120), its data gradient and its weight gradient,
 * for stride 1, padding 1, Cin == Cout == C in {16, 32, 64}, NHWC (channels_last) fp32:
 *   x, y, gy, gx: [N, H, W, C] physical;  w, gw: [C_out, 3, 3, C_in] physical (a channels_last [Cout, Cin, 3, 3]).
 * Implicit GEMM over shifted views of one staged tile (no im2col), accumulators in TMEM; every pointer 16-byte aligned.
 * mode: ALIGNQ_CONV_TF32 -- tf32 operands, one MMA pass (what cuDNN runs under torch's default allow_tf32 = True;
 * ~5e-4 relative to an fp32 convolution); ALIGNQ_CONV_TF32X3 -- operands split H + L, three passes, fp32 parity (1e-5).
 * Returns ALIGNQ_ERANGE for a shape / mode the kernels do not cover (e.g. C == 64 forward in TF32X3 mode: the split
 * weights do not fit in shared memory) -- the caller then uses its library convolution.
 * bwd_weight: ws = alignq_conv3x3_ws_bytes(C) bytes of scratch (split-K partials, summed in a fixed order);
 * accumulate != 0 adds into gw.                                                                                   */
enum { ALIGNQ_CONV_TF32 = 0, ALIGNQ_CONV_TF32X3 = 1 };
int alignq_conv3x3_fwd(const float* x, const float* w, float* y, int N, int H, int W, int C, int mode,
                       alignq_stream_t stream);
int alignq_conv3x3_bwd_data(const float* gy, const float* w, float* gx, int N, int H, int W, int C, int mode,
                            alignq_stream_t stream);
/* The data gradient with the reduce pass of the PRECEDING bn-act layer's backward in its epilogue (C = 16).  gx, the
 * gradient with respect to the convolution's input, is that layer's upstream gradient: the epilogue adds gy2 (nullable:
 * the gradient of a second consumer of the same tensor, e.g. the block shortcut, resnet.py:70-78) and WRITES THE SUM to
 * gx, forms g_z = gx [bn_y > 0 if relu] 2 ar phi(z) from bn_x (the bn-act layer's input; bn_y its output, i.e. this
 * convolution's forward input) and accumulates sum g_z, sum g_z xhat into the layer's workspace; the last CTA writes
 * ggamma / gbeta ([C], nullable) and the two means.  Follow it with alignq_bn_act_bwd_apply on the same workspace. */
int alignq_conv3x3_bwd_data_bnreduce(const float* gy, const float* w, float* gx, int N, int H, int W, int C, int mode,
                                     const float* bn_x, const float* bn_y, const float* gy2, const float* save_mean,
                                     const float* save_invstd, const float* gamma, const float* beta, float act_range,
                                     int relu, float* ggamma, float* gbeta, double* bn_ws, uint32_t* bn_counter,
                                     alignq_stream_t stream);
/* The forward convolution with the batch statistics of the FOLLOWING BatchNorm2d taken from its epilogue (C in {16, 32}):
 * save_mean / save_invstd [C] out, running statistics updated (nullable), *num_batches_tracked += 1 (nullable); bn_ws /
 * bn_counter are that BatchNorm layer's alignq_bn_act_* workspace (same accumulator layout, same last-block finalisation).
 * Follow it with alignq_bn_act_apply (the apply pass alone) instead of alignq_bn_act_fwd.                              */
int alignq_conv3x3_fwd_bnstats(const float* x, const float* w, float* y, int N, int H, int W, int C, int mode,
                               float* running_mean, float* running_var, float momentum, float bn_eps, float* save_mean,
                               float* save_invstd, double* bn_ws, uint32_t* bn_counter, int64_t* num_batches_tracked,
                               alignq_stream_t stream);
int alignq_bn_act_apply(const float* x, int64_t rows, int C, const float* gamma, const float* beta, const float* save_mean,
                        const float* save_invstd, int a_bit, float act_range, int variant, int relu, const float* residual,
                        float* y, alignq_stream_t stream);
size_t alignq_conv3x3_ws_bytes(int C);
int alignq_conv3x3_bwd_weight(const float* x, const float* gy, float* gw, int N, int H, int W, int C, int mode,
                              int accumulate, void* ws, size_t ws_bytes, alignq_stream_t stream);

/* ---- first-layer 3x3 convolution, Cin = 3 (conv0 of .../model/resnet.py:92 through quantization.py:116-120) ------------
 * x [N, H, W, 3], y / gy [N, H, W, Cout] NHWC fp32, w / gw [Cout][3][3][3] (channels_last weight), stride 1, padding 1,
 * Cout in {16, 32} (else ALIGNQ_ERANGE); direct fp32 FFMA kernels (exact fp32 products).
 * fwd: save_mean != NULL additionally produces the batch statistics of the following BatchNorm from the epilogue (same
 *      arguments and accumulator layout as alignq_conv3x3_fwd_bnstats; alignq_bn_act_apply follows); NULL: convolution only.
 * bwd_weight: ws = alignq_conv3x3_stem_ws_bytes(Cout) bytes, ZERO before the first call (the kernel re-arms it).
 * There is no data gradient (the image needs none). */
size_t alignq_conv3x3_stem_ws_bytes(int Cout);
int alignq_conv3x3_stem_fwd(const float* x, const float* w, float* y, int N, int H, int W, int Cout, float* running_mean,
                            float* running_var, float momentum, float bn_eps, float* save_mean, float* save_invstd,
                            double* bn_ws, uint32_t* bn_counter, int64_t* num_batches_tracked, alignq_stream_t stream);
int alignq_conv3x3_stem_bwd_weight(const float* x, const float* gy, float* gw, int N, int H, int W, int Cout, void* ws,
                                   size_t ws_bytes, alignq_stream_t stream);

/* ---- classifier head + loss: avg_pool2d -> view -> linear -> CrossEntropyLoss(mean) ------------------------------------
 * (cdf_alignment/resnet-20-cifar-10/model/resnet.py:108-110 and main.py:283-286) in one forward and one backward launch.
 * feat [B, HW, C] (NHWC / channels_last) fp32, W [K, C], bias [K] or NULL, target [B] int64 in [0, K); C % 4 == 0,
 * C <= 1024, K <= 1024 (else ALIGNQ_ERANGE).
 * fwd: logits [B, K], loss [1]; saves pooled [B, C] and glogits [B, K] = (softmax - onehot) / B for the backward;
 *      ws = alignq_head_ce_ws_bytes(B) bytes whose first 16 are ZERO before the first call (the kernel re-arms them).
 * bwd: gloss = device pointer to the upstream gradient of the loss (NULL: 1); g_feat [B, HW, C], g_W [K, C], g_b [K]
 *      (each may be NULL).  Deterministic (fixed summation orders). */
size_t alignq_head_ce_ws_bytes(int B);
int alignq_head_ce_fwd(const float* feat, const float* W, const float* bias, const int64_t* target, int B, int HW, int C, int K,
                       float* logits, float* loss, float* pooled, float* glogits, void* ws, size_t ws_bytes,
                       alignq_stream_t stream);
int alignq_head_ce_bwd(const float* W, const float* pooled, const float* glogits, const float* gloss, int B, int HW, int C, int K,
                       float* g_feat, float* g_W, float* g_b, alignq_stream_t stream);

/* ---- LMMD loss of the DSAN head (cdf_alignment_admm/dsan_office/utils/mmd.py:9-41) ------------------------------
 * total = cat(source, target) [n, d] fp32; W [n, n] = [[w_ss, -w_st], [-w_st^T, w_tt]] (the label weights of
 * utils/Weight.py:10-59, signed and assembled by the caller): loss = sum_ij W_ij K_ij with K the sum of kernel_num
 * Gaussian kernels of the pairwise squared distances (bandwidth = mean distance, a constant; fix_sigma > 0 overrides),
 * and 0 when K contains a NaN (mmd.py:33-34).  coef [n, n] out: what alignq_lmmd_bwd needs;
 * ws: alignq_lmmd_ws_bytes(n) bytes whose first 32 bytes are ZERO on entry (kept for the backward).
 * Backward: g_total = *gloss * d loss / d total (gloss a DEVICE scalar, NULL = 1).                                  */
size_t alignq_lmmd_ws_bytes(int n);
int alignq_lmmd_fwd(const float* total, int n, int d, const float* W, float kernel_mul, int kernel_num, float fix_sigma,
                    float* loss, float* coef, void* ws, size_t ws_bytes, alignq_stream_t stream);
int alignq_lmmd_bwd(const float* total, int n, int d, const float* coef, const float* gloss, const void* ws,
                    float* g_total, alignq_stream_t stream);

/* Data-parallel SyncBN inside the fused kernels (BASELINE.json north_star (3): "(sum, sum-of-squares) ... combined with
 * NCCL all-reduce ... so global statistics match the single-device reference"): the same kernels, cut where the ranks'
 * fp64 sums are exchanged.  The caller all-reduces `sums` ([2 C] doubles: per-channel sum and sum of squares forward;
 * sum g_z and sum g_z xhat backward) between the two calls of a pass; rows_global = rows summed over the ranks.
 * ggamma / gbeta are the LOCAL sums (they are averaged with the other parameter gradients).  Training mode only. */
int alignq_bn_act_sync_stats(const float* x, int64_t rows, int C, double* sums, double* ws, uint32_t* counter,
                             alignq_stream_t stream);
int alignq_bn_act_sync_apply(const float* x, int64_t rows, int64_t rows_global, int C, const double* sums,
                             const float* gamma, const float* beta, float* running_mean, float* running_var,
                             float momentum, float bn_eps, int a_bit, float act_range, int variant, int relu,
                             const float* residual, float* y, float* save_mean, float* save_invstd,
                             int64_t* num_batches_tracked, alignq_stream_t stream);
int alignq_bn_act_sync_bwd_reduce(const float* x, const float* y, const float* gy, int64_t rows, int C,
                                  const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_invstd, int a_bit, float act_range, int variant, int relu,
                                  double* sums, float* ggamma, float* gbeta, double* ws, uint32_t* counter,
                                  alignq_stream_t stream);
int alignq_bn_act_sync_bwd_apply(const float* x, const float* y, const float* gy, int64_t rows, int64_t rows_global,
                                 int C, const float* gamma, const float* beta, const float* save_mean,
                                 const float* save_invstd, int a_bit, float act_range, int variant, int relu,
                                 const double* sums, float* gx, float* g_residual, double* ws, alignq_stream_t stream);

/* The same exchange done INSIDE the kernels over NVLink peer memory -- no collective call, no extra launch: the last
 * block of the statistics (backward: reduce) kernel stores this rank's [2 C] fp64 sums into every rank's peer-mapped
 * buffer and releases a flag; every block of the apply kernel that follows waits for all ranks' flags in its own
 * buffer and adds the world's sums in rank order (bit-identical on every rank).
 * peer_bufs: DEVICE array of `world` (<= 8) base pointers, entry r = rank r's buffer of alignq_bn_act_peer_bytes()
 * bytes, zero-initialised, mapped into this process (e.g. torch symmetric memory); peer_seq: TWO zero-initialised
 * uint32 in local device memory (the exchange counter, advanced by the kernels: CUDA-graph replayable; and the word
 * on which the blocks of a consumer kernel wait for its block 0).  Every rank
 * must issue the same sequence of peer calls on ONE stream.  C <= 1024; rows_global = rows summed over the ranks.  */
size_t alignq_bn_act_peer_bytes(void);
int alignq_bn_act_fwd_peer(const float* x, int64_t rows, int64_t rows_global, int C, const float* gamma,
                           const float* beta, float* running_mean, float* running_var, float momentum, float bn_eps,
                           int a_bit, float act_range, int variant, int relu, const float* residual, float* y,
                           float* save_mean, float* save_invstd, double* ws, uint32_t* counter,
                           int64_t* num_batches_tracked, const void* const* peer_bufs, uint32_t* peer_seq, int rank,
                           int world, alignq_stream_t stream);
int alignq_bn_act_bwd_peer(const float* x, const float* y, const float* gy, int64_t rows, int64_t rows_global, int C,
                           const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                           int a_bit, float act_range, int variant, int relu, float* gx, float* g_residual,
                           float* ggamma, float* gbeta, double* ws, uint32_t* counter, const void* const* peer_bufs,
                           uint32_t* peer_seq, int rank, int world, alignq_stream_t stream);
int alignq_bn_act_bwd_peer_sum(const float* x, const float* y, const float* gy, const float* gy2, int64_t rows,
                               int64_t rows_global, int C, const float* gamma, const float* beta, const float* save_mean,
                               const float* save_invstd, int a_bit, float act_range, int variant, int relu, float* gx,
                               float* g_residual, float* ggamma, float* gbeta, double* ws, uint32_t* counter,
                               const void* const* peer_bufs, uint32_t* peer_seq, int rank, int world,
                               alignq_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* ALIGNQ_B200_H_ */
