/*
 * exposure_b200 -- C ABI of the B200-native Exposure hot path.
 *
 * The reference (yuanming-hu/exposure @ 7bb838a) has NO native code and NO FFI: every op
 * below replaces a chain of stock TensorFlow-1.6 kernels that the Python graph in
 * filters.py / critics.py / agent.py / net.py lowers to.  Each entry point cites the
 * reference lines whose arithmetic it replaces.  The reference-side binding a maintainer
 * would add (a ctypes stub inside Filter.process etc.) is shown in INTEGRATION.md.
 *
 * Conventions (tested in tests/test_cabi_*.py):
 *   - plain pointers and sizes only; all `float*` / `int*` are DEVICE pointers unless
 *     the name ends in `_host`;
 *   - the caller owns every buffer; the library never allocates, frees or synchronises;
 *     all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream);
 *   - images are NHWC contiguous fp32 `[B, H, W, 3]` (replay_memory.py:16-20 layout);
 *   - return value: 0 = EXP_OK, negative = error (see enum); exp_last_error() returns a
 *     thread-local human readable message.  No exceptions cross the ABI;
 *   - outputs are OVERWRITTEN, never accumulated into;
 *   - re-entrant / thread-safe given distinct streams and distinct workspaces.
 */
#ifndef EXPOSURE_B200_H_
#define EXPOSURE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  EXP_OK = 0,
  EXP_ERR_INVALID_ARG = -1, /* null pointer, bad size, bad filter id                */
  EXP_ERR_ALIGNMENT = -2,   /* pointer not aligned as documented                    */
  EXP_ERR_WORKSPACE = -3,   /* workspace too small                                  */
  EXP_ERR_CUDA = -4,        /* a CUDA runtime call failed (message has the reason)  */
  EXP_ERR_UNSUPPORTED = -5  /* shape outside what the kernel family supports        */
};

/* Filter ids.  0..7 = index in the shipped cfg.filters (config_example.py:22-25); 8, 9 = the two
 * Filter subclasses of filters.py that no shipped config lists. */
enum {
  EXP_FILTER_EXPOSURE = 0,   /* ExposureFilter              filters.py:170-182 */
  EXP_FILTER_GAMMA = 1,      /* GammaFilter                 filters.py:194-206 */
  EXP_FILTER_WB = 2,         /* ImprovedWhiteBalanceFilter  filters.py:215-238 */
  EXP_FILTER_SATPLUS = 3,    /* SaturationPlusFilter        filters.py:474-498 */
  EXP_FILTER_TONE = 4,       /* ToneFilter                  filters.py:298-322 */
  EXP_FILTER_CONTRAST = 5,   /* ContrastFilter              filters.py:404-419 */
  EXP_FILTER_WNB = 6,        /* WNBFilter                   filters.py:428-440 */
  EXP_FILTER_COLOR = 7,      /* ColorFilter                 filters.py:247-273 */
  EXP_NUM_FILTERS = 8,       /* length of the shipped cfg.filters = number of actions      */
  EXP_FILTER_LEVEL = 8,      /* LevelFilter                 filters.py:449-464 */
  EXP_FILTER_VIGNET = 9,     /* VignetFilter                filters.py:341-352 (process == img*0;
                                its own 5-parameter mask: exp_filter_masked_*)            */
  EXP_NUM_FILTER_KINDS = 10
};
#define EXP_MAX_FILTER_PARAMS 24 /* ColorFilter: 3 channels x cfg.curve_steps(8) */

/* Kernel variants of the per-pixel filter step (all produce identical results up to the
 * documented tolerance; AUTO picks the fastest legal one for the shape). */
enum {
  EXP_VARIANT_AUTO = 0,
  EXP_VARIANT_DIRECT = 1, /* register-resident float4 loads/stores                     */
  EXP_VARIANT_TMA = 2,    /* cp.async.bulk (TMA) -> shared-memory ring -> bulk store    */
  EXP_VARIANT_SCALAR = 3  /* one pixel per thread; any H*W, any 4-byte alignment        */
};
/* `options` of exp_filter_fwd / exp_filter_bwd = variant | flags.  EXP_OPT_LOGITS: `params`
 * holds the RAW regressor logits (the input of exp_filter_regress_fwd); the kernel applies
 * filter_param_regressor in its prologue and exp_filter_bwd returns dL/dlogits in `gparams`
 * (regressor fused into the filter step: no separate per-image launches). */
#define EXP_OPT_LOGITS 0x100
/* exp_filter_chain_fwd_bwd_uniform only: never take a compile-time instantiation of the chain (A/B switch of the
 * tests: the run-time kernel must give the same y / gx bit for bit). */
#define EXP_OPT_NO_STATIC_CHAIN 0x200

/* ---- library ------------------------------------------------------------------- */
int exp_version(void);               /* ABI version, currently 1                       */
const char* exp_last_error(void);    /* thread-local message of the last failure       */
int exp_num_filter_params(int filter_id); /* n of the table above, or EXP_ERR_INVALID_ARG */

/* cfg-driven ranges of the filter_param_regressors.  Replaces the cfg reads of filters.py:179
 * (cfg.exposure_range), :202 (cfg.gamma_range), :261 (cfg.color_curve_range, tanh_range(..., initial=1)) and
 * :309 (cfg.tone_curve_range).  Process-wide host state (the reference has ONE global cfg, util.load_config
 * util.py:326-329); every later launch carries a copy by value, so launches already enqueued or captured in a
 * CUDA graph keep the ranges they were launched with.  r == NULL restores the defaults of
 * config_example.py:27-33: 3.5, 3, (0.5, 2), (0.90, 1.10).  `curve_steps` is cfg.curve_steps
 * (config_example.py:27): the curve kernels are built for EXP_CURVE_STEPS knots, anything else returns
 * EXP_ERR_UNSUPPORTED instead of training a different curve.  (ImprovedWhiteBalanceFilter's range is the literal
 * 0.5 of filters.py:224 -- cfg.wb_range is never read by the reference.) */
#define EXP_CURVE_STEPS 8
typedef struct {
  float exposure_range;       /* p = tanh_range(-r, r)            */
  float gamma_range;          /* gamma = exp(tanh_range(-ln g, ln g)) */
  float tone_lo, tone_hi;     /* cfg.tone_curve_range             */
  float color_lo, color_hi;   /* cfg.color_curve_range, must contain 1 */
} exp_filter_ranges;
int exp_set_filter_ranges(const exp_filter_ranges* ranges, int curve_steps);
int exp_get_filter_ranges(exp_filter_ranges* ranges);
/* Programmatic dependent launch (process-wide; default OFF, environment EXPOSURE_PDL=1 turns it on):
 * the library's kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so that
 * on a stream -- and inside a CUDA graph captured from it -- a kernel's prologue overlaps the tail of
 * its predecessor.  Every kernel waits (griddepcontrol.wait) before its first global-memory access,
 * so results are identical with the switch on or off.  Off by default because the measured train
 * iteration is not launch-gap bound. */
int exp_set_pdl(int enable);

/* ---- filter_param_regressor (per image, tiny) -------------------------------------
 * logits [B, lstride] -> params [B, pstride] for the filter given by ids[b] (device
 * int32 [B]) or, when ids == NULL, by `uniform_id` for every image.
 * Replaces: ExposureFilter/GammaFilter/.../ColorFilter.filter_param_regressor
 * (filters.py:177-179, 201-203, 223-235, 256-262, 306-310, 411-413, 435-436, 481-482)
 * and util.tanh_range (util.py:281-294). */
int exp_filter_regress_fwd(const float* logits, int lstride, float* params, int pstride,
                           const int* ids, int uniform_id, int B, void* stream);
/* glogits [B, lstride] = d<gparams, params>/dlogits (entries >= n are written as 0). */
int exp_filter_regress_bwd(const float* logits, int lstride, const float* gparams, int pstride,
                           float* glogits, const int* ids, int uniform_id, int B, void* stream);

/* ---- Filter.process forward (the hot path) -----------------------------------------
 * y[b] = process_{id(b)}(x[b], params[b]) ; with cfg.masking == False this is exactly
 * Filter.apply's low_res_output / high_res_output (filters.py:62-99: lerp with
 * mask == ones(1,1,1,1)).  Also replaces the stack / one_hot / reduce_sum select of
 * agent.py:77,118-129 (only the selected filter is evaluated per image).
 * x, y: [B,H,W,3] fp32, 16-byte aligned for the DIRECT/TMA variants (else SCALAR).
 * x == y (in place) is allowed.  Algorithmic HBM traffic: 24 B / pixel.
 * ids[b] == -1 (pdf_sample's u == 0 quirk, pdf_sample_layer.py:5-10: an all-zero one-hot row) writes a BLACK
 * image y[b] = 0, and exp_filter_bwd writes gx[b] = 0, gparams[b, :] = 0: output buffers may be uninitialised. */
int exp_filter_fwd(const float* x, float* y, const float* params, int pstride,
                   const int* ids, int uniform_id, int B, int H, int W, int options,
                   void* stream);

/* ---- S filter steps in ONE pass (inference / high-resolution path) ----------------------
 * y[b] = f_{ids[S-1][b]}( ... f_{ids[0][b]}(x[b]) ... ) with params [S][B][pstride] and
 * ids int32 [S][B] (id -1 = black, the pdf_sample u==0 quirk).  Intermediates stay in
 * registers: 24 B/pixel for the whole episode instead of 24*S.  Replaces the per-step
 * high_res_output of Filter.apply (filters.py:89-96) driven by net.py:796-820 (one sess.run
 * per step on the full-resolution image).  S <= 8. */
int exp_filter_chain_fwd(const float* x, float* y, const float* params, int pstride, const int* ids,
                         int S, int B, int H, int W, int options, void* stream);

/* ---- whole chain forward + backward in ONE pass (SURVEY 8d "fully chain-fused variant") --------
 * y = f_{ids[S-1]}( ... f_{ids[0]}(x) ... ), gx = (dy/dx)^T gy, and for every step s the parameter
 * gradients gparams[s][b][:] = d<gy, y>/dparams[s][b] (or d/dlogits with EXP_OPT_LOGITS), computed from
 * ONE read of x and gy: 24 B/pixel in, 12 B/pixel out for gx plus 12 for y (both nullable: pass NULL to
 * skip the store) -- against 60 B/pixel/step for S separate exp_filter_fwd / exp_filter_bwd launches,
 * which this call equals in results (same per-pixel functions, same deterministic reduction scheme).
 * Step inputs are parked in shared memory, nothing is recomputed; the kernel is instruction-bound.
 * Use when the S filters and their parameters are known up front (a recorded episode, the 8-filter
 * chain of BASELINE configs[1] / [4]); the agent's step-by-step rollout keeps the per-step entry points
 * because each action depends on the previous step's output (agent.py:41-125).  No masking.
 * params / gparams: [S][B][pstride], ids int32 [S][B] (id -1 = black, zero gradients), S <= 8.
 * Of a gparams row only the first n = exp_num_filter_params(id) entries are written (all
 * EXP_MAX_FILTER_PARAMS for id -1): hand in a zero-filled buffer if the ids change between calls.
 * workspace: exp_filter_chain_fwd_bwd_workspace_bytes(); same zero-once / self-cleaning rule as
 * exp_filter_bwd, and the same buffer may be shared with it. */
size_t exp_filter_chain_fwd_bwd_workspace_bytes(int S, int B, int H, int W);
int exp_filter_chain_fwd_bwd(const float* x, const float* gy, float* y, float* gx, const float* params,
                             int pstride, const int* ids, int S, int B, int H, int W, float* gparams,
                             void* workspace, size_t workspace_bytes, int options, void* stream);
/* The same pass when EVERY image runs the same filter sequence: `ids_host` is a HOST array of S filter ids
 * (the benchmark chain of BASELINE configs[1] / [4]; a fixed retouching recipe applied to a whole batch).
 * Sequences with a compile-time instantiation -- currently the shipped cfg.filters order E,G,W,S+,T,Ct,BW,C
 * (config_example.py:22-25) on 16-byte aligned images with H*W % 4 == 0 -- run a kernel specialised for the
 * sequence (no per-step dispatch, parameter-gradient accumulators in registers across the tile loop, scale
 * steps recomputed instead of parked); every other sequence runs the kernel of exp_filter_chain_fwd_bwd.
 * y and gx are bit-identical between the two, gparams agree to reduction order.  Same workspace. */
int exp_filter_chain_fwd_bwd_uniform(const float* x, const float* gy, float* y, float* gx, const float* params,
                                     int pstride, const int* ids_host, int S, int B, int H, int W, float* gparams,
                                     void* workspace, size_t workspace_bytes, int options, void* stream);

/* Bytes of device workspace exp_filter_bwd needs for this shape: a fixed 256 KiB block of
 * per-image ticket counters followed by the partial-sum records of the per-image parameter
 * gradients.  The counter block must be zero-filled once before first use; every launch
 * leaves it zeroed, so one workspace can be reused across shapes and filters (but not by two
 * launches that may run concurrently). */
size_t exp_filter_bwd_workspace_bytes(int B, int H, int W);

/* ---- Filter.process backward ----------------------------------------------------------
 * Given gy = dL/dy, computes
 *   gparams[b, 0:n] = dL/dparams[b]   (what tf.gradients builds for filters.py process();
 *                                      deterministic two-stage reduction, fixed order)
 *   gx = dL/dx                        (nullable; needed only for the N-step chain of
 *                                      BASELINE.json -- the reference never differentiates
 *                                      w.r.t. the image, replay_memory.py:16 placeholder).
 * Outputs are recomputed from x, not re-read.  gx may alias gy (in place).
 * Algorithmic HBM traffic: 36 B / pixel with gx, 24 B / pixel without. */
int exp_filter_bwd(const float* x, const float* gy, float* gx, float* gparams,
                   const float* params, int pstride, const int* ids, int uniform_id,
                   int B, int H, int W, void* workspace, size_t workspace_bytes, int options,
                   void* stream);

/* ---- Filter.apply with a spatial mask (cfg.masking == True) -------------------------------
 * Replaces filters.py:62-99 (`lerp(img, self.process(img, p), self.get_mask(img, mask_p))`,
 * util.py:307-308) with Filter.get_mask (filters.py:110-148) or, for EXP_FILTER_VIGNET,
 * VignetFilter.get_mask (filters.py:354-396) evaluated per pixel in the same pass; the mask
 * image is never materialised unless `mask_out` asks for it.
 *   mask_logits [B, mstride>=6]: the RAW fc2 outputs `features[:, n:]` of extract_parameters
 *     (filters.py:43-44); tanh_range(-5, 5) (filters.py:123-125) is applied in-kernel.  NULL = zeros
 *     (the `specified_parameter` branch, filters.py:73-75).
 *   max_sharpness, min_strength: cfg.maximum_sharpness, cfg.minimum_strength
 *     (config_example.py:37-38);  masking: cfg.masking (0 -> mask == 1, i.e. exp_filter_fwd).
 *   mask_out (nullable) [B,H,W]: debug_info['mask'] / Filter.mask (filters.py:85-87).
 *   y == NULL with mask_out != NULL computes the mask alone (Filter.get_mask); params may be NULL.
 * Other arguments, variants (TMA is served by DIRECT) and EXP_OPT_LOGITS as in exp_filter_fwd. */
int exp_filter_masked_fwd(const float* x, float* y, float* mask_out, const float* params, int pstride,
                          const float* mask_logits, int mstride, const int* ids, int uniform_id, int B,
                          int H, int W, float max_sharpness, float min_strength, int masking, int options,
                          void* stream);
/* Backward of the above: gparams as exp_filter_bwd (the mask scales the filter's own gradient),
 * gmask_logits [B, mstride] = dL/dmask_logits (first 6 entries written; what tf.gradients builds
 * through get_mask), gx (nullable) = dL/dx including the path through the mask's luminance term.
 * Same workspace as exp_filter_bwd. */
int exp_filter_masked_bwd(const float* x, const float* gy, float* gx, float* gparams, float* gmask_logits,
                          const float* params, int pstride, const float* mask_logits, int mstride,
                          const int* ids, int uniform_id, int B, int H, int W, float max_sharpness,
                          float min_strength, int masking, void* workspace, size_t workspace_bytes,
                          int options, void* stream);

/* ======================================================================================
 * Policy / critic / value network primitives.
 * Replace ly.conv2d (kernel 4, stride 2, SAME, lrelu) and ly.fully_connected of
 * agent.py:11-37, 87-99, critics.py:6-38, 94-97, filters.py:28-44 and their tf.gradients,
 * including the WGAN-GP second-order term of net.py:174-194 (tangent modes).
 * Activations NHWC, conv weights HWIO [4,4,Cin,Cout], FC weights [in,out] -- the layouts
 * of the reference checkpoint.  IH, IW must be powers of two (cfg.source_img_size = 64).
 * lrelu(v) = 0.6 v + 0.4 |v| (util.py:225-229); its derivative is recovered from the sign
 * of the stored output (1, 0.2, or 0.6 at exactly 0).
 * ==================================================================================== */

/* GEMM engine of the conv / FC primitives: 0 (default) = TMA-fed tcgen05 tensor cores (kind::tf32 with the
 * 3xTF32 split, accumulators in TMEM; operands by cp.async.bulk.tensor im2col boxes) for every contraction
 * whose shape the engine takes, the exact-fp32 CUDA-core engine for the rest; 1 = the CUDA-core engine
 * everywhere (A/B switch of the tests).  Process-wide. */
int exp_set_gemm_backend(int backend);

/* y[B,IH/2,IW/2,Cout] = epi( conv4x4s2( concat(x[B,IH,IW,Cx], tile(vec[B,Cv])) - shift ) )
 *   `vec` (nullable when Cv == 0) is a per-image vector broadcast over the pixels: the
 *   states of util.enrich_image_input (util.py:31-36) and the 3 global statistics of
 *   critics.py:48-87, so the concatenated input tensor is never materialised; `shift` is the
 *   `net - 0.5` of agent.py:12 / critics.py:7 (0 for inner layers).
 *   mode 0: epi(v) = lrelu(v + bias[co])                      (forward)
 *   mode 1: epi(v) = v * lrelu'(mask_ref[...])  (no bias)     (forward-mode tangent, GP)
 *   y2 (nullable) additionally receives y * post_mul (tf.nn.dropout mask*2, agent.py:36). */
int exp_conv_fwd(const float* x, int Cx, const float* vec, int Cv, float shift, const float* W,
                 const float* bias, const float* mask_ref, const float* post_mul, float* y, float* y2,
                 int B, int IH, int IW, int Cout, int mode, void* stream);

/* dx[B,IH,IW,Cin] = conv4x4s2^T(dy[B,IH/2,IW/2,Cout]) (* lrelu'(a_in) when a_in != NULL):
 * Conv2DBackpropInput fused with the previous layer's activation derivative. */
int exp_conv_dgrad(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH,
                   int IW, int Cin, int Cout, void* stream);

/* gW[4,4,Cin,Cout] = Conv2DBackpropFilter(input as in exp_conv_fwd, dy); deterministic
 * split-K through `workspace` (size from exp_conv_wgrad_workspace_bytes); accumulate != 0
 * adds to gW instead of overwriting (gradient-penalty term, several losses). */
size_t exp_conv_wgrad_workspace_bytes(int B, int IH, int IW, int Cin, int Cout);
int exp_conv_wgrad(const float* x, int Cx, const float* vec, int Cv, float shift, const float* dy,
                   float* gW, int B, int IH, int IW, int Cout, int accumulate, void* workspace,
                   size_t workspace_bytes, void* stream);

/* First convolution layer on the tensor cores.  Its input is concat(image[3], per-image state
 * constants) - shift (util.py enrich_image_input, agent.py:17 / critics.py:51) with Cin = 6 / 14 /
 * 17 channels: not a multiple of 32, so it runs on a STAGING COPY
 *   xp[B][IH+2][IW+2][16]: zero border (ly.conv2d SAME padding made explicit), channels [0,Cin) =
 *   enriched input - shift, channels [Cin,16) = 0            (exp_conv1_pad_input)
 * and zero-padded weights Wp[4][4][16][Cout] (exp_conv1_pad_weights), both consumed by TMA im2col
 * boxes.  exp_conv1_fwd has the epilogue modes of exp_conv_fwd; exp_conv1_wgrad writes the
 * UNPADDED gradient gW[4][4][Cin][Cout] (deterministic split-K through `workspace`).
 * exp_conv1_supported(Cin, Cout) != 0 when this path can be used (Cin <= 16, Cout % 32 == 0). */
int exp_conv1_supported(int Cin, int Cout);
size_t exp_conv1_padded_input_elems(int B, int IH, int IW);
int exp_conv1_pad_input(const float* x, int Cx, const float* vec, int Cv, float shift, float* xp, int B,
                        int IH, int IW, void* stream);
int exp_conv1_pad_weights(const float* W, int Cin, int Cout, float* Wp, void* stream);
int exp_conv1_fwd(const float* xp, const float* Wp, const float* bias, const float* mask_ref,
                  const float* post_mul, float* y, float* y2, int B, int IH, int IW, int Cout, int mode,
                  void* stream);
size_t exp_conv1_wgrad_workspace_bytes(int B, int IH, int IW, int Cout);
int exp_conv1_wgrad(const float* xp, const float* dy, float* gW, int Cin, int B, int IH, int IW, int Cout,
                    int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* First convolution layer, constant channels split off (the product path for Cx = 3, Cout = 32; csrc/conv_first.cu).
 * Only the image channels vary over the pixels; the Cv tiled channels (util.py:31-36 enrich_image_input: agent states,
 * critics.py:51-87 statistics) are constants of an image, so their share of an output is one of <= 16 per-image values
 * per output channel (which taps fall inside the image depends only on first / last row and column).  Exact-fp32
 * CUDA-core kernels over the image itself, K = 48 instead of 16 (3 + Cv); same epilogue modes as exp_conv_fwd
 * (agent.py:21-33 / critics.py:13-19 forward; mode 1 = forward-mode tangent for the gradient penalty, net.py:181-194).
 *   exp_conv_first_fwd    y (and y2 = y * post_mul) [B, IH/2, IW/2, 32] from x [B,IH,IW,3], vec [B,Cv], W [4,4,3+Cv,32]
 *   exp_conv_first_wgrad  gW [4,4,3+Cv,32] (=|+=), deterministic two-pass reduction through `workspace`
 *   exp_conv_first_dgrad  dx_img [B,IH,IW,3] (nullable) = gradient w.r.t. the image channels, and gvec [B,Cv] (nullable)
 *                         = the gradient of each constant channel SUMMED over the pixels -- what tf.gradients hands to
 *                         the per-image producers of those channels (the statistics: exp_stats_bwd)
 * IH, IW even (powers of two for dx_img); y / y2 / mask_ref / post_mul / dy / W / workspace 16-byte aligned. */
int exp_conv_first_supported(int Cx, int Cv, int Cout);
int exp_conv_first_fwd(const float* x, const float* vec, int Cv, float shift, const float* W, const float* bias,
                       const float* mask_ref, const float* post_mul, float* y, float* y2, int B, int IH, int IW,
                       int mode, void* stream);
size_t exp_conv_first_wgrad_workspace_bytes(int B, int IH, int IW);
int exp_conv_first_wgrad(const float* x, const float* vec, int Cv, float shift, const float* dy, float* gW, int B,
                         int IH, int IW, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
int exp_conv_first_dgrad(const float* dy, const float* W, int Cv, float* dx_img, float* gvec, int B, int IH, int IW,
                         void* stream);

/* First layers with 16 < Cin <= 32 (value network: 17 channels): materialise the enriched input
 * out[B][IH][IW][32] = concat(x, tile(vec)) - shift, zero above Cin (exp_conv_enrich32) and the
 * weights padded to Wp[4][4][32][Cout] (exp_conv_pad_weights32); exp_conv_fwd / exp_conv_wgrad then
 * run on them with Cx = 32, Cv = 0, shift = 0 (the TMA-fed tensor-core path). */
int exp_conv_enrich32(const float* x, int Cx, const float* vec, int Cv, float shift, float* out, int B,
                      int IH, int IW, void* stream);
int exp_conv_pad_weights32(const float* W, int Cin, int Cout, float* Wp, void* stream);

/* FC layers take leading dimensions (row strides, in floats) so that several heads can share
 * one activation matrix: element (m,k) of x is x[m*ldx + k], etc.
 * y[M,N] = epi(x[M,K] W[K,N]); mode 0: lrelu(v+bias), 1: v*lrelu'(mask_ref) (tangent),
 * 2: v+bias, 3: v.  Deterministic split-K through `workspace`. */
size_t exp_fc_workspace_bytes(int M, int K, int N);
int exp_fc_fwd(const float* x, int ldx, const float* W, const float* bias, const float* mask_ref,
               int ldmask, float* y, int ldy, int M, int K, int N, int mode, void* workspace,
               size_t workspace_bytes, void* stream);
/* dx[M,K] = (accumulate ? dx : 0) + dy[M,N] W^T * lrelu'(mul_act) * mul_plain
 * (either multiplier may be NULL; mul_plain is the tf.nn.dropout mask*2 of agent.py:36). */
int exp_fc_dgrad(const float* dy, int ldy, const float* W, const float* mul_act, const float* mul_plain,
                 int ldmul, float* dx, int lddx, int M, int K, int N, int accumulate, void* stream);
/* gW[K,N] = (accumulate ? gW : 0) + x^T dy */
int exp_fc_wgrad(const float* x, int ldx, const float* dy, int ldy, float* gW, int M, int K, int N,
                 int accumulate, void* stream);
/* out[batch, cols] = column sums of a[batch, rows, cols] (bias gradients; per-image channel
 * sums of the layer-1 input gradient that feed exp_stats_bwd).  Long columns are reduced in
 * row chunks through `workspace` (deterministic two-stage; NULL workspace = single pass). */
size_t exp_colsum_workspace_bytes(int batch, int rows, int cols);
int exp_colsum(const float* a, int batch, int rows, int cols, float* out, void* workspace,
               size_t workspace_bytes, void* stream);

/* ======================================================================================
 * Per-image head math of the train step (all enqueue-only, no host round trip).
 * ==================================================================================== */

/* stats[B,3] = (mean luminance, population variance of luminance, mean saturation) of
 * critics.py:48-76 (the three channels the critic / value CNN input is enriched with). */
int exp_stats_fwd(const float* img, float* stats, int B, int H, int W, void* stream);
/* g_out[B,H,W,3] = (g_direct ? g_direct : 0) + J_stats^T g_stat   (tf.gradients through
 * tf.nn.moments / reduce_max / reduce_min / clip_by_value with TF's tie rules). */
int exp_stats_bwd(const float* img, const float* stats, const float* g_stat, const float* g_direct,
                  float* g_out, int B, int H, int W, void* stream);
/* dstat[B,3] = J_stats u : forward-mode tangent used by the gradient-penalty term. */
int exp_stats_jvp(const float* img, const float* stats, const float* u, float* dstat, int B, int H,
                  int W, void* stream);

/* Action-selection head, agent.py:100-122 + 208-252 + pdf_sample_layer.py:5-10:
 * logits[B,n] (selector_fc2 output), noise[B] (= z[:,0]), states[B,3+n] ->
 * pdf[B,n], ids[B] (int32; -1 reproduces the u==0 quirk of pdf_sample), surrogate[B],
 * entropy[B], penalty_head[B] (entropy + filter-usage penalties), new_states[B,3+n].
 * `progress` (replay_memory.py:33 placeholder, iter / max_iter_step) is a DEVICE scalar. */
int exp_policy_head_fwd(const float* logits, const float* noise, const float* states, int B,
                        int n_filters, int n_states, int is_train, int test_steps, float exploration,
                        float exploration_penalty, float filter_usage_penalty, const float* progress,
                        float* pdf, int* ids, float* surrogate, float* entropy, float* penalty_head,
                        float* new_states, void* stream);
int exp_policy_head_bwd(const float* logits, const int* ids, const float* g_surrogate,
                        const float* g_penalty, int B, int n_filters, float exploration,
                        float exploration_penalty, const float* progress, float* g_logits, void* stream);

/* pen[b] = mean(max(img-1,0)^2) (agent.py:247) and its backward
 * g_out = (g_in ? g_in : 0) + g_pen[b] * 2 max(img-1,0) / (H*W*3). */
int exp_overexposure_fwd(const float* img, float* pen, int B, int H, int W, void* stream);
int exp_overexposure_bwd(const float* img, const float* g_pen, const float* g_in, float* g_out, int B,
                         int H, int W, void* stream);

/* Reward / TD / advantage / losses of net.py:92-163 and the gradient seeds of g_loss and
 * v_loss: seeds[5][B] = d g_loss/d fake_logit, d g_loss/d new_value, d v_loss/d old_value,
 * d g_loss/d penalty, d g_loss/d surrogate; losses[2] = (g_loss, v_loss). */
int exp_rl_losses(const float* fake_logit, const float* fake_input_logit, const float* old_value,
                  const float* new_value, const float* penalty, const float* surrogate,
                  const float* new_states, int B, int n_states, float all_reward,
                  float critic_logit_multiplier, float discount_factor, float parameter_lr_mul,
                  int max_traj_len, int use_penalty, float* seeds, float* losses, void* stream);

/* WGAN-GP helpers, net.py:174-187: out = real + alpha[b] (fake - real) over n floats per
 * image; norm[b] = sqrt(1e-6 + sum g^2), u = g * lambda * 2 max(norm-1,0) / (B norm). */
int exp_interpolate(const float* real, const float* fake, const float* alpha, float* out, int B, int n,
                    void* stream);
int exp_gp_scale(const float* g, float* u, float* norm, float lambda, int B, int n, void* stream);

/* ---- fused bookkeeping kernels of the train step (csrc/train_glue.cu) ---------------------------------------
 * Each replaces a chain of element-wise / indexing launches of the TF graph (and of a tensor-library port of it).
 *
 * exp_critic_inputs: X[0:B] = real, X[B:2B] = fake, X[2B:3B] = real + alpha[b] (fake - real): the three batches the
 *   critic step scores (net.py:68-71, 174-179) as ONE batch; n floats per image (n % 4 == 0, 16-byte aligned).
 * exp_critic_scalars: from logits [3B] (real | fake | interpolated) and norm [B] (exp_gp_scale) writes out[0..4] =
 *   emd (net.py:164), gradient penalty lambda mean(max(norm-1,0)^2) (net.py:185-187), critic_gradient_norm =
 *   mean(norm), c_loss = -emd + gp (net.py:151,194), c_average (net.py:165) and, when ema_state != NULL, advances
 *   tf.train.ExponentialMovingAverage(decay, zero_debias=True) of c_average (net.py:119-120, 166-168, 268-269):
 *   ema_state = [debiased value, biased accumulator, local_step]. */
int exp_critic_inputs(const float* real, const float* fake, const float* alpha, float* X, int B, int n, void* stream);
int exp_critic_scalars(const float* logits, const float* norm, int B, float lambda, float* ema_state, float decay,
                       float* out, void* stream);

/* The fc2 layers of the n_heads filter heads (filters.py:39-44, one Filter per cfg.filters entry, agent.py:58-72) as
 * single launches.  Head j's weights [fc1, dims[j]] and biases [dims[j]] live at float offsets w_off[j] / b_off[j]
 * of the flat generator parameter buffer `params` (gradients at the same offsets of `grads`); dims[j] = npar[j]
 * filter parameters + nmask mask parameters (filters.py:43-44 split).  *_host arrays are HOST arrays of n_heads ints.
 *   exp_heads_fc2_fwd:  O[b, j, 0:dims[j]] = H[b, j*fc1 : (j+1)*fc1] W_j + bias_j, zeros up to ostride
 *                       (H [B, ldh] = the lrelu outputs of the heads' fc1 layers, one column block per head)
 *   exp_heads_select:   sel[b, 0:npar[id]] = O[b, id, 0:npar[id]] (zeros up to selstride), msel[b, 0:nmask] (nullable)
 *                       = O[b, id, npar[id] : npar[id]+nmask], id = ids[b]; id -1 -> zeros.  This is the one-hot select
 *                       of agent.py:113-125 applied to the head outputs instead of to 8 filtered images.
 *   exp_heads_fc2_bwd:  from gsel [B, selstride] = dL/dsel and gmsel [B, nmask] (nullable) writes
 *                       dH [B, ldh] = dL/d(fc1 pre-activation) (lrelu' applied; zero for the heads an image did not
 *                       select) and OVERWRITES the fc2 weight / bias gradients of every head in `grads`. */
int exp_heads_fc2_fwd(const float* params, const int* w_off_host, const int* b_off_host, const int* dims_host,
                      const int* npar_host, int n_heads, int fc1, int nmask, const float* H, int ldh, float* O,
                      int ostride, int B, void* stream);
int exp_heads_select(const float* O, int ostride, const int* ids, const int* npar_host, int n_heads, int nmask,
                     float* sel, int selstride, float* msel, int B, void* stream);
int exp_heads_fc2_bwd(const float* params, float* grads, const int* w_off_host, const int* b_off_host,
                      const int* dims_host, const int* npar_host, int n_heads, int fc1, int nmask, const float* H, int ldh,
                      const int* ids, const float* gsel, int selstride, const float* gmsel, float* dH, int B,
                      void* stream);

/* Up to 8 column sums in ONE launch: dst_t[cols_t] (=|+= when accumulate_t) sum over rows of src_t[rows_t, cols_t]
 * -- the bias gradients of a whole CNN backward (4 conv layers + 2 FC layers).  Deterministic: per-chunk partials,
 * the last block to arrive for a column block adds them in chunk order.  *_host: HOST arrays of n entries.
 * workspace: exp_colsum_multi_workspace_bytes() (0 = bad task list), zero-filled once: a fixed 4 KiB block of
 * self-cleaning ticket counters followed by the partial sums, so one workspace serves any sequence of task lists. */
size_t exp_colsum_multi_workspace_bytes(const int* rows_host, const int* cols_host, int n);
int exp_colsum_multi(const float* const* src_host, float* const* dst_host, const int* rows_host, const int* cols_host,
                     const int* accumulate_host, int n, void* workspace, size_t workspace_bytes, void* stream);

/* exp_stats_bwd fed directly with the layer-1 input gradient g_in [B,H,W,cin] (exp_conv_dgrad into the enriched
 * input): g_stat[b] = sum over pixels of the last three channels (the tiled statistics, critics.py:77-87),
 * g_out[B,H,W,3] = g_in[..., 0:3] + J_stats^T g_stat.  One launch instead of colsum + two slices + exp_stats_bwd. */
int exp_stats_bwd_gin(const float* img, const float* stats, const float* g_in, int cin, float* g_out, int B, int H,
                      int W, void* stream);

/* dst[0..n) = host_vals[0..n), n <= 16, passed to the kernel BY VALUE: the per-iteration scalars of a captured train
 * iteration (Adam's lr_t of every optimizer step, net.py:224; `progress`, agent.py:228-252) in one launch; the host
 * array may be reused as soon as the call returns. */
int exp_set_floats(float* dst, const float* host_vals, int n, void* stream);

/* ---- device-side replay memory (csrc/replay.cu, csrc/replay_logic.cuh) -----------------------------------------
 * Replaces the host list handling of replay_memory.py:187-273 (random.shuffle, list slicing, np.stack and a feed of
 * every image per sess.run) by index lists computed on the device.  Records live in ONE buffer of three regions
 * addressed by a flat index: [0, P) the pool (P = cfg.replay_memory_size), [P, P+B) the outputs of the last
 * generator step (B = cfg.batch_size), [P+B, 2P+2B) the fresh RAW records of this iteration (states == 0).
 *   exp_replay_draw_generator  get_next_fake_batch (230-246): batch_src int64 [B] = the first B non-terminated records
 *       of a random pool order (terminated ones met on the way are dropped; a pool that runs dry is rebuilt from fresh
 *       records, 64-75, 237-238); rest_src int32 [P] + ctl = the records that stay.
 *   exp_replay_replace         replace_memory (187-196) + fill_pool: new_pool_src int64 [P] = the remaining records, the
 *       generator outputs with step < max_traj_len (others with probability keep_prob), fresh records up to P.
 *   exp_replay_draw_critic     replay_fake_batch (249-273): batch_src int64 [B] = terminated records in random order,
 *       cycling when fewer than B exist; none at all sets ctl[4] (the reference asserts).
 * pool_states / new_states: float [P or B][n_states] (STOPPED = column 1, STEP = column 2, util.py:13-16).
 * ctl: int32 [exp_replay_ctl_words()], zero-filled once: 64-bit call counter of the Philox streams (a replayed CUDA
 * graph draws new numbers each time), hand-over between draw_generator and replace, error flag.  `seed` keys the
 * streams.  exp_gather_rows moves the records: dst[i, :] = src[idx[i], :] for rows of `row` floats.
 * exp_train_draws: the per-step random inputs of the train step in one launch -- uniform[n_uniform] ~ U[0,1) (z[:,0],
 * alpha) and mask[n_mask] = floor(keep + U) / keep (tf.nn.dropout, agent.py:36) -- from the same counter. */
int exp_replay_ctl_words(void);
int exp_replay_draw_generator(const float* pool_states, int n_states, int pool, int batch, unsigned long long seed,
                              int* ctl, long long* batch_src, int* rest_src, void* stream);
int exp_replay_replace(const float* new_states, int n_states, int pool, int batch, int max_traj_len, float keep_prob,
                       unsigned long long seed, int* ctl, const int* rest_src, long long* new_pool_src, void* stream);
int exp_replay_draw_critic(const float* pool_states, int n_states, int pool, int batch, unsigned long long seed,
                           int* ctl, long long* batch_src, void* stream);
int exp_gather_rows(const float* src, const long long* idx, float* dst, int n, int row, void* stream);
int exp_train_draws(unsigned long long seed, int* ctl, float* uniform, int n_uniform, float* mask, size_t n_mask,
                    float keep, void* stream);

/* ---- data-parallel optimizer step over NVLink peer memory (one process per GPU) -------------------------
 * The reference is single-GPU; this is the collective of SURVEY 8e / C1: ONE exchange per
 * optimizer step on the flat gradient buffer -- theta_g and theta_v together, both optimizers run in the same
 * sess.run (net.py:330-331); theta_c in each critic step (net.py:362) -- fused with the Adam update.
 * All-reduce (sum, fixed rank order) of `n` floats across `world` GPUs followed by Adam on the local replica
 * with grad_scale = 1/world, in ONE kernel: reduce-scatter and all-gather through peer loads over NVLink, two
 * flag barriers, no host involvement (the launch is a plain kernel node of the step's CUDA graph).
 *   grads_host / red_host / flags_host: HOST arrays of `world` DEVICE pointers -- entry q is rank q's gradient
 *     buffer [n], reduction scratch [n] and flag block (exp_dp_flag_bytes() bytes, zero-filled once), mapped into
 *     this process (cudaIpcOpenMemHandle; entry `rank` is this rank's own memory).  Every rank must pass the
 *     buffers in the same rank order and launch the same sequence of calls per flag block.
 *   n % (4 * world) == 0; elements [0, n_a) use lr_t = hyper_a[0], elements [n_a, n) hyper_b[0] (device scalars,
 *     as in exp_adam; n_a % 4 == 0; hyper_b may be NULL when n_a == n).
 * The replicas stay bit-identical: each element is reduced by one rank and read by all.  world <= exp_dp_max_world().
 *
 * IPC plumbing (host side, once per buffer): exp_dp_ipc_export returns the CUDA IPC handle
 * (exp_dp_ipc_handle_bytes() bytes) of the device allocation that contains `dev_ptr` and the offset of `dev_ptr`
 * inside it; the peers receive both (any host channel), exp_dp_ipc_open maps the allocation into the calling
 * process -- in the CURRENT device's context, with peer access to the owning GPU (cudaIpcMemLazyEnablePeerAccess)
 * -- and returns its base; the buffer is at base + offset.  One open per handle and process; exp_dp_ipc_close
 * unmaps.  These are the only calls of the library that touch the CUDA memory manager. */
size_t exp_dp_ipc_handle_bytes(void);
int exp_dp_ipc_export(const void* dev_ptr, void* handle_out, size_t* offset_out);
int exp_dp_ipc_open(const void* handle, void** base_out);
int exp_dp_ipc_close(void* base);
size_t exp_dp_flag_bytes(void);
int exp_dp_max_world(void);
int exp_dp_allreduce_adam(float* params, float* m, float* v, const float* const* grads_host, float* const* red_host,
                          unsigned* const* flags_host, int world, int rank, const float* hyper_a, size_t n_a,
                          const float* hyper_b, size_t n, float beta1, float beta2, float eps, void* stream);

/* Fused Adam over one flat buffer (tf.train.AdamOptimizer, config_example.py:158):
 * hyper[0] (device) = lr * sqrt(1-beta2^t) / (1-beta1^t); g is multiplied by grad_scale. */
int exp_adam(float* params, const float* grads, float* m, float* v, const float* hyper, float beta1,
             float beta2, float eps, float grad_scale, size_t n, void* stream);

/* ======================================================================================
 * Host-side helper (no device work): CRC-32C of `n` bytes, continuing from `crc` (0 to start).
 * The checksum of TensorFlow checkpoint bundles (table blocks, BundleEntryProto.crc32c), used by
 * exposure_b200/tf_bundle.py to write / verify files tf.train.Saver can restore (net.py:271,405).
 * ==================================================================================== */
unsigned int exp_crc32c(unsigned int crc, const void* data, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* EXPOSURE_B200_H_ */
