/* egaze.h -- C-ABI of libegaze.so: the B200-native (sm_100a) replacement for what PyTorch->cuDNN/oneDNN
 * executes underneath the reference's hot-path modules.  The reference (hyf015/egocentric-gaze-prediction) is pure
 * Python; its "FFI for this path" is the nn.Module surface.  The Python host mirror in
 * egocentric-gaze-prediction_b200/{models,floss.py,utils.py} keeps that surface and calls ONLY these entry points
 * through ctypes (egocentric-gaze-prediction_b200/egaze/_lib.py); INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns int: 0 = ok, < 0 = invalid argument / unsupported, > 0 = cudaError_t;
 *     egaze_last_error() returns the message of the calling thread's last failure.
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch caching allocator); the library never
 *     allocates or frees caller-visible memory.  `stream` is a cudaStream_t passed as void*.
 *   - launches are asynchronous and stream-ordered; no host synchronisation inside any entry point.
 *   - there is NO CPU fallback: a non-sm_100 device is an error (egaze_check_device).
 *   - "split" activations: two NHWC 16-bit planes hi and lo with x ~= hi + lo.  `fmt` / `*_f16` arguments name the element
 *     format: 0 = bf16 (16 significand bits in all, fp32's exponent range: gradients), 1 = fp16 (22 bits: the forward
 *     operands -- train-mode BatchNorm needs >= 19 for the 1e-3 parity gate).  lo may be NULL where
 *     documented.  An "xb" plane is bf16(x): the copy the weight-gradient GEMM reads when hi / lo are fp16 (kind::f16 tensor
 *     core instructions need both operands in the same format, and gradients stay bf16).
 */
#ifndef EGAZE_H_
#define EGAZE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------------------------------ */
int egaze_version(void);
int egaze_last_error(char* buf, int n);
int egaze_check_device(void);
int egaze_sm_count(int* out);

/* ---- layout (API tensors are NCHW fp32: reference SURVEY 8b "Tensor conventions") ---------------------------- */
/* x [N][C][H][W] fp32 -> hi/lo [N][H][W][Cp] bf16, channels C..Cp-1 zero.  (input side of utils.py:70 / model_SP.py:36-37) */
/* xb (optional): bf16(x) plane [N][H][W][Cp_xb] with its own channel stride (the weight-gradient GEMM reads 64-channel rows) */
int egaze_nchw_to_nhwc_split(const float* x, int N, int C, int H, int W, int Cp, void* hi, void* lo, void* xb, int Cp_xb,
                             int fmt, void* stream);
/* (hi[,lo]) or f32, NHWC with channel stride Cs -> out [N][C][H][W] fp32 (what forward hooks / callers see: AT.py:22,226) */
int egaze_nhwc_to_nchw(const void* hi, const void* lo, const float* f32, int N, int C, int H, int W, int Cs, int fmt,
                       float* out, void* stream);
int egaze_nchw_to_nhwc_f32(const float* x, int N, int C, int H, int W, float* out, void* stream);
int egaze_f32_to_split(const float* x, long long n, void* hi, void* lo, int fmt, void* stream);
/* nn.Conv2d weight OIHW fp32 -> packed split planes.  mode 0 (fprop): [9][Cout][cols_p>=Cin];
 * mode 1 (dgrad): taps flipped, [9][Cin][cols_p>=Cout].  fmt 1: fp16 planes of w * egaze_f16_weight_scale (the conv is then
 * called with acc_scale = 1 / that).  (weights of utils.py:70, model_SP.py:10,13-30) */
int egaze_f16_weight_scale(float* out);
/* mode 2 / 3: the SUB-PIXEL copies of a conv that follows nn.Upsample(scale_factor=2) (model_SP.py:16-17,20-21,24-25,27-28):
 * 16 planes [phase*4 + a*2 + b] of 2x2-tap weights pre-summed in fp32 -- output phase (py, px) of the upsample+conv pair is a
 * 2x2 conv of the LOW-resolution map with taps W2[py][a] = {py=0: w[0], w[1]+w[2]; py=1: w[0]+w[1], w[2]} (same along x);
 * mode 2 rows = Cout (forward), mode 3 rows = Cin (data gradient, transposed).  hi / lo: [16][rows][cols_p]. */
int egaze_pack_w3x3(const float* w_oihw, int Cout, int Cin, int cols_p, int mode, int fmt, void* hi, void* lo, void* stream);
/* egaze_pack_w3x3 for many weights in ONE launch.  jobs: device array of njobs 48-byte records
 * {const float* w; void* hi; void* lo; int Cout, Cin, rows, cols_p, mode, fmt;} (rows = Cout for mode 0, Cin for mode 1). */
int egaze_pack_w3x3_multi(const void* jobs, int njobs, void* stream);
/* wgrad accumulator [9][Cout_p][Cin_p] fp32 -> OIHW grad, gw = beta*gw + dw; clear != 0 zeroes the accumulator afterwards.
 * sub != 0: the accumulator holds the 16 sub-pixel planes [phase*4 + a*2 + b] (egaze_wgrad3x3_tc sub = 1); every 3x3 tap
 * gradient is the sum of the four planes it was pre-summed into (transpose of the egaze_pack_w3x3 mode-2 map). */
int egaze_unpack_wgrad(float* dwp, int Cout, int Cin, int Cout_p, int Cin_p, float beta, int clear, int sub, float* gw_oihw,
                       void* stream);

/* ---- 3x3 convolution, tcgen05 implicit GEMM (replaces nn.Conv2d(k=3,p=1): utils.py:70, model_SP.py:10,13-30) - */
/* Tile geometry the kernel will use for an (N,H,W) map; num_tiles sizes the BN-statistics workspace. */
int egaze_conv3x3_tiles(int N, int H, int W, int need_even, int* num_tiles, int* BH, int* BW);
/* BatchNorm-statistics workspace of egaze_conv3x3_tc: stats [partials][2][Cout], stats_cnt [partials][cnt_stride] (zeroed by
 * the caller); one (mean, M2, n) partial per persistent CTA. */
int egaze_conv3x3_stats_shape(int Cout, int precise, int* partials, int* cnt_stride, int* cnt_div);
/* Debug aid (tools/conv_prof.py): per-CTA cycle counters of the following egaze_conv3x3_tc launches go to buf
 * ([grid][16] int64 on the device); null switches it off.  Not used on the product path. */
int egaze_conv3x3_set_prof(void* buf);
/* y = epilogue(conv3x3(x, w)):  v = acc + bias; v = v*scale + shift; relu; 2x2 reduce (1 max = MaxPool2d utils.py:68,
 * 2 sum = grad of nn.Upsample); mask (zero where mask <= 0: ReLU backward; mask_ups: the mask tensor is stored 2x upsampled);
 * 2x nearest replicate (ups: model_SP.py:16,20,24,27).
 *   x_hi/x_lo : [N][H][W][Cin_p]             w_hi/w_lo : [9][Cout][Cin_p] (egaze_pack_w3x3); all four in the format in_f16
 *   operand modes (MMAs per product): x_hi,x_lo,w_hi,w_lo -> hi*hi + hi*lo + lo*hi (3);  x_lo NULL -> x_hi*[w_hi|w_lo] (2:
 *   the data-gradient mode);  x_lo and w_lo NULL -> 1
 *   out_f32 / out_hi / out_lo / out_xb : NHWC [N][Ho][Wo][Cout] (any non-NULL subset is written; out_hi / out_lo in the format
 *   out_f16, out_lo optional for bf16; out_xb = bf16(v), fp16 outputs only)
 *   acc_scale : the accumulators are multiplied by it first (1 / egaze_f16_weight_scale for fp16 weights, else 1)
 *   stats / stats_cnt : per-CTA (mean, M2, n) partials of (acc + bias) for BatchNorm batch statistics (egaze_conv3x3_stats_shape)
 *   colsum : optional [Cout] fp32, ACCUMULATED: += sum over all output pixels of the final (masked) values.  When the
 *            kernel computes a data gradient this is the bias gradient of the conv that produced the masked activation,
 *            so no separate reduction pass over the gradient tensor is needed.
 *   sub : sub-pixel decomposition of nn.Upsample(scale_factor=2) -> conv3x3 (model_SP.py:16-17,20-21,24-25,27-28), 16 instead of
 *            36 MACs per low-resolution pixel and weight.  1 = forward: x is the LOW-resolution [N][H][W][Cin_p] map, w the mode-2
 *            pack, the output is [N][2H][2W][Cout].  2 = data gradient: x is the PHASE-PLANAR output gradient
 *            [4N][H][W][Cin_p] (image (py*2+px)*N + n = dY[n, 2i+py, 2j+px]), w the mode-3 pack, the output is the gradient
 *            w.r.t. the low-resolution map [N][H][W][Cout] (the 2x2 sum of the Upsample gradient is part of the K loop).
 *            Cin_p % 64 == 0; no reduce / ups / stats.
 *   out_planar : store the [N][H][W][Cout] output phase-planar as [4N][H/2][W/2][Cout] (what a sub = 2 launch and the sub-pixel
 *            weight gradient read); H, W even.
 */
int egaze_conv3x3_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H, int W,
                     int Cin_p, int Cout, const float* bias, const float* scale, const float* shift, int relu,
                     int reduce, int ups, const void* mask, int mask_ups, float* out_f32, void* out_hi, void* out_lo,
                     void* out_xb, float* stats, float* stats_cnt, float* colsum, int in_f16, int out_f16, float acc_scale,
                     int sub, int out_planar, void* stream);
/* Plans: freeze one egaze_conv3x3_tc call (pointers, shapes, flags) with its encoded TMA descriptors (cuTensorMapEncodeTiled x4)
 * and tile configuration; run it with (plan, stream).  Valid while the buffers it names are alive; destroy with
 * egaze_plan_destroy.  ("egaze_plan_{create,destroy}: caches CUtensorMap descriptors per (ptr, shape)") */
int egaze_conv3x3_plan_create(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H, int W,
                              int Cin_p, int Cout, const float* bias, const float* scale, const float* shift, int relu,
                              int reduce, int ups, const void* mask, int mask_ups, float* out_f32, void* out_hi, void* out_lo,
                              void* out_xb, float* stats, float* stats_cnt, float* colsum, int in_f16, int out_f16,
                              float acc_scale, int sub, int out_planar, long long* plan);
int egaze_conv3x3_plan_run(long long plan, void* stream);
int egaze_plan_destroy(long long plan);
/* Weight gradient: dwp[9][Cout][Cin_p] (fp32, ACCUMULATED: zero first) += sum_pixels dY (x) X-window; tcgen05 GEMM with the
 * pixel axis as K, MN-major operands straight from NHWC.  Cin_p % 64 == 0, Cout % 64 == 0.  (loss.backward(): SP.py:136) */
/* sub != 0: weight gradient of the sub-pixel form: x is the LOW-resolution [N][H][W][Cin_p] input, dy the phase-planar output
 * gradient [4N][H][W][Cout], dwp the 16-plane accumulator [16][Cout][Cin_p] (-> egaze_unpack_wgrad sub = 1). */
int egaze_wgrad3x3_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int N, int H, int W,
                      int Cin_p, int Cout, float* dwp, int precise, int sub, void* stream);

/* ---- BatchNorm2d pieces (utils.py:72, model_SP.py:12, late_fusion.py:10-12) ---------------------------------- */
/* partial [T][2][C] (mean, M2); count behind (t, c) = cnt[t*cnt_stride + c/cnt_div].  num_batches_tracked (optional, device
 * int64 scalar): incremented by one, like nn.BatchNorm2d's forward in training mode. */
int egaze_bn_finalize(const float* partial, const float* cnt, int cnt_stride, int cnt_div, int T, int C, float eps,
                      float momentum, const float* gamma, const float* beta, float* running_mean, float* running_var, float* mean_out,
                      float* invstd_out, float* scale_out, float* shift_out, long long* num_batches_tracked, void* stream);
int egaze_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                  const float* conv_bias, float eps, int C, float* scale, float* shift, void* stream);
/* x [rows][C] fp32 -> partial [ceil(rows/128)][2][C], cnt [ceil(rows/128)] */
int egaze_col_stats(const float* x, long long rows, int C, float* partial, float* cnt, void* stream);
/* y = relu?(x*scale+shift) [-> MaxPool2d(2,2)] ; x NHWC fp32, outputs NHWC */
int egaze_bn_apply(const float* x, int N, int H, int W, int C, const float* scale, const float* shift, int relu,
                   int pool, float* out_f32, void* out_hi, void* out_lo, void* out_xb, int fmt, void* stream);
/* out[b] = max(x[b], x[b+B]) : Conv3d(1,3,3)+MaxPool3d((2,1,1)) second half (model_SP.py:11,43) */
int egaze_pairmax(const float* x, long long per_stream, float* out, void* stream);
/* backward pieces (autograd of the modules above; loss.backward() in SP.py:136, LF.py:98, spatialstream.py:140) */
/* partial: a PERSISTENT workspace of egaze_bn_bwd_blocks() * 2 * C floats, 8-byte aligned and zero before its first use; every
 * call leaves it zeroed again (fp64 accumulators, cleared by the call's own finalize kernel) */
int egaze_bn_bwd_blocks(int* nblk);
int egaze_bn_bwd_reduce(const float* raw, const float* g, int N, int H, int W, int C, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int pool, int relu, float* partial,
                        float* dgamma, float* dbeta, void* stream);
/* batch_stats = 0: the BatchNorm ran on its running statistics (eval mode): d(raw) = scale * gz, mean / invstd are the
 * running mean and 1/sqrt(running_var + eps).  out_hi / out_lo: bf16 planes (out_lo may be NULL). */
int egaze_bn_bwd_apply(const float* raw, const float* g, int N, int H, int W, int C, const float* scale,
                       const float* shift, const float* mean, const float* invstd, const float* dgamma,
                       const float* dbeta, int pool, int relu, int batch_stats, float* out_f32, void* out_hi, void* out_lo,
                       void* stream);
int egaze_pairmax_bwd(const float* x, const float* g, long long per_stream, void* hi, void* lo, void* stream);
int egaze_col_sum_split(const void* hi, const void* lo, long long rows, int C, float* out, void* stream);

/* ---- 1x1 conv to one channel + sigmoid (model_SP.py:30,32 ; late_fusion.py:13,15) ---------------------------- */
int egaze_head_fwd(const void* x_hi, const void* x_lo, int fmt, const float* w, const float* b, int C, int Cs, long long P,
                   float* out, float* logit_out, void* stream);
/* dx_hi / dx_lo: bf16 planes of the gradient w.r.t. x (dx_lo may be NULL) */
int egaze_head_bwd(const void* x_hi, const void* x_lo, int fmt, const float* w, int C, int Cs, long long P, const float* y,
                   const float* gy, int relu_mask, void* dx_hi, void* dx_lo, float* dw, float* db, void* stream);

/* ---- late fusion (models/late_fusion.py:6-23): cat(f, g) -> [conv3x3 + BN + ReLU] x3 (2->32->32->8) -> conv1x1 -> sigmoid ---
 * One call runs the whole network (csrc/lf.cu: warp-level tensor-core kernels with the BatchNorm affine + ReLU applied while
 * the next layer stages its input window, batch statistics in the conv epilogues).  All pointers are device pointers;
 * w / b / gamma / beta / run_mean / run_var / dw / dgamma / dbeta are HOST arrays of device pointers:
 *   w[4]  = fusion.{0,3,6,9}.weight (OIHW fp32: 32x2x3x3, 32x32x3x3, 8x32x3x3, 1x8x1x1), b[4] the biases (entries may be null)
 *   gamma[3], beta[3], run_mean[3], run_var[3] = fusion.{1,4,7}.* (run_mean / run_var may be null in training mode)
 * training != 0: batch statistics (running statistics updated exactly like nn.BatchNorm2d); 0: running statistics.
 * Workspaces (caller-allocated fp32): raw1, raw2 [B][H][W][32], raw3 [B][H][W][8] (raw conv outputs, kept for the backward;
 * raw3 may be null in eval mode), bn_ws [3][4][32] (per layer: mean, invstd, scale, shift -- outputs), scratch
 * (egaze_lf_scratch floats).  out: [B][1][H][W].  precise != 0: split bf16, 3 MMAs per product. */
int egaze_lf_scratch(int* fwd_floats, int* bwd_floats);
int egaze_lf_fwd(const float* f, const float* g, int B, int H, int W, const float* const* w, const float* const* b,
                 const float* const* gamma, const float* const* beta, float* const* run_mean, float* const* run_var,
                 int training, float eps, float momentum, float* raw1, float* raw2, float* raw3, float* bn_ws,
                 float* scratch, float* out, int precise, void* stream);
/* Backward of egaze_lf_fwd (loss.backward() in LF.py:98).  gout: gradient w.r.t. out.  g1, g2: [B][H][W][32] fp32 workspaces
 * (gradients w.r.t. the post-ReLU activations of layers 1 and 2).  dw[4]: OIHW gradients written directly (null entry = skip
 * that weight gradient); dbh [1]; dgamma[3] / dbeta[3]: BatchNorm parameter gradients (required: they are also the two sums
 * the BatchNorm backward needs).  gf / gg: optional [B][1][H][W] gradients of the inputs.  batch_stats = the `training` flag
 * of the forward.  Conv biases feeding a batch-statistics BatchNorm have an exactly-zero gradient (not computed here). */
int egaze_lf_bwd(const float* f, const float* g, int B, int H, int W, const float* const* w, const float* raw1,
                 const float* raw2, const float* raw3, const float* bn_ws, int batch_stats, const float* out,
                 const float* gout, float* g1, float* g2, float* scratch, float* const* dw, float* dbh,
                 float* const* dgamma, float* const* dbeta, float* gf, float* gg, int precise, void* stream);

/* ---- fused multi-tensor Adam (torch.optim.Adam semantics: SP.py:110-113,137; LF.py:77,99) that also rewrites the packed
 * operand copies of every conv weight it updates.  jobs: DEVICE array of njobs records (egaze_adam_job_bytes each):
 *   { float* w; const float* g; float* exp_avg; float* exp_avg_sq; float* step;          -- step: device scalar, incremented here
 *     void* p0_hi, *p0_lo;   -- forward copy  [9][rows0][cols0] in format fmt0 (fp16 copies hold w * f16_scale), or NULL
 *     void* p1_hi, *p1_lo;   -- data-gradient copy [9][rows1][cols1] bf16, taps flipped, or NULL
 *     long long n; int Co, Ci;  -- conv weight [Co][Ci][3][3]; Ci == 0: flat tensor of n elements (biases, BatchNorm, 1x1)
 *     int rows0, cols0, fmt0, rows1, cols1, pad; }
 * Two launches whatever the number of parameters. */
int egaze_adam_job_bytes(int* out);
int egaze_adam_multi(const void* jobs, int njobs, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float f16_scale, void* stream);

/* ---- floss (floss.py:9-41) ----------------------------------------------------------------------------------- */
int egaze_floss_centroid(const float* target, int B, int H, int W, double* centroid, void* stream);
int egaze_floss_weight(const double* centroid, int B, int H, int W, float* weights, void* stream);
int egaze_floss_fwd(const float* input, const float* target, const double* centroid, int B, int H, int W,
                    double* loss_acc, float* loss, void* stream);
int egaze_floss_bwd(const float* input, const float* target, const double* centroid, int B, int H, int W,
                    const float* grad_loss, float* grad_input, void* stream);

/* ---- AT glue (AT.py:25-39,58-66,236-241 ; run_spatialstream.py:85-104,136) ------------------------------------ */
int egaze_crop_mean(const float* feat_nchw, const int* gaze, int B, int C, int H, int W, int size, int down, float* out,
                    void* stream);
/* AT.crop_align_feature + mean (AT.py:41-56,239-241): bilinear x`up` (align_corners=True), (size*up)^2 crop, average */
int egaze_crop_align_mean(const float* feat_nchw, const int* gaze, int B, int C, int H, int W, int size, int up, float* out,
                          void* stream);
int egaze_weighted_map(const float* feat_nchw, const float* chn_weight, int B, int C, int HW, float* out, void* stream);
int egaze_bilinear_up(const float* x, int B, int h, int w, int scale, int align_corners, float* out, void* stream);

/* np.uint8(255 * x) / 255 element-wise (x in [0,1]): the quantisation of the maps the reference writes to image files between
 * its stages (AT.py:228-230,249-250 ; read back by data/lateDataset.py:21-34) -- the "bit-compatible" option of the
 * streaming pipeline */
int egaze_quant_u8(const float* x, long long n, float* out, void* stream);

/* ---- input pipeline on the device (replaces the per-sample CPU work of data/STdatas.py:50-73) ------------ */
/* nvJPEG decode of one JPEG held in HOST memory into DEVICE uint8: [H][W][3] BGR interleaved like cv2.imread (gray == 0) or
 * [H][W] (gray != 0, cv2.imread(.., 0)).  nvJPEG is dlopen'ed at first use; EGAZE_EUNSUPPORTED (-2) where it is missing. */
int egaze_jpeg_info(const void* jpeg, long long nbytes, int* width, int* height, int* components);
int egaze_jpeg_decode(const void* jpeg, long long nbytes, int gray, void* out, int H, int W, void* stream);
/* bgr_u8 [N][H][W][3] -> out [N][3][H][W] fp32 = ((u/255) - mean) / std in the reference's (BGR) channel order (STdatas.py:51-55) */
int egaze_image_norm(const void* bgr_u8, int N, int H, int W, float* out, void* stream);
/* Sliding flow window: ring [V][T][2][H][W] uint8 holds the last T decoded (flow_x, flow_y) frames of V videos, so every flow
 * frame is decoded / uploaded once instead of T times.  push: frames flow_x, flow_y [V][H][W] -> slot.  stack: out
 * [V][2T][H][W] fp32, channel 2k / 2k+1 = ((u/255) - 0.5) / 0.5 of flow_x / flow_y of the k-th most recent frame (newest = slot of
 * the latest push; count = frames pushed so far, capped at T: older positions repeat the oldest frame).  (STdatas.py:18-20,59-68) */
int egaze_flow_push(void* ring, const void* flow_x, const void* flow_y, int V, int T, int H, int W, int slot, void* stream);
int egaze_flow_stack(const void* ring, int V, int T, int H, int W, int newest, int count, float* out, void* stream);

/* ---- lstmnet (models/LSTMnet.py:15-37) ------------------------------------------------------------------------ */
int egaze_lstm_seq_fwd(const float* x, const float* h0, const float* c0, const float* const* w_ih,
                       const float* const* w_hh, const float* const* b_ih, const float* const* b_hh, const float* lin_w,
                       const float* lin_b, int T, int B, float* out, float* hn, float* cn, float* ws_h, float* ws_c,
                       float* ws_gates, void* stream);

/* BPTT of lstmnet.forward (loss.backward() in AT.trainLSTM, AT.py:138-142); all workspaces / outputs are caller-allocated:
 * xt, dz, dh_top [T][B][512]; dgates [2][T][B][2048]; tmp_x [B][512]; dh_next, dc_next [2][B][512]; dx0 [T][B][512] iff dinput;
 * dw_ih / dw_hh / db: arrays of 2 device pointers ([2048][512] / [2048]); dlin_w [512][512]; dlin_b [512]. */
int egaze_lstm_seq_bwd(const float* x, const float* h0, const float* c0, const float* const* w_ih,
                       const float* const* w_hh, const float* lin_w, int T, int B, const float* out, const float* gout,
                       const float* ghn, const float* gcn, const float* ws_h, const float* ws_c, const float* ws_gates,
                       float* xt, float* dz, float* dgates, float* dh_top, float* tmp_x, float* dh_next, float* dc_next,
                       float* dx0, float* const* dw_ih, float* const* dw_hh, float* const* db, float* dlin_w,
                       float* dlin_b, float* dinput, float* dh0, float* dc0, void* stream);

/* ---- validation metric on the device (replaces utils.computeAAEAUC, utils.py:96-140: scipy centre of mass + Gaussian
 * filter + AUC count per sample on the host).  out / tgt: [B][224][224] fp32; weights: [113] fp64 = scipy's sigma-14 Gaussian
 * kernel built on the host; res: [B][4] fp64 = (AAE deg, AUC, gaze row, col) */
int egaze_aae_auc(const float* out, const float* tgt, int B, int H, int W, const double* weights, double* res, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGAZE_H_ */
