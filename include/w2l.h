/*
 * w2l.h — C-ABI of the B200-native Wav2Lip compute core (libw2l.so).
 *
 * The reference (Rudrabha/Wav2Lip) has no FFI: its boundary is the Python module surface
 * `models` / `audio` that the scripts import by bare name.  Each entry point below replaces
 * one of those Python call targets; the reference-side binding (a ctypes-backed `models`
 * package with the same class names, ctor/forward signatures and state_dict keys) lives in
 * wav2lip_b200/models/ and is described in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only: pointers, sizes, ints.  No torch / C++ types cross this boundary.
 *  - every function returns 0 on success, a negative W2L_E* code on failure; the message is
 *    available through w2l_last_error() (thread local).  Nothing throws across the ABI.
 *  - "dev" pointers are CUDA device pointers on the context's device; tensors are fp32,
 *    contiguous, in the layout the reference's callers build (NCHW / 5-D B,C,T,H,W).
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *    asynchronous on that stream unless stated otherwise.
 *  - a context is bound to one device and is not thread safe (one context per GPU / stream).
 *  - there is NO CPU fallback: every compute entry point fails with W2L_ENODEV when no
 *    sm_100 device is usable.
 */
#ifndef W2L_H_
#define W2L_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2L_ABI_VERSION 1

/* error codes */
#define W2L_OK        0
#define W2L_EINVAL   -1   /* bad argument / shape */
#define W2L_ENODEV   -2   /* no usable sm_100 CUDA device */
#define W2L_ECUDA    -3   /* CUDA runtime / driver error (see w2l_last_error) */
#define W2L_ENOMEM   -4
#define W2L_ESTATE   -5   /* e.g. forward before weights were loaded */

/* networks */
#define W2L_NET_GENERATOR 0   /* models.Wav2Lip            /root/reference/models/wav2lip.py:8-125   */
#define W2L_NET_SYNCNET   1   /* models.SyncNet_color      /root/reference/models/syncnet.py:7-66    */
#define W2L_NET_DISC      2   /* models.Wav2Lip_disc_qual  /root/reference/models/wav2lip.py:127-184 */
#define W2L_NET_S3FD      3   /* face_detection s3fd       /root/reference/face_detection/detection/sfd/net_s3fd.py:22-129 */

/* block kinds — /root/reference/models/conv.py */
#define W2L_BLOCK_CONV_BN_RELU   0   /* Conv2d           conv.py:5-19  */
#define W2L_BLOCK_CONVT_BN_RELU  1   /* Conv2dTranspose  conv.py:33-44 */
#define W2L_BLOCK_CONV_LRELU     2   /* nonorm_Conv2d    conv.py:21-31 */
#define W2L_BLOCK_CONV_PLAIN     3   /* bare nn.Conv2d heads: wav2lip.py:84, :152 (followed by Sigmoid); S3FD's mbox convs */
#define W2L_BLOCK_CONV_RELU      4   /* F.relu(nn.Conv2d(x)): the S3FD backbone, net_s3fd.py:72-106 */

/* operand precision of the tensor-core path (accumulation is always fp32) */
#define W2L_PREC_F16  0   /* fp16 operands: 10-bit mantissa, same as TF32 (default) */
#define W2L_PREC_BF16 1   /* bf16 operands: for checkpoints whose activations exceed the fp16 range */
#define W2L_PREC_F32X 2   /* fp32-faithful: every activation and weight is carried as hi + lo (two fp16 values, ~22
                             significant bits) and each product as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo on the tensor
                             cores (3 MMAs, generic kernel only): ~1e-6 relative per block, ~1/4 of the throughput */

typedef struct w2l_ctx w2l_ctx;

/* One row of an architecture table: a conv block with the reference's own module path. */
typedef struct w2l_layer_info {
    char name[64];      /* e.g. "face_decoder_blocks.3.1" — state_dict prefix of the block */
    int32_t kind;       /* W2L_BLOCK_* */
    int32_t cin, cout;
    int32_t kh, kw, sh, sw, ph, pw;
    int32_t out_pad;    /* ConvTranspose2d output_padding */
    int32_t residual;   /* conv.py:16-18 */
    int32_t cout_real;  /* 0, or the parameter tensor's real output-channel count when `cout` is its 16-padded width (S3FD heads) */
} w2l_layer_info;

/* ---- introspection (host only; usable without a GPU) ---- */
int         w2l_abi_version(void);
const char* w2l_last_error(void);
int         w2l_net_num_layers(int net);
int         w2l_net_layer_info(int net, int index, w2l_layer_info* out);

/* ---- context ---- */
/* Replaces `Model().to(device)` (inference.py:169,178; wav2lip_train.py:356). */
int w2l_create(int device, int precision, w2l_ctx** out);
int w2l_destroy(w2l_ctx* ctx);

/* Replaces `model.load_state_dict(sd)` (inference.py:176): hands the context the fp32 device
 * tensors of one network by their reference state_dict names ("<block>.conv_block.0.weight",
 * "<block>.conv_block.1.running_var", "output_block.1.weight", ...).  Weights are re-packed
 * (fp16/bf16, tap-major K-major tiles; BatchNorm running stats folded into per-channel
 * scale/shift) on `stream`; the source tensors may be freed afterwards.  Re-callable.
 * Unknown names are ignored ("...num_batches_tracked"); a missing tensor is W2L_EINVAL. */
int w2l_load_weights(w2l_ctx* ctx, int net, int n_tensors, const char* const* names,
                     const void* const* dev_ptrs, const int64_t* numels, void* stream);

/* Replaces `Wav2Lip.forward(audio_sequences, face_sequences)` (wav2lip.py:87-125), eval mode.
 *   T == 0 : mel (B,1,80,16), face (B,6,96,96)      -> out (B,3,96,96)
 *   T  > 0 : mel (B,T,1,80,16), face (B,6,T,96,96)  -> out (B,3,T,96,96)   (t-major flatten inside) */
int w2l_generator_forward(w2l_ctx* ctx, const float* mel_dev, const float* face_dev, float* out_dev,
                          int B, int T, void* stream);

/* Same, with HOST buffers: H2D of the inputs, forward, D2H of the result, then a stream sync.
 * Buffers should be pinned for full PCIe bandwidth (pageable memory works, slower). */
int w2l_generator_forward_host(w2l_ctx* ctx, const float* mel_host, const float* face_host,
                               float* out_host, int B, int T);

/* Asynchronous form of the two host-buffer calls, for a serving loop that keeps the PCIe link and the GPU busy at
 * the same time (the reference's loop, inference.py:259-265, is copy -> forward -> copy, strictly serial):
 *   w2l_generator_submit_host / _submit_u8_host enqueue H2D -> forward -> D2H for one batch and return at once;
 *   at most two submissions are in flight (a third call first waits for the oldest);
 *   w2l_host_wait(ctx, keep) blocks until at most `keep` submissions are still in flight — their out_host buffers are
 *   complete when it returns, in submission order.  Host buffers must stay valid (and should be pinned) until retired.
 *       submit(batch 0); for k = 1..: submit(batch k); w2l_host_wait(ctx, 1); consume(batch k-1); ... w2l_host_wait(ctx, 0) */
int w2l_generator_submit_host(w2l_ctx* ctx, const float* mel_host, const float* face_host, float* out_host, int B, int T);
int w2l_generator_submit_u8_host(w2l_ctx* ctx, const float* mel_host, const uint8_t* faces_host, uint8_t* out_host, int N);
int w2l_host_wait(w2l_ctx* ctx, int keep_in_flight);

/* Scope row (f): the batch assembly around the generator call of inference.py, fused on the GPU.
 * Replaces inference.py:134-140 (mask the lower half, concat [masked | full] on channels, /255, NHWC->NCHW),
 * :259-263 (to device, forward) and :265,:269 (transpose, *255., astype(uint8)) — everything between the
 * cv2.resize of the crop (:126) and the cv2.resize of the prediction (:269):
 *   mel (N,1,80,16) fp32, faces (N,96,96,3) uint8 BGR crops -> out (N,96,96,3) uint8 BGR.
 * 8x fewer H2D and 4x fewer D2H bytes than the fp32 call. */
int w2l_generator_forward_u8(w2l_ctx* ctx, const float* mel_dev, const uint8_t* faces_dev, uint8_t* out_dev,
                             int N, void* stream);
int w2l_generator_forward_u8_host(w2l_ctx* ctx, const float* mel_host, const uint8_t* faces_host,
                                  uint8_t* out_host, int N);

/* Scope row (f2), the rest of the batch assembly: the two cv2.resize calls and the paste-back of inference.py, bit-identical
 * to OpenCV's fixed-point 8-bit INTER_LINEAR (11-bit coefficients; restated and pinned against cv2 in oracle/pipeline_oracle.py).
 *   frames (F,H,W,3) uint8 BGR on the device; boxes_host: N rows (frame index, y1, y2, x1, x2) in HOST memory (validated).
 *   w2l_crop_resize_u8    replaces `cv2.resize(face, (96, 96))` of every frames[f][y1:y2, x1:x2] (inference.py:102,:126)
 *                         -> crops (N,96,96,3) uint8, the input of w2l_generator_forward_u8.
 *   w2l_paste_u8          replaces `p = cv2.resize(p.astype(np.uint8), (x2-x1, y2-y1)); f[y1:y2, x1:x2] = p` on a copy of the
 *                         frame (inference.py:123, :267-271): pred (N,96,96,3) uint8 -> out_frames (N,H,W,3) uint8.
 *   w2l_lipsync_frames_u8 the whole inner loop of inference.py:120-140 + :259-271 in one call: crop + resize -> mask / concat
 *                         / 255 -> generator -> x255 -> uint8 -> resize -> paste.  mel (N,1,80,16) fp32. */
int w2l_crop_resize_u8(w2l_ctx* ctx, const uint8_t* frames_dev, int F, int H, int W, const int32_t* boxes_host, int N,
                       uint8_t* crops_dev, void* stream);
int w2l_paste_u8(w2l_ctx* ctx, const uint8_t* pred_dev, const uint8_t* frames_dev, int F, int H, int W,
                 const int32_t* boxes_host, int N, uint8_t* out_frames_dev, void* stream);
int w2l_lipsync_frames_u8(w2l_ctx* ctx, const float* mel_dev, const uint8_t* frames_dev, int F, int H, int W,
                          const int32_t* boxes_host, int N, uint8_t* out_frames_dev, void* stream);

/* Scope row (f4): the S3FD face detector's network (face_detection/detection/sfd/net_s3fd.py:22-129), the per-frame GPU
 * work of inference.py's face_detect (:73-100): 19 conv+ReLU layers (VGG16 backbone + fc6/fc7 + conv6/7), 5 max-pools,
 * 3 L2Norm layers and the 12 mbox heads, max-out of the first scale's background logits included.
 *   img (B,3,H,W) fp32, BGR minus (104,117,123) as detect.py:21-23 / :60-61 prepare it  ->  12 maps
 *   outs[2i] = cls_i (B,2,h_i,w_i), outs[2i+1] = reg_i (B,4,h_i,w_i), i = 0..5 (strides 4..128), raw logits as the module
 *   returns them (softmax / threshold / decode / NMS of detect.py:31-56 and bbox.py:44-64 stay on the host side).
 * w2l_s3fd_out_dims writes the six (h_i, w_i) pairs for an H x W input.  Weights: w2l_load_weights(ctx, W2L_NET_S3FD, ...)
 * with the module's own state_dict names ("conv1_1.weight", ..., "conv3_3_norm.weight", "conv7_2_mbox_loc.bias"). */
int w2l_s3fd_out_dims(int H, int W, int32_t* dims12);
int w2l_s3fd_forward(w2l_ctx* ctx, const float* img_dev, float* const* outs12_dev, int B, int H, int W, void* stream);

/* Replaces `SyncNet_color.forward(audio, face)` (syncnet.py:55-66):
 *   mel (B,1,80,16), face (B,15,48,96) -> audio_emb (B,512), face_emb (B,512), both L2-normalised. */
int w2l_syncnet_forward(w2l_ctx* ctx, const float* mel_dev, const float* face_dev,
                        float* audio_emb_dev, float* face_emb_dev, int B, void* stream);

/* Scope row (a8) + evaluation loops (wav2lip_train.py:262-292, hq_wav2lip_train.py eval): the expert-discriminator
 * call on generated frames, `get_sync_loss` (wav2lip_train.py:192-198) up to the embeddings:
 *   g[:, :, :, H/2:]  ->  cat([g[:, :, i] for i in range(syncnet_T)], dim=1)  ->  syncnet(mel, .)
 * with the slice and the channel stack done as addressing by the ingest kernel:
 *   mel (B,1,80,16), frames (B,3,T,96,96) with T == 5 -> audio_emb (B,512), face_emb (B,512). */
int w2l_syncnet_forward_frames(w2l_ctx* ctx, const float* mel_dev, const float* frames_dev, float* audio_emb_dev,
                               float* face_emb_dev, int B, int T, void* stream);

/* Replaces `cosine_loss(a, v, y)` (wav2lip_train.py:178-183, color_syncnet_train.py:133-138):
 *   d = F.cosine_similarity(a, v) (eps 1e-8);  loss = nn.BCELoss()(d.unsqueeze(1), y)  (mean; log clamped at -100).
 *   a, v (B,D) fp32; y (B) fp32 targets or NULL for all ones (get_sync_loss, :197); loss: 1 float on the device.
 *   Forward value only (the evaluation loops); deterministic summation order. */
int w2l_cosine_bce_loss(w2l_ctx* ctx, const float* a_dev, const float* v_dev, const float* y_dev, int B, int D,
                        float* loss_dev, void* stream);

/* Replaces `recon_loss = nn.L1Loss()` (wav2lip_train.py:191, :228, :281): loss = mean |x - y| over n fp32 elements
 * (16-byte aligned pointers); loss: 1 float on the device.  HBM-bound: 8 bytes read per element pair. */
int w2l_l1_loss(w2l_ctx* ctx, const float* x_dev, const float* y_dev, int64_t n, float* loss_dev, void* stream);

/* Replaces `Wav2Lip_disc_qual.forward(face_sequences)` (wav2lip.py:176-184):
 *   frames (B,3,T,96,96) -> prob (B*T,1), rows t-major (row = t*B + b). */
int w2l_disc_forward(w2l_ctx* ctx, const float* frames_dev, float* prob_dev, int B, int T, void* stream);

/* Replaces one `models.conv.{Conv2d,Conv2dTranspose,nonorm_Conv2d}.forward` (conv.py:15-19,29-31,
 * 42-44), eval mode — the operator-level entry used by the per-geometry parity tests.
 *   x (N,cin,H,W) fp32 -> y (N,cout,Hout,Wout) fp32.  weight is (cout,cin,kh,kw), or
 *   (cin,cout,kh,kw) for the transposed block; bn_* are NULL for nonorm/plain blocks. */
int w2l_conv_block_forward(w2l_ctx* ctx, const w2l_layer_info* spec,
                           const float* x_dev, int N, int H, int W,
                           const float* weight_dev, const float* bias_dev,
                           const float* bn_weight_dev, const float* bn_bias_dev,
                           const float* bn_mean_dev, const float* bn_var_dev,
                           float* y_dev, void* stream);

/* Debug/test aid: copy the output of block `layer` (index into the net's table) of the LAST
 * forward of `net` to y (N,cout,H,W) fp32.  *h,*w,*c receive the dims (any may be NULL). */
int w2l_debug_layer_output(w2l_ctx* ctx, int net, int layer, float* y_dev, int* n, int* c, int* h, int* w,
                           void* stream);

/* Replaces `audio.melspectrogram(wav)` (audio.py:45-51 with hparams.py:33-73):
 *   wav (L) fp32 -> mel (80, 1 + L/200) fp32 in [-4,4], row-major as numpy returns it.
 *   L >= 2; clips shorter than n_fft/2 = 400 samples reflect more than once, as np.pad(mode="reflect") does.
 *   w2l_melspectrogram_host first retires every asynchronous host submission of the context (w2l_host_wait(ctx, 0)):
 *   it shares their staging buffers. */
int w2l_melspectrogram(w2l_ctx* ctx, const float* wav_dev, int64_t n_samples, float* mel_dev, void* stream);
int w2l_melspectrogram_host(w2l_ctx* ctx, const float* wav_host, int64_t n_samples, float* mel_host);
/* number of frames for n_samples: 1 + n_samples/200 (librosa center=True) */
int64_t w2l_mel_num_frames(int64_t n_samples);

/* Scope row (f): replaces the mel chunking loop of inference.py:231-240 — chunk i = mel[:, s_i : s_i+16] with
 * s_i = int(i * 80./fps), the last chunk right-aligned; chunks out is (n_chunks,1,80,16) fp32, i.e. already the
 * `mel_batch` layout of inference.py:260.  w2l_mel_num_chunks gives n_chunks for a mel of n_frames columns. */
int64_t w2l_mel_num_chunks(int64_t n_frames, double fps);
int w2l_mel_chunks(w2l_ctx* ctx, const float* mel_dev, int64_t n_frames, double fps, float* chunks_dev,
                   int64_t n_chunks, void* stream);

/* ---- test aids ---- */
/* keep every block output of subsequent plans addressable (no buffer reuse) for w2l_debug_layer_output */
int w2l_set_debug(w2l_ctx* ctx, int keep_all_layer_outputs);
/* the kernel's own (80 x 401) Slaney mel filterbank, dense fp32, written to HOST memory */
int w2l_mel_basis_host(float* out_host);

/* ---- training step (scope row f1): wav2lip_train.py:210-231, color_syncnet_train.py:146-163, hq_wav2lip_train.py:213-255 ----
 * Train-mode forward (BatchNorm on batch statistics over the T*B flatten, conv.py:8-11 / wav2lip.py:93-94; running
 * averages updated with momentum 0.1) and backward through every block, as kernels: conv / dgrad on the tcgen05 conv
 * kernels, wgrad on a tcgen05 kernel whose K dimension is the pixel index, BatchNorm / ReLU / residual passes,
 * loss gradients, multi-tensor Adam, bucketed NCCL all-reduce of the gradients.  bf16 operands, fp32 accumulation,
 * fp32 master parameters and gradients (the context must be created with W2L_PREC_BF16).
 *
 * w2l_train_bind replaces `optimizer = optim.Adam([p for p in model.parameters() ...])` + the module's own tensors
 * (wav2lip_train.py:356-360): it hands the context, by reference state_dict name, the caller's fp32 device tensors —
 * parameters (value + gradient pointer; NULL gradient = frozen, wav2lip_train.py:188-189) and BatchNorm buffers
 * (running_mean / running_var, gradient NULL).  The pointers must stay valid; values are re-read (and the 16-bit weight
 * slabs re-packed) at every training forward, gradients are written by the backward, running averages are updated in
 * place.  Gradients laid out contiguously (one arena in state_dict order) are all-reduced as three large buckets. */
#define W2L_TRAIN_WGRAD           1   /* compute parameter gradients into the bound gradient tensors */
#define W2L_TRAIN_ACCUMULATE      2   /* add to the bound gradient tensors instead of overwriting (two backward() calls, hq_wav2lip_train.py:248-253) */
#define W2L_TRAIN_INPUT_GRAD      4   /* the plan also produces dL/d(input frames) (expert / discriminator inside a generator step) */
#define W2L_TRAIN_NO_STAT_UPDATE  8   /* leave running_mean / running_var untouched */
int w2l_train_bind(w2l_ctx* ctx, int net, int n_tensors, const char* const* names, void* const* value_ptrs,
                   void* const* grad_ptrs, const int64_t* numels);

/* Train-mode forward; keeps the tape (block inputs, pre-BatchNorm outputs, outputs) for w2l_train_backward.
 *   W2L_NET_GENERATOR: in0 = mel, in1 = face (4-D with T == 0 or 5-D with T > 0, as w2l_generator_forward), out0 = g.
 *                      out0 must stay valid until the backward (the head's backward re-reads it).
 *   W2L_NET_SYNCNET  : in0 = mel (B,1,80,16); in1 = face (B,15,48,96) [T == 0] or frames (B,3,5,96,96) [T == 5];
 *                      out0 = audio_embedding, out1 = face_embedding (B,512), L2-normalised.
 *   W2L_NET_DISC     : in0 = frames (B,3,T,96,96), out0 = prob (B*T,1) t-major. */
int w2l_train_forward(w2l_ctx* ctx, int net, const float* in0_dev, const float* in1_dev, float* out0_dev, float* out1_dev,
                      int B, int T, int flags, void* stream);

/* Backward of the last w2l_train_forward of `net`.  d0 (/ d1) = dL/d out0 (/ out1) fp32, same shapes; with
 * W2L_TRAIN_WGRAD the bound gradient tensors receive dL/dparameter (`loss.backward()`, wav2lip_train.py:230);
 * dinput (needs W2L_TRAIN_INPUT_GRAD at the forward) receives dL/d in1 (SyncNet: frames (B,3,5,96,96) — zero in the
 * upper half — or face (B,15,48,96)) or dL/d in0 (disc frames), fp32. */
int w2l_train_backward(w2l_ctx* ctx, int net, const float* d0_dev, const float* d1_dev, float* dinput_dev, int flags,
                       void* stream);

/* Replaces `optimizer.step()` (torch.optim.Adam, no weight decay / amsgrad; wav2lip_train.py:231, :357-360) on every
 * bound tensor of `net` that has a gradient: one multi-tensor kernel; moments live in the context. */
int w2l_adam_step(w2l_ctx* ctx, int net, float lr, float beta1, float beta2, float eps, void* stream);

/* One whole iteration of wav2lip_train.py:210-231 on the bound generator and (frozen, train-mode as in the scripts,
 * :187-189) expert: g = model(indiv_mels, x); sync_loss = get_sync_loss(mel, g) if syncnet_wt > 0; l1 = L1(g, gt);
 * loss = syncnet_wt*sync + (1-syncnet_wt)*l1; backward; [gradient all-reduce, overlapped]; Adam(lr, (0.9,0.999), 1e-8).
 *   indiv_mels (B,T,1,80,16), x (B,6,T,96,96), mel (B,1,80,16), gt (B,3,T,96,96); losses_dev: 4 floats on the device
 *   [sync_loss, l1, 0, loss] or NULL. */
int w2l_wav2lip_train_step(w2l_ctx* ctx, const float* indiv_mels_dev, const float* x_dev, const float* mel_dev,
                           const float* gt_dev, int B, int T, float syncnet_wt, float lr, float* losses_dev, void* stream);
/* copies the generator output g (B,3,T,96,96) of the last fused step into out_dev (n floats) */
int w2l_train_last_output(w2l_ctx* ctx, float* out_dev, int64_t n, void* stream);
/* algorithmic forward FLOPs of the last training plan of `net` (2 x true MACs of its convs) */
double w2l_train_flops(w2l_ctx* ctx, int net);

/* Per-stage CUDA-event times of the last training plan of `net` (run one forward + backward first): rows
 * "<block> fwd | bn | bwd_bn | dgrad | wgrad" with the stage's algorithmic FLOPs; returns the number of rows (<= cap). */
int w2l_train_profile(w2l_ctx* ctx, int net, int iters, int cap, float* ms_out, double* flop_out, char (*names_out)[64],
                      void* stream);

/* Data-parallel training: the gradient all-reduce is the one collective of the system.  NCCL is
 * resolved at run time from the process (torch loads libnccl.so.2); rank 0 creates the 128-byte unique id, the host
 * side broadcasts it (torch.distributed), every rank calls w2l_comm_init.  w2l_wav2lip_train_step then averages the
 * gradients over the ranks (ncclAvg) in three buckets launched on a side stream as the backward completes them. */
int w2l_comm_unique_id(w2l_ctx* ctx, char* id128);
int w2l_comm_init(w2l_ctx* ctx, const char* id128, int rank, int world);

/* One conv.py block in train mode, forward + backward (operator-level entry of the per-geometry gradient tests):
 *   x (N,cin,H,W), dy (N,cout,Ho,Wo) fp32 -> y, dx (same layouts; dy/dx/dw may be NULL), dw in the parameter's own
 *   layout, db, dgamma, dbeta; bn_mean / bn_var (running averages) are updated in place when given. */
int w2l_conv_block_train(w2l_ctx* ctx, const w2l_layer_info* spec, const float* x_dev, int N, int H, int W,
                         float* weight_dev, float* bias_dev, float* bn_weight_dev, float* bn_bias_dev, float* bn_mean_dev,
                         float* bn_var_dev, const float* dy_dev, float* y_dev, float* dx_dev, float* dw_dev, float* db_dev,
                         float* dgamma_dev, float* dbeta_dev, void* stream);

/* Range guard of the fp16 modes (W2L_PREC_F16 / _F32X): every epilogue that stores an activation sets a sticky
 * per-device flag when the rounded value leaves the fp16 range (|v| > 65504 -> inf, or NaN).  Synchronises `stream`,
 * writes the flag to *flag (0 = every activation stored so far was finite) and, if `clear`, resets it.  A checkpoint
 * that trips it must run in W2L_PREC_BF16 (fp32's exponent range).  bf16 contexts never set it. */
int w2l_f16_overflow(w2l_ctx* ctx, int clear, int* flag, void* stream);

/* ---- instrumentation ---- */
/* kernels launched by this library since the context was created (all streams) */
int64_t w2l_launch_count(const w2l_ctx* ctx);
/* bytes of device memory currently held by the context (weights + activation arenas) */
int64_t w2l_device_bytes(const w2l_ctx* ctx);
/* Time the conv kernels of the last-built plan of `net` individually: runs every launch `iters`
 * times with CUDA events on `stream`, the L2 flushed before each timed launch (cold cache, as in the step), and writes
 * per-launch mean milliseconds and flop counts.
 * Returns the number of launches written (<= cap). */
int w2l_profile_plan(w2l_ctx* ctx, int net, int iters, int cap, float* ms_out, double* flop_out,
                     char (*names_out)[64], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2L_H_ */
