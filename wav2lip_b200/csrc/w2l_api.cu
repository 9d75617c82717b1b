// w2l_api.cu — the C-ABI of include/w2l.h (libw2l.so).  The host side behind it lives in the host_*.h / host_*.cuh
// headers included below (one translation unit): data model, kernel launch tables, op builders, weight packing, plans,
// mel tables.  No torch, no CPU compute path.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/w2l.h"
#include "aux_kernels.cuh"
#include "conv_patch.cuh"
#include "conv_rowstack.cuh"
#include "conv_swap.cuh"
#include "conv_tcgen05.cuh"
#include "convt_fused.cuh"
#include "mel.cuh"
#include "netspec.h"
#include "resize.cuh"
#include "train_kernels.cuh"
#include "wgrad_tcgen05.cuh"

using namespace w2l;

#include "host_types.h"
#include "host_launch.cuh"
#include "host_ops.cuh"
#include "host_weights.cuh"
#include "host_plans.cuh"
#include "host_mel_tables.h"
#include "host_train.cuh"

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int w2l_abi_version(void) { return W2L_ABI_VERSION; }
const char* w2l_last_error(void) { return g_err.c_str(); }

int w2l_net_num_layers(int net) {
    const std::vector<Layer>* L = net_layers(net);
    return L ? (int)L->size() : fail(W2L_EINVAL, "unknown net %d", net);
}

int w2l_net_layer_info(int net, int index, w2l_layer_info* out) {
    const std::vector<Layer>* Ls = net_layers(net);
    if (!Ls || !out || index < 0 || index >= (int)Ls->size()) return fail(W2L_EINVAL, "bad net/index %d/%d", net, index);
    const Layer& L = (*Ls)[index];
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%s", L.name.c_str());
    out->kind = L.kind; out->cin = L.cin; out->cout = L.cout; out->kh = L.kh; out->kw = L.kw;
    out->sh = L.sh; out->sw = L.sw; out->ph = L.ph; out->pw = L.pw; out->out_pad = L.out_pad; out->residual = L.residual ? 1 : 0;
    out->cout_real = L.cout_real;
    return W2L_OK;
}

/* product's own mel filterbank, dense (80 x 401) fp32, host memory — for the parity tests */
int w2l_mel_basis_host(float* out) {
    if (!out) return fail(W2L_EINVAL, "null output");
    std::vector<float> d;
    build_mel_basis(&d);
    memcpy(out, d.data(), d.size() * 4);
    return W2L_OK;
}

int w2l_create(int device, int precision, w2l_ctx** out) {
    if (!out) return fail(W2L_EINVAL, "null out");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(W2L_ENODEV, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return fail(W2L_EINVAL, "device %d out of range (%d devices)", device, ndev);
    if (precision != W2L_PREC_F16 && precision != W2L_PREC_BF16 && precision != W2L_PREC_F32X) return fail(W2L_EINVAL, "unknown precision %d", precision);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(W2L_ENODEV, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    if (!get_encode_fn()) return fail(W2L_ENODEV, "cuTensorMapEncodeTiled not found in the driver");
    DeviceGuard g(device);
    w2l_ctx* ctx = new w2l_ctx();
    ctx->device = device;
    ctx->bf16 = precision == W2L_PREC_BF16;
    ctx->x2 = precision == W2L_PREC_F32X;
    ctx->num_sms = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { w2l_destroy(ctx); return fail(W2L_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    int r = init_mel_tables(ctx);
    if (r != W2L_OK) { const std::string msg = g_err; w2l_destroy(ctx); g_err = msg; return r; }
    {
        // A/B switches, read once per context: W2L_DISABLE_<NAME>=1 turns one specialised path off (tests/test_gpu_variants.py)
        auto enabled = [](const char* name) { const char* v = getenv(name); return !(v && v[0] == '1'); };
        ctx->use_patch = enabled("W2L_DISABLE_HALO");
        ctx->use_fold = enabled("W2L_DISABLE_FOLD");
        ctx->use_fold_s2 = enabled("W2L_DISABLE_FOLDS2");
        ctx->use_bn256 = enabled("W2L_DISABLE_BN256");
        ctx->use_mt2 = enabled("W2L_DISABLE_MT2");
        ctx->use_swap = enabled("W2L_DISABLE_SWAP");
        ctx->use_rounds = enabled("W2L_DISABLE_ROUNDS");
        ctx->use_wg_stream = enabled("W2L_DISABLE_WGSTREAM");
        ctx->use_aux_stream = enabled("W2L_DISABLE_AUXSTREAM");
        ctx->use_tma_epi = enabled("W2L_DISABLE_TMAEPI");
        ctx->use_ctfused = enabled("W2L_DISABLE_CTFUSED");
        ctx->use_rowstack = enabled("W2L_DISABLE_ROWSTACK");
        ctx->use_side = enabled("W2L_DISABLE_SIDESTREAM");
        ctx->use_pdl = enabled("W2L_DISABLE_PDL");
        ctx->use_mel_v2 = enabled("W2L_DISABLE_MELV2");
        if (ctx->x2) {  // the split-operand mode runs on the generic kernel with the direct epilogue only
            ctx->use_patch = ctx->use_fold = ctx->use_fold_s2 = ctx->use_ctfused = ctx->use_tma_epi = false;
        }
        if (ctx->use_fold) {
            // the folded first layers need a tensor map whose pixel stride (16 B) is smaller than its inner extent
            // (128 B): probe once that the driver encodes such overlapping windows
            Act probe;
            probe.base = (uint16_t*)ctx->mel_tw; probe.N = 2; probe.H = 16; probe.W = 16; probe.Cs = 8; probe.C = 64; probe.Wp = 24;
            CUtensorMap tm;
            if (encode_act_map(ctx, &tm, probe, 64, 8, 8, 2, 1, 1, "probe") != W2L_OK) { ctx->use_fold = false; g_err.clear(); }
        }
    }
    *out = ctx;
    return W2L_OK;
}

int w2l_destroy(w2l_ctx* ctx) {
    if (!ctx) return W2L_OK;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    free_train_state(ctx);
    for (auto& kv : ctx->plans) free_plan(kv.second.get());
    for (int i = 0; i < 3; ++i) if (ctx->s3fd_l2w[i]) cudaFree(ctx->s3fd_l2w[i]);
    for (int n = 0; n < 4; ++n) {
        for (auto& lw : ctx->nets[n].layers) free_layer(lw);
        if (ctx->nets[n].head_w) cudaFree(ctx->nets[n].head_w);
        if (ctx->nets[n].head_b) cudaFree(ctx->nets[n].head_b);
    }
    for (int i = 0; i < 6; ++i) if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (int i = 0; i < 2; ++i) { if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]); if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]); if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]); }
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->boxes_dev) cudaFree(ctx->boxes_dev);
    if (ctx->crops_dev) cudaFree(ctx->crops_dev);
    if (ctx->preds_dev) cudaFree(ctx->preds_dev);
    if (ctx->mel_tw) cudaFree(ctx->mel_tw);
    if (ctx->mel_bvals) cudaFree(ctx->mel_bvals);
    if (ctx->mel_boff) cudaFree(ctx->mel_boff);
    if (ctx->mel_bstart) cudaFree(ctx->mel_bstart);
    if (ctx->mel_blen) cudaFree(ctx->mel_blen);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return W2L_OK;
}

int w2l_set_debug(w2l_ctx* ctx, int keep_all_layer_outputs) {
    if (!ctx) return fail(W2L_EINVAL, "null ctx");
    ctx->keep_all = keep_all_layer_outputs != 0;
    return W2L_OK;
}

int w2l_load_weights(w2l_ctx* ctx, int net, int n_tensors, const char* const* names, const void* const* dev_ptrs,
                     const int64_t* numels, void* stream) {
    if (!ctx || !names || !dev_ptrs || !numels) return fail(W2L_EINVAL, "null argument");
    const std::vector<Layer>* Ls = net_layers(net);
    if (!Ls) return fail(W2L_EINVAL, "unknown net %d", net);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    TensorMap tm;
    for (int i = 0; i < n_tensors; ++i) {
        std::string nm = names[i];
        if (nm.rfind("module.", 0) == 0) nm = nm.substr(7);  // DataParallel-era checkpoints, inference.py:174-175
        tm[nm] = TensorRef{(const float*)dev_ptrs[i], numels[i]};
    }
    CK(cudaDeviceSynchronize());  // queued forwards (asynchronous host submissions included) still read the old weights
    drop_plans(ctx, net);
    NetW& nw = ctx->nets[net];
    nw.loaded = false;
    nw.layers.resize(Ls->size());
    // which blocks see a 1x1 input (GEMM form of the transposed conv): generator decoder stage 1
    for (size_t i = 0; i < Ls->size(); ++i) {
        const Layer& L = (*Ls)[i];
        const float *W, *b, *gm, *be, *m, *v;
        CKR(fetch_block_tensors(tm, L, &W, &b, &gm, &be, &m, &v));
        const bool hw1 = (net == W2L_NET_GENERATOR && L.name == "face_decoder_blocks.1.0");
        // blocks fed directly by the ingest kernel (caller tensors): the only ones with a tiny Cin
        bool first = L.name == "face_encoder_blocks.0.0" || L.name == "audio_encoder.0" || L.name == "face_encoder.0" ||
                     (net == W2L_NET_S3FD && L.name == "conv1_1");
        // the generator's 16->32 stride-2 block reads a dense zero-bordered copy of the first block's output (written by
        // the patch kernel's second TMA store) through the same overlapping-window trick
        if (net == W2L_NET_GENERATOR && L.name == "face_encoder_blocks.1.0" && ctx->use_patch && ctx->use_fold && ctx->use_fold_s2) first = true;
        CKR(load_layer(ctx, &nw.layers[i], L, W, b, gm, be, m, v, hw1, first, st));
    }
    if (nw.head_w) { cudaFree(nw.head_w); nw.head_w = nullptr; }
    if (nw.head_b) { cudaFree(nw.head_b); nw.head_b = nullptr; }
    if (net == W2L_NET_GENERATOR || net == W2L_NET_DISC) {
        const char* wn = net == W2L_NET_GENERATOR ? "output_block.1.weight" : "binary_pred.0.weight";
        const char* bn = net == W2L_NET_GENERATOR ? "output_block.1.bias" : "binary_pred.0.bias";
        const int64_t wcount = net == W2L_NET_GENERATOR ? 96 : 512, bcount = net == W2L_NET_GENERATOR ? 3 : 1;
        const float *hw, *hb;
        CKR(need(tm, wn, wcount, &hw));
        CKR(need(tm, bn, bcount, &hb));
        void* p;
        CKR(dev_alloc(&p, wcount * 4)); nw.head_w = (float*)p;
        CKR(dev_alloc(&p, bcount * 4)); nw.head_b = (float*)p;
        CK(cudaMemcpyAsync(nw.head_w, hw, wcount * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(nw.head_b, hb, bcount * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (net == W2L_NET_S3FD) {   // L2Norm weights (net_s3fd.py:12-14, :64-66)
        const char* names3[3] = {"conv3_3_norm.weight", "conv4_3_norm.weight", "conv5_3_norm.weight"};
        const int64_t n3[3] = {256, 512, 512};
        for (int i = 0; i < 3; ++i) {
            const float* w;
            CKR(need(tm, names3[i], n3[i], &w));
            if (!ctx->s3fd_l2w[i]) { void* p; CKR(dev_alloc(&p, 512 * 4)); ctx->s3fd_l2w[i] = (float*)p; }
            CK(cudaMemcpyAsync(ctx->s3fd_l2w[i], w, n3[i] * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    CK(cudaStreamSynchronize(st));  // the caller may free / mutate the fp32 sources after we return
    nw.loaded = true;
    return W2L_OK;
}

int w2l_generator_forward(w2l_ctx* ctx, const float* mel, const float* face, float* out, int B, int T, void* stream) {
    if (!ctx || !mel || !face || !out) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || T < 0) return fail(W2L_EINVAL, "bad batch B=%d T=%d", B, T);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_GENERATOR, B, T, &pl));
    return run_plan(ctx, pl, mel, face, out, nullptr, (cudaStream_t)stream);
}

// ---- host-buffer entry points: a two-slot software pipeline over three streams --------------------------------
// Every submission (a whole call, or one chunk of a synchronous call) goes H2D (s_h2d) -> kernels (ctx->stream) ->
// D2H (s_d2h) through device staging slot seq & 1, so the copies of one submission overlap the kernels of its
// neighbours.  At most two submissions are in flight.
static int host_drain(w2l_ctx* ctx, int keep) {
    while (ctx->host_inflight > keep) {
        const long long oldest = ctx->host_seq - ctx->host_inflight;
        CK(cudaEventSynchronize(ctx->ev_out[oldest & 1]));
        ctx->host_inflight--;
    }
    return W2L_OK;
}

static int host_submit(w2l_ctx* ctx, int B, int T, const void* mel_h, size_t mel_bytes, const void* face_h, size_t face_bytes,
                       void* out_h, size_t out_bytes, bool u8) {
    if (ctx->host_inflight >= 2) CKR(host_drain(ctx, 1));
    const int sl = (int)(ctx->host_seq & 1);
    if (ctx->stage_bytes[0 + sl] < mel_bytes || ctx->stage_bytes[2 + sl] < face_bytes || ctx->stage_bytes[4 + sl] < out_bytes) {
        CKR(host_drain(ctx, 0));  // growing a staging buffer frees the old one
        CKR(ensure_stage(ctx, 0 + sl, mel_bytes));
        CKR(ensure_stage(ctx, 2 + sl, face_bytes));
        CKR(ensure_stage(ctx, 4 + sl, out_bytes));
    }
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_GENERATOR, B, T, &pl));
    // (waiting on an event that was never recorded is a no-op)
    CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[sl], 0));   // the kernels that read staging_in[sl] two submissions ago
    CK(cudaMemcpyAsync(ctx->stage[0 + sl], mel_h, mel_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaMemcpyAsync(ctx->stage[2 + sl], face_h, face_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaEventRecord(ctx->ev_in[sl], ctx->s_h2d));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[sl], 0));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[sl], 0));   // staging_out[sl] drained
    CKR(run_plan(ctx, pl, ctx->stage[0 + sl], ctx->stage[2 + sl], ctx->stage[4 + sl], nullptr, ctx->stream, u8));
    CK(cudaEventRecord(ctx->ev_done[sl], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[sl], 0));
    CK(cudaMemcpyAsync(out_h, ctx->stage[4 + sl], out_bytes, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CK(cudaEventRecord(ctx->ev_out[sl], ctx->s_d2h));
    ctx->host_seq++;
    ctx->host_inflight++;
    return W2L_OK;
}

static int host_chunks(int B) {
    int n = B >= 64 ? 2 : 1;  // fewer, larger chunks: small batches run the low-resolution layers inefficiently
    if (const char* ev = getenv("W2L_HOST_CHUNKS")) n = std::max(1, std::min(atoi(ev), B));
    return n;
}

int w2l_generator_forward_host(w2l_ctx* ctx, const float* mel_h, const float* face_h, float* out_h, int B, int T) {
    if (!ctx || !mel_h || !face_h || !out_h) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || T < 0) return fail(W2L_EINVAL, "bad batch B=%d T=%d", B, T);
    DeviceGuard g(ctx->device);
    // A synchronous call cuts the batch along B (whole T-windows, so every chunk is itself a legal call) so that the
    // copies of one chunk overlap the kernels of the other.
    const int tt = T > 0 ? T : 1;
    const int cb = (B + host_chunks(B) - 1) / host_chunks(B);
    const size_t per_b_mel = (size_t)tt * 1280 * 4, per_b_face = (size_t)tt * 6 * 9216 * 4, per_b_out = (size_t)tt * 3 * 9216 * 4;
    for (int b0 = 0; b0 < B; b0 += cb) {
        const int bc = std::min(cb, B - b0);
        CKR(host_submit(ctx, bc, T, (const char*)mel_h + b0 * per_b_mel, bc * per_b_mel, (const char*)face_h + b0 * per_b_face,
                        bc * per_b_face, (char*)out_h + b0 * per_b_out, bc * per_b_out, false));
    }
    return host_drain(ctx, 0);
}

int w2l_generator_submit_host(w2l_ctx* ctx, const float* mel_h, const float* face_h, float* out_h, int B, int T) {
    if (!ctx || !mel_h || !face_h || !out_h) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || T < 0) return fail(W2L_EINVAL, "bad batch B=%d T=%d", B, T);
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)B * (T > 0 ? T : 1);
    return host_submit(ctx, B, T, mel_h, n * 1280 * 4, face_h, n * 6 * 9216 * 4, out_h, n * 3 * 9216 * 4, false);
}

int w2l_generator_submit_u8_host(w2l_ctx* ctx, const float* mel_h, const uint8_t* faces_h, uint8_t* out_h, int N) {
    if (!ctx || !mel_h || !faces_h || !out_h) return fail(W2L_EINVAL, "null argument");
    if (N <= 0) return fail(W2L_EINVAL, "bad batch %d", N);
    DeviceGuard g(ctx->device);
    return host_submit(ctx, N, 0, mel_h, (size_t)N * 1280 * 4, faces_h, (size_t)N * 96 * 96 * 3, out_h, (size_t)N * 96 * 96 * 3, true);
}

int w2l_host_wait(w2l_ctx* ctx, int keep_in_flight) {
    if (!ctx) return fail(W2L_EINVAL, "null argument");
    if (keep_in_flight < 0) keep_in_flight = 0;
    DeviceGuard g(ctx->device);
    return host_drain(ctx, keep_in_flight);
}

int w2l_generator_forward_u8(w2l_ctx* ctx, const float* mel, const uint8_t* faces, uint8_t* out, int N, void* stream) {
    if (!ctx || !mel || !faces || !out) return fail(W2L_EINVAL, "null argument");
    if (N <= 0) return fail(W2L_EINVAL, "bad batch %d", N);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_GENERATOR, N, 0, &pl));
    return run_plan(ctx, pl, mel, faces, out, nullptr, (cudaStream_t)stream, true);
}

int w2l_generator_forward_u8_host(w2l_ctx* ctx, const float* mel_h, const uint8_t* faces_h, uint8_t* out_h, int N) {
    if (!ctx || !mel_h || !faces_h || !out_h) return fail(W2L_EINVAL, "null argument");
    if (N <= 0) return fail(W2L_EINVAL, "bad batch %d", N);
    DeviceGuard g(ctx->device);
    const int cb = (N + host_chunks(N) - 1) / host_chunks(N);
    const size_t per_mel = 1280 * 4, per_face = 96 * 96 * 3, per_out = 96 * 96 * 3;
    for (int b0 = 0; b0 < N; b0 += cb) {
        const int bc = std::min(cb, N - b0);
        CKR(host_submit(ctx, bc, 0, (const char*)mel_h + b0 * per_mel, bc * per_mel, faces_h + b0 * per_face, bc * per_face,
                        out_h + b0 * per_out, bc * per_out, true));
    }
    return host_drain(ctx, 0);
}


// ---- scope row f2: the two cv2.resize calls and the paste around the generator call (inference.py:126, :269-271) ----
static int upload_boxes(w2l_ctx* ctx, const int32_t* boxes, int N, int F, int H, int W, cudaStream_t st) {
    for (int n = 0; n < N; ++n) {
        const int32_t* b = boxes + 5 * n;
        if (b[0] < 0 || b[0] >= F || b[1] < 0 || b[2] > H || b[1] >= b[2] || b[3] < 0 || b[4] > W || b[3] >= b[4])
            return fail(W2L_EINVAL, "box %d = (frame %d, y %d:%d, x %d:%d) is empty or outside the %d frames of %dx%d", n, b[0], b[1], b[2], b[3], b[4], F, H, W);
    }
    if (ctx->box_cap < N) {
        CK(cudaDeviceSynchronize());
        if (ctx->boxes_dev) cudaFree(ctx->boxes_dev);
        void* p = nullptr;
        CKR(dev_alloc(&p, (size_t)N * 5 * 4));
        ctx->boxes_dev = (int*)p; ctx->box_cap = N;
    }
    CK(cudaMemcpyAsync(ctx->boxes_dev, boxes, (size_t)N * 5 * 4, cudaMemcpyHostToDevice, st));
    return W2L_OK;
}

int w2l_crop_resize_u8(w2l_ctx* ctx, const uint8_t* frames, int F, int H, int W, const int32_t* boxes_host, int N, uint8_t* crops,
                       void* stream) {
    if (!ctx || !frames || !boxes_host || !crops) return fail(W2L_EINVAL, "null argument");
    if (F <= 0 || H <= 0 || W <= 0 || N <= 0) return fail(W2L_EINVAL, "bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    CKR(upload_boxes(ctx, boxes_host, N, F, H, W, st));
    const long long total = (long long)N * 96 * 96;
    crop_resize_kernel<<<(int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16), 256, 0, st>>>(frames, H, W, ctx->boxes_dev, N, 96, crops);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_paste_u8(w2l_ctx* ctx, const uint8_t* pred, const uint8_t* frames, int F, int H, int W, const int32_t* boxes_host, int N,
                 uint8_t* out_frames, void* stream) {
    if (!ctx || !pred || !frames || !boxes_host || !out_frames) return fail(W2L_EINVAL, "null argument");
    if (F <= 0 || H <= 0 || W <= 0 || N <= 0) return fail(W2L_EINVAL, "bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    CKR(upload_boxes(ctx, boxes_host, N, F, H, W, st));
    const long long total = (long long)N * H * W;
    paste_kernel<<<(int)std::min<long long>((total + 255) / 256, ctx->num_sms * 32), 256, 0, st>>>(pred, 96, frames, H, W, ctx->boxes_dev, N, out_frames);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_lipsync_frames_u8(w2l_ctx* ctx, const float* mel, const uint8_t* frames, int F, int H, int W, const int32_t* boxes_host, int N,
                          uint8_t* out_frames, void* stream) {
    if (!ctx || !mel || !frames || !boxes_host || !out_frames) return fail(W2L_EINVAL, "null argument");
    if (F <= 0 || H <= 0 || W <= 0 || N <= 0) return fail(W2L_EINVAL, "bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cb = (size_t)N * 96 * 96 * 3;
    if (ctx->crop_cap < cb) {
        CK(cudaDeviceSynchronize());
        if (ctx->crops_dev) cudaFree(ctx->crops_dev);
        if (ctx->preds_dev) cudaFree(ctx->preds_dev);
        void* p = nullptr;
        CKR(dev_alloc(&p, cb)); ctx->crops_dev = (uint8_t*)p;
        CKR(dev_alloc(&p, cb)); ctx->preds_dev = (uint8_t*)p;
        ctx->crop_cap = cb;
    }
    CKR(w2l_crop_resize_u8(ctx, frames, F, H, W, boxes_host, N, ctx->crops_dev, stream));
    CKR(w2l_generator_forward_u8(ctx, mel, ctx->crops_dev, ctx->preds_dev, N, stream));
    const long long total = (long long)N * H * W;
    paste_kernel<<<(int)std::min<long long>((total + 255) / 256, ctx->num_sms * 32), 256, 0, st>>>(ctx->preds_dev, 96, frames, H, W, ctx->boxes_dev, N, out_frames);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}


// ---- scope row f4: S3FD network (face_detection/detection/sfd/net_s3fd.py:22-129) ----
int w2l_s3fd_out_dims(int H, int W, int32_t* dims12) {
    if (!dims12 || H < 32 || W < 32) return fail(W2L_EINVAL, "S3FD needs an image of at least 32 x 32");
    int hs[6], ws[6];
    s3fd_dims(H, W, hs, ws);
    for (int i = 0; i < 6; ++i) { dims12[2 * i] = hs[i]; dims12[2 * i + 1] = ws[i]; }
    return W2L_OK;
}

int w2l_s3fd_forward(w2l_ctx* ctx, const float* img, float* const* outs, int B, int H, int W, void* stream) {
    if (!ctx || !img || !outs) return fail(W2L_EINVAL, "null argument");
    for (int i = 0; i < 12; ++i) if (!outs[i]) return fail(W2L_EINVAL, "null output %d", i);
    if (B <= 0 || H < 32 || W < 32) return fail(W2L_EINVAL, "bad shape B=%d H=%d W=%d", B, H, W);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_S3FD, B, 0, &pl, H, W));
    return run_plan(ctx, pl, img, nullptr, nullptr, nullptr, (cudaStream_t)stream, false, outs);
}

int w2l_syncnet_forward(w2l_ctx* ctx, const float* mel, const float* face, float* a_emb, float* v_emb, int B, void* stream) {
    if (!ctx || !mel || !face || !a_emb || !v_emb) return fail(W2L_EINVAL, "null argument");
    if (B <= 0) return fail(W2L_EINVAL, "bad batch %d", B);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_SYNCNET, B, 0, &pl));
    return run_plan(ctx, pl, mel, face, a_emb, v_emb, (cudaStream_t)stream);
}

int w2l_syncnet_forward_frames(w2l_ctx* ctx, const float* mel, const float* frames, float* a_emb, float* v_emb, int B, int T,
                               void* stream) {
    if (!ctx || !mel || !frames || !a_emb || !v_emb) return fail(W2L_EINVAL, "null argument");
    if (B <= 0) return fail(W2L_EINVAL, "bad batch %d", B);
    if (T != 5) return fail(W2L_EINVAL, "SyncNet_color takes syncnet_T = 5 frames (15 channels), got T=%d", T);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_SYNCNET, B, T, &pl));
    return run_plan(ctx, pl, mel, frames, a_emb, v_emb, (cudaStream_t)stream);
}

static int ensure_scratch(w2l_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return W2L_OK;
    CK(cudaDeviceSynchronize());
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr; ctx->scratch_bytes = 0;
    void* p = nullptr;
    CKR(dev_alloc(&p, bytes));
    ctx->scratch = (float*)p; ctx->scratch_bytes = bytes;
    return W2L_OK;
}

int w2l_cosine_bce_loss(w2l_ctx* ctx, const float* a_emb, const float* v_emb, const float* y, int B, int D, float* loss, void* stream) {
    if (!ctx || !a_emb || !v_emb || !loss) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || D <= 0) return fail(W2L_EINVAL, "bad shape B=%d D=%d", B, D);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    CKR(ensure_scratch(ctx, (size_t)std::max(B, 4096) * 4));
    cosine_bce_terms_kernel<<<(B + 3) / 4, 128, 0, st>>>(a_emb, v_emb, y, ctx->scratch, B, D);
    sum_scale_kernel<<<1, 1024, 0, st>>>(ctx->scratch, loss, B, 1.0f / (float)B);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_l1_loss(w2l_ctx* ctx, const float* x, const float* y, int64_t n, float* loss, void* stream) {
    if (!ctx || !x || !y || !loss) return fail(W2L_EINVAL, "null argument");
    if (n <= 0) return fail(W2L_EINVAL, "bad element count");
    if ((((uintptr_t)x) | ((uintptr_t)y)) & 15) return fail(W2L_EINVAL, "inputs must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::min<long long>(std::max<long long>((n / 4 + 255) / 256, 1), (long long)ctx->num_sms * 8);
    CKR(ensure_scratch(ctx, (size_t)std::max(blocks, 4096) * 4));
    l1_partial_kernel<<<blocks, 256, 0, st>>>(x, y, ctx->scratch, (long long)n);
    sum_scale_kernel<<<1, 1024, 0, st>>>(ctx->scratch, loss, blocks, (float)(1.0 / (double)n));
    ctx->launches += 2;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_disc_forward(w2l_ctx* ctx, const float* frames, float* prob, int B, int T, void* stream) {
    if (!ctx || !frames || !prob) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || T <= 0) return fail(W2L_EINVAL, "bad batch B=%d T=%d", B, T);
    DeviceGuard g(ctx->device);
    Plan* pl;
    CKR(get_plan(ctx, W2L_NET_DISC, B, T, &pl));
    return run_plan(ctx, pl, frames, nullptr, prob, nullptr, (cudaStream_t)stream);
}

int w2l_conv_block_forward(w2l_ctx* ctx, const w2l_layer_info* spec, const float* x, int N, int H, int W,
                           const float* weight, const float* bias, const float* bn_w, const float* bn_b,
                           const float* bn_m, const float* bn_v, float* y, void* stream) {
    if (!ctx || !spec || !x || !weight || !y) return fail(W2L_EINVAL, "null argument");
    if (N <= 0 || H <= 0 || W <= 0) return fail(W2L_EINVAL, "bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    Layer L;
    L.name = spec->name[0] ? spec->name : "block";
    L.kind = spec->kind; L.cin = spec->cin; L.cout = spec->cout; L.kh = spec->kh; L.kw = spec->kw;
    L.sh = spec->sh; L.sw = spec->sw; L.ph = spec->ph; L.pw = spec->pw; L.out_pad = spec->out_pad; L.residual = spec->residual != 0;
    if (L.cout % 16 != 0) return fail(W2L_EINVAL, "cout must be a multiple of 16");
    if (L.kh * L.kw > kMaxTaps) return fail(W2L_EINVAL, "kernel too large");
    int Ho, Wo;
    conv_out_dims(L, H, W, &Ho, &Wo);
    if (Ho <= 0 || Wo <= 0) return fail(W2L_EINVAL, "empty output");
    if (L.residual && (L.cin != L.cout || Ho != H || Wo != W)) return fail(W2L_EINVAL, "residual needs same shape");
    const bool saved_fold = ctx->use_fold;
    if (L.residual) ctx->use_fold = false;  // the residual is read from the block input: keep it in the plain NHWC layout
    // a private one-block "network"
    NetW scratch;
    scratch.layers.resize(1);
    int r = load_layer(ctx, &scratch.layers[0], L, weight, bias, bn_w, bn_b, bn_m, bn_v, H == 1 && W == 1, true, st);
    Plan pl;
    pl.net = W2L_NET_DISC; pl.N = N; pl.B = N; pl.T = 0;
    pl.x2 = ctx->x2;
    Act in, out;
    if (r == W2L_OK) r = plan_input_act(&pl, &in, N, H, W, L.cin, scratch.layers[0], L);
    if (r == W2L_OK) r = plan_act(&pl, &out, N, Ho, Wo, L.cout);
    if (r == W2L_OK) {
        add_ingest(&pl, "ingest.x", 0, in, N, L.cin, (long long)L.cin * H * W, (long long)H * W, 0, 0, W);
        r = emit_block(ctx, &pl, scratch, 0, L, in, out, L.residual ? &in : nullptr);
        if (r == W2L_OK) {
            Plan* lp = ctx->last_plan[W2L_NET_DISC];
            r = run_plan(ctx, &pl, x, nullptr, nullptr, nullptr, st);  // (pl.net only labels the plan)
            ctx->last_plan[W2L_NET_DISC] = lp;
        }
    }
    if (r == W2L_OK) {
        const long long total = (long long)N * L.cout * Ho * Wo;
        const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
        if (ctx->bf16) export_kernel<true><<<blocks, 256, 0, st>>>(out.base, y, N, Ho, Wo, L.cout, out.Cs, 0, out.lo_off);
        else export_kernel<false><<<blocks, 256, 0, st>>>(out.base, y, N, Ho, Wo, L.cout, out.Cs, 0, out.lo_off);
        ctx->launches++;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) r = fail(W2L_ECUDA, "conv block failed: %s", cudaGetErrorString(e));
    }
    free_plan(&pl);
    free_layer(scratch.layers[0]);
    ctx->use_fold = saved_fold;
    return r;
}

int w2l_debug_layer_output(w2l_ctx* ctx, int net, int layer, float* y, int* n, int* c, int* h, int* w, void* stream) {
    if (!ctx || net < 0 || net > 3) return fail(W2L_EINVAL, "bad argument");
    Plan* pl = ctx->last_plan[net];
    if (!pl) return fail(W2L_ESTATE, "no forward has run for net %d", net);
    auto it = pl->layer_out.find(layer);
    if (it == pl->layer_out.end()) return fail(W2L_EINVAL, "layer %d has no materialised output (fused head?)", layer);
    const Act& a = it->second;
    if (n) *n = a.N;
    if (c) *c = a.C;
    if (h) *h = a.H;
    if (w) *w = a.W;
    if (!y) return W2L_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)a.N * a.C * a.H * a.W;
    const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
    const uint16_t* src = a.f32 ? (const uint16_t*)((const float*)a.base + a.c_off) : a.ptr();
    if (ctx->bf16) export_kernel<true><<<blocks, 256, 0, st>>>(src, y, a.N, a.H, a.W, a.C, a.Cs, a.f32 ? 1 : 0, a.f32 ? 0 : a.lo_off);
    else export_kernel<false><<<blocks, 256, 0, st>>>(src, y, a.N, a.H, a.W, a.C, a.Cs, a.f32 ? 1 : 0, a.f32 ? 0 : a.lo_off);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int64_t w2l_mel_num_frames(int64_t n_samples) { return n_samples < 0 ? 0 : 1 + n_samples / MEL_HOP; }

int w2l_melspectrogram(w2l_ctx* ctx, const float* wav, int64_t n_samples, float* mel, void* stream) {
    if (!ctx || !wav || !mel) return fail(W2L_EINVAL, "null argument");
    if (n_samples < 2) return fail(W2L_EINVAL, "need at least 2 samples (got %lld)", (long long)n_samples);
    DeviceGuard g(ctx->device);
    MelParams p;
    p.wav = wav; p.L = n_samples; p.mel = mel; p.F = w2l_mel_num_frames(n_samples);
    p.tw = ctx->mel_tw; p.bvals = ctx->mel_bvals; p.boff = ctx->mel_boff; p.bstart = ctx->mel_bstart; p.blen = ctx->mel_blen;
    if (ctx->use_mel_v2) {   // FFT in registers (mel.cuh, round 2); W2L_DISABLE_MELV2=1 selects the shared-memory version
        const long long blocks = (p.F + MEL2_FPB - 1) / MEL2_FPB;
        mel_kernel_v2<<<(unsigned)blocks, MEL2_THREADS, kMel2SmemBytes, (cudaStream_t)stream>>>(p);
    } else {
        const long long blocks = (p.F + MEL_FPB - 1) / MEL_FPB;
        mel_kernel<<<(unsigned)blocks, MEL_FPB * MEL_TPF, kMelSmemBytes, (cudaStream_t)stream>>>(p);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_melspectrogram_host(w2l_ctx* ctx, const float* wav_h, int64_t n_samples, float* mel_h) {
    if (!ctx || !wav_h || !mel_h) return fail(W2L_EINVAL, "null argument");
    if (n_samples < 2) return fail(W2L_EINVAL, "need at least 2 samples (got %lld)", (long long)n_samples);
    DeviceGuard g(ctx->device);
    CKR(host_drain(ctx, 0));  // the staging buffers below are slot 0 of the asynchronous host pipeline
    const int64_t F = w2l_mel_num_frames(n_samples);
    CKR(ensure_stage(ctx, 0, (size_t)n_samples * 4));
    CKR(ensure_stage(ctx, 4, (size_t)F * MEL_BANDS * 4));
    CK(cudaMemcpyAsync(ctx->stage[0], wav_h, (size_t)n_samples * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKR(w2l_melspectrogram(ctx, (const float*)ctx->stage[0], n_samples, (float*)ctx->stage[4], ctx->stream));
    CK(cudaMemcpyAsync(mel_h, ctx->stage[4], (size_t)F * MEL_BANDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return W2L_OK;
}

int64_t w2l_mel_num_chunks(int64_t n_frames, double fps) {
    if (n_frames < 16 || !(fps > 0)) return 0;
    const double mult = 80.0 / fps;  // inference.py:232
    int64_t i = 0;
    while ((int64_t)((double)i * mult) + 16 <= n_frames) ++i;  // :235-239: the first i that overruns becomes the last chunk
    return i + 1;
}

int w2l_mel_chunks(w2l_ctx* ctx, const float* mel, int64_t n_frames, double fps, float* chunks, int64_t n_chunks, void* stream) {
    if (!ctx || !mel || !chunks) return fail(W2L_EINVAL, "null argument");
    if (n_frames < 16) return fail(W2L_EINVAL, "mel shorter than one 16-frame chunk");
    if (n_chunks != w2l_mel_num_chunks(n_frames, fps)) return fail(W2L_EINVAL, "n_chunks %lld does not match w2l_mel_num_chunks = %lld", (long long)n_chunks, (long long)w2l_mel_num_chunks(n_frames, fps));
    DeviceGuard g(ctx->device);
    const long long total = (long long)n_chunks * 1280;
    const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
    mel_chunk_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mel, n_frames, 80.0 / fps, (int)n_chunks, chunks);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_f16_overflow(w2l_ctx* ctx, int clear, int* flag, void* stream) {
    if (!ctx || !flag) return fail(W2L_EINVAL, "null argument");
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaStreamSynchronize(ctx->s_side));
    CK(cudaStreamSynchronize(ctx->stream));
    int v = 0;
    CK(cudaMemcpyFromSymbol(&v, g_f16_overflow, sizeof(int)));
    *flag = v;
    if (clear && v) { const int z = 0; CK(cudaMemcpyToSymbol(g_f16_overflow, &z, sizeof(int))); }
    return W2L_OK;
}


// ================================================================================================
// training
// ================================================================================================
int w2l_train_bind(w2l_ctx* ctx, int net, int n_tensors, const char* const* names, void* const* value_ptrs, void* const* grad_ptrs,
                   const int64_t* numels) {
    if (!ctx || !names || !value_ptrs || !numels) return fail(W2L_EINVAL, "null argument");
    if (net < 0 || net > 2) return fail(W2L_EINVAL, "unknown net %d", net);
    DeviceGuard g(ctx->device);
    TrainState* ts = train_state(ctx);
    CK(cudaDeviceSynchronize());
    // plans bake the bound pointers: drop the ones of this net
    for (auto it = ts->plans.begin(); it != ts->plans.end();) {
        if (it->second->net == net) { free_train_plan(it->second.get()); it = ts->plans.erase(it); }
        else ++it;
    }
    ts->last[net] = nullptr;
    AdamSlot& a = ts->adam[net];
    for (float* p : a.m) cudaFree(p);
    for (float* p : a.v) cudaFree(p);
    if (a.dev) cudaFree(a.dev);
    a = AdamSlot();
    ts->bound[net].clear();
    for (int i = 0; i < n_tensors; ++i) {
        std::string nm = names[i];
        if (nm.rfind("module.", 0) == 0) nm = nm.substr(7);
        ts->bound[net][nm] = ParamRef{(float*)value_ptrs[i], grad_ptrs ? (float*)grad_ptrs[i] : nullptr, (long long)numels[i]};
    }
    ts->is_bound[net] = true;
    return ensure_train_scratch(ctx, 1, 1);
}

int w2l_train_forward(w2l_ctx* ctx, int net, const float* in0, const float* in1, float* out0, float* out1, int B, int T, int flags,
                      void* stream) {
    if (!ctx || !in0 || !out0) return fail(W2L_EINVAL, "null argument");
    if (net < 0 || net > 2 || B <= 0 || T < 0) return fail(W2L_EINVAL, "bad argument net=%d B=%d T=%d", net, B, T);
    if (net != W2L_NET_DISC && !in1) return fail(W2L_EINVAL, "null argument");
    if (net == W2L_NET_SYNCNET && (!out1 || (T != 0 && T != 5))) return fail(W2L_EINVAL, "SyncNet_color: T must be 0 (stacked faces) or 5 (frames)");
    if (net == W2L_NET_DISC && T <= 0) return fail(W2L_EINVAL, "Wav2Lip_disc_qual takes (B,3,T,96,96)");
    DeviceGuard g(ctx->device);
    TrainPlan* tp;
    const bool wg = net == W2L_NET_GENERATOR ? true : (flags & TRAIN_WGRAD) != 0;
    const bool ig = net == W2L_NET_GENERATOR ? false : (flags & TRAIN_INPUT_GRAD) != 0;
    CKR(get_train_plan(ctx, net, B, T, wg, ig, &tp));
    return train_forward(ctx, tp, in0, in1, out0, out1, flags, (cudaStream_t)stream);
}

int w2l_train_backward(w2l_ctx* ctx, int net, const float* d0, const float* d1, float* dinput, int flags, void* stream) {
    if (!ctx || !d0) return fail(W2L_EINVAL, "null argument");
    if (net < 0 || net > 2) return fail(W2L_EINVAL, "unknown net %d", net);
    DeviceGuard g(ctx->device);
    TrainState* ts = train_state(ctx);
    TrainPlan* tp = ts->last[net];
    if (!tp) return fail(W2L_ESTATE, "backward of net %d before a training forward", net);
    cudaStream_t st = (cudaStream_t)stream;
    if (net == W2L_NET_SYNCNET && !d1) return fail(W2L_EINVAL, "null argument");
    if (dinput && !tp->input_grad) return fail(W2L_ESTATE, "the forward was not run with W2L_TRAIN_INPUT_GRAD");
    if (net == W2L_NET_GENERATOR) flags |= TRAIN_WGRAD;
    CKR(train_backward(ctx, tp, d0, d1, flags, st));
    if (dinput) {
        if (net == W2L_NET_SYNCNET && tp->T == 0) {
            const long long total = (long long)tp->N * 15 * 48 * 96;
            const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
            if (ctx->bf16) export_grad_kernel<true><<<blocks, 256, 0, st>>>(tp->dface_in.ptr(), tp->dface_in.Cs, dinput, tp->N, 48, 96, 15);
            else export_grad_kernel<false><<<blocks, 256, 0, st>>>(tp->dface_in.ptr(), tp->dface_in.Cs, dinput, tp->N, 48, 96, 15);
        } else {
            GenLossGradParams lp;
            memset(&lp, 0, sizeof(lp));
            lp.dg = dinput; lp.B = tp->B; lp.T = tp->T;
            if (net == W2L_NET_SYNCNET) lp.dsync = tp->dface_in.ptr(); else lp.ddisc = tp->dframes_in.ptr();
            const long long total = (long long)tp->B * 3 * tp->T * 9216;
            const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
            if (ctx->bf16) gen_loss_grad_kernel<true><<<blocks, 256, 0, st>>>(lp);
            else gen_loss_grad_kernel<false><<<blocks, 256, 0, st>>>(lp);
        }
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return W2L_OK;
}

int w2l_adam_step(w2l_ctx* ctx, int net, float lr, float beta1, float beta2, float eps, void* stream) {
    if (!ctx || net < 0 || net > 2) return fail(W2L_EINVAL, "bad argument");
    DeviceGuard g(ctx->device);
    TrainState* ts = train_state(ctx);
    if (!ts->is_bound[net]) return fail(W2L_ESTATE, "adam: net %d is not bound", net);
    return adam_step(ctx, net, lr, beta1, beta2, eps, 1.0f, (cudaStream_t)stream);
}

int w2l_comm_unique_id(w2l_ctx* ctx, char* id128) {
    if (!ctx || !id128) return fail(W2L_EINVAL, "null argument");
    TrainState* ts = train_state(ctx);
    NcclGetUniqueIdFn f = (NcclGetUniqueIdFn)nccl_sym(ts, "ncclGetUniqueId");
    if (!f) return fail(W2L_ENODEV, "NCCL not found in the process (libnccl.so.2)");
    const int rc = f(id128);
    if (rc != 0) return fail(W2L_ECUDA, "ncclGetUniqueId failed (%d)", rc);
    return W2L_OK;
}

int w2l_comm_init(w2l_ctx* ctx, const char* id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(W2L_EINVAL, "bad argument");
    DeviceGuard g(ctx->device);
    TrainState* ts = train_state(ctx);
    if (ts->comm) return fail(W2L_ESTATE, "communicator already initialised");
    ts->rank = rank; ts->world = world;
    if (world == 1) return W2L_OK;
    NcclCommInitRankFn init = (NcclCommInitRankFn)nccl_sym(ts, "ncclCommInitRank");
    ts->all_reduce = (NcclAllReduceFn)nccl_sym(ts, "ncclAllReduce");
    ts->comm_destroy = (NcclCommDestroyFn)nccl_sym(ts, "ncclCommDestroy");
    ts->err_string = (NcclGetErrorStringFn)nccl_sym(ts, "ncclGetErrorString");
    if (!init || !ts->all_reduce) return fail(W2L_ENODEV, "NCCL not found in the process (libnccl.so.2)");
    NcclId id;
    memcpy(id.bytes, id128, 128);
    const int rc = init(&ts->comm, world, id, rank);
    if (rc != 0) { ts->comm = nullptr; return fail(W2L_ECUDA, "ncclCommInitRank failed: %s", ts->err_string ? ts->err_string(rc) : "?"); }
    CK(cudaStreamCreateWithFlags(&ts->s_comm, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ts->ev_bucket, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ts->ev_comm, cudaEventDisableTiming));
    return W2L_OK;
}

/* one iteration of wav2lip_train.py:210-231 on the bound generator (+ frozen expert), everything on `stream` */
int w2l_wav2lip_train_step(w2l_ctx* ctx, const float* indiv_mels, const float* x, const float* mel, const float* gt, int B, int T,
                           float syncnet_wt, float lr, float* losses_dev, void* stream) {
    if (!ctx || !indiv_mels || !x || !gt) return fail(W2L_EINVAL, "null argument");
    if (B <= 0 || T <= 0) return fail(W2L_EINVAL, "bad batch B=%d T=%d", B, T);
    if (syncnet_wt > 0.0f && (!mel || T != 5)) return fail(W2L_EINVAL, "the sync loss needs mel and T == 5 (syncnet_T)");
    DeviceGuard g(ctx->device);
    TrainState* ts = train_state(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    CKR(ensure_train_scratch(ctx, B, T));
    TrainPlan *gp, *sp = nullptr;
    CKR(get_train_plan(ctx, W2L_NET_GENERATOR, B, T, true, false, &gp));
    if (syncnet_wt > 0.0f) CKR(get_train_plan(ctx, W2L_NET_SYNCNET, B, T, false, true, &sp));
    float* L = ts->loss_dev;   // [0] sync, [1] l1, [2] perceptual, [3] total
    CK(cudaMemsetAsync(L, 0, 4 * 4, st));
    CKR(train_forward(ctx, gp, indiv_mels, x, ts->g_buf, nullptr, 0, st));
    const long long numel = (long long)B * 3 * T * 9216;
    GenLossGradParams lp;
    memset(&lp, 0, sizeof(lp));
    if (sp) {
        // get_sync_loss (:192-198); the scripts leave the frozen expert in train mode (:187-189): batch statistics, running
        // averages move
        CKR(train_forward(ctx, sp, mel, ts->g_buf, ts->a_emb, ts->v_emb, 0, st));
        CKR(w2l_cosine_bce_loss(ctx, ts->a_emb, ts->v_emb, nullptr, B, 512, L + 0, st));
        cosine_bce_bwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(ts->a_emb, ts->v_emb, nullptr, syncnet_wt, ts->da, ts->dv, B, 512);
        ctx->launches++;
        CKR(train_backward(ctx, sp, ts->da, ts->dv, 0, st));
        lp.dsync = sp->dface_in.ptr();
    }
    CKR(w2l_l1_loss(ctx, ts->g_buf, gt, numel, L + 1, st));
    lp.g = ts->g_buf; lp.gt = gt; lp.dg = ts->dg_buf; lp.l1_scale = (1.0f - syncnet_wt) / (float)numel; lp.B = B; lp.T = T;
    {
        const int blocks = (int)std::min<long long>((numel + 255) / 256, ctx->num_sms * 16);
        if (ctx->bf16) gen_loss_grad_kernel<true><<<blocks, 256, 0, st>>>(lp);
        else gen_loss_grad_kernel<false><<<blocks, 256, 0, st>>>(lp);
        ctx->launches++;
    }
    CKR(generator_backward_dp(ctx, gp, ts->dg_buf, st));
    CKR(adam_step(ctx, W2L_NET_GENERATOR, lr, 0.9f, 0.999f, 1e-8f, 1.0f, st));
    combine_losses_kernel<<<1, 32, 0, st>>>(L, syncnet_wt, 0.0f);
    ctx->launches++;
    if (losses_dev) CK(cudaMemcpyAsync(losses_dev, L, 4 * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    return W2L_OK;
}

/* generator output of the last fused step (B,3,T,96,96) fp32, device pointer owned by the context (tests / logging) */
int w2l_train_last_output(w2l_ctx* ctx, float* out, int64_t n, void* stream) {
    if (!ctx || !out || !ctx->train || !ctx->train->g_buf) return fail(W2L_ESTATE, "no fused training step has run");
    if (n <= 0 || (size_t)n > ctx->train->g_cap) return fail(W2L_EINVAL, "bad element count");
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(out, ctx->train->g_buf, (size_t)n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return W2L_OK;
}

double w2l_train_flops(w2l_ctx* ctx, int net) {
    if (!ctx || !ctx->train || net < 0 || net > 2 || !ctx->train->last[net]) return 0.0;
    return ctx->train->last[net]->fwd_flops;
}


/* Per-stage CUDA-event times of the last training plan of `net` (after a forward + backward have run, so every buffer holds
 * real data): for each block "<name> fwd / stats+apply / bwd_bn / dgrad / wgrad", `iters` back-to-back repetitions each.
 * Returns the number of rows written (<= cap). */
int w2l_train_profile(w2l_ctx* ctx, int net, int iters, int cap, float* ms_out, double* flop_out, char (*names_out)[64], void* stream) {
    if (!ctx || net < 0 || net > 2 || iters <= 0 || !ctx->train) return fail(W2L_EINVAL, "bad argument");
    TrainPlan* tp = ctx->train->last[net];
    if (!tp) return fail(W2L_ESTATE, "no training forward has run for net %d", net);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int k = 0;
    const bool pdl = ctx->use_pdl;
    ctx->use_pdl = false;
    auto timed = [&](const std::string& name, double flops, const std::function<int()>& fn) -> int {
        if (k >= cap) return W2L_OK;
        CKR(fn());
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) CKR(fn());
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_out) ms_out[k] = ms / iters;
        if (flop_out) flop_out[k] = flops;
        if (names_out) snprintf(names_out[k], 64, "%s", name.c_str());
        ++k;
        return W2L_OK;
    };
    int r = W2L_OK;
    for (TBlock& b : tp->blocks) {
        double f = 0;
        for (size_t i = b.fwd0; i < b.fwd1; ++i) f += tp->pl.ops[i].flops;
        // forward conv only, then the statistics + normalise passes (running averages untouched)
        r = timed(b.L.name + " fwd", f, [&]() -> int { for (size_t i = b.fwd0; i < b.fwd1; ++i) CKR(launch_conv(ctx, tp->pl.ops[i], st)); return W2L_OK; });
        if (r != W2L_OK) break;
        if (b.bn) {
            TBlock c = b; c.fwd0 = c.fwd1 = 0;
            r = timed(b.L.name + " bn", 0, [&]() -> int { return block_forward(ctx, tp, c, false, st); });
            if (r != W2L_OK) break;
        }
        {
            TBlock c = b; c.dg0 = c.dg1 = 0; c.wg.on = false;
            r = timed(b.L.name + " bwd_bn", 0, [&]() -> int { return block_backward(ctx, tp, c, false, false, st); });
            if (r != W2L_OK) break;
        }
        if (b.dg1 > b.dg0) {
            double fd = 0;
            for (size_t i = b.dg0; i < b.dg1; ++i) fd += tp->pl.ops[i].flops;
            r = timed(b.L.name + " dgrad", fd, [&]() -> int { for (size_t i = b.dg0; i < b.dg1; ++i) CKR(launch_conv(ctx, tp->pl.ops[i], st)); return W2L_OK; });
            if (r != W2L_OK) break;
        }
        if (b.wg.on) {
            r = timed(b.L.name + " wgrad", b.wg.flops, [&]() -> int { return launch_wgrad(ctx, tp, b, false, st); });
            if (r != W2L_OK) break;
        }
    }
    ctx->use_pdl = pdl;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return r == W2L_OK ? k : r;
}

/* One block, train mode, forward + backward (the operator-level entry of the per-geometry gradient tests):
 *   y = block(x) with batch statistics; given dy: dx, dw, db, dgamma, dbeta; running stats updated in place. */
int w2l_conv_block_train(w2l_ctx* ctx, const w2l_layer_info* spec, const float* x, int N, int H, int W, float* weight, float* bias,
                         float* bn_weight, float* bn_bias, float* bn_mean, float* bn_var, const float* dy, float* y, float* dx,
                         float* dw, float* db, float* dgamma, float* dbeta, void* stream) {
    if (!ctx || !spec || !x || !weight || !y) return fail(W2L_EINVAL, "null argument");
    if (!ctx->bf16) return fail(W2L_ESTATE, "training runs with bf16 operands: create the context with W2L_PREC_BF16");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    Layer L;
    L.name = "block";
    L.kind = spec->kind; L.cin = spec->cin; L.cout = spec->cout; L.kh = spec->kh; L.kw = spec->kw;
    L.sh = spec->sh; L.sw = spec->sw; L.ph = spec->ph; L.pw = spec->pw; L.out_pad = spec->out_pad; L.residual = spec->residual != 0;
    if (L.cout % 16 != 0) return fail(W2L_EINVAL, "cout must be a multiple of 16");
    int Ho, Wo;
    conv_out_dims(L, H, W, &Ho, &Wo);
    if (Ho <= 0 || Wo <= 0) return fail(W2L_EINVAL, "empty output");
    TrainState* ts = train_state(ctx);
    CKR(ensure_train_scratch(ctx, 1, 1));
    const int slot = W2L_NET_DISC;
    std::map<std::string, ParamRef> saved;
    saved.swap(ts->bound[slot]);
    const long long wn = (long long)L.cin * L.cout * L.kh * L.kw;
    ts->bound[slot]["block.conv_block.0.weight"] = ParamRef{weight, dw, wn};
    ts->bound[slot]["block.conv_block.0.bias"] = ParamRef{bias, db, L.cout};
    const bool bn = L.kind == W2L_BLOCK_CONV_BN_RELU || L.kind == W2L_BLOCK_CONVT_BN_RELU;
    if (bn) {
        ts->bound[slot]["block.conv_block.1.weight"] = ParamRef{bn_weight, dgamma, L.cout};
        ts->bound[slot]["block.conv_block.1.bias"] = ParamRef{bn_bias, dbeta, L.cout};
        if (bn_mean) ts->bound[slot]["block.conv_block.1.running_mean"] = ParamRef{bn_mean, nullptr, L.cout};
        if (bn_var) ts->bound[slot]["block.conv_block.1.running_var"] = ParamRef{bn_var, nullptr, L.cout};
    }
    TrainPlan tp;
    tp.net = slot; tp.N = N; tp.B = N; tp.T = 0;
    const bool s_fold = ctx->use_fold, s_rs = ctx->use_rowstack;
    ctx->use_fold = false; ctx->use_rowstack = false;
    Act xin, dxin, yv, dyv, none;
    size_t ws_need = 0;
    int r = tp_act(&tp, &xin, N, H, W, round_up(L.cin, 16));
    if (r == W2L_OK) r = tp_act(&tp, &dxin, N, H, W, round_up(L.cin, 16));
    if (r == W2L_OK) r = tp_act(&tp, &yv, N, Ho, Wo, L.cout);
    if (r == W2L_OK) r = tp_act(&tp, &dyv, N, Ho, Wo, L.cout);
    if (r == W2L_OK) {
        add_train_ingest(&tp, "ingest.x", 0, xin, N, L.cin, (long long)L.cin * H * W, (long long)H * W, 0, 0, W);
        add_train_ingest(&tp, "ingest.dy", 1, dyv, N, L.cout, (long long)L.cout * Ho * Wo, (long long)Ho * Wo, 0, 0, Wo);
        r = add_train_block(ctx, &tp, slot, 0, L, xin, yv, dyv, dx ? dxin : none, none, dw != nullptr,
                            L.kind == W2L_BLOCK_CONVT_BN_RELU && H == 1 && W == 1, &ws_need);
    }
    ctx->use_fold = s_fold; ctx->use_rowstack = s_rs;
    if (r == W2L_OK && ws_need) { void* p = nullptr; r = plan_alloc(&tp.pl, &p, ws_need); tp.wg_ws = (float*)p; }
    if (r == W2L_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) r = fail(W2L_ECUDA, "plan build failed: %s", cudaGetErrorString(e));
    }
    if (r == W2L_OK) r = repack_weights(ctx, &tp, st);
    if (r == W2L_OK) r = launch_ingest(ctx, tp.pl.ops[tp.ingest[0]], x, st);
    if (r == W2L_OK) r = block_forward(ctx, &tp, tp.blocks[0], true, st);
    if (r == W2L_OK) {
        const long long total = (long long)N * L.cout * Ho * Wo;
        const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
        if (ctx->bf16) export_kernel<true><<<blocks, 256, 0, st>>>(yv.base, y, N, Ho, Wo, L.cout, yv.Cs, 0, 0);
        else export_kernel<false><<<blocks, 256, 0, st>>>(yv.base, y, N, Ho, Wo, L.cout, yv.Cs, 0, 0);
        ctx->launches++;
    }
    if (r == W2L_OK && dy) {
        r = launch_ingest(ctx, tp.pl.ops[tp.ingest[1]], dy, st);
        if (r == W2L_OK) r = block_backward(ctx, &tp, tp.blocks[0], dw != nullptr, false, st);
        if (r == W2L_OK && dx) {
            const long long total = (long long)N * L.cin * H * W;
            const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
            if (ctx->bf16) export_grad_kernel<true><<<blocks, 256, 0, st>>>(dxin.ptr(), dxin.Cs, dx, N, H, W, L.cin);
            else export_grad_kernel<false><<<blocks, 256, 0, st>>>(dxin.ptr(), dxin.Cs, dx, N, H, W, L.cin);
            ctx->launches++;
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (r == W2L_OK && e != cudaSuccess) r = fail(W2L_ECUDA, "conv block train failed: %s", cudaGetErrorString(e));
    free_train_plan(&tp);
    ts->bound[slot].swap(saved);
    return r;
}

int64_t w2l_launch_count(const w2l_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t w2l_device_bytes(const w2l_ctx* ctx) {
    if (!ctx) return 0;
    size_t b = ctx->weight_bytes;
    for (auto& kv : ctx->plans) b += kv.second->bytes;
    for (int i = 0; i < 6; ++i) b += ctx->stage_bytes[i];
    return (int64_t)b;
}

int w2l_profile_plan(w2l_ctx* ctx, int net, int iters, int cap, float* ms_out, double* flop_out, char (*names_out)[64], void* stream) {
    if (!ctx || net < 0 || net > 3 || iters <= 0) return fail(W2L_EINVAL, "bad argument");
    Plan* pl = ctx->last_plan[net];
    if (!pl) return fail(W2L_ESTATE, "no forward has run for net %d", net);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    // Cold-cache timing: the 126 MB L2 is flushed (a 256 MB buffer is overwritten) before EVERY timed launch, so that layers
    // whose tensors fit the L2 are not timed warm (the timing rule: flush L2 between timed iterations).
    const size_t flush_bytes = (size_t)256 << 20;
    void* flush = nullptr;
    if (cudaMalloc(&flush, flush_bytes) != cudaSuccess) { cudaGetLastError(); flush = nullptr; }
    int k = 0;
    int r = W2L_OK;
    for (Op& op : pl->ops) {
        if (op.type != OP_CONV || k >= cap) continue;
        if (op.head && op.cp.ep.head_out == nullptr && op.pp.ep.head_out == nullptr && op.cp.ep.head_out_u8 == nullptr) continue;
        if ((r = launch_conv(ctx, op, st, false)) != W2L_OK) break;  // warm the instruction cache / attributes
        float total = 0.0f;
        for (int i = 0; i < iters && r == W2L_OK; ++i) {
            if (flush) cudaMemsetAsync(flush, i, flush_bytes, st);
            cudaEventRecord(e0, st);
            r = launch_conv(ctx, op, st, false);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) { r = fail(W2L_ECUDA, "profile: %s", cudaGetErrorString(cudaGetLastError())); break; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
        }
        if (r != W2L_OK) break;
        if (ms_out) ms_out[k] = total / iters;
        if (flop_out) flop_out[k] = op.flops;
        if (names_out) snprintf(names_out[k], 64, "%s", op.name.c_str());
        ++k;
    }
    if (flush) cudaFree(flush);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return r == W2L_OK ? k : r;
}

}  // extern "C"
