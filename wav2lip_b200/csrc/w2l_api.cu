// w2l_api.cu — host side of libw2l.so: the C-ABI of include/w2l.h, the per-batch-size execution plans
// (buffers, TMA tensor maps, kernel parameters) and the launch loops.  No torch, no CPU compute path.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/w2l.h"
#include "aux_kernels.cuh"
#include "conv_patch.cuh"
#include "conv_rowstack.cuh"
#include "conv_tcgen05.cuh"
#include "convt_fused.cuh"
#include "mel.cuh"
#include "netspec.h"

using namespace w2l;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(W2L_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKR(expr)               \
    do {                        \
        int r_ = (expr);        \
        if (r_ != W2L_OK) return r_; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// specs (built once, host only)
// ------------------------------------------------------------------------------------------------
static const GeneratorSpec& gen_spec() { static GeneratorSpec s = build_generator_spec(); return s; }
static const SyncnetSpec& sync_spec() { static SyncnetSpec s = build_syncnet_spec(); return s; }
static const DiscSpec& disc_spec() { static DiscSpec s = build_disc_spec(); return s; }
static const std::vector<Layer>* net_layers(int net) {
    switch (net) {
        case W2L_NET_GENERATOR: return &gen_spec().layers;
        case W2L_NET_SYNCNET: return &sync_spec().layers;
        case W2L_NET_DISC: return &disc_spec().layers;
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// ------------------------------------------------------------------------------------------------
// tensors in HBM
// ------------------------------------------------------------------------------------------------
// Activations are NHWC, 16-bit (fp16 or bf16), channel pitch Cs; a view may select a channel slice
// [c_off, c_off + C) of a wider buffer (the skip-concat buffers of the decoder).
struct Act {
    uint16_t* base = nullptr;  // start of the buffer (not of the slice)
    int N = 0, H = 0, W = 0;
    int Cs = 0;     // channel pitch of the buffer
    int c_off = 0;  // first channel of this view
    int C = 0;      // channels of this view
    bool f32 = false;
    int Wp = 0;     // row pitch in pixels (0 = W); > W only for the zero-bordered first-layer inputs
    int x_off = 0;  // left border of those inputs
    int lo_off = 0;   // split-operand mode: channel distance from the hi plane to the lo plane of the same pixel
    int wstride = 1;  // folded views: pixels between consecutive windows (= the conv's horizontal stride)
    int nwin = 0;     // folded views: number of windows per row (= output width); 0 = W
    int pitch() const { return Wp ? Wp : W; }
    uint16_t* ptr() const { return base + c_off; }
    Act slice(int off, int c) const { Act a = *this; a.c_off = c_off + off; a.C = c; return a; }
};

struct PackedW {
    uint16_t* w = nullptr;  // [ntaps][cout_pad][cin_pad]
    int ntaps = 0, cout_pad = 0, cin_pad = 0;
    int nslabs = 0;                   // weight slabs stored: ntaps, or 2*ntaps (hi then lo) in the split-operand mode
    std::vector<signed char> dx, dy;  // input offset of each tap relative to (out * stride)
    int py = 0, px = 0;               // output phase (transposed conv)
    // "kw folded into K" form for tiny-Cin first layers: one K row = kw taps x Cp channels (zero padded to kfold)
    bool fold = false;
    int Cp = 0, kfold = 0, win = 0;   // channel pitch of the input, folded K per filter row, pixels spanned by a window
};

struct LayerW {
    std::vector<PackedW> ph;  // 1 for conv, 4 for stride-2 convT, 1 (as GEMM) for the 1x1->3x3 convT
    float* scale = nullptr;
    float* shift = nullptr;
    int n_scale = 0;
    bool gemm_convT = false;
    bool has_all_taps = false;  // ph.back() holds all 9 taps of a stride-2 transposed conv (fused 4-phase kernel)
    bool loaded = false;
};

struct NetW {
    std::vector<LayerW> layers;
    float* head_w = nullptr;  // generator output_block.1 (3x32) / disc binary_pred (512)
    float* head_b = nullptr;
    bool loaded = false;
};

enum OpType { OP_CONV = 0, OP_INGEST = 1, OP_L2NORM = 2, OP_DISC_HEAD = 3 };

struct Op {
    int type = OP_CONV;
    std::string name;
    // conv
    ConvParams cp;
    int BN = 0, BK = 0, MT = 1;
    bool head = false;
    int grid = 0;
    double flops = 0;  // algorithmic (true MACs*2), not padded
    bool patch = false;  // conv_patch_kernel instead of conv_igemm_kernel
    PatchParams pp;
    int dyn_smem = 0;
    bool ctf = false;   // convt_fused_kernel
    ConvTParams tp;
    bool rowstack = false;  // conv_rowstack_kernel
    int rs_shape = 0;       // 0: output block (C=32, S=2, 3x3, head)   1: folded 7-row first block (C=16, S=3)
    RowStackParams rs;
    // ingest
    IngestParams ip;
    int ingest_src = 0;  // which caller tensor: 0 = mel / frames, 1 = face
    // l2norm / disc head
    const void* aux_in = nullptr;
    int aux_rows = 0, aux_dim = 0;
    int aux_out = 0;  // which caller output
    int aux_pitch = 0, aux_lo = 0;
    int lane = 0;          // 1: runs on the context's side stream (the audio encoder, concurrently with the face encoder)
    bool join_side = false;  // wait for the side stream before this op
};

struct Plan {
    int net = 0, B = 0, T = 0, N = 0;
    std::vector<Op> ops;
    std::vector<void*> allocs;
    size_t bytes = 0;
    std::map<int, Act> layer_out;  // layer index -> activation view (debug export)
    long long last_used = 0;       // LRU stamp
    bool x2 = false;               // split-operand precision: activations carry hi and lo planes
    bool has_side = false;         // some ops run on the side stream
};

struct w2l_ctx {
    int device = 0;
    bool bf16 = false;
    bool x2 = false;        // W2L_PREC_F32X: split fp16 operands (hi + lo), generic kernel only
    int num_sms = 148;
    bool keep_all = false;  // debug: no buffer reuse, every layer output stays readable
    bool use_patch = true;   // W2L_DISABLE_HALO=1 turns the patch kernel off (A/B testing)
    bool use_bn256 = true;  // W2L_DISABLE_BN256=1
    bool use_mt2 = true;    // W2L_DISABLE_MT2=1
    bool use_tma_epi = true;  // W2L_DISABLE_TMAEPI=1
    bool use_fold_s2 = true;  // W2L_DISABLE_FOLDS2=1
    bool use_ctfused = true;  // W2L_DISABLE_CTFUSED=1
    bool use_fold = true;   // W2L_DISABLE_FOLD=1 / driver rejects overlapping-stride tensor maps
    bool use_pdl = true;      // W2L_DISABLE_PDL=1
    bool use_rowstack = true;  // W2L_DISABLE_ROWSTACK=1
    NetW nets[3];
    std::map<std::string, std::unique_ptr<Plan>> plans;
    Plan* last_plan[3] = {nullptr, nullptr, nullptr};
    int64_t launches = 0;
    long long plan_clock = 0;
    size_t weight_bytes = 0;
    // host-buffer entry points: compute stream + copy streams, double-buffered device staging
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_side = nullptr;   // audio-encoder lane of the generator plan
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool use_side = true;            // W2L_DISABLE_SIDESTREAM=1
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    void* stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* scratch = nullptr;  // partial sums of the loss kernels
    size_t scratch_bytes = 0;
    long long host_seq = 0;   // host-buffer submissions so far (staging slot = seq & 1)
    int host_inflight = 0;    // submitted and not yet retired by host_drain
    size_t stage_bytes[6] = {0, 0, 0, 0, 0, 0};
    // mel tables
    double2* mel_tw = nullptr;
    float* mel_bvals = nullptr;
    int* mel_boff = nullptr;
    int* mel_bstart = nullptr;
    int* mel_blen = nullptr;
};

static int dev_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(W2L_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return W2L_OK;
}

static int plan_alloc(Plan* pl, void** p, size_t bytes) {
    CKR(dev_alloc(p, bytes));
    pl->allocs.push_back(*p);
    pl->bytes += bytes;
    return W2L_OK;
}

static int plan_act(Plan* pl, Act* a, int N, int H, int W, int C, bool f32 = false) {
    void* p = nullptr;
    const bool planes = pl->x2 && !f32;
    const size_t bytes = (size_t)N * H * W * C * (f32 ? 4 : 2) * (planes ? 2 : 1);
    CKR(plan_alloc(pl, &p, bytes));
    a->base = (uint16_t*)p;
    a->N = N; a->H = H; a->W = W; a->Cs = planes ? 2 * C : C; a->c_off = 0; a->C = C; a->f32 = f32;
    a->lo_off = planes ? C : 0;
    return W2L_OK;
}

// Buffer the ingest kernel fills for the first block of a chain. Folded first layers read it through an
// overlapping-window tensor map: channel pitch Cp, rows padded with pw zero pixels on the left and enough
// on the right for the last window; the view handed to the conv is (C = kfold, W windows).
static int plan_input_act(Plan* pl, Act* a, int N, int H, int W, int cin, const LayerW& lw, const Layer& L) {
    const PackedW& w = lw.ph[0];
    if (!w.fold) return plan_act(pl, a, N, H, W, ((cin + 15) / 16) * 16);
    const int Wout = (W + 2 * L.pw - L.kw) / L.sw + 1;
    const int Wp = (std::max(W + L.pw, (Wout - 1) * L.sw + w.win) + 1) / 2 * 2;
    void* p = nullptr;
    const size_t bytes = ((size_t)N * H * Wp * w.Cp + w.kfold) * 2;  // + one window of slack at the very end
    CKR(plan_alloc(pl, &p, bytes));
    CK(cudaMemset(p, 0, bytes));
    a->base = (uint16_t*)p;
    a->N = N; a->H = H; a->W = W; a->Cs = w.Cp; a->c_off = 0; a->C = w.kfold; a->f32 = false;
    a->Wp = Wp; a->x_off = L.pw;
    a->wstride = L.sw; a->nwin = Wout;
    return W2L_OK;
}

static void free_plan(Plan* pl) {
    for (void* p : pl->allocs) cudaFree(p);
    pl->allocs.clear();
}

// ------------------------------------------------------------------------------------------------
// conv kernel dispatch
// ------------------------------------------------------------------------------------------------
typedef void (*ConvKernelFn)(const ConvParams);
struct ConvKernelEntry { int BN, BK; bool bf16, head; ConvKernelFn fn; int smem; uint64_t attr_set; int mt; int threads; };

#define W2L_CONV_ENTRY(BN_, BK_)                                                                                               \
    {BN_, BK_, false, false, conv_igemm_kernel<BN_, BK_, false, false>, ConvCfg<BN_, BK_>::kSmemBytes, 0, 1, 256},         \
    {BN_, BK_, true, false, conv_igemm_kernel<BN_, BK_, true, false>, ConvCfg<BN_, BK_>::kSmemBytes, 0, 1, 256}
#define W2L_CONV_ENTRY_MT2(BN_, BK_)                                                                                           \
    {BN_, BK_, false, false, conv_igemm_kernel<BN_, BK_, false, false, 2>, ConvCfg<BN_, BK_, 2>::kSmemBytes, 0, 2, 384},   \
    {BN_, BK_, true, false, conv_igemm_kernel<BN_, BK_, true, false, 2>, ConvCfg<BN_, BK_, 2>::kSmemBytes, 0, 2, 384}

static ConvKernelEntry g_conv_kernels[] = {
    W2L_CONV_ENTRY(16, 16), W2L_CONV_ENTRY(16, 32), W2L_CONV_ENTRY(16, 64),
    W2L_CONV_ENTRY(32, 16), W2L_CONV_ENTRY(32, 32), W2L_CONV_ENTRY(32, 64),
    W2L_CONV_ENTRY(64, 16), W2L_CONV_ENTRY(64, 32), W2L_CONV_ENTRY(64, 64),
    W2L_CONV_ENTRY(128, 16), W2L_CONV_ENTRY(128, 32), W2L_CONV_ENTRY(128, 64),
    W2L_CONV_ENTRY(256, 64),
    W2L_CONV_ENTRY_MT2(128, 64), W2L_CONV_ENTRY_MT2(64, 64), W2L_CONV_ENTRY_MT2(64, 32),
    {32, 16, false, true, conv_igemm_kernel<32, 16, false, true>, ConvCfg<32, 16>::kSmemBytes, 0, 1, 256},
    {32, 16, true, true, conv_igemm_kernel<32, 16, true, true>, ConvCfg<32, 16>::kSmemBytes, 0, 1, 256},
};

static ConvKernelEntry* find_conv_kernel(int BN, int BK, bool bf16, bool head, int mt = 1) {
    for (auto& e : g_conv_kernels)
        if (e.BN == BN && e.BK == BK && e.bf16 == bf16 && e.head == head && e.mt == mt) return &e;
    return nullptr;
}

typedef void (*PatchKernelFn)(const PatchParams);
struct PatchKernelEntry { int BN, BK; bool bf16, head; PatchKernelFn fn; uint64_t attr_set; };
#define W2L_PATCH_ENTRY(BN_, BK_)                                                    \
    {BN_, BK_, false, false, conv_patch_kernel<BN_, BK_, false, false>, 0},    \
    {BN_, BK_, true, false, conv_patch_kernel<BN_, BK_, true, false>, 0}
static PatchKernelEntry g_patch_kernels[] = {
    W2L_PATCH_ENTRY(16, 16), W2L_PATCH_ENTRY(16, 32), W2L_PATCH_ENTRY(16, 64),
    W2L_PATCH_ENTRY(32, 16), W2L_PATCH_ENTRY(32, 32), W2L_PATCH_ENTRY(32, 64),
    W2L_PATCH_ENTRY(64, 16), W2L_PATCH_ENTRY(64, 32), W2L_PATCH_ENTRY(64, 64),
    {32, 16, false, true, conv_patch_kernel<32, 16, false, true>, 0},
    {32, 16, true, true, conv_patch_kernel<32, 16, true, true>, 0},
};
static PatchKernelEntry* find_patch_kernel(int BN, int BK, bool bf16, bool head) {
    for (auto& e : g_patch_kernels)
        if (e.BN == BN && e.BK == BK && e.bf16 == bf16 && e.head == head) return &e;
    return nullptr;
}

typedef void (*RsKernelFn)(const RowStackParams);
struct RsKernelEntry { int shape; bool bf16; RsKernelFn fn; uint64_t attr_set; };
static RsKernelEntry g_rs_kernels[] = {
    {0, false, conv_rowstack_kernel<32, 2, 3, 3, 16, false, true>, 0},
    {0, true, conv_rowstack_kernel<32, 2, 3, 3, 16, true, true>, 0},
    {1, false, conv_rowstack_kernel<16, 3, 7, 1, 0, false, false>, 0},
    {1, true, conv_rowstack_kernel<16, 3, 7, 1, 0, true, false>, 0},
    {2, false, conv_rowstack_kernel<32, 3, 7, 1, 0, false, false>, 0},
    {2, true, conv_rowstack_kernel<32, 3, 7, 1, 0, true, false>, 0},
};
using RsCfg0 = RowStackCfg<32, 2, 3, 3, 16>;   // generator output block (+ head)
using RsCfg1 = RowStackCfg<16, 3, 7, 1, 0>;    // generator first block (6 -> 16, 7x7 folded)
using RsCfg2 = RowStackCfg<32, 3, 7, 1, 0>;    // disc first block (3 -> 32, 7x7 folded, LeakyReLU)
struct RsShape { int PW, PH, tile_h, R, ndx, fixed, per_stage; };
static RsShape rs_shape(int shape) {
    switch (shape) {
        case 0: return {RsCfg0::PW, RsCfg0::PH, RsCfg0::kTileH, 1, 3, RsCfg0::smem_bytes(0, true), RsCfg0::kStageStride};
        case 1: return {RsCfg1::PW, RsCfg1::PH, RsCfg1::kTileH, 3, 1, RsCfg1::smem_bytes(0, false), RsCfg1::kStageStride};
        default: return {RsCfg2::PW, RsCfg2::PH, RsCfg2::kTileH, 3, 1, RsCfg2::smem_bytes(0, false), RsCfg2::kStageStride};
    }
}

typedef void (*CtKernelFn)(const ConvTParams);
struct CtKernelEntry { int BK; bool bf16; CtKernelFn fn; uint64_t attr_set; };
static CtKernelEntry g_ct_kernels[] = {
    {32, false, convt_fused_kernel<32, false>, 0}, {32, true, convt_fused_kernel<32, true>, 0},
    {64, false, convt_fused_kernel<64, false>, 0}, {64, true, convt_fused_kernel<64, true>, 0},
};
constexpr int kCtSmemMax = 227 * 1024;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember it per device (one bit
// each), so that contexts on several GPUs of one process all get it
static int ensure_smem_attr(uint64_t* mask, int device, const void* fn, int bytes) {
    const uint64_t bit = 1ull << (device & 63);
    if (*mask & bit) return W2L_OK;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    *mask |= bit;
    return W2L_OK;
}

// One launch, optionally with programmatic stream serialization (the kernels call griddepcontrol.wait before they
// touch global memory, so their prologue overlaps the previous kernel's tail).
template <typename P>
static cudaError_t launch_k(void (*fn)(const P), int grid, int block, size_t smem, cudaStream_t st, const P& p, bool pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, fn, p);
}

static int launch_conv(w2l_ctx* ctx, const Op& op, cudaStream_t st, bool pdl = true) {
    pdl = pdl && ctx->use_pdl;
    if (op.ctf) {
        CtKernelEntry* e = nullptr;
        for (auto& k : g_ct_kernels) if (k.BK == op.BK && k.bf16 == ctx->bf16) e = &k;
        if (!e) return fail(W2L_EINVAL, "no fused convT kernel for BK=%d", op.BK);
        CKR(ensure_smem_attr(&e->attr_set, ctx->device, (const void*)e->fn, kCtSmemMax));
        CK(launch_k(e->fn, op.grid, kCtThreads, op.dyn_smem, st, op.tp, pdl));
        ctx->launches++;
        return W2L_OK;
    }
    if (op.rowstack) {
        RsKernelEntry* e = nullptr;
        for (auto& k : g_rs_kernels) if (k.shape == op.rs_shape && k.bf16 == ctx->bf16) e = &k;
        if (!e) return fail(W2L_EINVAL, "no row-stack kernel for shape %d", op.rs_shape);
        CKR(ensure_smem_attr(&e->attr_set, ctx->device, (const void*)e->fn, kSmemBudget + kSmemExtra));
        CK(launch_k(e->fn, op.grid, kRsThreads, op.dyn_smem, st, op.rs, pdl));
        ctx->launches++;
        return W2L_OK;
    }
    if (op.patch) {
        PatchKernelEntry* e = find_patch_kernel(op.BN, op.BK, ctx->bf16, op.head);
        if (!e) return fail(W2L_EINVAL, "no patch kernel for BN=%d BK=%d head=%d", op.BN, op.BK, (int)op.head);
        CKR(ensure_smem_attr(&e->attr_set, ctx->device, (const void*)e->fn, kSmemBudget + kSmemExtra));
        CK(launch_k(e->fn, op.grid, kPatchThreads, op.dyn_smem, st, op.pp, pdl));
        ctx->launches++;
        return W2L_OK;
    }
    ConvKernelEntry* e = find_conv_kernel(op.BN, op.BK, ctx->bf16, op.head, op.MT);
    if (!e) return fail(W2L_EINVAL, "no conv kernel for BN=%d BK=%d head=%d MT=%d", op.BN, op.BK, (int)op.head, op.MT);
    CKR(ensure_smem_attr(&e->attr_set, ctx->device, (const void*)e->fn, e->smem));
    CK(launch_k(e->fn, op.grid, e->threads, (size_t)e->smem, st, op.cp, pdl));
    ctx->launches++;
    return W2L_OK;
}

// ------------------------------------------------------------------------------------------------
// building one conv launch
// ------------------------------------------------------------------------------------------------
static int pick_bk(int cin_pad) { return (cin_pad % 64 == 0) ? 64 : (cin_pad % 32 == 0) ? 32 : 16; }
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// The 128-row tile is a (bw x bh x bn) box of output pixels; choose the box with the least padding waste.
static void pick_box(int W, int H, int N, int sx, int sy, int* bw, int* bh, int* bn) {
    double best = 1e30;
    int b_w = 1, b_h = 1, b_n = 1;
    for (int w = 1; w <= std::min(W, kTileM); ++w) {
        if (w * sx > 256) break;
        for (int h = 1; h <= std::min(H, kTileM / w); ++h) {
            if (h * sy > 256) break;
            int n = std::min(kTileM / (w * h), std::max(N, 1));
            if (n < 1) continue;
            if (n > 256) n = 256;
            const double tiles = (double)((W + w - 1) / w) * ((H + h - 1) / h) * ((N + n - 1) / n);
            // prefer wide boxes (longer contiguous runs) on ties
            const double cost = tiles - 1e-6 * w - 1e-9 * h;
            if (cost < best) { best = cost; b_w = w; b_h = h; b_n = n; }
        }
    }
    *bw = b_w; *bh = b_h; *bn = b_n;
}

struct ConvArgs {
    std::string name;
    Act in, out;
    const PackedW* w = nullptr;
    int sx = 1, sy = 1;        // input stride per logical output pixel
    int Hl = 0, Wl = 0;        // logical output grid handled by this launch
    int osy = 1, osx = 1;      // output pixel = logical * os + phase
    int phy = 0, phx = 0;
    const Act* res = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    int ch_off = 0;            // offset into scale/shift
    int act = ACT_RELU;
    int cout = 0;              // channels produced
    double macs_per_pixel = 0; // true MACs per logical output pixel (for flop accounting)
    // fused head
    bool head = false;
    const float* head_w = nullptr;
    const float* head_b = nullptr;
    int head_B = 1, head_T = 1;
};

static void fill_epi(EpiParams* e, const ConvArgs& a) {
    memset(e, 0, sizeof(*e));
    e->Wout = a.Wl; e->Hout = a.Hl; e->N = a.in.N;
    e->act = a.act;
    e->out_f32 = a.out.f32 ? 1 : 0;
    const long long oCs = a.out.Cs;
    const long long Wfull = a.out.W;
    if (!a.head) {
        const long long base_off = ((long long)a.phy * Wfull + a.phx) * oCs + a.out.c_off;
        e->out = a.out.f32 ? (void*)((float*)a.out.base + base_off) : (void*)(a.out.base + base_off);
        e->out_sn = (long long)a.out.H * Wfull * oCs;
        e->out_sy = (long long)a.osy * Wfull * oCs;
        e->out_sx = (long long)a.osx * oCs;
    }
    if (a.res) {
        e->res = a.res->ptr();
        e->res_sn = (long long)a.res->H * a.res->W * a.res->Cs;
        e->res_sy = (long long)a.res->W * a.res->Cs;
        e->res_sx = a.res->Cs;
    }
    e->scale = a.scale + a.ch_off;
    e->shift = a.shift + a.ch_off;
    e->head_w = a.head_w; e->head_b = a.head_b; e->head_out = nullptr; e->head_B = a.head_B; e->head_T = a.head_T;
    e->x2 = (a.out.lo_off > 0 || (a.res && a.res->lo_off > 0)) ? 1 : 0;
    e->out_lo_off = a.out.lo_off;
    e->res_lo_off = a.res ? a.res->lo_off : 0;
}

static int encode_act_map(w2l_ctx* ctx, CUtensorMap* tm, const Act& in, int BK, int bx, int by, int bn, int sx, int sy,
                          const char* name) {
    EncodeTiledFn enc = get_encode_fn();
    const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[4] = {(cuuint64_t)(in.lo_off + in.C), (cuuint64_t)(in.nwin ? in.nwin : in.W), (cuuint64_t)in.H, (cuuint64_t)in.N};
    cuuint64_t strides[3] = {(cuuint64_t)in.Cs * in.wstride * 2, (cuuint64_t)in.pitch() * in.Cs * 2, (cuuint64_t)in.H * in.pitch() * in.Cs * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
    CUresult r = enc(tm, dt, 4, in.ptr(), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(A) failed with %d (dims %d,%d,%d,%d box %d,%d,%d,%d)", name, (int)r,
                    in.C, in.W, in.H, in.N, BK, bx, by, bn);
    return W2L_OK;
}

static int encode_w_map(w2l_ctx* ctx, CUtensorMap* tm, const PackedW& w, int BK, int BN, const char* name) {
    EncodeTiledFn enc = get_encode_fn();
    const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[3] = {(cuuint64_t)w.cin_pad, (cuuint64_t)w.cout_pad, (cuuint64_t)(w.nslabs ? w.nslabs : w.ntaps)};
    cuuint64_t strides[2] = {(cuuint64_t)w.cin_pad * 2, (cuuint64_t)w.cin_pad * w.cout_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, dt, 3, w.w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(B) failed with %d", name, (int)r);
    return W2L_OK;
}

// Few-channel stride-1 layers: one input patch per tile + resident weights (conv_patch.cuh)
struct PatchGeom { int ox, oy, PW, PH, BK, patch_bytes, patch_stride, wbytes, stg_bytes, res_tap; };

static bool patch_eligible(const w2l_ctx* ctx, const ConvArgs& a, PatchGeom* g) {
    if (!ctx->use_patch) return false;
    const PackedW& w = *a.w;
    if (a.sx != 1 || a.sy != 1 || w.ntaps > kPatchMaxTaps) return false;
    if (a.cout != 16 && a.cout != 32 && a.cout != 64) return false;
    if (w.cout_pad != a.cout) return false;
    if (a.head && a.cout != 32) return false;
    if (a.Wl < kPatchTileW || a.Hl < kPatchTileW) return false;
    if (a.out.f32) return false;
    const double tiles = (double)((a.Wl + kPatchTileW - 1) / kPatchTileW) * ((a.Hl + kPatchTileH - 1) / kPatchTileH);
    if ((double)a.Wl * a.Hl / (tiles * kTileM) < 0.6) return false;
    int mnx = 127, mxx = -127, mny = 127, mxy = -127;
    for (int t = 0; t < w.ntaps; ++t) {
        mnx = std::min(mnx, (int)w.dx[t]); mxx = std::max(mxx, (int)w.dx[t]);
        mny = std::min(mny, (int)w.dy[t]); mxy = std::max(mxy, (int)w.dy[t]);
    }
    g->ox = mnx; g->oy = mny;
    g->PW = kPatchTileW + (mxx - mnx); g->PH = kPatchTileH + (mxy - mny);
    g->BK = pick_bk(w.cin_pad);
    g->patch_bytes = g->PW * g->PH * g->BK * 2;
    g->patch_stride = (g->patch_bytes + 1023) / 1024 * 1024;
    g->wbytes = w.ntaps * w.cin_pad * a.cout * 2;
    g->stg_bytes = 2 * ((kTileM * a.cout * 2 + 1023) / 1024 * 1024);  // the kernel always carves two staging tiles
    if (g->PW > 256 || g->PH > 256) return false;
    const int need_stages = a.res ? 3 : 2;  // the epilogue holds the patch of a residual block a little longer
    if (g->wbytes + g->stg_bytes + need_stages * (w.cin_pad / g->BK) * g->patch_stride > kSmemBudget) return false;
    g->res_tap = -1;
    if (a.res) {
        // the patch kernel takes the residual from the input patch in shared memory: it must BE the block input
        if (a.res->base != a.in.base || a.res->c_off != a.in.c_off || a.res->Cs != a.in.Cs) return false;
        if (w.cin_pad != a.cout || w.cin_pad != g->BK) return false;
        for (int t = 0; t < w.ntaps; ++t)
            if (w.dx[t] == 0 && w.dy[t] == 0) g->res_tap = t;
        if (g->res_tap < 0) return false;
    }
    return true;
}

static int make_patch_op(w2l_ctx* ctx, Plan* pl, const ConvArgs& a, const PatchGeom& g) {
    Op op;
    op.type = OP_CONV;
    op.name = a.name + (a.w->fold ? " [fold+patch]" : " [patch]");
    op.patch = true;
    op.head = a.head;
    const PackedW& w = *a.w;
    const int BK = g.BK, BN = a.cout;
    op.BN = BN; op.BK = BK;
    PatchParams& h = op.pp;
    memset(&h, 0, sizeof(h));
    CKR(encode_act_map(ctx, &h.tmA, a.in, BK, g.PW, g.PH, 1, 1, 1, a.name.c_str()));
    CKR(encode_w_map(ctx, &h.tmB, w, BK, BN, a.name.c_str()));
    h.tiles_x = (a.Wl + kPatchTileW - 1) / kPatchTileW;
    h.tiles_y = (a.Hl + kPatchTileH - 1) / kPatchTileH;
    h.kc = w.cin_pad / BK;
    h.PW = g.PW; h.PH = g.PH; h.ox = g.ox; h.oy = g.oy;
    h.ntaps = w.ntaps;
    h.patch_bytes = g.patch_bytes; h.patch_stride = g.patch_stride;
    for (int t = 0; t < w.ntaps; ++t) h.tap_row[t] = (w.dy[t] - g.oy) * g.PW + (w.dx[t] - g.ox);
    h.stages = std::min(kPatchMaxStages, (kSmemBudget - g.wbytes - g.stg_bytes) / (h.kc * g.patch_stride));
    op.dyn_smem = g.wbytes + h.stages * h.kc * g.patch_stride + g.stg_bytes + kSmemExtra;
    if (op.dyn_smem > kSmemBudget + kSmemExtra || h.stages < 2) return fail(W2L_EINVAL, "%s: patch kernel smem plan %d B / %d stages", a.name.c_str(), op.dyn_smem, h.stages);
    fill_epi(&h.ep, a);
    h.res_row = g.res_tap >= 0 ? h.tap_row[g.res_tap] : -1;
    h.pair = h.stages >= 3 ? 1 : 0;  // two tiles in flight + at least one being prefetched
    if (!a.head) {
        // TMA-store view of the output: the BN-channel slice, with this launch's pixel strides (transposed-conv phases
        // interleave), box = one 8 x 16 tile; out-of-range pixels of ragged tiles are clipped by the TMA unit
        EncodeTiledFn enc = get_encode_fn();
        const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const CUtensorMapSwizzle sw = BN == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BN == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        cuuint64_t dims[4] = {(cuuint64_t)BN, (cuuint64_t)a.Wl, (cuuint64_t)a.Hl, (cuuint64_t)a.in.N};
        cuuint64_t strides[3] = {(cuuint64_t)h.ep.out_sx * 2, (cuuint64_t)h.ep.out_sy * 2, (cuuint64_t)h.ep.out_sn * 2};
        cuuint32_t box[4] = {(cuuint32_t)BN, (cuuint32_t)kPatchTileW, (cuuint32_t)kPatchTileH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (a.out.f32) return fail(W2L_EINVAL, "%s: patch kernel stores 16-bit outputs only", a.name.c_str());
        CUresult r = enc(&h.tmO, dt, 4, h.ep.out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(out) failed with %d", a.name.c_str(), (int)r);
    } else {
        h.tmO = h.tmA;  // never used by the head variant; keep the descriptor valid for the prefetch
    }
    // constant-bank copies of the folded BatchNorm and the head (plan-build time only)
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.cscale, a.scale + a.ch_off, (size_t)BN * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.cshift, a.shift + a.ch_off, (size_t)BN * 4, cudaMemcpyDeviceToHost));
    if (a.head) {
        CK(cudaMemcpy(h.chead_w, a.head_w, 96 * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h.chead_b, a.head_b, 3 * 4, cudaMemcpyDeviceToHost));
    }
    const long long total = (long long)h.tiles_x * h.tiles_y * a.in.N;
    op.grid = (int)std::min<long long>(total, ctx->num_sms);
    op.flops = 2.0 * a.macs_per_pixel * (double)a.Wl * a.Hl * a.in.N;
    pl->ops.push_back(op);
    return W2L_OK;
}

// The narrowest 96 x 96 layers: S output rows per GEMM row (conv_rowstack.cuh).  Returns the shape id or -1.
static int rowstack_eligible(const w2l_ctx* ctx, const ConvArgs& a, int* tap_of) {
    if (!ctx->use_rowstack || !ctx->use_patch || ctx->x2) return -1;
    const PackedW& w = *a.w;
    if (a.sx != 1 || a.sy != 1 || a.osx != 1 || a.osy != 1 || a.out.f32 || a.res) return -1;
    int shape = -1;
    if (a.head && a.cout == 32 && w.cin_pad == 80 && w.cout_pad == 32 && w.ntaps == 9 && !w.fold && !a.in.nwin && a.in.wstride == 1) shape = 0;
    if (!a.head && a.cout == 16 && w.cout_pad == 16 && w.fold && w.ntaps == 7 && w.cin_pad == 64 && a.in.wstride == 1) shape = 1;
    if (!a.head && a.cout == 32 && w.cout_pad == 32 && w.fold && w.ntaps == 7 && w.cin_pad == 64 && a.in.wstride == 1) shape = 2;
    if (shape < 0) return -1;
    const RsShape sh = rs_shape(shape);
    const int tile_h = sh.tile_h, R = sh.R, ndx = sh.ndx, ty = 2 * R + 1;
    if (a.Wl % kRsTileW != 0 || a.Hl % tile_h != 0) return -1;   // 96 x 96 here; ragged tiles would waste the pipe
    // no batch-size threshold: the kernel choice (and with it the fp32 summation order) must not depend on N, so that a
    // crop's result is bit-identical whatever batch it travels in (tests/test_gpu_nets.py)
    for (int i = 0; i < ndx * ty; ++i) tap_of[i] = -1;
    for (int t = 0; t < w.ntaps; ++t) {
        const int dx = w.dx[t], dy = w.dy[t];
        if (dy < -R || dy > R || (ndx == 1 ? dx != 0 : (dx < -1 || dx > 1))) return -1;
        tap_of[(ndx == 1 ? 0 : dx + 1) * ty + (R - dy)] = t;
    }
    for (int i = 0; i < ndx * ty; ++i) if (tap_of[i] < 0) return -1;
    return shape;
}

static int make_rowstack_op(w2l_ctx* ctx, Plan* pl, const ConvArgs& a, int shape, const int* tap_of) {
    Op op;
    op.type = OP_CONV;
    op.name = a.name + (shape == 0 ? " [rowstack x2]" : " [fold+rowstack x3]");
    const RsShape sh = rs_shape(shape);
    op.rowstack = true;
    op.rs_shape = shape;
    op.head = a.head;
    const PackedW& w = *a.w;
    const int C = a.cout;
    op.BN = C; op.BK = 64;
    RowStackParams& h = op.rs;
    memset(&h, 0, sizeof(h));
    const int PW = sh.PW, PH = sh.PH, tile_h = sh.tile_h;
    CKR(encode_act_map(ctx, &h.tmA0, a.in, 64, PW, PH, 1, 1, 1, a.name.c_str()));
    CKR(encode_w_map(ctx, &h.tmB0, w, 64, C, a.name.c_str()));
    if (shape == 0) {
        CKR(encode_act_map(ctx, &h.tmA1, a.in, 16, PW, PH, 1, 1, 1, a.name.c_str()));
        CKR(encode_w_map(ctx, &h.tmB1, w, 16, C, a.name.c_str()));
    } else {
        h.tmA1 = h.tmA0; h.tmB1 = h.tmB0;
    }
    h.tiles_x = a.Wl / kRsTileW;
    h.tiles_y = a.Hl / tile_h;
    h.ox = shape == 0 ? -1 : 0;   // folded inputs: the window already starts at the leftmost tap
    h.oy = -sh.R;
    for (int i = 0; i < sh.ndx * (2 * sh.R + 1); ++i) h.tap_of[i] = tap_of[i];
    const int fixed = sh.fixed, per_stage = sh.per_stage;
    h.stages = std::min(kRsMaxStages, (kSmemBudget + kSmemExtra - fixed) / per_stage);
    op.dyn_smem = fixed + h.stages * per_stage;
    if (h.stages < 2) return fail(W2L_EINVAL, "%s: row-stack kernel smem plan %d B / %d stages", a.name.c_str(), op.dyn_smem, h.stages);
    fill_epi(&h.ep, a);
    if (!a.head) {
        EncodeTiledFn enc = get_encode_fn();
        const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)a.Wl, (cuuint64_t)a.Hl, (cuuint64_t)a.in.N};
        cuuint64_t strides[3] = {(cuuint64_t)h.ep.out_sx * 2, (cuuint64_t)h.ep.out_sy * 2, (cuuint64_t)h.ep.out_sn * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)kRsTileW, (cuuint32_t)tile_h, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&h.tmO, dt, 4, h.ep.out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(out) failed with %d", a.name.c_str(), (int)r);
    } else {
        h.tmO = h.tmA0;
    }
    h.tmO2 = h.tmO;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.cscale, a.scale + a.ch_off, (size_t)C * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.cshift, a.shift + a.ch_off, (size_t)C * 4, cudaMemcpyDeviceToHost));
    if (a.head) {
        CK(cudaMemcpy(h.chead_w, a.head_w, 96 * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h.chead_b, a.head_b, 3 * 4, cudaMemcpyDeviceToHost));
    }
    const long long total = (long long)h.tiles_x * h.tiles_y * a.in.N;
    op.grid = (int)std::min<long long>(total, ctx->num_sms);
    op.flops = 2.0 * a.macs_per_pixel * (double)a.Wl * a.Hl * a.in.N;
    pl->ops.push_back(op);
    return W2L_OK;
}

static int make_conv_op(w2l_ctx* ctx, Plan* pl, const ConvArgs& a) {
    if (!get_encode_fn()) return fail(W2L_ENODEV, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    const PackedW& w = *a.w;
    if (a.in.C != w.cin_pad) return fail(W2L_EINVAL, "%s: input view has %d channels, weights packed for %d", a.name.c_str(), a.in.C, w.cin_pad);
    if (a.cout % 16 != 0) return fail(W2L_EINVAL, "%s: cout %d not a multiple of 16", a.name.c_str(), a.cout);
    int rs_taps[21];
    const int rs_shape = rowstack_eligible(ctx, a, rs_taps);
    if (rs_shape >= 0) return make_rowstack_op(ctx, pl, a, rs_shape, rs_taps);
    PatchGeom geom;
    if (patch_eligible(ctx, a, &geom)) return make_patch_op(ctx, pl, a, geom);
    Op op;
    op.type = OP_CONV;
    op.name = a.name + (w.fold ? " [fold]" : "");
    const int BK = pick_bk(w.cin_pad);
    int bw, bh, bn;
    pick_box(a.Wl, a.Hl, a.in.N, a.sx, a.sy, &bw, &bh, &bn);
    const int tiles_x = (a.Wl + bw - 1) / bw, tiles_y = (a.Hl + bh - 1) / bh, tiles_n = (a.in.N + bn - 1) / bn;
    const int m_tiles = tiles_x * tiles_y * tiles_n;
    int BN = 16;
    for (int cand : {128, 64, 32, 16})
        if (a.cout % cand == 0) { BN = cand; break; }
    // 256-wide tiles halve the A-operand traffic (L2 -> smem and smem -> tensor core) per FLOP; worth it once
    // there are enough tiles to fill the machine several times over
    if (ctx->use_bn256 && BK == 64 && a.cout % 256 == 0 && (long long)m_tiles * (a.cout / 256) >= 3LL * ctx->num_sms) BN = 256;
    if (a.head) BN = 32;
    else
        while (BN > 32 && m_tiles * (a.cout / BN) < ctx->num_sms && a.cout % (BN / 2) == 0) BN /= 2;
    if (w.cout_pad % BN != 0) return fail(W2L_EINVAL, "%s: cout_pad %d vs BN %d", a.name.c_str(), w.cout_pad, BN);
    op.BN = BN; op.BK = BK; op.head = a.head;
    // two M tiles per CTA (shared weight slab, two accumulators) once there is plenty of work
    const int n_tiles_ = a.cout / BN;
    if (ctx->use_mt2 && !a.head && find_conv_kernel(BN, BK, ctx->bf16, false, 2) &&
        (long long)((m_tiles + 1) / 2) * n_tiles_ >= 2LL * ctx->num_sms)
        op.MT = 2;
    if (op.MT == 2) op.name += " [2M]";

    ConvParams& p = op.cp;
    memset(&p, 0, sizeof(p));
    CKR(encode_act_map(ctx, &p.tmA, a.in, BK, bw * a.sx, bh * a.sy, bn, a.sx, a.sy, a.name.c_str()));
    CKR(encode_w_map(ctx, &p.tmB, w, BK, BN, a.name.c_str()));
    p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.tiles_n = tiles_n; p.n_tiles = a.cout / BN;
    p.bw = bw; p.bh = bh; p.bn = bn;
    p.sx = a.sx; p.sy = a.sy;
    p.ntaps = w.ntaps; p.kc_per_tap = w.cin_pad / BK;
    p.stage_tx_bytes = (unsigned)(op.MT * bw * bh * bn * BK * 2 + BN * BK * 2);
    fill_epi(&p.ep, a);
    // staged epilogue (TMA residual load + TMA store) for 16-bit outputs; the head / fp32 outputs keep direct stores
    p.tma_epi = 0;
    if (ctx->use_tma_epi && !ctx->x2 && !a.head && !a.out.f32) {
        EncodeTiledFn enc = get_encode_fn();
        const int EW = BN < 64 ? BN : 64;
        const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const CUtensorMapSwizzle esw = EW == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : EW == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        cuuint64_t dims[4] = {(cuuint64_t)a.cout, (cuuint64_t)a.Wl, (cuuint64_t)a.Hl, (cuuint64_t)a.in.N};
        cuuint32_t box[4] = {(cuuint32_t)EW, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        cuuint32_t es[4] = {1, 1, 1, 1};
        cuuint64_t os[3] = {(cuuint64_t)p.ep.out_sx * 2, (cuuint64_t)p.ep.out_sy * 2, (cuuint64_t)p.ep.out_sn * 2};
        CUresult r = enc(&p.tmO, dt, 4, p.ep.out, dims, os, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, esw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(out) failed with %d", a.name.c_str(), (int)r);
        if (a.res) {
            cuuint64_t rs[3] = {(cuuint64_t)p.ep.res_sx * 2, (cuuint64_t)p.ep.res_sy * 2, (cuuint64_t)p.ep.res_sn * 2};
            r = enc(&p.tmR, dt, 4, const_cast<void*>(p.ep.res), dims, rs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, esw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(res) failed with %d", a.name.c_str(), (int)r);
        } else {
            p.tmR = p.tmO;
        }
        p.tma_epi = 1;
        p.epi_box_bytes = (unsigned)(bw * bh * bn * EW * 2);
    }
    if (w.ntaps > kMaxTaps) return fail(W2L_EINVAL, "%s: too many taps", a.name.c_str());
    if (ctx->x2) {
        // split operands: x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (the dropped x_lo*w_lo term is ~2^-22 relative)
        if (a.in.lo_off <= 0 || w.nslabs != 2 * w.ntaps) return fail(W2L_ESTATE, "%s: split-operand mode needs hi/lo planes", a.name.c_str());
        int k = 0;
        for (int t = 0; t < w.ntaps; ++t)
            for (int v = 0; v < 3; ++v, ++k) {
                p.dx[k] = w.dx[t]; p.dy[k] = w.dy[t];
                p.a_lo[k] = (v == 1) ? 1 : 0;
                p.b_slab[k] = (unsigned char)((v == 2) ? w.ntaps + t : t);
            }
        p.ntaps = 3 * w.ntaps;
        p.a_lo_off = a.in.lo_off;
    } else {
        for (int t = 0; t < w.ntaps; ++t) { p.dx[t] = w.dx[t]; p.dy[t] = w.dy[t]; p.a_lo[t] = 0; p.b_slab[t] = (unsigned char)t; }
    }
    const int total = ((m_tiles + op.MT - 1) / op.MT) * p.n_tiles;
    op.grid = std::min(total, ctx->num_sms);
    op.flops = 2.0 * a.macs_per_pixel * (double)a.Wl * a.Hl * a.in.N;
    pl->ops.push_back(op);
    return W2L_OK;
}

// Conv2dTranspose k3 s2 p1 op1 with 64 output channels: all four phases in one launch (convt_fused.cuh)
static int make_convt_fused_op(w2l_ctx* ctx, Plan* pl, const Layer& L, const LayerW& lw, const Act& in, const Act& out) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(W2L_ENODEV, "cuTensorMapEncodeTiled is not available");
    const PackedW& w = lw.ph.back();  // all 9 taps, grouped by input shift (load_layer)
    Op op;
    op.type = OP_CONV;
    op.name = L.name + " [fused 4-phase]";
    op.ctf = true;
    const int BK = pick_bk(w.cin_pad);
    op.BK = BK; op.BN = kCtBN;
    ConvTParams& t = op.tp;
    memset(&t, 0, sizeof(t));
    CKR(encode_act_map(ctx, &t.tmA, in, BK, kCtPW, kCtPH, 1, 1, 1, L.name.c_str()));
    {
        const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        cuuint64_t dims[3] = {(cuuint64_t)w.cin_pad, (cuuint64_t)w.cout_pad, 9};
        cuuint64_t strides[2] = {(cuuint64_t)w.cin_pad * 2, (cuuint64_t)w.cin_pad * w.cout_pad * 2};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)kCtBN, 9};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&t.tmB, dt, 3, w.w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(B9) failed with %d", L.name.c_str(), (int)r);
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                uint16_t* base = out.base + ((long long)py * out.W + px) * out.Cs + out.c_off;
                cuuint64_t od[4] = {(cuuint64_t)kCtBN, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N};
                cuuint64_t os[3] = {(cuuint64_t)2 * out.Cs * 2, (cuuint64_t)2 * out.W * out.Cs * 2, (cuuint64_t)out.H * out.W * out.Cs * 2};
                cuuint32_t ob[4] = {(cuuint32_t)kCtBN, 8, 16, 1};
                cuuint32_t oe[4] = {1, 1, 1, 1};
                CUresult r2 = enc(&t.tmO[py * 2 + px], dt, 4, base, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r2 != CUDA_SUCCESS) return fail(W2L_ECUDA, "%s: cuTensorMapEncodeTiled(out phase) failed with %d", L.name.c_str(), (int)r2);
            }
    }
    t.tiles_x = (in.W + 7) / 8; t.tiles_y = (in.H + 15) / 16; t.N = in.N;
    t.kc = w.cin_pad / BK;
    t.patch_bytes = kCtPW * kCtPH * BK * 2;
    t.patch_stride = (t.patch_bytes + 1023) / 1024 * 1024;
    const int stage_bytes = t.patch_stride + 9 * kCtBN * BK * 2;
    const int fixed = 2 * kTileM * kCtBN * 2 + kSmemExtra;
    t.stages = std::min(kCtMaxStages, (kCtSmemMax - fixed) / stage_bytes);
    if (t.stages < 2) return fail(W2L_EINVAL, "%s: fused convT does not fit shared memory", L.name.c_str());
    op.dyn_smem = t.stages * stage_bytes + fixed;
    t.act = ACT_RELU;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(t.cscale, lw.scale, kCtBN * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t.cshift, lw.shift, kCtBN * 4, cudaMemcpyDeviceToHost));
    const long long units = (long long)t.tiles_x * t.tiles_y * in.N;
    op.grid = (int)std::min<long long>(units, ctx->num_sms);
    op.flops = 2.0 * (double)L.cin * L.cout * 9 * (double)in.W * in.H * in.N;
    pl->ops.push_back(op);
    return W2L_OK;
}

// Emit the launches of one block (conv / convT) of a spec table.
static int emit_block(w2l_ctx* ctx, Plan* pl, const NetW& nw, int li, const Layer& L, const Act& in, const Act& out,
                      const Act* res, bool head = false, int head_B = 1, int head_T = 1) {
    const LayerW& lw = nw.layers[li];
    ConvArgs a;
    a.in = in; a.out = out; a.res = res;
    a.scale = lw.scale; a.shift = lw.shift;
    a.act = (L.kind == W2L_BLOCK_CONV_LRELU) ? ACT_LRELU : (L.kind == W2L_BLOCK_CONV_PLAIN ? ACT_NONE : ACT_RELU);
    a.cout = L.cout;
    a.head = head; a.head_w = nw.head_w; a.head_b = nw.head_b; a.head_B = head_B; a.head_T = head_T;
    if (L.kind != W2L_BLOCK_CONVT_BN_RELU) {
        a.name = L.name;
        a.w = &lw.ph[0];
        a.sx = lw.ph[0].fold ? 1 : L.sw; a.sy = L.sh;  // folded first layers: the tensor map already strides the windows
        a.Hl = out.H; a.Wl = out.W;
        a.macs_per_pixel = (double)L.cin * L.cout * L.kh * L.kw;
        return make_conv_op(ctx, pl, a);
    }
    if (lw.gemm_convT) {
        // 1x1 -> kh x kw transposed conv == GEMM with kh*kw*cout output columns landing NHWC-contiguous
        Act o = out;
        o.H = 1; o.W = 1; o.Cs = out.Cs * out.H * out.W; o.C = L.cout * L.kh * L.kw;
        if (out.c_off != 0 || out.C != out.Cs) return fail(W2L_EINVAL, "%s: gemm convT needs a dense output", L.name.c_str());
        a.name = L.name;
        a.out = o;
        a.w = &lw.ph[0];
        a.Hl = 1; a.Wl = 1;
        a.cout = L.cout * L.kh * L.kw;
        a.macs_per_pixel = (double)L.cin * L.cout * L.kh * L.kw;
        return make_conv_op(ctx, pl, a);
    }
    if (lw.has_all_taps && ctx->use_ctfused && in.W >= 8 && in.H >= 8 &&
        (double)in.W * in.H / ((double)((in.W + 7) / 8) * ((in.H + 15) / 16) * kTileM) >= 0.6 && !out.f32)
        return make_convt_fused_op(ctx, pl, L, lw, in, out);
    const size_t nph = lw.ph.size() - (lw.has_all_taps ? 1 : 0);
    for (size_t i = 0; i < nph; ++i) {
        const PackedW& w = lw.ph[i];
        ConvArgs b = a;
        b.name = L.name + ".ph" + std::to_string(w.py) + std::to_string(w.px);
        b.w = &w;
        b.osy = L.sh; b.osx = L.sw; b.phy = w.py; b.phx = w.px;
        b.Hl = (out.H - w.py + L.sh - 1) / L.sh;
        b.Wl = (out.W - w.px + L.sw - 1) / L.sw;
        b.macs_per_pixel = (double)L.cin * L.cout * w.ntaps;
        CKR(make_conv_op(ctx, pl, b));
    }
    return W2L_OK;
}

static void conv_out_dims(const Layer& L, int H, int W, int* Ho, int* Wo) {
    if (L.kind == W2L_BLOCK_CONVT_BN_RELU) {
        *Ho = (H - 1) * L.sh - 2 * L.ph + L.kh + L.out_pad;
        *Wo = (W - 1) * L.sw - 2 * L.pw + L.kw + L.out_pad;
    } else {
        *Ho = (H + 2 * L.ph - L.kh) / L.sh + 1;
        *Wo = (W + 2 * L.pw - L.kw) / L.sw + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
struct TensorRef { const float* p; int64_t n; };
typedef std::map<std::string, TensorRef> TensorMap;

static int need(const TensorMap& tm, const std::string& name, int64_t numel, const float** out) {
    auto it = tm.find(name);
    if (it == tm.end()) return fail(W2L_EINVAL, "missing tensor '%s'", name.c_str());
    if (it->second.n != numel) return fail(W2L_EINVAL, "tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)it->second.n, (long long)numel);
    *out = it->second.p;
    return W2L_OK;
}

static int pack_taps(w2l_ctx* ctx, PackedW* pw, const float* src, int cout, int cin, int kh, int kw, bool transposed,
                     const std::vector<std::pair<int, int>>& rs, int cout_pad_to, cudaStream_t st, uint16_t* dst_override = nullptr) {
    PackParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.src = src;
    pp.ntaps = (int)rs.size();
    pp.cout = cout; pp.cin = cin;
    pp.cin_pad = round_up(cin, 16);
    pp.cout_pad = round_up(cout, cout_pad_to);
    if (transposed) { pp.s_ci = (long long)cout * kh * kw; pp.s_co = (long long)kh * kw; }
    else { pp.s_co = (long long)cin * kh * kw; pp.s_ci = (long long)kh * kw; }
    pp.s_r = kw; pp.s_s = 1;
    for (size_t t = 0; t < rs.size(); ++t) { pp.r[t] = (signed char)rs[t].first; pp.s[t] = (signed char)rs[t].second; }
    const size_t n = (size_t)pp.ntaps * pp.cout_pad * pp.cin_pad;
    const int planes = (ctx->x2 && !dst_override) ? 2 : 1;
    if (dst_override) pp.dst = dst_override;
    else {
        void* d = nullptr;
        CKR(dev_alloc(&d, n * 2 * planes));
        ctx->weight_bytes += n * 2 * planes;
        pp.dst = (uint16_t*)d;
        pw->w = pp.dst;
        pw->ntaps = pp.ntaps; pw->cout_pad = pp.cout_pad; pw->cin_pad = pp.cin_pad;
        pw->nslabs = pp.ntaps * planes;
    }
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
    for (int pl_ = 0; pl_ < planes; ++pl_) {  // hi slabs, then (split-operand mode) the lo slabs w - fp16(w)
        pp.lo = pl_;
        if (ctx->bf16) pack_w_kernel<true><<<blocks, 256, 0, st>>>(pp);
        else pack_w_kernel<false><<<blocks, 256, 0, st>>>(pp);
        ctx->launches++;
        pp.dst += n;
    }
    CK(cudaGetLastError());
    return W2L_OK;
}

static void free_layer(LayerW& lw) {
    for (auto& p : lw.ph) if (p.w) cudaFree(p.w);
    lw.ph.clear();
    if (lw.scale) cudaFree(lw.scale);
    if (lw.shift) cudaFree(lw.shift);
    lw.scale = lw.shift = nullptr;
    lw.loaded = false;
    lw.has_all_taps = false;
    lw.gemm_convT = false;
}

// Pack one block's parameters. in_hw1: the block is applied to a 1x1 input (enables the GEMM form of convT).
static int load_layer(w2l_ctx* ctx, LayerW* lw, const Layer& L, const float* W, const float* bias, const float* gamma,
                      const float* beta, const float* mean, const float* var, bool in_hw1, bool first_layer, cudaStream_t st) {
    free_layer(*lw);
    const int pad_to = 16;
    int reps = 1;
    if (first_layer && ctx->use_fold && L.kind != W2L_BLOCK_CONVT_BN_RELU && L.cin <= 16 && L.kw >= 3 && (L.sw == 1 || L.sw == 2)) {
        // tiny-Cin first layer: fold the kw horizontal taps into K (one K row per filter row r)
        PackedW pw;
        pw.fold = true;
        pw.Cp = L.cin <= 8 ? 8 : 16;
        const int raw = L.kw * pw.Cp;
        pw.kfold = raw <= 32 ? 32 : round_up(raw, 64);
        pw.win = pw.kfold / pw.Cp;
        pw.ntaps = L.kh; pw.cin_pad = pw.kfold; pw.cout_pad = round_up(L.cout, pad_to);
        for (int r = 0; r < L.kh; ++r) { pw.dy.push_back((signed char)(r - L.ph)); pw.dx.push_back(0); }
        const size_t n = (size_t)pw.ntaps * pw.cout_pad * pw.kfold;
        void* d = nullptr;
        CKR(dev_alloc(&d, n * 2));
        ctx->weight_bytes += n * 2;
        pw.w = (uint16_t*)d;
        PackFoldParams fp;
        fp.src = W; fp.dst = pw.w; fp.kh = L.kh; fp.kw = L.kw; fp.cout = L.cout; fp.cin = L.cin;
        fp.cout_pad = pw.cout_pad; fp.kfold = pw.kfold; fp.Cp = pw.Cp;
        const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
        if (ctx->bf16) pack_fold_kernel<true><<<blocks, 256, 0, st>>>(fp);
        else pack_fold_kernel<false><<<blocks, 256, 0, st>>>(fp);
        ctx->launches++;
        CK(cudaGetLastError());
        lw->ph.push_back(pw);
    } else if (L.kind != W2L_BLOCK_CONVT_BN_RELU) {
        std::vector<std::pair<int, int>> rs;
        PackedW pw;
        for (int r = 0; r < L.kh; ++r)
            for (int s = 0; s < L.kw; ++s) { rs.push_back({r, s}); pw.dy.push_back((signed char)(r - L.ph)); pw.dx.push_back((signed char)(s - L.pw)); }
        CKR(pack_taps(ctx, &pw, W, L.cout, L.cin, L.kh, L.kw, false, rs, pad_to, st));
        lw->ph.push_back(pw);
    } else if (in_hw1 && !ctx->x2 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0) {
        // out[n, y, x, co] = sum_ci in[n, ci] * W[ci, co, y, x]  -> GEMM with columns (y, x, co)
        lw->gemm_convT = true;
        reps = L.kh * L.kw;
        PackedW pw;
        pw.ntaps = 1; pw.cin_pad = round_up(L.cin, 16); pw.cout_pad = round_up(L.cout, pad_to) * reps;
        pw.dx.push_back(0); pw.dy.push_back(0);
        void* d = nullptr;
        const size_t n = (size_t)pw.cout_pad * pw.cin_pad;
        CKR(dev_alloc(&d, n * 2));
        ctx->weight_bytes += n * 2;
        pw.w = (uint16_t*)d;
        if (L.cout % pad_to != 0) return fail(W2L_EINVAL, "%s: gemm convT needs cout %% 16 == 0", L.name.c_str());
        for (int r = 0; r < L.kh; ++r)
            for (int s = 0; s < L.kw; ++s) {
                std::vector<std::pair<int, int>> rs = {{r, s}};
                CKR(pack_taps(ctx, nullptr, W, L.cout, L.cin, L.kh, L.kw, true, rs, pad_to, st,
                              pw.w + (size_t)(r * L.kw + s) * L.cout * pw.cin_pad));
            }
        lw->ph.push_back(pw);
    } else {
        // transposed conv: oy = iy*s - p + r.  Output phase py uses the taps r == (py + p) mod s at input row y + (py + p - r)/s
        for (int py = 0; py < L.sh; ++py)
            for (int px = 0; px < L.sw; ++px) {
                std::vector<std::pair<int, int>> rs;
                PackedW pw;
                pw.py = py; pw.px = px;
                for (int r = 0; r < L.kh; ++r) {
                    if ((py + L.ph - r) % L.sh != 0) continue;
                    for (int s = 0; s < L.kw; ++s) {
                        if ((px + L.pw - s) % L.sw != 0) continue;
                        rs.push_back({r, s});
                        pw.dy.push_back((signed char)((py + L.ph - r) / L.sh));
                        pw.dx.push_back((signed char)((px + L.pw - s) / L.sw));
                    }
                }
                if (rs.empty()) return fail(W2L_EINVAL, "%s: empty transposed-conv phase", L.name.c_str());
                CKR(pack_taps(ctx, &pw, W, L.cout, L.cin, L.kh, L.kw, true, rs, pad_to, st));
                lw->ph.push_back(pw);
            }
        if (!ctx->x2 && L.cout == kCtBN && L.kh == 3 && L.kw == 3 && L.sh == 2 && L.sw == 2 && L.ph == 1 && L.pw == 1 && L.out_pad == 1) {
            // all nine taps for the fused four-phase kernel, grouped by the input shift (dy,dx) they read and, inside a
            // group, in the accumulator's phase order [00 | 01 | 11 | 10] (convt_fused.cuh): tap (r,s) belongs to phase
            // ((r+1)&1, (s+1)&1) and reads in[y + (r==0), x + (s==0)]
            std::vector<std::pair<int, int>> rs = {{1, 1}, {1, 2}, {2, 2}, {2, 1},   // shift (0,0): phases 00 01 11 10
                                                   {1, 0}, {2, 0},                   // shift (0,1): phases 01 11
                                                   {0, 2}, {0, 1},                   // shift (1,0): phases 11 10
                                                   {0, 0}};                          // shift (1,1): phase 11
            PackedW pw;
            for (int t = 0; t < 9; ++t) { pw.dy.push_back(0); pw.dx.push_back(0); }
            CKR(pack_taps(ctx, &pw, W, L.cout, L.cin, L.kh, L.kw, true, rs, pad_to, st));
            lw->ph.push_back(pw);
            lw->has_all_taps = true;
        }
    }
    const int n_pad = round_up(L.cout, pad_to) * reps;
    void* sc = nullptr; void* sh = nullptr;
    CKR(dev_alloc(&sc, (size_t)n_pad * 4));
    CKR(dev_alloc(&sh, (size_t)n_pad * 4));
    lw->scale = (float*)sc; lw->shift = (float*)sh; lw->n_scale = n_pad;
    fold_bn_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(bias, gamma, beta, mean, var, 1e-5f, L.cout, reps, n_pad, lw->scale, lw->shift);
    ctx->launches++;
    CK(cudaGetLastError());
    lw->loaded = true;
    return W2L_OK;
}

static int fetch_block_tensors(const TensorMap& tm, const Layer& L, const float** W, const float** b, const float** g,
                               const float** be, const float** m, const float** v) {
    const int64_t wn = (int64_t)L.cin * L.cout * L.kh * L.kw;
    CKR(need(tm, L.name + ".conv_block.0.weight", wn, W));
    CKR(need(tm, L.name + ".conv_block.0.bias", L.cout, b));
    *g = *be = *m = *v = nullptr;
    if (L.kind == W2L_BLOCK_CONV_BN_RELU || L.kind == W2L_BLOCK_CONVT_BN_RELU) {
        CKR(need(tm, L.name + ".conv_block.1.weight", L.cout, g));
        CKR(need(tm, L.name + ".conv_block.1.bias", L.cout, be));
        CKR(need(tm, L.name + ".conv_block.1.running_mean", L.cout, m));
        CKR(need(tm, L.name + ".conv_block.1.running_var", L.cout, v));
    }
    return W2L_OK;
}

static void drop_plans(w2l_ctx* ctx, int net) {
    for (auto it = ctx->plans.begin(); it != ctx->plans.end();) {
        if (it->second->net == net) { free_plan(it->second.get()); it = ctx->plans.erase(it); }
        else ++it;
    }
    ctx->last_plan[net] = nullptr;
}

// ------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------
struct TmpPool {  // two ping-pong temporaries per chain, grown on demand
    Act slot[2];
    size_t cap[2] = {0, 0};
    int next = 0;
};

static int tmp_act(w2l_ctx* ctx, Plan* pl, TmpPool* tp, Act* a, int N, int H, int W, int C, const uint16_t* avoid) {
    int s = tp->next;
    if (tp->slot[s].base != nullptr && tp->slot[s].base == avoid) s ^= 1;
    const size_t need_b = (size_t)N * H * W * C * 2 * (pl->x2 ? 2 : 1);
    if (ctx->keep_all || tp->cap[s] < need_b) {
        CKR(plan_act(pl, &tp->slot[s], N, H, W, C));
        tp->cap[s] = need_b;
    }
    Act v = tp->slot[s];
    v.N = N; v.H = H; v.W = W; v.Cs = pl->x2 ? 2 * C : C; v.c_off = 0; v.C = C; v.lo_off = pl->x2 ? C : 0;
    *a = v;
    tp->next = s ^ 1;
    return W2L_OK;
}

static void add_ingest(Plan* pl, const char* name, int src_id, const Act& dst, int B, int C, long long sB, long long sC,
                       long long sT, int y_off, int Wsrc) {
    Op op;
    op.type = OP_INGEST;
    op.name = name;
    op.ingest_src = src_id;
    IngestParams& ip = op.ip;
    ip.src = nullptr; ip.dst = dst.base;
    ip.N = dst.N; ip.B = B; ip.C = C; ip.H = dst.H; ip.W = dst.W;
    ip.Cpad = dst.lo_off > 0 ? dst.lo_off : dst.Cs;  // logical (padded) channels; Cs is the pixel pitch
    ip.Cpix = dst.Cs;
    ip.Wp = dst.pitch(); ip.x_off = dst.x_off;
    ip.lo_off = dst.lo_off;
    ip.sB = sB; ip.sC = sC; ip.sT = sT; ip.y_off = y_off; ip.Wsrc = Wsrc;
    ip.cgrp = 0; ip.sG = 0;
    pl->ops.push_back(op);
}

// a straight chain of blocks (encoders): ping-pong temporaries, optional final destination
static int emit_chain(w2l_ctx* ctx, Plan* pl, int net, const std::vector<Layer>& layers, const std::vector<int>& idx,
                      Act x, TmpPool* tp, const Act* final_dst, Act* result) {
    for (size_t k = 0; k < idx.size(); ++k) {
        const Layer& L = layers[idx[k]];
        int Ho, Wo;
        conv_out_dims(L, x.H, x.W, &Ho, &Wo);
        Act out;
        if (k + 1 == idx.size() && final_dst) {
            out = *final_dst;
            if (out.H != Ho || out.W != Wo || out.C != L.cout) return fail(W2L_EINVAL, "%s: destination shape mismatch (%dx%dx%d vs %dx%dx%d)", L.name.c_str(), out.H, out.W, out.C, Ho, Wo, L.cout);
        } else {
            CKR(tmp_act(ctx, pl, tp, &out, x.N, Ho, Wo, L.cout, x.base));
        }
        CKR(emit_block(ctx, pl, ctx->nets[net], idx[k], L, x, out, L.residual ? &x : nullptr));
        pl->layer_out[idx[k]] = out;
        x = out;
    }
    if (result) *result = x;
    return W2L_OK;
}

static int build_generator_plan(w2l_ctx* ctx, Plan* pl) {
    const GeneratorSpec& g = gen_spec();
    const int N = pl->N, B = pl->B, T = pl->T;
    Act faceIn, melIn;
    const NetW& nw = ctx->nets[W2L_NET_GENERATOR];
    CKR(plan_input_act(pl, &faceIn, N, 96, 96, 6, nw.layers[g.face_enc[0][0]], g.layers[g.face_enc[0][0]]));
    CKR(plan_input_act(pl, &melIn, N, 80, 16, 1, nw.layers[g.audio_enc[0]], g.layers[g.audio_enc[0]]));
    if (T > 0) {
        add_ingest(pl, "ingest.mel", 0, melIn, B, 1, (long long)T * 1280, 1280, 1280, 0, 16);
        add_ingest(pl, "ingest.face", 1, faceIn, B, 6, (long long)6 * T * 9216, (long long)T * 9216, 9216, 0, 96);
    } else {
        add_ingest(pl, "ingest.mel", 0, melIn, N, 1, 1280, 1280, 0, 0, 16);
        add_ingest(pl, "ingest.face", 1, faceIn, N, 6, 6 * 9216, 9216, 0, 0, 96);
    }
    // skip-concat buffers D[k]: [decoder output | encoder feature] at resolution hw[k]   (wav2lip.py:108)
    const int hw[7] = {1, 3, 6, 12, 24, 48, 96};
    const int dec_c[7] = {512, 512, 512, 384, 256, 128, 64};
    const int skip_c[7] = {512, 512, 256, 128, 64, 32, 16};
    Act D[7];
    for (int k = 0; k < 7; ++k) CKR(plan_act(pl, &D[k], N, hw[k], hw[k], dec_c[k] + skip_c[k]));

    // audio encoder -> (N,1,1,512)
    Act AE;
    CKR(plan_act(pl, &AE, N, 1, 1, 512));
    TmpPool tpa;
    const size_t audio_first = 0;  // ingest.mel is op 0; ingest.face (op 1) stays on the main lane
    CKR(emit_chain(ctx, pl, W2L_NET_GENERATOR, g.layers, g.audio_enc, melIn, &tpa, &AE, nullptr));
    // the audio encoder (small, latency-bound launches) runs on a side stream while the face encoder runs on the main one
    pl->ops[audio_first].lane = 1;
    for (size_t i = 2; i < pl->ops.size(); ++i) pl->ops[i].lane = 1;
    pl->has_side = true;

    // face encoder: stage i ends in the skip half of D[6-i] and the next stage reads it from there
    TmpPool tpe;
    Act x = faceIn;
    for (int i = 0; i < 7; ++i) {
        Act dst = D[6 - i].slice(dec_c[6 - i], skip_c[6 - i]);
        CKR(emit_chain(ctx, pl, W2L_NET_GENERATOR, g.layers, g.face_enc[i], x, &tpe, &dst, &x));
        if (i == 0 && nw.layers[g.face_enc[1][0]].ph[0].fold) {
            // The 16->32 stride-2 block gathers every other pixel of a 16-channel slice of D[6]: 32-byte TMA rows, the
            // slowest layer per FLOP. Give it a dense zero-bordered copy of the first block's output instead (second
            // TMA store of the same staged tile), read through the overlapping-window map with the 3 horizontal taps
            // folded into K.
            Op& prev = pl->ops.back();
            if ((!prev.patch && !prev.rowstack) || prev.head) return fail(W2L_ESTATE, "folded stride-2 block needs the patch kernel on the first block");
            const Layer& L1 = g.layers[g.face_enc[1][0]];
            Act e0;
            CKR(plan_input_act(pl, &e0, N, 96, 96, L1.cin, nw.layers[g.face_enc[1][0]], L1));
            EncodeTiledFn enc = get_encode_fn();
            const CUtensorMapDataType dt = ctx->bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
            cuuint64_t od[4] = {16, 96, 96, (cuuint64_t)N};
            cuuint64_t os[3] = {(cuuint64_t)e0.Cs * 2, (cuuint64_t)e0.Wp * e0.Cs * 2, (cuuint64_t)96 * e0.Wp * e0.Cs * 2};
            cuuint32_t ob[4] = {16, (cuuint32_t)kPatchTileW, (cuuint32_t)(prev.rowstack ? RsCfg1::kTileH : kPatchTileH), 1};
            cuuint32_t oe[4] = {1, 1, 1, 1};
            CUresult r = enc(prev.rowstack ? &prev.rs.tmO2 : &prev.pp.tmO2, dt, 4, e0.base + (size_t)e0.x_off * e0.Cs, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(W2L_ECUDA, "cuTensorMapEncodeTiled(dense copy) failed with %d", (int)r);
            prev.pp.has_out2 = 1;
            prev.rs.has_out2 = 1;
            x = e0;
        }
    }
    // decoder
    TmpPool tpd;
    x = AE;
    const size_t dec_first = pl->ops.size();
    for (int k = 0; k < 7; ++k) {
        Act dst = D[k].slice(0, dec_c[k]);
        CKR(emit_chain(ctx, pl, W2L_NET_GENERATOR, g.layers, g.face_dec[k], x, &tpd, &dst, nullptr));
        x = D[k];
    }
    // output block with the fused 1x1 + sigmoid head; writes the caller's fp32 tensor
    const Layer& L = g.layers[g.output_block0];
    Act none;
    none.N = N; none.H = 96; none.W = 96; none.Cs = 32; none.C = 32;
    CKR(emit_block(ctx, pl, ctx->nets[W2L_NET_GENERATOR], g.output_block0, L, x, none, nullptr, true, T > 0 ? B : N, T > 0 ? T : 1));
    pl->ops[dec_first].join_side = true;  // the decoder's first block consumes the audio embedding
    return W2L_OK;
}

static int build_syncnet_plan(w2l_ctx* ctx, Plan* pl) {
    const SyncnetSpec& s = sync_spec();
    const int N = pl->N;
    Act faceIn, melIn, fe, ae;
    const NetW& nw = ctx->nets[W2L_NET_SYNCNET];
    CKR(plan_input_act(pl, &faceIn, N, 48, 96, 15, nw.layers[s.face_enc[0]], s.layers[s.face_enc[0]]));
    CKR(plan_input_act(pl, &melIn, N, 80, 16, 1, nw.layers[s.audio_enc[0]], s.layers[s.audio_enc[0]]));
    CKR(plan_act(pl, &fe, N, 1, 1, 512, true));
    CKR(plan_act(pl, &ae, N, 1, 1, 512, true));
    add_ingest(pl, "ingest.mel", 0, melIn, N, 1, 1280, 1280, 0, 0, 16);
    if (pl->T > 0) {
        // face input = generated / ground-truth frames (B,3,T,96,96): lower half, the T frames stacked on channels
        // (c' = 3 t + c) — wav2lip_train.py:193-194 as addressing
        const int T = pl->T;
        add_ingest(pl, "ingest.frames", 1, faceIn, N, 3 * T, (long long)3 * T * 9216, (long long)T * 9216, 0, 48, 96);
        pl->ops.back().ip.cgrp = 3; pl->ops.back().ip.sG = 9216;
    } else {
        add_ingest(pl, "ingest.face", 1, faceIn, N, 15, 15 * 4608, 4608, 0, 0, 96);
    }
    TmpPool tpf, tpa;
    // the two encoders are independent until the embeddings: the audio one (short launches, issued first) runs on the
    // side stream while the face encoder runs on the main one
    CKR(emit_chain(ctx, pl, W2L_NET_SYNCNET, s.layers, s.audio_enc, melIn, &tpa, &ae, nullptr));
    pl->ops[0].lane = 1;  // ingest.mel
    for (size_t i = 2; i < pl->ops.size(); ++i) pl->ops[i].lane = 1;
    pl->has_side = true;
    CKR(emit_chain(ctx, pl, W2L_NET_SYNCNET, s.layers, s.face_enc, faceIn, &tpf, &fe, nullptr));
    const size_t join_at = pl->ops.size();
    for (int which = 0; which < 2; ++which) {
        Op op;
        op.type = OP_L2NORM;
        op.name = which == 0 ? "l2norm.audio" : "l2norm.face";
        op.aux_in = which == 0 ? ae.base : fe.base;
        op.aux_rows = N; op.aux_dim = 512; op.aux_out = which;
        pl->ops.push_back(op);
    }
    pl->ops[join_at].join_side = true;
    return W2L_OK;
}

static int build_disc_plan(w2l_ctx* ctx, Plan* pl) {
    const DiscSpec& d = disc_spec();
    const int N = pl->N, B = pl->B, T = pl->T;
    Act in, feat;
    CKR(plan_input_act(pl, &in, N, 48, 96, 3, ctx->nets[W2L_NET_DISC].layers[0], d.layers[0]));
    CKR(plan_act(pl, &feat, N, 1, 1, 512));
    // (B,3,T,96,96): t-major flatten + rows 48..95   (wav2lip.py:155-161)
    add_ingest(pl, "ingest.frames", 0, in, B, 3, (long long)3 * T * 9216, (long long)T * 9216, 9216, 48, 96);
    std::vector<int> idx;
    for (size_t i = 0; i < d.layers.size(); ++i) idx.push_back((int)i);
    TmpPool tp;
    CKR(emit_chain(ctx, pl, W2L_NET_DISC, d.layers, idx, in, &tp, &feat, nullptr));
    Op op;
    op.type = OP_DISC_HEAD;
    op.name = "binary_pred";
    op.aux_in = feat.base; op.aux_rows = N; op.aux_dim = 512; op.aux_out = 0;
    op.aux_pitch = feat.Cs; op.aux_lo = feat.lo_off;
    pl->ops.push_back(op);
    return W2L_OK;
}

static int get_plan(w2l_ctx* ctx, int net, int B, int T, Plan** out) {
    char key[64];
    snprintf(key, sizeof(key), "%d:%d:%d:%d", net, B, T, (int)ctx->keep_all);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { it->second->last_used = ++ctx->plan_clock; *out = it->second.get(); return W2L_OK; }
    if (!ctx->nets[net].loaded) return fail(W2L_ESTATE, "weights of net %d not loaded", net);
    // keep at most a few plans per net alive (activation arenas are large): evict the least recently used
    for (;;) {
        int count = 0;
        auto lru = ctx->plans.end();
        for (auto p = ctx->plans.begin(); p != ctx->plans.end(); ++p)
            if (p->second->net == net) {
                ++count;
                if (lru == ctx->plans.end() || p->second->last_used < lru->second->last_used) lru = p;
            }
        if (count < 6) break;
        CK(cudaDeviceSynchronize());  // the plan's buffers may still be in use by queued launches
        if (ctx->last_plan[net] == lru->second.get()) ctx->last_plan[net] = nullptr;
        free_plan(lru->second.get());
        ctx->plans.erase(lru);
    }
    std::unique_ptr<Plan> pl(new Plan());
    pl->net = net; pl->B = B; pl->T = T;
    pl->x2 = ctx->x2;
    pl->N = (net == W2L_NET_SYNCNET) ? B : (T > 0 ? B * T : B);
    int r = W2L_OK;
    if (net == W2L_NET_GENERATOR) r = build_generator_plan(ctx, pl.get());
    else if (net == W2L_NET_SYNCNET) r = build_syncnet_plan(ctx, pl.get());
    else r = build_disc_plan(ctx, pl.get());
    if (r != W2L_OK) { free_plan(pl.get()); return r; }
    pl->last_used = ++ctx->plan_clock;
    *out = pl.get();
    ctx->plans[key] = std::move(pl);
    return W2L_OK;
}

static int run_plan(w2l_ctx* ctx, Plan* pl, const void* in0, const void* in1, void* out0, void* out1, cudaStream_t st,
                    bool u8 = false) {
    const bool side = pl->has_side && ctx->use_side;
    cudaStream_t main_st = st;
    if (side) {
        CK(cudaEventRecord(ctx->ev_fork, main_st));
        CK(cudaStreamWaitEvent(ctx->s_side, ctx->ev_fork, 0));
    }
    for (Op& op : pl->ops) {
        if (side && op.join_side) {
            CK(cudaEventRecord(ctx->ev_join, ctx->s_side));
            CK(cudaStreamWaitEvent(main_st, ctx->ev_join, 0));
        }
        st = (side && op.lane == 1) ? ctx->s_side : main_st;
        switch (op.type) {
            case OP_INGEST: {
                if (u8 && op.ingest_src == 1) {  // uint8 crops: mask + concat + /255 fused into the ingest
                    IngestU8Params up;
                    up.src = (const unsigned char*)in1; up.dst = op.ip.dst;
                    up.N = op.ip.N; up.H = op.ip.H; up.W = op.ip.W; up.Cpad = op.ip.Cpad; up.Wp = op.ip.Wp; up.x_off = op.ip.x_off; up.lo_off = op.ip.lo_off; up.Cpix = op.ip.Cpix;
                    const long long tot = (long long)up.N * up.H * up.W;
                    const int blk = (int)std::min<long long>((tot + 255) / 256, ctx->num_sms * 16);
                    if (ctx->bf16) ingest_u8_kernel<true><<<blk, 256, 0, st>>>(up);
                    else ingest_u8_kernel<false><<<blk, 256, 0, st>>>(up);
                    ctx->launches++;
                    break;
                }
                IngestParams ip = op.ip;
                ip.src = (const float*)(op.ingest_src == 0 ? in0 : in1);
                const long long total = (long long)ip.N * ip.H * ip.W;
                const bool vec4 = ip.lo_off == 0 && ((ip.W | ip.Wsrc) & 3) == 0 && ((ip.sB | ip.sC | ip.sT | ip.sG) & 3) == 0 &&
                                  (((uintptr_t)ip.src) & 15) == 0;
                if (vec4) {
                    const int blocks = (int)std::min<long long>((total / 4 + 255) / 256, ctx->num_sms * 16);
                    if (ctx->bf16) ingest4_kernel<true><<<blocks, 256, 0, st>>>(ip);
                    else ingest4_kernel<false><<<blocks, 256, 0, st>>>(ip);
                } else {
                    const int blocks = (int)std::min<long long>((total + 255) / 256, ctx->num_sms * 16);
                    if (ctx->bf16) ingest_kernel<true><<<blocks, 256, 0, st>>>(ip);
                    else ingest_kernel<false><<<blocks, 256, 0, st>>>(ip);
                }
                ctx->launches++;
                break;
            }
            case OP_CONV: {
                if (op.head) {
                    op.cp.ep.head_out = u8 ? nullptr : (float*)out0; op.pp.ep.head_out = op.cp.ep.head_out; op.rs.ep.head_out = op.cp.ep.head_out;
                    op.cp.ep.head_out_u8 = u8 ? (unsigned char*)out0 : nullptr; op.pp.ep.head_out_u8 = op.cp.ep.head_out_u8; op.rs.ep.head_out_u8 = op.cp.ep.head_out_u8;
                }
                CKR(launch_conv(ctx, op, st));
                break;
            }
            case OP_L2NORM: {
                float* o = (float*)(op.aux_out == 0 ? out0 : out1);
                l2norm_kernel<<<(op.aux_rows + 3) / 4, 128, 0, st>>>((const float*)op.aux_in, o, op.aux_rows, op.aux_dim);
                ctx->launches++;
                break;
            }
            case OP_DISC_HEAD: {
                const NetW& nw = ctx->nets[W2L_NET_DISC];
                if (ctx->bf16) disc_head_kernel<true><<<(op.aux_rows + 3) / 4, 128, 0, st>>>((const uint16_t*)op.aux_in, nw.head_w, nw.head_b, (float*)out0, op.aux_rows, op.aux_dim, op.aux_pitch, op.aux_lo);
                else disc_head_kernel<false><<<(op.aux_rows + 3) / 4, 128, 0, st>>>((const uint16_t*)op.aux_in, nw.head_w, nw.head_b, (float*)out0, op.aux_rows, op.aux_dim, op.aux_pitch, op.aux_lo);
                ctx->launches++;
                break;
            }
        }
    }
    CK(cudaGetLastError());
    ctx->last_plan[pl->net] = pl;
    return W2L_OK;
}

// ------------------------------------------------------------------------------------------------
// mel tables (host, double precision) — librosa 0.7.0 filters.mel(16000, 800, 80, 55, 7600), Slaney
// ------------------------------------------------------------------------------------------------
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
static void build_mel_basis(std::vector<float>* dense) {
    const int nm = MEL_BANDS, nb = MEL_BINS;
    dense->assign((size_t)nm * nb, 0.0f);
    std::vector<double> mel_f(nm + 2);
    const double m0 = hz_to_mel(55.0), m1 = hz_to_mel(7600.0);
    const double step = (m1 - m0) / (nm + 1);
    for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(i == nm + 1 ? m1 : m0 + i * step);
    for (int i = 0; i < nm; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < nb; ++k) {
            const double f = 8000.0 * k / (nb - 1);
            const double lower = -(mel_f[i] - f) / fd0, upper = (mel_f[i + 2] - f) / fd1;
            const float w32 = (float)std::max(0.0, std::min(lower, upper));
            (*dense)[(size_t)i * nb + k] = (float)((double)w32 * enorm);
        }
    }
}

static int init_mel_tables(w2l_ctx* ctx) {
    std::vector<double2> tw(MEL_TW_TOTAL);
    const long double kTwoPi = 2.0L * 3.141592653589793238462643383279502884L;
    for (int m = 0; m <= 400; ++m) {  // post-pass / window table: exp(-2 pi i m / 800)
        const long double a = -kTwoPi * m / MEL_NFFT;
        tw[MEL_TW_POST + m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    auto fill_pass = [&](int base, int R, int Ns) {  // T[r-1][k] = exp(-2 pi i r k / (Ns R))
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                const long double a = -kTwoPi * (long double)(r * k) / (long double)(Ns * R);
                tw[base + (r - 1) * Ns + k] = make_double2((double)cosl(a), (double)sinl(a));
            }
    };
    fill_pass(MEL_TW_P2, 5, 5);
    fill_pass(MEL_TW_P3, 4, 25);
    fill_pass(MEL_TW_P4, 4, 100);
    std::vector<float> dense;
    build_mel_basis(&dense);
    std::vector<float> vals;
    std::vector<int> off(MEL_BANDS), start(MEL_BANDS), len(MEL_BANDS);
    for (int i = 0; i < MEL_BANDS; ++i) {
        int a = -1, b = -1;
        for (int k = 0; k < MEL_BINS; ++k)
            if (dense[(size_t)i * MEL_BINS + k] != 0.0f) { if (a < 0) a = k; b = k; }
        off[i] = (int)vals.size();
        start[i] = a < 0 ? 0 : a;
        len[i] = a < 0 ? 0 : b - a + 1;
        for (int k = 0; k < len[i]; ++k) vals.push_back(dense[(size_t)i * MEL_BINS + start[i] + k]);
    }
    void* p;
    CKR(dev_alloc(&p, tw.size() * sizeof(double2))); ctx->mel_tw = (double2*)p;
    CKR(dev_alloc(&p, vals.size() * 4)); ctx->mel_bvals = (float*)p;
    CKR(dev_alloc(&p, MEL_BANDS * 4)); ctx->mel_boff = (int*)p;
    CKR(dev_alloc(&p, MEL_BANDS * 4)); ctx->mel_bstart = (int*)p;
    CKR(dev_alloc(&p, MEL_BANDS * 4)); ctx->mel_blen = (int*)p;
    CK(cudaMemcpy(ctx->mel_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mel_bvals, vals.data(), vals.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mel_boff, off.data(), MEL_BANDS * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mel_bstart, start.data(), MEL_BANDS * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mel_blen, len.data(), MEL_BANDS * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMelSmemBytes));
    return W2L_OK;
}

static int ensure_stage(w2l_ctx* ctx, int i, size_t bytes) {
    if (ctx->stage_bytes[i] >= bytes) return W2L_OK;
    if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0;
    CKR(dev_alloc(&ctx->stage[i], bytes));
    ctx->stage_bytes[i] = bytes;
    return W2L_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

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
    if (e != cudaSuccess) { delete ctx; return fail(W2L_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    int r = init_mel_tables(ctx);
    if (r != W2L_OK) { delete ctx; return r; }
    {
        const char* e1 = getenv("W2L_DISABLE_HALO");
        const char* e2 = getenv("W2L_DISABLE_FOLD");
        ctx->use_patch = !(e1 && e1[0] == '1');
        ctx->use_fold = !(e2 && e2[0] == '1');
        const char* e3 = getenv("W2L_DISABLE_BN256");
        ctx->use_bn256 = !(e3 && e3[0] == '1');
        const char* e7 = getenv("W2L_DISABLE_FOLDS2");
        ctx->use_fold_s2 = !(e7 && e7[0] == '1');
        const char* e6 = getenv("W2L_DISABLE_TMAEPI");
        ctx->use_tma_epi = !(e6 && e6[0] == '1');
        const char* e5 = getenv("W2L_DISABLE_MT2");
        ctx->use_mt2 = !(e5 && e5[0] == '1');
        const char* e9 = getenv("W2L_DISABLE_ROWSTACK");
        ctx->use_rowstack = !(e9 && e9[0] == '1');
        const char* e11 = getenv("W2L_DISABLE_PDL");
        ctx->use_pdl = !(e11 && e11[0] == '1');
        const char* e8 = getenv("W2L_DISABLE_SIDESTREAM");
        ctx->use_side = !(e8 && e8[0] == '1');
        const char* e4 = getenv("W2L_DISABLE_CTFUSED");
        ctx->use_ctfused = !(e4 && e4[0] == '1');
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
    for (auto& kv : ctx->plans) free_plan(kv.second.get());
    for (int n = 0; n < 3; ++n) {
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
        bool first = L.name == "face_encoder_blocks.0.0" || L.name == "audio_encoder.0" || L.name == "face_encoder.0";
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
    if (!ctx || net < 0 || net > 2) return fail(W2L_EINVAL, "bad argument");
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
    if (n_samples <= MEL_NFFT / 2) return fail(W2L_EINVAL, "need more than %d samples for reflect padding (got %lld)", MEL_NFFT / 2, (long long)n_samples);
    DeviceGuard g(ctx->device);
    MelParams p;
    p.wav = wav; p.L = n_samples; p.mel = mel; p.F = w2l_mel_num_frames(n_samples);
    p.tw = ctx->mel_tw; p.bvals = ctx->mel_bvals; p.boff = ctx->mel_boff; p.bstart = ctx->mel_bstart; p.blen = ctx->mel_blen;
    const long long blocks = (p.F + MEL_FPB - 1) / MEL_FPB;
    mel_kernel<<<(unsigned)blocks, MEL_FPB * MEL_TPF, kMelSmemBytes, (cudaStream_t)stream>>>(p);
    ctx->launches++;
    CK(cudaGetLastError());
    return W2L_OK;
}

int w2l_melspectrogram_host(w2l_ctx* ctx, const float* wav_h, int64_t n_samples, float* mel_h) {
    if (!ctx || !wav_h || !mel_h) return fail(W2L_EINVAL, "null argument");
    if (n_samples <= MEL_NFFT / 2) return fail(W2L_EINVAL, "need more than %d samples for reflect padding (got %lld)", MEL_NFFT / 2, (long long)n_samples);
    DeviceGuard g(ctx->device);
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

int64_t w2l_launch_count(const w2l_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t w2l_device_bytes(const w2l_ctx* ctx) {
    if (!ctx) return 0;
    size_t b = ctx->weight_bytes;
    for (auto& kv : ctx->plans) b += kv.second->bytes;
    for (int i = 0; i < 6; ++i) b += ctx->stage_bytes[i];
    return (int64_t)b;
}

int w2l_profile_plan(w2l_ctx* ctx, int net, int iters, int cap, float* ms_out, double* flop_out, char (*names_out)[64], void* stream) {
    if (!ctx || net < 0 || net > 2 || iters <= 0) return fail(W2L_EINVAL, "bad argument");
    Plan* pl = ctx->last_plan[net];
    if (!pl) return fail(W2L_ESTATE, "no forward has run for net %d", net);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int k = 0;
    for (Op& op : pl->ops) {
        if (op.type != OP_CONV || k >= cap) continue;
        if (op.head && op.cp.ep.head_out == nullptr && op.pp.ep.head_out == nullptr && op.cp.ep.head_out_u8 == nullptr) continue;
        CKR(launch_conv(ctx, op, st, false));  // warm
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) CKR(launch_conv(ctx, op, st, false));
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_out) ms_out[k] = ms / iters;
        if (flop_out) flop_out[k] = op.flops;
        if (names_out) snprintf(names_out[k], 64, "%s", op.name.c_str());
        ++k;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return k;
}

}  // extern "C"
