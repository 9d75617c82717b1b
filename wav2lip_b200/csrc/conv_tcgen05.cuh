// conv_tcgen05.cuh — the generic fused conv-block kernel of the Wav2Lip hot path for sm_100a, plus the PTX
// wrappers and the epilogue helpers shared by the specialised kernels (conv_patch.cuh, convt_fused.cuh).
//
// One persistent, warp-specialised kernel computes, for every block kind of
// /root/reference/models/conv.py (Conv2d :5-19, nonorm_Conv2d :21-31 and, phase by phase,
// Conv2dTranspose :33-44):
//
//     y = act( scale[c] * conv(x, w)[., c] + shift[c]  (+ residual) )
//
// as an implicit GEMM on the 5th-generation tensor cores:
//     M = output pixels (a 128-row tile = a bw x bh x bn box of the NHWC output grid),
//     N = output channels (BN per tile), K = filter taps x input channels (BK per step).
//
//   warp 0  : TMA producer. For each K step one 4-D tiled TMA load per M tile brings the input box of the
//             current filter tap (start coordinate = tile origin * stride + tap offset; out-of-bounds
//             coordinates are zero-filled by the TMA unit, which IS the conv zero padding) and one
//             3-D TMA load brings the [BN x BK] weight slice of that tap. Both land K-major with the
//             hardware 128/64/32-byte swizzle that the UMMA shared-memory descriptors expect.
//   warp 1  : MMA issuer. One lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//             accumulating in TMEM (fp32); tcgen05.commit releases smem stages / publishes the unit.
//             With MT=2 a CTA works on two M tiles that share every weight slab and alternates the MMAs
//             between their accumulators (less operand traffic per FLOP, dependent-issue latency hidden).
//   warp 2  : TMEM allocator (2 accumulator stages x MT x BN columns, so the epilogue of unit i overlaps
//             the main loop of unit i+1).
//   warps 4+: one epilogue group of 4 warps per M tile. tcgen05.ld the accumulator (lane = GEMM row = output
//             pixel), apply the folded BatchNorm scale/shift (conv bias folded in), the residual add
//             (conv.py:16-18: after BN, before ReLU), ReLU / LeakyReLU(0.01) and write NHWC 16-bit straight
//             into the channel slice of the consumer's buffer (so torch.cat of wav2lip.py:108 never exists).
//             Staged mode (tma_epi): the residual tile arrives by TMA into a swizzled shared-memory tile,
//             is combined in place and leaves by ONE TMA tensor store per 64 channels (which also clips
//             ragged tiles); direct mode (fp32 outputs, fused generator head wav2lip.py:84-85): per-thread
//             global accesses.
//
// Everything a launch needs is in ConvParams (a __grid_constant__), built once per plan on the host.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace w2l {

constexpr int kTileM = 128;
constexpr int kMaxTaps = 49;              // filter taps of one launch
constexpr int kMaxSteps = 3 * kMaxTaps;   // K-loop entries: x3 in the split-operand (fp32-faithful) precision mode
constexpr int kConvThreads = 256;
constexpr int kSmemBudget = 224 * 1024;  // patch kernels: weights + patch ring + staging; barriers live in the extra KB
constexpr int kSmemExtra = 2048;         // 1024 alignment slack + barriers
constexpr int kSmemMax = 227 * 1024;     // dynamic shared memory limit per CTA on sm_100

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

// What the epilogue needs, shared by every conv kernel variant.
struct EpiParams {
    int Wout, Hout, N;      // logical output grid of this launch (rows outside are masked)
    int act;
    int out_f32;
    void* out;              // element (n,y,x,c) at out + n*out_sn + y*out_sy + x*out_sx + c  (elements)
    long long out_sn, out_sy, out_sx;
    const void* res;        // nullptr = no residual; same addressing
    long long res_sn, res_sy, res_sx;
    const float* scale;     // [n_tiles*BN]
    const float* shift;
    // fused generator head (kHead kernels only): out = sigmoid(W[3x32] relu(y) + b), fp32 NCHW/5-D
    const float* head_w;
    const float* head_b;
    // split-operand (fp32-faithful) mode: every value is stored as hi + lo (two fp16 planes, lo_off channels apart)
    int x2;
    int out_lo_off, res_lo_off;
    float* head_out;
    unsigned char* head_out_u8;  // if set: (N,H,W,3) uint8 = trunc(sigmoid * 255.f), inference.py:265,269
    int head_B, head_T;     // n = t*head_B + b ; T=1,B=N for the 4-D call
};

struct alignas(64) ConvParams {
    CUtensorMap tmA;  // activations, dims (C, W, H, N), box (BK, bw*sx, bh*sy, bn), elem strides (1,sx,sy,1)
    CUtensorMap tmB;  // weights, dims (Cin_pad, Cout_pad, taps), box (BK, BN, 1)
    CUtensorMap tmO;  // output slice (Cout, Wl, Hl, N) with the launch's pixel strides, box (EW, bw, bh, bn)   [tma_epi]
    CUtensorMap tmR;  // residual, same geometry                                                              [tma_epi]
    int tma_epi;             // 1: epilogue stages through swizzled smem and uses TMA for the residual and the output
    unsigned epi_box_bytes;  // bytes of one epilogue box (bw*bh*bn rows x EW channels)
    // M tiling of the logical output grid
    int tiles_x, tiles_y, tiles_n, n_tiles;
    int bw, bh, bn;
    int sx, sy;
    // K loop
    int ntaps, kc_per_tap;
    unsigned stage_tx_bytes;  // bytes both TMA loads of one stage deliver
    EpiParams ep;
    // K-loop entries ("steps"): one per filter tap, or three per tap in the split-operand mode
    // (x_hi*w_hi, x_lo*w_hi, x_hi*w_lo).  a_lo selects the lo plane of the activations, b_slab the weight slab.
    int a_lo_off;                    // channel offset of the lo plane inside a pixel (0 in the 16-bit modes)
    signed char dx[kMaxSteps];
    signed char dy[kMaxSteps];
    unsigned char a_lo[kMaxSteps];
    unsigned char b_slab[kMaxSteps];
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W2L_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra W2L_DONE_%=;\n\t"
        "bra W2L_WAIT_%=;\n\t"
        "W2L_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Programmatic dependent launch: the next kernel in the stream may start its prologue (barrier init, TMEM
// allocation, descriptor prefetch) while this one drains; it touches global memory only after pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory: rows of BK*2 bytes, 8-row groups, hardware swizzle = row bytes.
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
    constexpr uint64_t kLayout = (BK == 64) ? 2 : (BK == 32) ? 4 : 6;  // SWIZZLE_128B / 64B / 32B
    constexpr uint64_t kSBO = (8 * BK * 2) >> 4;                       // bytes between 8-row groups >> 4
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);  // start address      [0,14)
    d |= static_cast<uint64_t>(1) << 16;                  // leading byte offset [16,30) (unused, K-major swizzled)
    d |= kSBO << 32;                                      // stride byte offset  [32,46)
    d |= static_cast<uint64_t>(1) << 46;                  // descriptor version  [46,48) = 1 on sm_100
    d |= kLayout << 61;                                   // layout type         [61,64)
    return d;
}

template <int BN, bool kBF16>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    // cute::UMMA::InstrDescriptor: c_format [4,6)=1 (F32); a_format [7,10), b_format [10,13): 0=F16 1=BF16;
    // a_major [15], b_major [16] = 0 (K-major); n_dim [17,23) = N>>3; m_dim [24,29) = M>>4.
    return (1u << 4) | ((kBF16 ? 1u : 0u) << 7) | ((kBF16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
           (static_cast<uint32_t>(kTileM >> 4) << 24);
}

// MT = number of 128-row M tiles a CTA works on at once (sharing each weight slab): MT = 2 halves the weight
// traffic per FLOP and alternates MMAs between two independent accumulators.
template <int BN, int BK, int MT = 1>
struct ConvCfg {
    static constexpr int kATile = kTileM * BK * 2;
    static constexpr int kABytes = MT * kATile;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEW = BN < 64 ? BN : 64;                 // channels per epilogue pass
    static constexpr int kStgTile = kTileM * kEW * 2;             // staging tile of one epilogue group
    static constexpr int kStgBytes = MT * ((kStgTile + 1023) / 1024 * 1024);
    static constexpr int kStagesRaw = (kSmemMax - kSmemExtra - kStgBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + kSmemExtra;
    static constexpr int kAccCols = 2 * MT * BN;
    static constexpr int kTmemCols = (kAccCols <= 32) ? 32 : (kAccCols <= 64) ? 64 : (kAccCols <= 128) ? 128 : (kAccCols <= 256) ? 256 : 512;
    static constexpr int kThreads = 128 + 128 * MT;  // 4 control warps + 4 epilogue warps per M tile
    static_assert(kAccCols <= 512, "two accumulator stages must fit TMEM");
    static_assert(kStages >= 2, "need a pipeline");
};

// Sticky range flag of the fp16 modes: set by any epilogue that rounds a value beyond the fp16 range (|v| > 65504 ->
// inf, or a NaN) when it stores an activation.  One instance per device (module-scope __device__ variable); read and
// cleared through w2l_f16_overflow().  bf16 has fp32's exponent range and never sets it.
__device__ int g_f16_overflow = 0;

// `live` = the value belongs to a real output pixel.  (GEMM rows beyond a partial tile box are computed from stale shared
// memory — any bit pattern, NaN included — and never stored; they must not raise the flag.)
template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float a, float b, bool live = true) {
    if constexpr (kBF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        const uint32_t u = *reinterpret_cast<uint32_t*>(&h);
        // exponent field all ones (inf / NaN) in either half: adding 0x0400 to the magnitude bits carries into bit 15
        if (live && (((u & 0x7FFF7FFFu) + 0x04000400u) & 0x80008000u)) g_f16_overflow = 1;
        return u;
    }
}
template <bool kBF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    if constexpr (kBF16) {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    } else {
        return __half22float2(*reinterpret_cast<__half2*>(&u));
    }
}

// One accumulator tile (this thread's row = one output pixel, BN columns) -> epilogue math -> global memory.
template <int BN, bool kBF16, bool kHead>
__device__ __forceinline__ void epilogue_tile(const EpiParams& e, uint32_t taddr, bool valid, int n, int y, int x, int nt) {
    constexpr int CH = (BN >= 32) ? 32 : 16;  // columns per tcgen05.ld batch
    const long long o_off = (long long)n * e.out_sn + (long long)y * e.out_sy + (long long)x * e.out_sx;
    const long long r_off = (long long)n * e.res_sn + (long long)y * e.res_sy + (long long)x * e.res_sx;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t v[CH];
        tmem_ld16(taddr + c0, v);
        if constexpr (CH == 32) tmem_ld16(taddr + c0 + 16, v + 16);
        tmem_ld_wait();
        if (valid) {
            const int cg = nt * BN + c0;  // first output channel of this batch
            float f[CH];
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(e.scale + cg + j));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(e.shift + cg + j));
                f[j + 0] = fmaf(__uint_as_float(v[j + 0]), sc.x, sh.x);
                f[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
                f[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z);
                f[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
            }
            if (e.res != nullptr) {
                const uint4* rp = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const uint16_t*>(e.res) + r_off + cg);
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) {
                    const uint4 r = __ldg(rp + j);
                    const float2 a = unpack2<kBF16>(r.x), b = unpack2<kBF16>(r.y);
                    const float2 c = unpack2<kBF16>(r.z), d = unpack2<kBF16>(r.w);
                    f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
                    f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
                }
                if (e.x2) {  // lo plane of the residual
#pragma unroll
                    for (int j = 0; j < CH / 8; ++j) {
                        const uint4 r = __ldg(rp + (e.res_lo_off >> 3) + j);
                        const float2 a = unpack2<kBF16>(r.x), b = unpack2<kBF16>(r.y);
                        const float2 c = unpack2<kBF16>(r.z), d = unpack2<kBF16>(r.w);
                        f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
                        f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
                    }
                }
            }
            if (e.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.0f);
            } else if (e.act == ACT_LRELU) {
#pragma unroll
                for (int j = 0; j < CH; ++j) f[j] = f[j] > 0.0f ? f[j] : 0.01f * f[j];
            }
            if constexpr (kHead) {
                // wav2lip.py:84-85: Conv2d(32,3,1) + Sigmoid on the fp32 block output still in registers
                const int hb = n % e.head_B, ht = n / e.head_B;
                const long long plane = (long long)e.Hout * e.Wout;
#pragma unroll
                for (int oc = 0; oc < 3; ++oc) {
                    float s = __ldg(e.head_b + oc);
#pragma unroll
                    for (int j = 0; j < 32; ++j) s = fmaf(f[j], __ldg(e.head_w + oc * 32 + j), s);
                    s = 1.0f / (1.0f + __expf(-s));
                    if (e.head_out_u8 != nullptr)
                        e.head_out_u8[(((long long)n * e.Hout + y) * e.Wout + x) * 3 + oc] = (unsigned char)__fmul_rn(s, 255.0f);
                    else
                        e.head_out[(((long long)hb * 3 + oc) * e.head_T + ht) * plane + (long long)y * e.Wout + x] = s;
                }
            } else if (e.out_f32) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + o_off + cg);
#pragma unroll
                for (int j = 0; j < CH / 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + o_off + cg);
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) {
                    uint4 o;
                    o.x = pack2<kBF16>(f[8 * j + 0], f[8 * j + 1]);
                    o.y = pack2<kBF16>(f[8 * j + 2], f[8 * j + 3]);
                    o.z = pack2<kBF16>(f[8 * j + 4], f[8 * j + 5]);
                    o.w = pack2<kBF16>(f[8 * j + 6], f[8 * j + 7]);
                    op[j] = o;
                    if (e.x2) {  // lo = fp16(v - fp16(v)): together the two planes carry ~22 significant bits
                        const float2 h0 = unpack2<kBF16>(o.x), h1 = unpack2<kBF16>(o.y);
                        const float2 h2 = unpack2<kBF16>(o.z), h3 = unpack2<kBF16>(o.w);
                        uint4 l;
                        l.x = pack2<kBF16>(f[8 * j + 0] - h0.x, f[8 * j + 1] - h0.y);
                        l.y = pack2<kBF16>(f[8 * j + 2] - h1.x, f[8 * j + 3] - h1.y);
                        l.z = pack2<kBF16>(f[8 * j + 4] - h2.x, f[8 * j + 5] - h2.y);
                        l.w = pack2<kBF16>(f[8 * j + 6] - h3.x, f[8 * j + 7] - h3.y);
                        op[(e.out_lo_off >> 3) + j] = l;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int BK, bool kBF16, bool kHead, int MT = 1>
__global__ void __launch_bounds__(ConvCfg<BN, BK, MT>::kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
    pdl_launch_dependents();
    using Cfg = ConvCfg<BN, BK, MT>;
    constexpr int kStages = Cfg::kStages;
    static_assert(!kHead || BN == 32, "fused head expects the 32-channel output block");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms need 1024-B alignment
    const uint32_t stg_base = smem_base + kStages * Cfg::kStageBytes;
    const uint32_t bar_base = stg_base + Cfg::kStgBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    auto res_bar = [&](int g) { return bar_base + 8u * (2 * kStages + 4 + g); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.tma_epi) {
            tma_prefetch_desc(&p.tmO);
            tma_prefetch_desc(&p.tmR);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4 * MT);  // one arrive per epilogue warp
        }
        for (int g = 0; g < MT; ++g) mbar_init(res_bar(g), 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    pdl_wait();  // everything above overlaps the previous kernel's tail; global memory is touched only below
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int m_units = (m_tiles + MT - 1) / MT;  // a unit = MT consecutive M tiles x one N tile (a tile index past the
    const int total_tiles = m_units * p.n_tiles;  // end decodes to n >= N: its loads are zero-filled, its rows masked)
    const int k_steps = p.ntaps * p.kc_per_tap;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int mu = tile / p.n_tiles;
                int x_in0[MT], y_in0[MT], n0[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int m = mu * MT + mt;
                    const int tx = m % p.tiles_x;
                    const int ty = (m / p.tiles_x) % p.tiles_y;
                    const int tn = m / (p.tiles_x * p.tiles_y);
                    x_in0[mt] = tx * p.bw * p.sx;
                    y_in0[mt] = ty * p.bh * p.sy;
                    n0[mt] = tn * p.bn;
                }
                for (int t = 0; t < p.ntaps; ++t) {
                    for (int kc = 0; kc < p.kc_per_tap; ++kc) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
                        const uint32_t b_dst = a_dst + Cfg::kABytes;
                        mbar_arrive_expect_tx(full_bar(stage), p.stage_tx_bytes);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
                            tma_load_4d(a_dst + mt * Cfg::kATile, &p.tmA, full_bar(stage), kc * BK + (p.a_lo[t] ? p.a_lo_off : 0),
                                        x_in0[mt] + p.dx[t], y_in0[mt] + p.dy[t], n0[mt]);
                        tma_load_3d(b_dst, &p.tmB, full_bar(stage), kc * BK, nt * BN, p.b_slab[t]);
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc<BN, kBF16>();
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (MT * BN);
            for (int ks = 0; ks < k_steps; ++ks) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                    const uint64_t bdesc = make_kmajor_desc<BK>(a_addr + Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 32 B (16 elements) along K inside the swizzle atom: +2 in the >>4 address field
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
                            tc_mma_f16(tmem_d + mt * BN, make_kmajor_desc<BK>(a_addr + mt * Cfg::kATile) + 2u * k,
                                       bdesc + 2u * k, idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));  // frees the smem stage when these MMAs retire
                    if (ks == k_steps - 1) tc_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int q = (warp - 4) & 3;      // TMEM lane quarter this warp may access (warp id % 4)
        const int mt = (warp - 4) >> 2;    // which of the unit's M tiles this epilogue group drains
        const int row = q * 32 + lane;     // GEMM row == pixel index inside the tile box
        const int rows_valid = p.bw * p.bh * p.bn;
        const int px = row % p.bw;
        const int py = (row / p.bw) % p.bh;
        const int pn = row / (p.bw * p.bh);
        if (!kHead && p.tma_epi) {
            // ---- staged epilogue: residual in by TMA, result out by TMA, one swizzled smem tile per group ----
            constexpr int EW = Cfg::kEW;
            constexpr int kPasses = BN / EW;
            constexpr uint32_t kRowB = EW * 2;
            constexpr uint32_t kSwz = (kRowB == 128) ? 7u : (kRowB == 64) ? 3u : 1u;
            const uint32_t stg = stg_base + mt * ((Cfg::kStgTile + 1023) / 1024 * 1024);
            const bool leader = (q == 0 && lane == 0);
            const uint32_t bar_id = 1 + mt;
            const bool has_res = p.ep.res != nullptr;
            const EpiParams& e = p.ep;
            auto tile_origin = [&](int tile_, int* nt_, int* x0_, int* y0_, int* n0_) {
                *nt_ = tile_ % p.n_tiles;
                const int m_ = (tile_ / p.n_tiles) * MT + mt;
                *x0_ = (m_ % p.tiles_x) * p.bw;
                *y0_ = ((m_ / p.tiles_x) % p.tiles_y) * p.bh;
                *n0_ = (m_ / (p.tiles_x * p.tiles_y)) * p.bn;
            };
            uint32_t rphase = 0;
            if (has_res && leader && static_cast<int>(blockIdx.x) < total_tiles) {
                int nt0, x0, y0, n0;
                tile_origin(blockIdx.x, &nt0, &x0, &y0, &n0);
                mbar_arrive_expect_tx(res_bar(mt), p.epi_box_bytes);
                tma_load_4d(stg, &p.tmR, res_bar(mt), nt0 * BN, x0, y0, n0);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                int nt, x0, y0, n0;
                tile_origin(tile, &nt, &x0, &y0, &n0);
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BN) + mt * BN;
#pragma unroll 1
                for (int ps = 0; ps < kPasses; ++ps) {
                    uint32_t v[EW];
#pragma unroll
                    for (int c0 = 0; c0 < EW; c0 += 16) tmem_ld16(taddr + ps * EW + c0, v + c0);
                    tmem_ld_wait();
                    if (ps == kPasses - 1) {  // accumulator fully read: release the TMEM stage
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(acc));
                    }
                    if (has_res) {
                        mbar_wait(res_bar(mt), rphase);
                        rphase ^= 1u;
                    }
                    const int cg = nt * BN + ps * EW;
#pragma unroll
                    for (int j = 0; j < EW / 8; ++j) {
                        float f[8];
                        const float4 s0 = __ldg(reinterpret_cast<const float4*>(e.scale + cg + 8 * j));
                        const float4 s1 = __ldg(reinterpret_cast<const float4*>(e.scale + cg + 8 * j + 4));
                        const float4 h0 = __ldg(reinterpret_cast<const float4*>(e.shift + cg + 8 * j));
                        const float4 h1 = __ldg(reinterpret_cast<const float4*>(e.shift + cg + 8 * j + 4));
                        f[0] = fmaf(__uint_as_float(v[8 * j + 0]), s0.x, h0.x);
                        f[1] = fmaf(__uint_as_float(v[8 * j + 1]), s0.y, h0.y);
                        f[2] = fmaf(__uint_as_float(v[8 * j + 2]), s0.z, h0.z);
                        f[3] = fmaf(__uint_as_float(v[8 * j + 3]), s0.w, h0.w);
                        f[4] = fmaf(__uint_as_float(v[8 * j + 4]), s1.x, h1.x);
                        f[5] = fmaf(__uint_as_float(v[8 * j + 5]), s1.y, h1.y);
                        f[6] = fmaf(__uint_as_float(v[8 * j + 6]), s1.z, h1.z);
                        f[7] = fmaf(__uint_as_float(v[8 * j + 7]), s1.w, h1.w);
                        uint32_t a = stg + row * kRowB + j * 16;
                        a ^= ((a >> 7) & kSwz) << 4;
                        if (has_res) {
                            uint32_t r0, r1, r2, r3;
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
                            const float2 a0 = unpack2<kBF16>(r0), a1 = unpack2<kBF16>(r1);
                            const float2 a2 = unpack2<kBF16>(r2), a3 = unpack2<kBF16>(r3);
                            f[0] += a0.x; f[1] += a0.y; f[2] += a1.x; f[3] += a1.y;
                            f[4] += a2.x; f[5] += a2.y; f[6] += a3.x; f[7] += a3.y;
                        }
                        if (e.act == ACT_RELU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.0f);
                        } else if (e.act == ACT_LRELU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.0f ? f[i] : 0.01f * f[i];
                        }
                        const bool live = row < rows_valid;
                        const uint32_t o0 = pack2<kBF16>(f[0], f[1], live), o1 = pack2<kBF16>(f[2], f[3], live);
                        const uint32_t o2 = pack2<kBF16>(f[4], f[5], live), o3 = pack2<kBF16>(f[6], f[7], live);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (leader) {
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&p.tmO)), "r"(stg), "r"(cg), "r"(x0), "r"(y0), "r"(n0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the store has read the tile
                        if (has_res) {  // fetch the residual of the next pass / next tile into the (now free) tile
                            int nnt = nt, nx0 = x0, ny0 = y0, nn0 = n0, nps = ps + 1;
                            bool more = true;
                            if (nps == kPasses) {
                                nps = 0;
                                const int ntile = tile + static_cast<int>(gridDim.x);
                                more = ntile < total_tiles;
                                if (more) tile_origin(ntile, &nnt, &nx0, &ny0, &nn0);
                            }
                            if (more) {
                                mbar_arrive_expect_tx(res_bar(mt), p.epi_box_bytes);
                                tma_load_4d(stg, &p.tmR, res_bar(mt), nnt * BN + nps * EW, nx0, ny0, nn0);
                            }
                        }
                    }
                    // everyone else must not touch the tile again before the leader is past wait_group.read
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
            }
            if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                const int nt = tile % p.n_tiles;
                const int m = (tile / p.n_tiles) * MT + mt;
                const int tx = m % p.tiles_x;
                const int ty = (m / p.tiles_x) % p.tiles_y;
                const int tn = m / (p.tiles_x * p.tiles_y);
                const int x = tx * p.bw + px;
                const int y = ty * p.bh + py;
                const int n = tn * p.bn + pn;
                const bool valid = (row < rows_valid) && (x < p.ep.Wout) && (y < p.ep.Hout) && (n < p.ep.N);

                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BN) + mt * BN;
                epilogue_tile<BN, kBF16, kHead>(p.ep, taddr, valid, n, y, x, nt);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace w2l
