// conv_halo.cuh — "patch" variant of the implicit-GEMM conv for layers with few channels, where the generic
// kernel (conv_tcgen05.cuh) is bound by L2->SM operand traffic because it re-loads the input box for every
// filter tap (profiles/r1_v1_ncu_full_conv_summary.txt).  Used for
//   * 3x3 / stride 1 / pad 1 blocks with Cout <= 64 (conv.py:5-19 at wav2lip.py:16-22,40-45,79-83; the
//     80->32 output block with the fused 1x1+sigmoid head, wav2lip.py:83-85),
//   * the output phases of the last transposed conv (conv.py:33-44 at wav2lip.py:79), and
//   * the first layers with tiny Cin whose kw horizontal taps are folded into K (7x7 / 3x3 on 1..15 channels).
//
// A persistent CTA
//   * keeps ALL weights of the launch resident in shared memory (taps x Cin x Cout x 2 B <= ~100 KB), loaded
//     once by TMA, and
//   * per output tile (8 wide x 16 tall pixels of one image = 128 GEMM rows) loads ONE input patch per channel
//     chunk (PW x PH pixels x BK channels, e.g. 10 x 18 for a 3x3 conv; out-of-bounds zero-filled by TMA =
//     the conv padding).
// Every tap is a shifted VIEW of that patch: in the K-major swizzled layout a pixel is one shared-memory row,
// the 8 pixels of an output row are 8 consecutive rows (one UMMA 8-row group) and the next output row starts
// exactly one patch row (PW pixels) further, so the UMMA descriptor of tap t is
//     start = patch + tap_row[t] * row_bytes,     stride-byte-offset = PW * row_bytes.
// The hardware swizzle is a function of the shared-memory ADDRESS bits (TMA writes and UMMA reads apply the
// same XOR), so group starts need not be aligned to the 8-row swizzle atom — verified on B200 by the parity
// tests.  Operand traffic per tile drops from taps x (A + B) to one patch.
//
// With K this short the epilogue, not the main loop, is the critical path, so there are TWO epilogue warp
// groups (one per TMEM accumulator stage, alternating tiles), the residual is prefetched into registers
// before the accumulator is ready, and scale/shift/head weights come from the constant bank (kernel params).
#pragma once

#include "conv_tcgen05.cuh"

namespace w2l {

constexpr int kHaloW = 8, kHaloH = 16;  // output tile (pixels)
constexpr int kHaloThreads = 384;       // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue A, 8-11 epilogue B
constexpr int kHaloMaxTaps = 9;
constexpr int kHaloMaxStages = 8;

struct alignas(64) HaloParams {
    CUtensorMap tmA;  // activations (C, W, H, N), box (BK, PW, PH, 1)
    CUtensorMap tmB;  // weights (Cin_pad, Cout_pad, taps), box (BK, BN, 1)
    int tiles_x, tiles_y;   // tiles per image
    int kc;                 // channel chunks of BK
    int stages;             // depth of the patch ring
    int PW, PH;             // patch size in pixels
    int ox, oy;             // patch origin relative to the tile origin (-1,-1 for a padded 3x3)
    int ntaps;
    int patch_bytes;        // PW*PH*BK*2 (TMA transaction size)
    int patch_stride;       // ring slot size (patch_bytes rounded up to 1024)
    int tap_row[kHaloMaxTaps];  // first patch row (pixel index) of each tap's view
    EpiParams ep;
    float cscale[64], cshift[64];  // folded BatchNorm, constant bank
    float chead_w[96], chead_b[4]; // fused generator head
};

template <int BN, int BK, bool kBF16, bool kHead>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_patch_kernel(const __grid_constant__ HaloParams p) {
    constexpr int kSlab = BN * BK * 2;  // one (tap, chunk) weight slab
    constexpr int kRowBytes = BK * 2;
    constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : 256;
    static_assert(BN <= 64, "resident-weight variant is for narrow layers");
    static_assert(!kHead || BN == 32, "fused head expects the 32-channel output block");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int kc = p.kc;
    const int stages = p.stages;
    const int ntaps = p.ntaps;
    const uint32_t w_base = smem_base;
    const uint32_t a_base = w_base + static_cast<uint32_t>(ntaps * kc) * kSlab;
    const uint32_t bar_base = a_base + static_cast<uint32_t>(stages * kc) * p.patch_stride;  // a stage = all chunks of one tile
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kHaloMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kHaloMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kHaloMaxStages + 2 + a); };
    const uint32_t w_bar = bar_base + 8u * (2 * kHaloMaxStages + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kHaloMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_img * p.ep.N;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(ntaps * kc) * kSlab);
            for (int tap = 0; tap < ntaps; ++tap)
                for (int c = 0; c < kc; ++c)
                    tma_load_3d(w_base + (tap * kc + c) * kSlab, &p.tmB, w_bar, c * BK, 0, tap);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img;
                const int r = tile - n * tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(kc) * p.patch_bytes);
                for (int c = 0; c < kc; ++c)
                    tma_load_4d(a_base + (stage * kc + c) * p.patch_stride, &p.tmA, full_bar(stage), c * BK,
                                tx * kHaloW + p.ox, ty * kHaloH + p.oy, n);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc<BN, kBF16>();
        // descriptor halves: hi = [SBO | version | layout], lo = [start address >> 4 | LBO = 1]
        constexpr uint32_t kLayout = (BK == 64) ? 2u : (BK == 32) ? 4u : 6u;
        const uint32_t a_hi = ((static_cast<uint32_t>(p.PW) * kRowBytes) >> 4) | (1u << 14) | (kLayout << 29);
        constexpr uint32_t b_hi = ((8u * kRowBytes) >> 4) | (1u << 14) | (kLayout << 29);
        uint32_t tap_off[kHaloMaxTaps];
#pragma unroll
        for (int t = 0; t < kHaloMaxTaps; ++t) tap_off[t] = (t < ntaps ? p.tap_row[t] : 0) * kRowBytes;
        mbar_wait(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int c = 0; c < kc; ++c) {
                    const uint32_t patch = a_base + (stage * kc + c) * p.patch_stride;
#pragma unroll
                    for (int tap = 0; tap < kHaloMaxTaps; ++tap) {
                        if (tap < ntaps) {
                            const uint32_t a_lo = ((patch + tap_off[tap]) >> 4) | 0x10000u;
                            const uint32_t b_lo = ((w_base + (tap * kc + c) * kSlab) >> 4) | 0x10000u;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + 2u * k);
                                const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2u * k);
                                tc_mma_f16(tmem_d, adesc, bdesc, idesc, (c | tap | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                }
                tc_commit(empty_bar(stage));
                tc_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp >= 4) {
        // =============================== epilogue (two groups, one per accumulator stage) ===============================
        const int grp = (warp - 4) >> 2;   // 0: tiles 0,2,4..  1: tiles 1,3,5..
        const int q = (warp - 4) & 3;      // TMEM lane quarter = warp id % 4
        const int row = q * 32 + lane;
        const int py = row >> 3, px = row & 7;  // GEMM row -> pixel inside the 8 x 16 tile
        const EpiParams& e = p.ep;
        constexpr int CH = (BN >= 32) ? 32 : 16;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            const int acc = grp;
            const uint32_t acc_phase = (it >> 1) & 1u;
            const int n = tile / tiles_per_img;
            const int r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int x = tx * kHaloW + px, y = ty * kHaloH + py;
            const bool valid = (x < e.Wout) && (y < e.Hout);
            const long long o_off = (long long)n * e.out_sn + (long long)y * e.out_sy + (long long)x * e.out_sx;
            // residual prefetch: issued before the accumulator is ready, so its L2/HBM latency hides behind the MMAs
            uint4 rv[BN / 8];
            if (e.res != nullptr && valid) {
                const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(e.res) +
                                                                 (long long)n * e.res_sn + (long long)y * e.res_sy +
                                                                 (long long)x * e.res_sx);
#pragma unroll
                for (int j = 0; j < BN / 8; ++j) rv[j] = __ldg(rp + j);
            } else {
#pragma unroll
                for (int j = 0; j < BN / 8; ++j) rv[j] = make_uint4(0u, 0u, 0u, 0u);
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += CH) {
                uint32_t v[CH];
                tmem_ld16(taddr + c0, v);
                if constexpr (CH == 32) tmem_ld16(taddr + c0 + 16, v + 16);
                tmem_ld_wait();
                if (c0 + CH >= BN) {  // accumulator fully read: hand the TMEM stage back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (valid) {
                    float f[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) f[j] = fmaf(__uint_as_float(v[j]), p.cscale[c0 + j], p.cshift[c0 + j]);
#pragma unroll
                    for (int j = 0; j < CH / 8; ++j) {
                        const uint4 rr = rv[c0 / 8 + j];
                        const float2 a = unpack2<kBF16>(rr.x), b = unpack2<kBF16>(rr.y);
                        const float2 c = unpack2<kBF16>(rr.z), d = unpack2<kBF16>(rr.w);
                        f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
                        f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
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
                            float s = p.chead_b[oc];
#pragma unroll
                            for (int j = 0; j < 32; ++j) s = fmaf(f[j], p.chead_w[oc * 32 + j], s);
                            s = 1.0f / (1.0f + __expf(-s));
                            e.head_out[(((long long)hb * 3 + oc) * e.head_T + ht) * plane + (long long)y * e.Wout + x] = s;
                        }
                    } else if (e.out_f32) {
                        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + o_off + c0);
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + o_off + c0);
#pragma unroll
                        for (int j = 0; j < CH / 8; ++j) {
                            uint4 o;
                            o.x = pack2<kBF16>(f[8 * j + 0], f[8 * j + 1]);
                            o.y = pack2<kBF16>(f[8 * j + 2], f[8 * j + 3]);
                            o.z = pack2<kBF16>(f[8 * j + 4], f[8 * j + 5]);
                            o.w = pack2<kBF16>(f[8 * j + 6], f[8 * j + 7]);
                            op[j] = o;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace w2l
