// conv_rowstack.cuh — "row-stack" variant of the patch kernel (conv_patch.cuh) for the narrowest, highest-resolution
// stride-1 layers of the generator, both at 96 x 96:
//   * output_block.0  (80 -> 32, 3x3, with the fused 1x1 + sigmoid head, wav2lip.py:83-85):  S = 2 rows per GEMM row
//   * face_encoder_blocks.0.0 (6 -> 16, 7x7, wav2lip.py:16; kw folded into K):               S = 3 rows per GEMM row
//
// Why: with the pixels as the M operand an M=128, K=16 tcgen05.mma costs >= 64 cycles however small N is (the A
// operand is read from shared memory at ~64 B/clk), so a layer with Cout = 32 (16) can use at most 25 % (12.5 %) of the
// tensor pipe — DESIGN.md section 3.  This kernel widens N without adding work per output: GEMM row r stands for the
// S vertically adjacent output pixels (x, S*q + j), j < S, and accumulator columns [j*C, (j+1)*C) belong to pixel j.
// An A view that starts at input row S*q + v (v = -R .. R+S-1, R = the filter's vertical radius) is the input row of
// filter row dy = v - j for pixel j, so ONE instruction with the weight slabs of dy = v-jmin .. v-jmax side by side
// as its B operand serves up to S output rows.  With the resident slabs stored in order of decreasing dy those
// windows are contiguous, nothing is stored twice.  Per (dx, 16 channels) that is TY+S-1 instructions per S*128
// pixels instead of S*TY:   3x3, S=2: 4 instead of 6;    7 rows, S=3: 9 instead of 21.
// The A views are the same shifted views of one haloed input patch as in conv_patch.cuh, with the stride between
// 8-row groups set to S patch rows.
//
// The dependent-issue latency on one accumulator (~90 cycles) is hidden by splitting every tile's instructions over
// two accumulators that the epilogue adds; each starts with a full-width view (TY > S guarantees two of them).
// TMEM: 2 accumulators x S*C columns x 2 tiles in flight.
//
// Input channels come in one 64-wide chunk (128-byte rows, SWIZZLE_128B) plus, for the 80-channel concat buffer of
// the output block, one 16-wide chunk (32-byte rows, SWIZZLE_32B) — 2 TMA loads per tile instead of 5.
//
// (The same scheme on the 64 -> 64 blocks of decoder stage 6 was measured and dropped: N = 128 from shared memory
// needs 128 B/clk of operand bandwidth, all an SM has, and ran 12 % slower than the patch kernel.)
#pragma once

#include "conv_tcgen05.cuh"

namespace w2l {

constexpr int kRsTileW = 8;
constexpr int kRsThreads = 384;              // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue A, 8-11 epilogue B
constexpr int kRsMaxStages = 4;

struct alignas(64) RowStackParams {
    CUtensorMap tmA0;  // activations, box (64, PW, PH, 1), SWIZZLE_128B
    CUtensorMap tmA1;  // same tensor, box (16, PW, PH, 1), SWIZZLE_32B — channels [64, 80)
    CUtensorMap tmB0;  // weights (Cin_pad, Cout, taps), box (64, C, 1)
    CUtensorMap tmB1;  // box (16, C, 1)
    CUtensorMap tmO;   // output channel slice (C, W, H, N), box (C, 8, 16*S, 1)
    CUtensorMap tmO2;  // optional second destination of the same tile (dense zero-bordered copy for a folded consumer)
    int has_out2;
    int tiles_x, tiles_y;
    int stages;        // depth of the patch ring
    int ox, oy;        // patch origin relative to the tile origin
    int tap_of[21];    // packed-weight tap index of slab (dxi * TY + dyi); dyi = 0..TY-1 <-> dy = R .. -R
    EpiParams ep;
    float cscale[64], cshift[64];
    float chead_w[96], chead_b[4];
};

template <int C, int S, int TY, int NDX, int BK1>
struct RowStackCfg {
    static constexpr int R = (TY - 1) / 2;
    static constexpr int NV = TY + S - 1;                    // A views per (dx, k step)
    static constexpr int PW = kRsTileW + NDX - 1, PH = 16 * S + TY - 1;
    static constexpr int kTileH = 16 * S;
    static constexpr int kRow0 = 128, kRow1 = BK1 * 2;
    static constexpr int kSlab0 = C * kRow0, kSlab1 = C * kRow1;
    static constexpr int kNSlab = NDX * TY;
    static constexpr int kW0 = (kNSlab * kSlab0 + 1023) / 1024 * 1024, kW1 = (kNSlab * kSlab1 + 1023) / 1024 * 1024;
    static constexpr int kPatch0 = PW * PH * kRow0, kPatch1 = PW * PH * kRow1;
    static constexpr int kStride0 = (kPatch0 + 1023) / 1024 * 1024, kStride1 = (kPatch1 + 1023) / 1024 * 1024;
    static constexpr int kStageStride = kStride0 + kStride1;
    static constexpr int kStgTile = (kRsTileW * kTileH * C * 2 + 1023) / 1024 * 1024;  // one staged output tile
    static constexpr int kAccCols = S * C;
    static constexpr int kTmemRaw = 4 * kAccCols;
    static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : kTmemRaw <= 64 ? 64 : kTmemRaw <= 128 ? 128 : kTmemRaw <= 256 ? 256 : 512;
    static constexpr int smem_bytes(int stages, bool head) {
        return kW0 + kW1 + stages * kStageStride + (head ? 0 : 2 * kStgTile) + kSmemExtra;
    }
    static_assert(TY > S, "need two full-width views to initialise the two accumulators");
    static_assert(kTmemRaw <= 512, "accumulators exceed TMEM");
    static_assert(kNSlab <= 21, "tap table too small");
};

template <int C, int S, int TY, int NDX, int BK1, bool kBF16, bool kHead>
__global__ void __launch_bounds__(kRsThreads, 1) conv_rowstack_kernel(const __grid_constant__ RowStackParams p) {
    pdl_launch_dependents();
    using Cfg = RowStackCfg<C, S, TY, NDX, BK1>;
    static_assert(C == 16 || C == 32, "row-stack kernel: Cout is 16 or 32");
    static_assert(BK1 == 0 || BK1 == 16, "second channel chunk is absent or 16 wide");
    static_assert(!kHead || (C == 32 && S == 2), "fused head expects the 32-channel output block");
    constexpr int PW = Cfg::PW, R = Cfg::R, NV = Cfg::NV;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = p.stages;
    const uint32_t w0_base = smem_base;
    const uint32_t w1_base = w0_base + Cfg::kW0;
    const uint32_t a_base = w1_base + Cfg::kW1;
    const uint32_t stg_base = a_base + static_cast<uint32_t>(stages) * Cfg::kStageStride;
    const uint32_t bar_base = stg_base + (kHead ? 0u : 2u * Cfg::kStgTile);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kRsMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kRsMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kRsMaxStages + 2 + a); };
    const uint32_t w_bar = bar_base + 8u * (2 * kRsMaxStages + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kRsMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA0);
        tma_prefetch_desc(&p.tmB0);
        if (BK1) { tma_prefetch_desc(&p.tmA1); tma_prefetch_desc(&p.tmB1); }
        if (!kHead) tma_prefetch_desc(&p.tmO);
        if (!kHead && p.has_out2) tma_prefetch_desc(&p.tmO2);
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
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    pdl_wait();  // everything above overlaps the previous kernel's tail; global memory is touched only below
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_img * p.ep.N;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(Cfg::kNSlab) * (Cfg::kSlab0 + Cfg::kSlab1));
            for (int s = 0; s < Cfg::kNSlab; ++s) {
                tma_load_3d(w0_base + s * Cfg::kSlab0, &p.tmB0, w_bar, 0, 0, p.tap_of[s]);
                if (BK1) tma_load_3d(w1_base + s * Cfg::kSlab1, &p.tmB1, w_bar, 64, 0, p.tap_of[s]);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img;
                const int r = tile - n * tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(Cfg::kPatch0 + Cfg::kPatch1));
                const uint32_t dst = a_base + stage * Cfg::kStageStride;
                tma_load_4d(dst, &p.tmA0, full_bar(stage), 0, tx * kRsTileW + p.ox, ty * Cfg::kTileH + p.oy, n);
                if (BK1) tma_load_4d(dst + Cfg::kStride0, &p.tmA1, full_bar(stage), 64, tx * kRsTileW + p.ox, ty * Cfg::kTileH + p.oy, n);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc_base = (1u << 4) | ((kBF16 ? 1u : 0u) << 7) | ((kBF16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(kTileM >> 4) << 24);
        mbar_wait(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int ts = it & 1;
            mbar_wait(tempty_bar(ts), ((it >> 1) & 1u) ^ 1u);
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t acc0 = tmem_base + ts * 2 * Cfg::kAccCols;
#pragma unroll
                for (int ch = 0; ch < (BK1 ? 2 : 1); ++ch) {
                    const uint32_t rowb = ch ? Cfg::kRow1 : Cfg::kRow0;
                    const uint32_t layout = ch ? 6u : 2u;
                    const uint32_t patch = a_base + stage * Cfg::kStageStride + (ch ? Cfg::kStride0 : 0);
                    const uint32_t wb = ch ? w1_base : w0_base;
                    const uint32_t slab = ch ? Cfg::kSlab1 : Cfg::kSlab0;
                    const int ksteps = ch ? BK1 / 16 : 4;
                    // descriptor halves: hi = [SBO | version | layout], lo = [start address >> 4 | LBO = 1]
                    const uint64_t a_hi = static_cast<uint64_t>(((S * PW * rowb) >> 4) | (1u << 14) | (layout << 29)) << 32;  // next 8-pixel group: S rows down
                    const uint64_t b_hi = static_cast<uint64_t>(((8u * rowb) >> 4) | (1u << 14) | (layout << 29)) << 32;
#pragma unroll
                    for (int dxi = 0; dxi < NDX; ++dxi) {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint32_t accum = (ch | dxi | k) != 0 ? 1u : 0u;
#pragma unroll
                            for (int vi = 0; vi < NV; ++vi) {
                                // full-width views first (v = S-1-R .. R), then the partial ones on either side
                                constexpr int kFull = TY - S + 1;
                                const int v = vi < kFull ? (S - 1 - R + vi) : ((vi - kFull) < (S - 1) ? (-R + (vi - kFull)) : (R + 1 + (vi - kFull) - (S - 1)));
                                const int jmin = v - R > 0 ? v - R : 0;
                                const int jmax = v + R < S - 1 ? v + R : S - 1;
                                const int width = (jmax - jmin + 1) * C;
                                const uint32_t a_lo = ((patch + ((v + R) * PW + dxi) * rowb) >> 4) | 0x10000u;
                                const uint32_t b_lo = ((wb + (dxi * TY + (R - (v - jmin))) * slab) >> 4) | 0x10000u;
                                const uint32_t idesc = idesc_base | (static_cast<uint32_t>(width >> 3) << 17);
                                const uint32_t d = acc0 + (vi & 1) * Cfg::kAccCols + jmin * C;
                                tc_mma_f16(d, a_hi | (a_lo + 2u * k), b_hi | (b_lo + 2u * k), idesc, vi < 2 ? accum : 1u);
                            }
                        }
                    }
                }
                tc_commit(empty_bar(stage));
                tc_commit(tfull_bar(ts));
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp >= 4) {
        // =============================== epilogue (two groups alternating tiles) ===============================
        const int grp = (warp - 4) >> 2;
        const int q = (warp - 4) & 3;      // TMEM lane quarter = warp id % 4
        const int row = q * 32 + lane;     // GEMM row = S vertically adjacent pixels
        const int pr = row >> 3, px = row & 7;
        const EpiParams& e = p.ep;
        constexpr uint32_t kOutRow = C * 2;
        constexpr uint32_t kOutSwz = (kOutRow == 128) ? 7u : (kOutRow == 64) ? 3u : 1u;
        const uint32_t stg = stg_base + grp * Cfg::kStgTile;
        const bool leader = (q == 0 && lane == 0);
        const uint32_t bar_id = 1 + grp;
        for (int u = 0;; ++u) {
            const int it = 2 * u + grp;
            const int tile = blockIdx.x + it * gridDim.x;
            if (tile >= total_tiles) break;
            const int n = tile / tiles_per_img;
            const int r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;

            mbar_wait(tfull_bar(grp), u & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * 2 * Cfg::kAccCols;

            if constexpr (kHead) {
                float f[S][C];
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    uint32_t v0[C], v1[C];
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        tmem_ld16(tacc + j * C + c0, v0 + c0);
                        tmem_ld16(tacc + Cfg::kAccCols + j * C + c0, v1 + c0);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        f[j][c] = fmaxf(fmaf(__uint_as_float(v0[c]) + __uint_as_float(v1[c]), p.cscale[c], p.cshift[c]), 0.0f);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(grp));
                // wav2lip.py:84-85: Conv2d(32,3,1) + Sigmoid on the fp32 block output still in registers
                const int x = tx * kRsTileW + px;
                const int hb = n % e.head_B, ht = n / e.head_B;
                const long long plane = (long long)e.Hout * e.Wout;
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const int y = ty * Cfg::kTileH + S * pr + j;
                    if (x < e.Wout && y < e.Hout) {
#pragma unroll
                        for (int oc = 0; oc < 3; ++oc) {
                            float s = p.chead_b[oc];
#pragma unroll
                            for (int c = 0; c < 32; ++c) s = fmaf(f[j][c], p.chead_w[oc * 32 + c], s);
                            s = 1.0f / (1.0f + __expf(-s));
                            if (e.head_out_u8 != nullptr)
                                e.head_out_u8[(((long long)n * e.Hout + y) * e.Wout + x) * 3 + oc] = (unsigned char)__fmul_rn(s, 255.0f);
                            else
                                e.head_out[(((long long)hb * 3 + oc) * e.head_T + ht) * plane + (long long)y * e.Wout + x] = s;
                        }
                    }
                }
            } else {
                if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has read the buffer
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                for (int j = 0; j < S; ++j) {
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16(tacc + j * C + c0, v0);
                        tmem_ld16(tacc + Cfg::kAccCols + j * C + c0, v1);
                        tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            f[c] = fmaf(__uint_as_float(v0[c]) + __uint_as_float(v1[c]), p.cscale[c0 + c], p.cshift[c0 + c]);
                        if (e.act == ACT_RELU) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) f[c] = fmaxf(f[c], 0.0f);
                        } else if (e.act == ACT_LRELU) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) f[c] = f[c] > 0.0f ? f[c] : 0.01f * f[c];
                        }
                        // staged tile: pixel-major rows of C 16-bit channels in the hardware swizzle pattern of tmO
                        const uint32_t srow = stg + ((S * pr + j) * kRsTileW + px) * kOutRow + c0 * 2;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t a = srow + h * 16;
                            a ^= ((a >> 7) & kOutSwz) << 4;
                            const uint32_t o0 = pack2<kBF16>(f[8 * h + 0], f[8 * h + 1]);
                            const uint32_t o1 = pack2<kBF16>(f[8 * h + 2], f[8 * h + 3]);
                            const uint32_t o2 = pack2<kBF16>(f[8 * h + 4], f[8 * h + 5]);
                            const uint32_t o3 = pack2<kBF16>(f[8 * h + 6], f[8 * h + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(grp));  // accumulators drained
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&p.tmO)), "r"(stg), "r"(0), "r"(tx * kRsTileW), "r"(ty * Cfg::kTileH), "r"(n)
                                 : "memory");
                    if (p.has_out2)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&p.tmO2)), "r"(stg), "r"(0), "r"(tx * kRsTileW), "r"(ty * Cfg::kTileH), "r"(n)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (!kHead && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace w2l
