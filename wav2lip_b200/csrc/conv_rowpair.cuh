// conv_rowpair.cuh — "row-pair" variant of the patch kernel (conv_patch.cuh) for the three heaviest narrow layers of
// the generator: face_decoder_blocks.6.1 / 6.2 (64 -> 64, residual, wav2lip.py:80-81) and output_block.0 (80 -> 32 with
// the fused 1x1 + sigmoid head, wav2lip.py:83-85), all 3x3 / stride 1 / pad 1 at 96 x 96.
//
// Why: with the pixels as the M operand an M=128, K=16 tcgen05.mma costs >= 64 cycles however small N is (the A
// operand is read from shared memory at ~64 B/clk), so a layer with Cout = 64 (32) can use at most 50 % (25 %) of the
// tensor pipe — DESIGN.md section 3.  This kernel doubles N without doubling the work per output: GEMM row r stands
// for the PAIR of output pixels (x, 2q) and (x, 2q+1), and the accumulator's columns [0,C) belong to the upper pixel,
// [C,2C) to the lower one.  An A view that starts at input row 2q+v (v = -1..2) feeds
//      v = -1 : upper pixel, filter row dy=-1                       (N = C,  columns [0,C))
//      v =  0 : upper pixel dy=0  and lower pixel dy=-1             (N = 2C, weights [W(dy=0) | W(dy=-1)])
//      v = +1 : upper pixel dy=+1 and lower pixel dy=0              (N = 2C, weights [W(dy=+1) | W(dy=0)])
//      v = +2 : lower pixel dy=+1                                   (N = C,  columns [C,2C))
// With the resident weight slabs ordered dy = +1, 0, -1 the two-slab windows are contiguous, so nothing is stored
// twice.  4 instructions per (dx, 16 channels) now produce 2 x 128 pixels instead of 6: 1.5x fewer tensor-pipe cycles
// per output.  The A views are the same shifted views of one haloed input patch as in conv_patch.cuh, with the
// stride between 8-row groups set to TWO patch rows.
//
// The dependent-issue latency on one accumulator (~90 cycles) is hidden by splitting every tile's MMAs over two
// accumulators (views {0,-1} and {+1,+2}) that the epilogue adds.  TMEM: 2 accumulators x 2C columns x 2 tiles
// in flight = 8C columns (512 for C = 64).
//
// Input channels come in one 64-wide chunk (128-byte rows, SWIZZLE_128B) plus, for the 80-channel concat buffer of
// the output block, one 16-wide chunk (32-byte rows, SWIZZLE_32B) — 2 TMA loads per tile instead of 5.
#pragma once

#include "conv_tcgen05.cuh"

namespace w2l {

constexpr int kRpTileW = 8, kRpTileH = 32;   // output tile: 8 x 32 pixels = 128 row pairs
constexpr int kRpPW = 10, kRpPH = 34;        // input patch with the 1-pixel halo
constexpr int kRpThreads = 384;              // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue A, 8-11 epilogue B
constexpr int kRpMaxStages = 4;

struct alignas(64) RowPairParams {
    CUtensorMap tmA0;  // activations (C, W, H, N), box (64, 10, 34, 1), SWIZZLE_128B
    CUtensorMap tmA1;  // same tensor, box (16, 10, 34, 1), SWIZZLE_32B — channels [64, 80)
    CUtensorMap tmB0;  // weights (Cin_pad, Cout, taps), box (64, C, 1)
    CUtensorMap tmB1;  // box (16, C, 1)
    CUtensorMap tmO;   // output channel slice (C, W, H, N), box (C, 8, 32, 1)
    int tiles_x, tiles_y;
    int stages;        // depth of the patch ring
    int tap_of[9];     // packed-weight tap index of slab (dxi * 3 + dyi); dyi = 0,1,2 <-> dy = +1,0,-1; dxi <-> dx = dxi - 1
    int has_res;       // the residual is the block's own input: centre of the patch
    EpiParams ep;
    float cscale[64], cshift[64];
    float chead_w[96], chead_b[4];
};

template <int C, int BK1>
struct RowPairCfg {
    static constexpr int kRow0 = 128, kRow1 = BK1 * 2;
    static constexpr int kSlab0 = C * kRow0, kSlab1 = C * kRow1;
    static constexpr int kW0 = 9 * kSlab0, kW1 = (9 * kSlab1 + 1023) / 1024 * 1024;
    static constexpr int kPatch0 = kRpPW * kRpPH * kRow0, kPatch1 = kRpPW * kRpPH * kRow1;
    static constexpr int kStride0 = (kPatch0 + 1023) / 1024 * 1024, kStride1 = (kPatch1 + 1023) / 1024 * 1024;
    static constexpr int kStageStride = kStride0 + kStride1;
    static constexpr int kStgTile = kRpTileW * kRpTileH * C * 2;  // one staged output tile
    static constexpr int kTmemCols = 8 * C;
    static constexpr int smem_bytes(int stages, bool head) {
        return kW0 + kW1 + stages * kStageStride + (head ? 0 : 2 * kStgTile) + kSmemExtra;
    }
};

template <int C, int BK1, bool kBF16, bool kHead>
__global__ void __launch_bounds__(kRpThreads, 1) conv_rowpair_kernel(const __grid_constant__ RowPairParams p) {
    pdl_launch_dependents();
    using Cfg = RowPairCfg<C, BK1>;
    static_assert(C == 32 || C == 64, "row-pair kernel: Cout is 32 or 64");
    static_assert(BK1 == 0 || BK1 == 16, "second channel chunk is absent or 16 wide");
    static_assert(!kHead || C == 32, "fused head expects the 32-channel output block");
    constexpr int PW = kRpPW;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = p.stages;
    const uint32_t w0_base = smem_base;
    const uint32_t w1_base = w0_base + Cfg::kW0;
    const uint32_t a_base = w1_base + Cfg::kW1;
    const uint32_t stg_base = a_base + static_cast<uint32_t>(stages) * Cfg::kStageStride;
    const uint32_t bar_base = stg_base + (kHead ? 0u : 2u * Cfg::kStgTile);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kRpMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kRpMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kRpMaxStages + 2 + a); };
    const uint32_t w_bar = bar_base + 8u * (2 * kRpMaxStages + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kRpMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA0);
        tma_prefetch_desc(&p.tmB0);
        if (BK1) { tma_prefetch_desc(&p.tmA1); tma_prefetch_desc(&p.tmB1); }
        if (!kHead) tma_prefetch_desc(&p.tmO);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), p.has_res ? 5 : 1);  // MMA commit (+ the 4 epilogue warps that read the residual)
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
            mbar_arrive_expect_tx(w_bar, 9u * (Cfg::kSlab0 + Cfg::kSlab1));
            for (int s = 0; s < 9; ++s) {
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
                tma_load_4d(dst, &p.tmA0, full_bar(stage), 0, tx * kRpTileW - 1, ty * kRpTileH - 1, n);
                if (BK1) tma_load_4d(dst + Cfg::kStride0, &p.tmA1, full_bar(stage), 64, tx * kRpTileW - 1, ty * kRpTileH - 1, n);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc2 = make_idesc<2 * C, kBF16>();
        constexpr uint32_t idesc1 = make_idesc<C, kBF16>();
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
                const uint32_t acc0 = tmem_base + ts * 4 * C;   // views 0 and -1
                const uint32_t acc1 = acc0 + 2 * C;             // views +1 and +2
#pragma unroll
                for (int ch = 0; ch < (BK1 ? 2 : 1); ++ch) {
                    const uint32_t rowb = ch ? Cfg::kRow1 : Cfg::kRow0;
                    const uint32_t layout = ch ? 6u : 2u;
                    const uint32_t patch = a_base + stage * Cfg::kStageStride + (ch ? Cfg::kStride0 : 0);
                    const uint32_t wb = ch ? w1_base : w0_base;
                    const uint32_t slab = ch ? Cfg::kSlab1 : Cfg::kSlab0;
                    const int ksteps = ch ? BK1 / 16 : 4;
                    // descriptor halves: hi = [SBO | version | layout], lo = [start address >> 4 | LBO = 1]
                    const uint32_t a_hi = ((2u * PW * rowb) >> 4) | (1u << 14) | (layout << 29);  // next 8-pixel group: two rows down
                    const uint32_t b_hi = ((8u * rowb) >> 4) | (1u << 14) | (layout << 29);
#pragma unroll
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        const uint32_t a_m1 = ((patch + (0 * PW + dxi) * rowb) >> 4) | 0x10000u;
                        const uint32_t a_0 = ((patch + (1 * PW + dxi) * rowb) >> 4) | 0x10000u;
                        const uint32_t a_p1 = ((patch + (2 * PW + dxi) * rowb) >> 4) | 0x10000u;
                        const uint32_t a_p2 = ((patch + (3 * PW + dxi) * rowb) >> 4) | 0x10000u;
                        const uint32_t b_p1 = ((wb + (dxi * 3 + 0) * slab) >> 4) | 0x10000u;   // [W(+1) | W(0)]
                        const uint32_t b_0 = ((wb + (dxi * 3 + 1) * slab) >> 4) | 0x10000u;    // [W(0) | W(-1)]
                        const uint32_t b_m1 = ((wb + (dxi * 3 + 2) * slab) >> 4) | 0x10000u;   // W(-1)
                        for (int k = 0; k < ksteps; ++k) {
                            const uint32_t accum = (ch | dxi | k) != 0 ? 1u : 0u;
                            const uint64_t ah = static_cast<uint64_t>(a_hi) << 32, bh = static_cast<uint64_t>(b_hi) << 32;
                            tc_mma_f16(acc0, ah | (a_0 + 2u * k), bh | (b_0 + 2u * k), idesc2, accum);
                            tc_mma_f16(acc1, ah | (a_p1 + 2u * k), bh | (b_p1 + 2u * k), idesc2, accum);
                            tc_mma_f16(acc0, ah | (a_m1 + 2u * k), bh | (b_m1 + 2u * k), idesc1, 1u);
                            tc_mma_f16(acc1 + C, ah | (a_p2 + 2u * k), bh | (b_p1 + 2u * k), idesc1, 1u);
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
        const int row = q * 32 + lane;     // GEMM row = pixel pair
        const int pr = row >> 3, px = row & 7;
        const EpiParams& e = p.ep;
        constexpr uint32_t kOutRow = C * 2;
        constexpr uint32_t kOutSwz = (kOutRow == 128) ? 7u : 3u;
        const uint32_t stg = stg_base + grp * Cfg::kStgTile;
        const bool leader = (q == 0 && lane == 0);
        const uint32_t bar_id = 1 + grp;
        for (int u = 0;; ++u) {
            const int it = 2 * u + grp;
            const int tile = blockIdx.x + it * gridDim.x;
            if (tile >= total_tiles) break;
            const int stage = it % stages;
            const int n = tile / tiles_per_img;
            const int r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;

            // residual = the block's input = centre of the patch, fetched while the tensor pipe is still busy with this tile
            uint32_t resv[kHead ? 1 : 2][kHead ? 1 : C / 2];
            if constexpr (!kHead) {
                if (p.has_res) {
                    mbar_wait(full_bar(stage), (it / stages) & 1u);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t prow = a_base + stage * Cfg::kStageStride + ((2 * pr + j + 1) * PW + px + 1) * Cfg::kRow0;
#pragma unroll
                        for (int c8 = 0; c8 < C / 8; ++c8) {
                            uint32_t a = prow + c8 * 16;
                            a ^= ((a >> 7) & 7u) << 4;
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(resv[j][4 * c8]), "=r"(resv[j][4 * c8 + 1]), "=r"(resv[j][4 * c8 + 2]), "=r"(resv[j][4 * c8 + 3])
                                         : "r"(a));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar(stage));
                }
            }

            mbar_wait(tfull_bar(grp), u & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * 4 * C;

            if constexpr (kHead) {
                float f[2][C];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t v0[C], v1[C];
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        tmem_ld16(tacc + j * C + c0, v0 + c0);
                        tmem_ld16(tacc + 2 * C + j * C + c0, v1 + c0);
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
                const int x = tx * kRpTileW + px;
                const int hb = n % e.head_B, ht = n / e.head_B;
                const long long plane = (long long)e.Hout * e.Wout;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int y = ty * kRpTileH + 2 * pr + j;
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
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16(tacc + j * C + c0, v0);
                        tmem_ld16(tacc + 2 * C + j * C + c0, v1);
                        tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            f[c] = fmaf(__uint_as_float(v0[c]) + __uint_as_float(v1[c]), p.cscale[c0 + c], p.cshift[c0 + c]);
                        if (p.has_res) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float2 rr = unpack2<kBF16>(resv[j][c0 / 2 + c]);
                                f[2 * c] += rr.x;
                                f[2 * c + 1] += rr.y;
                            }
                        }
                        if (e.act == ACT_RELU) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) f[c] = fmaxf(f[c], 0.0f);
                        } else if (e.act == ACT_LRELU) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) f[c] = f[c] > 0.0f ? f[c] : 0.01f * f[c];
                        }
                        // staged tile: pixel-major rows of C 16-bit channels in the hardware swizzle pattern of tmO
                        const uint32_t srow = stg + ((2 * pr + j) * kRpTileW + px) * kOutRow + c0 * 2;
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
                                 ::"l"(reinterpret_cast<uint64_t>(&p.tmO)), "r"(stg), "r"(0), "r"(tx * kRpTileW), "r"(ty * kRpTileH), "r"(n)
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
