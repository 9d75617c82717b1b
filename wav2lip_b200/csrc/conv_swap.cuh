// conv_swap.cuh — the generic implicit GEMM with the operand roles swapped, for layers whose output-channel tile is 128.
//
// Same block as conv_igemm_kernel (/root/reference/models/conv.py:5-19 Conv2d, :33-44 Conv2dTranspose phase by phase):
//     y = act( scale[c] * conv(x, w)[., c] + shift[c]  (+ residual) )
// but the tcgen05 instruction is issued as  D[cout, pixel] (M = 128 output channels, N = 256 output pixels, K = 16):
//     A = the [128 x BK] weight slab of the current filter tap (what conv_igemm_kernel<128,64,.,.,2> uses as B),
//     B = the TWO 128-pixel input boxes of the unit, adjacent in shared memory = one 256-row K-major operand.
// Why: an M=128,N=128 instruction needs 4 KB (A) + 4 KB (B) of shared memory per 64 tensor-core cycles = 128 B/clk, all an
// SM's shared-memory bandwidth; the generic kernel's two N=128 instructions per K step sit exactly on that limit
// (66 % tensor pipe measured, profiles/r1_final_ncu_full_conv_summary.txt).  One M=128,N=256 instruction does the same
// FLOPs from 4 + 8 = 12 KB per 128 cycles = 96 B/clk (the operand economy of the BN=256 tiles, DESIGN.md section 3 fact 2,
// for layers that only have 128 or 384 output channels).  L2 -> SM traffic per FLOP is unchanged.
//
// The accumulator is therefore channel-major: TMEM lane = output channel, column = pixel.  The epilogue thread owns ONE
// channel (scale / shift are per-thread constants) and walks the 128 pixels of its tile through the same staged tiles as
// conv_igemm_kernel's TMA epilogue — [128 pixels x 64 channels], SWIZZLE_128B, residual in by TMA, combined in place,
// result out by one TMA tensor store (which also clips ragged tiles) — only with the roles of lane and loop swapped: a
// warp's 32 lanes touch 32 consecutive channels of one pixel (64 contiguous bytes, conflict-free).  ~6-8 instructions
// per output element; a first version with per-element global accesses and a pixel->address table was 4x slower than
// the main loop (issue- and latency-bound), profiles/r2_swap_v1_launch_profile.txt.
//
// Reuses ConvParams unchanged (tmA box = one 128-pixel tile, tmB box = [128 x 64] slab, tmO / tmR boxes = 64 channels of
// one pixel tile).  16-bit outputs only (no fp32 / split-operand planes, no fused head): the host selects it accordingly.
#pragma once

#include "conv_tcgen05.cuh"

namespace w2l {

struct SwapCfg {
    static constexpr int BK = 64;
    static constexpr int kATile = kTileM * BK * 2;          // one 128-pixel box: 16 KB
    static constexpr int kPixBytes = 2 * kATile;            // the unit's two boxes = the 256-row B operand
    static constexpr int kWBytes = 128 * BK * 2;            // weight slab = the A operand
    static constexpr int kStageBytes = kPixBytes + kWBytes; // 48 KB
    static constexpr int kStages = 3;
    static constexpr int kStgTile = kTileM * 64 * 2;        // [128 pixels x 64 channels] 16-bit, SWIZZLE_128B
    static constexpr int kStgBytes = 4 * kStgTile;          // (pixel tile g, channel half h)
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + kSmemExtra;
    static constexpr int kThreads = 384;
    static_assert(kSmemBytes <= kSmemMax, "shared memory budget");
};

template <bool kBF16>
__device__ __forceinline__ float cvt16(uint16_t u) {
    if constexpr (kBF16) return __uint_as_float(static_cast<uint32_t>(u) << 16);
    else return __half2float(__ushort_as_half(u));
}
// 32 pixels of one channel: v = accumulator columns, rowp = the staged tile at the first of the 32 pixel rows, offk[r & 7] =
// this thread's (swizzled) byte offset inside a pixel row.  act(f) = max(f, slope * f): slope 0 = ReLU, 0.01 = LeakyReLU,
// 1 = none.  `mag` collects the largest stored fp16 magnitude bits (range guard, checked once per tile by the caller).
template <bool kBF16, bool kRes>
__device__ __forceinline__ void swap_epi_chunk(const uint32_t (&v)[32], uint8_t* rowp, const int (&offk)[8], float sc, float sh,
                                               float slope, uint32_t& mag) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint16_t* const ep = reinterpret_cast<uint16_t*>(rowp + j * 128 + offk[j & 7]);
        float f = fmaf(__uint_as_float(v[j]), sc, sh);
        if constexpr (kRes) f += cvt16<kBF16>(*ep);
        f = fmaxf(f, slope * f);
        if constexpr (kBF16) {
            *ep = __bfloat16_as_ushort(__float2bfloat16_rn(f));
        } else {
            const uint16_t u = __half_as_ushort(__float2half_rn(f));
            mag = max(mag, static_cast<uint32_t>(u & 0x7FFFu));
            *ep = u;
        }
    }
}

template <bool kBF16>
__global__ void __launch_bounds__(SwapCfg::kThreads, 1) conv_swap_kernel(const __grid_constant__ ConvParams p) {
    pdl_launch_dependents();
    using Cfg = SwapCfg;
    constexpr int kStages = Cfg::kStages;
    constexpr int BK = Cfg::BK;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;
    uint8_t* const smem_al = smem_raw + (smem_base - smem_raw_u32);
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
    const int rows_valid = p.bw * p.bh * p.bn;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmO);
        tma_prefetch_desc(&p.tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);  // one arrive per epilogue warp
        }
        for (int g = 0; g < 2; ++g) mbar_init(res_bar(g), 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    if (warp >= 4 && rows_valid < kTileM) {
        // pixel rows a smaller-than-128 box never writes: zero them once in every stage, so that the dead GEMM columns
        // compute 0 (never a stale NaN bit pattern that would trip the fp16 range guard)
        const int t = threadIdx.x - 128;
        const int dead16 = (kTileM - rows_valid) * 8;           // 16-byte chunks per tile
        for (int i = t; i < kStages * 2 * dead16; i += 256) {
            const int tile_i = i / dead16, c = i % dead16;
            uint8_t* dst = smem_al + (tile_i >> 1) * Cfg::kStageBytes + (tile_i & 1) * Cfg::kATile + rows_valid * 128 + c * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    pdl_wait();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int m_units = (m_tiles + 1) / 2;        // a unit = two consecutive pixel tiles x one 128-channel tile (a tile index
    const int total_tiles = m_units * p.n_tiles;  // past the end decodes to n >= N: its loads are zero-filled, its stores clipped)
    const int k_steps = p.ntaps * p.kc_per_tap;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int mu = tile / p.n_tiles;
                int x_in0[2], y_in0[2], n0[2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int m = mu * 2 + mt;
                    x_in0[mt] = (m % p.tiles_x) * p.bw * p.sx;
                    y_in0[mt] = ((m / p.tiles_x) % p.tiles_y) * p.bh * p.sy;
                    n0[mt] = (m / (p.tiles_x * p.tiles_y)) * p.bn;
                }
                for (int t = 0; t < p.ntaps; ++t) {
                    for (int kc = 0; kc < p.kc_per_tap; ++kc) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t pix_dst = smem_base + stage * Cfg::kStageBytes;
                        mbar_arrive_expect_tx(full_bar(stage), p.stage_tx_bytes);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            tma_load_4d(pix_dst + mt * Cfg::kATile, &p.tmA, full_bar(stage), kc * BK, x_in0[mt] + p.dx[t],
                                        y_in0[mt] + p.dy[t], n0[mt]);
                        tma_load_3d(pix_dst + Cfg::kPixBytes, &p.tmB, full_bar(stage), kc * BK, nt * 128, p.b_slab[t]);
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc<256, kBF16>();   // M = 128 (channels), N = 256 (pixels), both operands K-major
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * 256;
            for (int ks = 0; ks < k_steps; ++ks) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t pix = smem_base + stage * Cfg::kStageBytes;
                    const uint64_t wdesc = make_kmajor_desc<BK>(pix + Cfg::kPixBytes);
                    const uint64_t pdesc = make_kmajor_desc<BK>(pix);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma_f16(tmem_d, wdesc + 2u * k, pdesc + 2u * k, idesc, (ks | k) != 0 ? 1u : 0u);
                    tc_commit(empty_bar(stage));
                    if (ks == k_steps - 1) tc_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue (channel-major accumulator) ===============================
        const int q = (warp - 4) & 3;      // TMEM lane quarter = 32 output channels
        const int g = (warp - 4) >> 2;     // which pixel tile of the unit (columns [128 g, 128 g + 128))
        const int h = q >> 1;              // which 64-channel half (= which staged tile / TMA box) this warp's channels are in
        const EpiParams& e = p.ep;
        const bool has_res = e.res != nullptr;
        const bool leader = (q == 0 && lane == 0);
        const uint32_t bar_id = 1 + g;
        const uint32_t stg_g = stg_base + g * 2 * Cfg::kStgTile;                   // two tiles: h = 0, 1
        uint8_t* const stg_ptr = smem_al + (stg_g - smem_base) + h * Cfg::kStgTile;
        // element (pixel r, channel c of the half) lives at r*128 + ((c/8) ^ (r & 7))*16 + (c % 8)*2; c = (q & 1)*32 + lane
        const int chunk = (q & 1) * 4 + (lane >> 3);
        int offk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) offk[k] = ((chunk ^ k) << 4) + (lane & 7) * 2;
        auto tile_origin = [&](int tile_, int* nt_, int* x0_, int* y0_, int* n0_) {
            *nt_ = tile_ % p.n_tiles;
            const int m_ = (tile_ / p.n_tiles) * 2 + g;
            *x0_ = (m_ % p.tiles_x) * p.bw;
            *y0_ = ((m_ / p.tiles_x) % p.tiles_y) * p.bh;
            *n0_ = (m_ / (p.tiles_x * p.tiles_y)) * p.bn;
        };
        const float slope = e.act == ACT_RELU ? 0.0f : (e.act == ACT_LRELU ? 0.01f : 1.0f);
        uint32_t mag = 0;
        uint32_t rphase = 0;
        if (has_res && leader && static_cast<int>(blockIdx.x) < total_tiles) {
            int nt0, x0, y0, n0;
            tile_origin(blockIdx.x, &nt0, &x0, &y0, &n0);
            mbar_arrive_expect_tx(res_bar(g), 2 * p.epi_box_bytes);
            tma_load_4d(stg_g, &p.tmR, res_bar(g), nt0 * 128, x0, y0, n0);
            tma_load_4d(stg_g + Cfg::kStgTile, &p.tmR, res_bar(g), nt0 * 128 + 64, x0, y0, n0);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            int nt, x0, y0, n0;
            tile_origin(tile, &nt, &x0, &y0, &n0);
            const int ch = nt * 128 + q * 32 + lane;
            const float sc = __ldg(e.scale + ch), sh = __ldg(e.shift + ch);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (has_res) {
                mbar_wait(res_bar(g), rphase);
                rphase ^= 1u;
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + g * 128;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32];
                tmem_ld16(taddr + c0, v);
                tmem_ld16(taddr + c0 + 16, v + 16);
                tmem_ld_wait();
                if (c0 == 96) {  // accumulator fully read: release the TMEM stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                uint8_t* const rowp = stg_ptr + c0 * 128;
                if (has_res) swap_epi_chunk<kBF16, true>(v, rowp, offk, sc, sh, slope, mag);
                else swap_epi_chunk<kBF16, false>(v, rowp, offk, sc, sh, slope, mag);
            }
            if (!kBF16 && mag >= 0x7C00u) g_f16_overflow = 1;   // inf / NaN after rounding: the fp16 range guard
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (leader) {
                const int cg = nt * 128;
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&p.tmO)), "r"(stg_g), "r"(cg), "r"(x0), "r"(y0), "r"(n0) : "memory");
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&p.tmO)), "r"(stg_g + Cfg::kStgTile), "r"(cg + 64), "r"(x0), "r"(y0), "r"(n0) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stores have read the tiles
                if (has_res) {  // fetch the next unit's residual into the (now free) tiles
                    const int ntile = tile + static_cast<int>(gridDim.x);
                    if (ntile < total_tiles) {
                        int nnt, nx0, ny0, nn0;
                        tile_origin(ntile, &nnt, &nx0, &ny0, &nn0);
                        mbar_arrive_expect_tx(res_bar(g), 2 * p.epi_box_bytes);
                        tma_load_4d(stg_g, &p.tmR, res_bar(g), nnt * 128, nx0, ny0, nn0);
                        tma_load_4d(stg_g + Cfg::kStgTile, &p.tmR, res_bar(g), nnt * 128 + 64, nx0, ny0, nn0);
                    }
                }
            }
            // nobody touches the tiles again before the leader is past wait_group.read
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace w2l
