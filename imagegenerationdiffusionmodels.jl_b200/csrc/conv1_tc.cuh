// First convolution of the sampler on tensor cores (sm_100a).
//
// down1.conv1 = Conv((3,3), 1+128 => 64, pad=1) on cat(x, tile(t_emb)) (/root/reference/src/train_brain.jl:111,164-168)
// with the embedding channels folded into per-border-class constants (kernels.cuh, conv1_kernel).  What is left is a
// K = 9 contraction per output -- far too thin for an implicit GEMM on FP16 inputs alone, and the CUDA-core kernel
// spends ~125 instructions per (pixel, 4 channels).  Here the contraction runs on tcgen05 with both FP32 operands
// split into two BF16 parts (BF16 keeps the FP32 exponent range, so the low parts never go subnormal -- an FP16
// split lost the low weight parts that way and was only FP16-accurate on B200):
//     x*w  ~=  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo          (16 significand bits per operand; the dropped
//                                                           x_lo*w_lo term and the residuals are ~2^-17 relative,
//                                                           60x below the FP16 rounding of the stored activation)
// so one output row is a K = 27 (padded to 32) dot product:  A row = [x_hi(9) | x_lo(9) | x_hi(9) | 0(5)],
// B row(co) = [w_hi(9) | w_hi(9) | w_lo(9) | 0(5)], FP32 accumulation in TMEM, w = scale[co] * Wimg[tap][co].
// Four builder warps write the A tile (128 positions x 64 B used of a 128-byte swizzled row) straight into shared
// memory from the FP32 image -- no im2col buffer in HBM --, one thread issues two M128 N64 K16 MMAs per tile, and
// two sets of four epilogue warps add the (timestep, border class) constant and the BatchNorm shift in FP32, apply
// ReLU and write the tile, halo rows as zeros, with a TMA store.
// Used when one timestep is shared by the whole batch (the reverse-diffusion loop, generate_images.jl:196-208);
// per-image timesteps and training keep the CUDA-core kernel.
#pragma once
#include "conv_tc.cuh"

namespace ddpm {
namespace tc {

constexpr int C1_THREADS = 32 * 13;   // warps 0..3 A-tile builders, warp 4 MMA issuer / TMEM owner, warps 5..8 and 9..12 epilogue sets
constexpr int C1_STAGES = 4;

struct C1Params {
    const float* x;        // [N][32][32] FP32 sample
    const float* Wimg;     // [9][64] image-channel weights, tap-major
    const float* Ecls_t;   // [9][64] embedding constants of this timestep per border class
    const float* scale;    // [64] inference BatchNorm scale (nullptr = 1)
    const float* shift;    // [64] inference BatchNorm shift / bias (nullptr = 0)
    int relu;
    Geo g;                 // 32x32 output geometry
    int num_tiles;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {      // both values are already BF16-representable
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split2(float v, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(v));
    lo = __bfloat162float(__float2bfloat16_rn(v - hi));
}

template <typename TOut>
__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmO, const C1Params p) {
    constexpr int WP = WP_32, HS = 33, H = 32, W = 32;
    constexpr int ACC_BUFS = 4, NOUT = 64;
    constexpr uint32_t A_STAGE_BYTES = TC_BM * 128;
    constexpr uint32_t B_BYTES = NOUT * 128;
    constexpr uint32_t O_TILE = TC_BM * NOUT * 2;
    constexpr uint32_t IDESC = make_idesc(1u, TC_BM, NOUT);          // BF16 operands

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_b = smem_u32(smem);
    const uint32_t s_a = s_b + B_BYTES;
    const uint32_t s_o = s_a + C1_STAGES * A_STAGE_BYTES;
    const uint32_t s_bar = s_o + 2 * O_TILE;
    auto bar_afull = [&](int s) { return s_bar + 8u * s; };
    auto bar_aempty = [&](int s) { return s_bar + 8u * (C1_STAGES + s); };
    auto bar_accfull = [&](int b) { return s_bar + 8u * (2 * C1_STAGES + b); };
    auto bar_accempty = [&](int b) { return s_bar + 8u * (2 * C1_STAGES + ACC_BUFS + b); };
    uint8_t* misc = smem + B_BYTES + C1_STAGES * A_STAGE_BYTES + 2 * O_TILE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (2 * C1_STAGES + 2 * ACC_BUFS));
    float* s_E = reinterpret_cast<float*>(misc + 256);                // [9][64]: Ecls*scale + shift per border class

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmO);
        for (int s = 0; s < C1_STAGES; ++s) { mbar_init(bar_afull(s), 4); mbar_init(bar_aempty(s), 1); }
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<256>(smem_u32(tmem_slot));
    pdl_launch_dependents();
    pdl_wait();                               // nothing above reads global memory (programmatic dependent launch)
    for (int i = threadIdx.x; i < 9 * 64; i += C1_THREADS) {
        const int c = i & 63;
        s_E[i] = p.Ecls_t[i] * (p.scale ? p.scale[c] : 1.f) + (p.shift ? p.shift[c] : 0.f);
    }
    if (threadIdx.x < NOUT) {
        // B row of output channel c: [w_hi | w_hi | w_lo | 0], 16-byte chunks XOR-swizzled by the row index
        const int c = threadIdx.x;
        const float sc = p.scale ? p.scale[c] : 1.f;
        float hi[9], lo[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            split2(p.Wimg[k * 64 + c] * sc, hi[k], lo[k]);
        }
        const float col[32] = {hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7], hi[8], hi[0], hi[1], hi[2], hi[3],
                               hi[4], hi[5], hi[6], hi[7], hi[8], lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7],
                               lo[8], 0.f,   0.f,   0.f,   0.f,   0.f};
        const uint32_t rbase = s_b + (uint32_t)c * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((uint32_t)j ^ ((uint32_t)c & 7u)) << 4)),
                         "r"(pack_h2(col[8 * j + 0], col[8 * j + 1])), "r"(pack_h2(col[8 * j + 2], col[8 * j + 3])),
                         "r"(pack_h2(col[8 * j + 4], col[8 * j + 5])), "r"(pack_h2(col[8 * j + 6], col[8 * j + 7]))
                         : "memory");
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int npos = (int)p.g.npos;

    if (warp < 4) {
        // ================= A-tile builders: thread r writes row r of every tile this CTA owns =================
        const int r = threadIdx.x;
        // The nine taps of output row r are nine entries of ONE window of the zero-padded position space:
        //   tap(dy,dx) = Xp[r + (WP+1) + dy*WP + dx],   Xp[j] = padded image value at position tile*128 - (WP+1) + j, j < 128 + 2*(WP+1).
        // The first version fetched them with 9 bounds-checked global loads per thread (1152 per tile); ncu's source
        // view showed the builders latency-bound on exactly those loads with every other warp idle on its barrier.
        // Now the 128 builder threads fetch the 198 window entries once (<= 2 loads per thread, two tiles ahead),
        // publish them in a double-buffered shared-memory window and read their taps from there.
        float* s_xp = s_E + 9 * 64;                               // [2][208] window buffers
        auto load_window = [&](int tile, float v[2]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int jdx = r + q * TC_BM;
                const int pos = tile * TC_BM - (WP + 1) + jdx;
                const int pr = pos / WP, pc = pos - pr * WP;
                const int n = pr / HS, prr = pr - n * HS;
                const bool in = jdx < TC_BM + 2 * (WP + 1) && tile < p.num_tiles && pos >= 0 && pos < npos && prr != 0 && pc >= 1 &&
                                pc <= W && n < p.g.N;
                v[q] = in ? __ldg(p.x + ((long long)n * H + (prr - 1)) * W + (pc - 1)) : 0.f;
            }
        };
        auto load_taps_from_window = [&](const float w2[2], int it, float v[9]) {
            float* xp = s_xp + (it & 1) * 208;
            xp[r] = w2[0];
            if (r + TC_BM < 208) xp[r + TC_BM] = w2[1];
            named_bar_sync(4, 128);                                // the four builder warps only
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = xp[r + (WP + 1) + (k / 3 - 1) * WP + (k % 3 - 1)];
        };
        float ws[3][2];
        load_window(blockIdx.x, ws[0]);
        load_window(blockIdx.x + gridDim.x, ws[1]);
        auto pack_tile = [&](const float v[9], int it) {
            const int stage = it % C1_STAGES;
            const uint32_t phase = (uint32_t)(it / C1_STAGES) & 1u;
            float hi[9], lo[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) split2(v[k], hi[k], lo[k]);
            const float col[32] = {hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7], hi[8], lo[0], lo[1], lo[2], lo[3],
                                   lo[4], lo[5], lo[6], lo[7], lo[8], hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7],
                                   hi[8], 0.f,   0.f,   0.f,   0.f,   0.f};
            mbar_wait(bar_aempty(stage), phase ^ 1u);
            const uint32_t rbase = s_a + (uint32_t)stage * A_STAGE_BYTES + (uint32_t)r * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4)),
                             "r"(pack_h2(col[8 * j + 0], col[8 * j + 1])), "r"(pack_h2(col[8 * j + 2], col[8 * j + 3])),
                             "r"(pack_h2(col[8 * j + 4], col[8 * j + 5])), "r"(pack_h2(col[8 * j + 6], col[8 * j + 7]))
                             : "memory");
            }
            fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull(stage));
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles;) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (tile < p.num_tiles) {
                    load_window(tile + 2 * gridDim.x, ws[(u + 2) % 3]);     // in flight while this tile and the next are packed
                    float v[9];
                    load_taps_from_window(ws[u], it, v);
                    pack_tile(v, it);
                    tile += gridDim.x; ++it;
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo_base = ((s_a & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo_base = ((s_b & 0x3FFFFu) >> 4) | (1u << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int stage = it % C1_STAGES, buf = it % ACC_BUFS;
            const uint32_t phase = (uint32_t)(it / C1_STAGES) & 1u, acc_phase = (uint32_t)(it / ACC_BUFS) & 1u;
            mbar_wait(bar_accempty(buf), acc_phase ^ 1u);
            mbar_wait(bar_afull(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = a_lo_base + stage * (A_STAGE_BYTES >> 4);
                const uint32_t d_tmem = tmem_base + buf * NOUT;
                umma_f16_lh(d_tmem, a_lo, b_lo_base, DESC_HI, IDESC, 0u);
                umma_f16_lh(d_tmem, a_lo + (32 >> 4), b_lo_base + (32 >> 4), DESC_HI, IDESC, 1u);
                umma_commit(bar_aempty(stage));
                umma_commit(bar_accfull(buf));
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue sets (alternate tiles): TMEM -> +E[cls] -> ReLU -> staging tile -> TMA store =================
        const int eset = (warp >= 9) ? 1 : 0;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const bool store_thread = (threadIdx.x == (eset ? 9 * 32 : 5 * 32));
        const uint32_t stage_o = s_o + (uint32_t)eset * O_TILE;
        for (int it = eset, tile = blockIdx.x + eset * gridDim.x; tile < p.num_tiles; it += 2, tile += 2 * gridDim.x) {
            const int buf = it % ACC_BUFS;
            const uint32_t acc_phase = (uint32_t)(it / ACC_BUFS) & 1u;
            const int pos = tile * TC_BM + row;
            const int pr = pos / WP, pc = pos - pr * WP;
            const int n = pr / HS, prr = pr - n * HS;
            const int h = prr - 1, w = pc - 1;
            const bool valid = pos < npos && prr != 0 && pc >= 1 && pc <= W && n < p.g.N;
            const int cls = (h == 0 ? 0 : (h == H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
            const float* erow = s_E + (valid ? cls : 4) * 64;
            mbar_wait(bar_accfull(buf), acc_phase);
            tc_fence_after();
            if (store_thread) tma_store_wait_read<0>();       // this set's previous TMA store has drained the staging tile
            named_bar_sync(1 + eset, 128);
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * NOUT;
            uint32_t racc[4][16];
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_ld16(taddr + q * 16, racc[q]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty(buf));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 e = *reinterpret_cast<const float4*>(erow + q * 16 + 4 * j4);
                    v[4 * j4 + 0] = __uint_as_float(racc[q][4 * j4 + 0]) + e.x;
                    v[4 * j4 + 1] = __uint_as_float(racc[q][4 * j4 + 1]) + e.y;
                    v[4 * j4 + 2] = __uint_as_float(racc[q][4 * j4 + 2]) + e.z;
                    v[4 * j4 + 3] = __uint_as_float(racc[q][4 * j4 + 3]) + e.w;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = valid ? (p.relu ? fmaxf(v[j], 0.f) : v[j]) : 0.f;
                uint4 qa, qb;
                pack16<TOut>(v, qa, qb);
                const uint32_t rbase = stage_o + (uint32_t)row * 128;
                const uint32_t ch = (uint32_t)(q * 2);
                const uint32_t sw = (uint32_t)row & 7u;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((ch ^ sw) << 4)), "r"(qa.x), "r"(qa.y),
                             "r"(qa.z), "r"(qa.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((ch + 1) ^ sw) << 4)), "r"(qb.x), "r"(qb.y),
                             "r"(qb.z), "r"(qb.w) : "memory");
            }
            fence_proxy_async();
            named_bar_sync(1 + eset, 128);
            if (store_thread) {
                tma_store_2d(&tmO, stage_o, 0, tile * TC_BM + p.g.guard);
                tma_store_commit();
            }
        }
        if (store_thread) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Second version (default, option conv1_tc = 2): the (timestep, border class) constant rides on the contraction too.
//
// ncu (profiles/r2/conv1_tc_r2_ncu.txt) showed the kernel above latency-bound, not write-bound: DRAM at 24 % of peak,
// 0.44 instructions per cycle and scheduler, 1,630 cycles per 128-position tile -- the four builder warps (one tile at
// a time, ~200 dependent instructions each) and the epilogue (position decode, 64 FP32 adds from a per-class table, 64
// max, 64 selects: ~520 instructions per warp and tile) were each a serial chain.  Here
//   * the A row also carries a one-hot of its border class three times and the B row the three BF16 parts of
//     E[cls][co]*scale + shift (24 significand bits = the FP32 value exactly): K = 18 + 9 + 27 = 54 -> 64, four K=16
//     MMAs (the tensor pipe is idle anyway).  Rows of halo positions are all-zero, so the accumulator already IS the
//     pre-activation (exactly 0 for halo rows) and the epilogue is TMEM load -> convert -> packed ReLU -> staging
//     tile -> TMA store, without any position arithmetic;
//   * the window entries are split into (hi, lo) BF16 pairs ONCE, by the thread that loads them, together with the
//     entry's border class; a row's K columns 0..17 are then its nine window words as they are, columns 18..26 five
//     byte permutes, the rest a 64-byte table row picked by the class;
//   * two builder sets (warps 0..3 / 4..7) work on alternate tiles.
// K layout of an A row:  [ (x_hi, x_lo) x 9 taps | x_hi x 9 | onehot(cls) x 3 | 0 x 10 ]
//          of a B row:   [ (w_hi, w_hi) x 9      | w_lo x 9 | E_hi(9 cls), E_mid(9), E_lo(9) | 0 x 10 ]
constexpr int C1F_THREADS = 32 * 17;  // warps 0..3 / 4..7 builder sets, warp 8 MMA issuer / TMEM owner, warps 9..12 / 13..16 epilogue sets

// max(v, 0) on eight packed 16-bit values
template <typename T>
__device__ __forceinline__ void relu_packed(uint4& q) {
    using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
    T2* h = reinterpret_cast<T2*>(&q);
    uint32_t zero_bits = 0u;
    const T2 z = *reinterpret_cast<T2*>(&zero_bits);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

template <typename TOut>
__global__ void __launch_bounds__(C1F_THREADS, 1)
conv1f_tc_kernel(const __grid_constant__ CUtensorMap tmO, const C1Params p) {
    constexpr int WP = WP_32, HS = 33, H = 32, W = 32;
    constexpr int ACC_BUFS = 4, NOUT = 64, WIN = 208;                  // window: 128 + 2*(WP+1) = 196 entries, padded
    constexpr uint32_t A_STAGE_BYTES = TC_BM * 128;
    constexpr uint32_t B_BYTES = NOUT * 128;
    constexpr uint32_t O_TILE = TC_BM * NOUT * 2;
    constexpr uint32_t IDESC = make_idesc(1u, TC_BM, NOUT);          // BF16 operands

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_b = smem_u32(smem);
    const uint32_t s_a = s_b + B_BYTES;
    const uint32_t s_o = s_a + C1_STAGES * A_STAGE_BYTES;
    const uint32_t s_bar = s_o + 2 * O_TILE;
    auto bar_afull = [&](int s) { return s_bar + 8u * s; };
    auto bar_aempty = [&](int s) { return s_bar + 8u * (C1_STAGES + s); };
    auto bar_accfull = [&](int b) { return s_bar + 8u * (2 * C1_STAGES + b); };
    auto bar_accempty = [&](int b) { return s_bar + 8u * (2 * C1_STAGES + ACC_BUFS + b); };
    uint8_t* misc = smem + B_BYTES + C1_STAGES * A_STAGE_BYTES + 2 * O_TILE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (2 * C1_STAGES + 2 * ACC_BUFS));
    // misc + 256: one-hot table [10][16 words]; + 1024: zero words [128]; + 1536: windows [2 sets][2 buffers][WIN] words;
    // + 1536 + 3328: border-class codes [2][2][WIN] bytes
    const uint32_t s_T = s_bar + 256, s_zero = s_bar + 1024, s_win = s_bar + 1536, s_code = s_win + 2 * 2 * WIN * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmO);
        for (int s = 0; s < C1_STAGES; ++s) { mbar_init(bar_afull(s), 4); mbar_init(bar_aempty(s), 1); }
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), 4); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<256>(smem_u32(tmem_slot));
    // one-hot table: row cls (9 = halo / out of range: all zero) holds the 32-bit words 12..27 (K columns 24..55) of an A row
    // with the image parts left out: 1.0 (BF16 0x3F80) at columns 27+cls, 36+cls and 45+cls
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 160) {
        const int i = threadIdx.x - 128, cls = i >> 4, w = 12 + (i & 15);
        uint32_t word = 0;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            const int col = 2 * w + hlf;
            if (cls < 9 && (col == 27 + cls || col == 36 + cls || col == 45 + cls)) word |= 0x3F80u << (16 * hlf);
        }
        sts32(s_T + 4u * i, word);
    }
    if (threadIdx.x >= 320 && threadIdx.x < 320 + 128) sts32(s_zero + 4u * (threadIdx.x - 320), 0u);
    // A columns 56..63 (16-byte chunk 7) are never written by the builders: zero them once in every stage
    for (int i = threadIdx.x; i < C1_STAGES * TC_BM; i += C1F_THREADS) {
        const uint32_t stg = (uint32_t)i / TC_BM, rr = (uint32_t)i % TC_BM;
        sts128(s_a + stg * A_STAGE_BYTES + rr * 128 + ((7u ^ (rr & 7u)) << 4), 0u, 0u, 0u, 0u);
    }
    pdl_launch_dependents();
    pdl_wait();                               // nothing above reads global memory (programmatic dependent launch)
    if (threadIdx.x < NOUT) {
        const int c = threadIdx.x;
        const float sc = p.scale ? p.scale[c] : 1.f, sh = p.shift ? p.shift[c] : 0.f;
        float hi[9], lo[9], e0[9], e1[9], e2[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            split2(p.Wimg[k * 64 + c] * sc, hi[k], lo[k]);
            const float ev = p.Ecls_t[k * 64 + c] * sc + sh;          // the FP32 value the first version adds in its epilogue
            e0[k] = __bfloat162float(__float2bfloat16_rn(ev));
            const float r1 = ev - e0[k];                               // exact
            e1[k] = __bfloat162float(__float2bfloat16_rn(r1));
            e2[k] = __bfloat162float(__float2bfloat16_rn(r1 - e1[k])); // three parts = 24 significand bits: exact
        }
        auto colv = [&](int k) -> float {
            return k < 18 ? hi[k >> 1] : k < 27 ? lo[k - 18] : k < 36 ? e0[k - 27] : k < 45 ? e1[k - 36] : k < 54 ? e2[k - 45] : 0.f;
        };
        const uint32_t rbase = s_b + (uint32_t)c * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            sts128(rbase + (((uint32_t)j ^ ((uint32_t)c & 7u)) << 4), pack_h2(colv(8 * j + 0), colv(8 * j + 1)),
                   pack_h2(colv(8 * j + 2), colv(8 * j + 3)), pack_h2(colv(8 * j + 4), colv(8 * j + 5)),
                   pack_h2(colv(8 * j + 6), colv(8 * j + 7)));
    }
    fence_proxy_async();                      // B rows, zeroed A chunks: generic-proxy writes -> tensor core reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int npos = (int)p.g.npos;

    if (warp < 8) {
        // ================= A-tile builders: thread r of set b writes row r of tiles b, b+2, ... of this CTA =================
        const int set = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t sw = (uint32_t)r & 7u;
        // The nine taps of output row r are nine entries of ONE window of the zero-padded position space:
        //   tap(dy,dx) = Xp[r + (WP+1) + dy*WP + dx],   Xp[j] = padded image value at position tile*128 - (WP+1) + j, j < 128 + 2*(WP+1);
        // entry r + 35 is the row's own position.  The 128 threads of a set fetch the 198 entries once (<= 2 loads per
        // thread, two of the set's tiles ahead), split them and publish (hi | lo << 16, border-class code).
        auto load_window = [&](int tile, float v[2], int code[2]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int jdx = r + q * TC_BM;
                const int pos = tile * TC_BM - (WP + 1) + jdx;
                const unsigned upos = (unsigned)max(pos, 0);
                const unsigned pr = upos / WP, pc = upos - pr * WP;
                const unsigned n = pr / HS, prr = pr - n * HS;
                const bool in = jdx < TC_BM + 2 * (WP + 1) && tile < p.num_tiles && pos >= 0 && pos < npos && prr != 0 && pc >= 1 &&
                                pc <= W && (int)n < p.g.N;
                v[q] = in ? __ldg(p.x + ((long long)n * H + (prr - 1)) * W + (pc - 1)) : 0.f;
                code[q] = in ? (int)((prr == 1 ? 0 : (prr == H ? 2 : 1)) * 3 + (pc == 1 ? 0 : (pc == W ? 2 : 1))) : 9;
            }
        };
        auto publish = [&](const float v[2], const int code[2], int its) {
            const uint32_t wb = s_win + (uint32_t)((set * 2 + (its & 1)) * WIN) * 4u, cb = s_code + (uint32_t)((set * 2 + (its & 1)) * WIN);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 0 || r + TC_BM < WIN) {
                    float hi, lo;
                    split2(v[q], hi, lo);
                    sts32(wb + 4u * (uint32_t)(r + q * TC_BM), pack_h2(hi, lo));
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(cb + (uint32_t)(r + q * TC_BM)), "r"(code[q]) : "memory");
                }
            }
        };
        auto build_row = [&](int its) {
            const int it = 2 * its + set;
            const int stage = it % C1_STAGES;
            const uint32_t phase = (uint32_t)(it / C1_STAGES) & 1u;
            const uint32_t wb = s_win + (uint32_t)((set * 2 + (its & 1)) * WIN) * 4u, cb = s_code + (uint32_t)((set * 2 + (its & 1)) * WIN);
            named_bar_sync(4 + set, 128);                          // the four warps of this set only
            uint32_t code;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(code) : "r"(cb + (uint32_t)(r + WP + 1)));
            // halo / out-of-range rows read their taps from a block of zeros: the whole A row becomes zero
            const uint32_t tb = (code != 9u) ? wb + 4u * (uint32_t)r : s_zero;
            uint32_t t[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) t[k] = lds32(tb + 4u * (uint32_t)((WP + 1) + (k / 3 - 1) * WP + (k % 3 - 1)));
            const uint4 t3 = lds128(s_T + code * 64u), t4 = lds128(s_T + code * 64u + 16u), t5 = lds128(s_T + code * 64u + 32u),
                        t6 = lds128(s_T + code * 64u + 48u);
            const uint32_t h01 = __byte_perm(t[0], t[1], 0x5410), h23 = __byte_perm(t[2], t[3], 0x5410),
                           h45 = __byte_perm(t[4], t[5], 0x5410), h67 = __byte_perm(t[6], t[7], 0x5410);
            mbar_wait(bar_aempty(stage), phase ^ 1u);
            const uint32_t rbase = s_a + (uint32_t)stage * A_STAGE_BYTES + (uint32_t)r * 128;
            sts128(rbase + ((0u ^ sw) << 4), t[0], t[1], t[2], t[3]);
            sts128(rbase + ((1u ^ sw) << 4), t[4], t[5], t[6], t[7]);
            sts128(rbase + ((2u ^ sw) << 4), t[8], h01, h23, h45);
            sts128(rbase + ((3u ^ sw) << 4), h67, (t[8] & 0xFFFFu) | t3.y, t3.z, t3.w);
            sts128(rbase + ((4u ^ sw) << 4), t4.x, t4.y, t4.z, t4.w);
            sts128(rbase + ((5u ^ sw) << 4), t5.x, t5.y, t5.z, t5.w);
            sts128(rbase + ((6u ^ sw) << 4), t6.x, t6.y, t6.z, t6.w);
            fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull(stage));
        };
        const int tstep = 2 * (int)gridDim.x;                      // tile stride of one builder set
        float wv[3][2];
        int wc[3][2];
        int tile = blockIdx.x + set * gridDim.x;
        load_window(tile, wv[0], wc[0]);
        load_window(tile + tstep, wv[1], wc[1]);
        int its = 0;
        while (tile < p.num_tiles) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (tile < p.num_tiles) {
                    load_window(tile + 2 * tstep, wv[(u + 2) % 3], wc[(u + 2) % 3]);     // in flight while two tiles are built
                    publish(wv[u], wc[u], its);
                    build_row(its);
                    tile += tstep; ++its;
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo_base = ((s_a & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo_base = ((s_b & 0x3FFFFu) >> 4) | (1u << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int stage = it % C1_STAGES, buf = it % ACC_BUFS;
            const uint32_t phase = (uint32_t)(it / C1_STAGES) & 1u, acc_phase = (uint32_t)(it / ACC_BUFS) & 1u;
            mbar_wait(bar_accempty(buf), acc_phase ^ 1u);
            mbar_wait(bar_afull(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = a_lo_base + stage * (A_STAGE_BYTES >> 4);
                const uint32_t d_tmem = tmem_base + buf * NOUT;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_f16_lh(d_tmem, a_lo + ((ks * 32) >> 4), b_lo_base + ((ks * 32) >> 4), DESC_HI, IDESC, ks ? 1u : 0u);
                umma_commit(bar_aempty(stage));
                umma_commit(bar_accfull(buf));
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue sets (alternate tiles): TMEM -> convert -> packed ReLU -> staging tile -> TMA store =================
        const int eset = (warp >= 13) ? 1 : 0;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const bool store_thread = (threadIdx.x == (eset ? 13 * 32 : 9 * 32));
        const uint32_t stage_o = s_o + (uint32_t)eset * O_TILE;
        const uint32_t rbase = stage_o + (uint32_t)row * 128;
        const uint32_t sw = (uint32_t)row & 7u;
        for (int it = eset, tile = blockIdx.x + eset * gridDim.x; tile < p.num_tiles; it += 2, tile += 2 * gridDim.x) {
            const int buf = it % ACC_BUFS;
            const uint32_t acc_phase = (uint32_t)(it / ACC_BUFS) & 1u;
            mbar_wait(bar_accfull(buf), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * NOUT;
            uint32_t racc[4][16];
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_ld16(taddr + q * 16, racc[q]);
            if (store_thread) tma_store_wait_read<0>();       // this set's previous TMA store has drained the staging tile
            named_bar_sync(1 + eset, 128);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty(buf));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(racc[q][j]);
                uint4 qa, qb;
                pack16<TOut>(v, qa, qb);
                if (p.relu) { relu_packed<TOut>(qa); relu_packed<TOut>(qb); }      // == rounding max(v, 0): rounding is monotone and keeps 0
                const uint32_t ch = (uint32_t)(q * 2);
                sts128(rbase + ((ch ^ sw) << 4), qa.x, qa.y, qa.z, qa.w);
                sts128(rbase + (((ch + 1) ^ sw) << 4), qb.x, qb.y, qb.z, qb.w);
            }
            fence_proxy_async();
            named_bar_sync(1 + eset, 128);
            if (store_thread) {
                tma_store_2d(&tmO, stage_o, 0, tile * TC_BM + p.g.guard);
                tma_store_commit();
            }
        }
        if (store_thread) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// Launch: a1 = relu((x (*) Wimg + Ecls[t]) * scale + shift) on the padded 32x32 layout, one shared timestep
template <typename TOut>
bool conv1_shared_t(cudaStream_t st, const float* x, const float* Wimg, const float* Ecls_t, const float* scale,
                    const float* shift, int relu, TOut* out, const Geo& g, bool fold = true) {
    if (!available()) return false;
    if constexpr (sizeof(TOut) != 2) {
        return false;
    } else {
    if (g.Wp != WP_32 || g.Hs != 33 || g.npos + 2 * TC_BM >= (1ll << 31)) return false;
    C1Params p{};
    p.x = x; p.Wimg = Wimg; p.Ecls_t = Ecls_t; p.scale = scale; p.shift = shift; p.relu = relu; p.g = g;
    p.num_tiles = cdiv(g.npos, TC_BM);
    CUtensorMap o = make_map_2d<TOut>(out - (size_t)g.guard * 64, (uint64_t)g.alloc_positions(), 64, TC_BM);
    // both versions: B rows, A stages, two staging tiles, then barriers + tables + windows (version 2: 1536 + 4*208*5 bytes)
    constexpr size_t smem = 1024 + 64 * 128 + (size_t)C1_STAGES * TC_BM * 128 + 2 * TC_BM * 64 * 2 + 1536 + 4 * 208 * 5 + 256;
    static_assert(smem >= 1024 + 64 * 128 + (size_t)C1_STAGES * TC_BM * 128 + 2 * TC_BM * 64 * 2 + 256 + 9 * 64 * 4 + 2 * 208 * 4 + 256, "version 1 layout");
    auto kern = fold ? conv1f_tc_kernel<TOut> : conv1_tc_kernel<TOut>;
    ensure_smem_attr(kern, smem);
    int ctas = state().num_sms;
    if (ctas > p.num_tiles) ctas = p.num_tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(fold ? C1F_THREADS : C1_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = state().pdl ? 1 : 0;
    DDPM_CUDA(cudaLaunchKernelEx(&cfg, kern, o, p));
    DDPM_LAUNCH_CHECK();
    return true;
    }
}

}  // namespace tc
}  // namespace ddpm
