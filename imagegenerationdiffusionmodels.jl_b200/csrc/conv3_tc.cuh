// tcgen05 3x3 convolution, second formulation: "kernel-row MMAs with column taps packed into N".
//
// The first formulation (conv_tc.cuh) issues one M128 x N64 x K16 MMA per (tap, K-step): 36 per tile.  Each
// re-reads a 4 KB A window and a 2 KB B tile from shared memory, 113 B/clk -- the SS-mode operand bandwidth
// limit -- so the tensor pipe runs at 53 instead of 32 cycles per instruction (profiles/README.md).
//
// Here the three taps of one kernel row (dx = -1, 0, +1) share ONE A window: with the window of kernel row dy,
//      D[q][ dxi*64 + co ] += sum_ci x[q + dy*Wp][ci] * Wcc[co][dy, dx][ci]           (N = 3*64 = 192)
// is the contribution of input row q to output position q - dx.  One tile therefore needs 3 (dy) x 4 (K-steps)
// = 12 MMAs of N = 192 (96 cycles each: the same 1152 math cycles) but only 12 x (4 KB + 6 KB) = 120 KB of
// operand reads instead of 216 KB: 107 B/clk, below the shared-memory limit.  The epilogue recombines
//      out[p] = D[p-1][0:64] + D[p][64:128] + D[p+1][128:192]
// with two warp shuffles per element (TMEM lanes are rows), a 2 KB shared-memory exchange for the rows at the
// warp boundaries, and tiles that overlap by two rows (126 outputs per 128-row tile).
#pragma once
#include "conv_tc.cuh"

namespace ddpm {
namespace tc {

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int C3_BOUT = 126;      // output rows per tile (D rows 1..126)
constexpr int C3_THREADS = 384;   // warp 0 TMA, warps 1 & 2 MMA issuers, warp 3 TMEM owner, warps 4..11 epilogue

struct C3Params {
    void* out;
    int out_cs;
    Geo g;
    const float* shift;
    int relu;
    int num_tiles;
    int chunk1_src1;
    // EPI == 2: fused final 1x1 conv + reverse-diffusion update (see conv_tc.cuh)
    float* x;
    const float* z;
    const float* wf;
    const float* bf;
    float sig, sqa, sqp, sqv;
    int final_clamp;
    long long* dbg;   // optional per-CTA role cycle counters (see conv_tc.cuh)
};

template <int CHUNKS, int WP, int STAGES, int EPI, int TMAST, typename TIn, typename TOut>
__global__ void __launch_bounds__(C3_THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const C3Params p) {
    constexpr int R = ((TC_BM + 2 * WP + 7) / 8) * 8;          // slab rows: positions base-Wp .. base+128+Wp
    constexpr uint32_t A_STAGE_BYTES = R * 128;
    constexpr uint32_t W_TILE_BYTES = 192 * 128;               // one (dy, chunk) tile [3*64 rows (dx, co)][64 K]
    constexpr uint32_t W_BYTES = 3 * CHUNKS * W_TILE_BYTES;
    constexpr uint32_t O_BYTES = TMAST ? 2u * TC_BM * 128u : 0u;
    constexpr int ACC_BUFS = 2;
    constexpr bool DUAL = (STAGES % (2 * CHUNKS)) == 0;
    constexpr uint32_t IDESC = make_idesc(IsBf16<TIn>::v, TC_BM, 192);
    constexpr int EPI_T0 = 128;                                // first epilogue thread (warp 4)
    constexpr int EPI_THREADS = 256;                           // warps 4..11: two sets of four warps, 32 channels each

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_w = smem_u32(smem);
    const uint32_t s_a = s_w + W_BYTES;
    const uint32_t s_o = s_a + STAGES * A_STAGE_BYTES;
    const uint32_t s_bar = s_o + O_BYTES;
    auto bar_w = [&]() { return s_bar; };
    auto bar_afull = [&](int s) { return s_bar + 8u * (1 + s); };
    auto bar_aempty = [&](int s) { return s_bar + 8u * (1 + STAGES + s); };
    auto bar_accfull = [&](int b) { return s_bar + 8u * (1 + 2 * STAGES + b); };
    auto bar_accempty = [&](int b) { return s_bar + 8u * (1 + 2 * STAGES + ACC_BUFS + b); };
    uint8_t* misc = smem + W_BYTES + STAGES * A_STAGE_BYTES + O_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (1 + 2 * STAGES + 2 * ACC_BUFS));
    float* s_shift = reinterpret_cast<float*>(misc + 256);     // [64]
    float* s_wf = s_shift + 64;                                // [64]
    float* s_dot = s_wf + 64;                                  // [128] partial eps_hat of epilogue set 1 (EPI == 2)
    float* s_xchg = reinterpret_cast<float*>(misc + 2048);     // [2 parity][4 lane groups][2: blk0 of lane 31 | blk2 of lane 0][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.y;
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_kernel0 = p.dbg ? clock64() : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0); prefetch_tmap(&tmW);
        if (p.chunk1_src1) prefetch_tmap(&tmA1);
        if (TMAST) prefetch_tmap(&tmO);
        mbar_init(bar_w(), 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_afull(s), 1); mbar_init(bar_aempty(s), 1); }
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), 8); }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(smem_u32(tmem_slot));
    if (threadIdx.x >= EPI_T0 && threadIdx.x < EPI_T0 + 64) {
        const int c = threadIdx.x - EPI_T0;
        s_shift[c] = p.shift ? p.shift[n_blk * 64 + c] : 0.f;
        if (EPI == 2) s_wf[c] = p.wf[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(bar_w(), W_BYTES);
            for (int d = 0; d < 3; ++d)
                for (int c = 0; c < CHUNKS; ++c)
                    tma_load_2d(s_w + (d * CHUNKS + c) * W_TILE_BYTES, &tmW, (d * CHUNKS + c) * 64, n_blk * 192, bar_w());
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int row0 = tile * C3_BOUT - 1 - WP + p.g.guard;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                long long t0 = p.dbg ? clock64() : 0;
                mbar_wait(bar_aempty(stage), phase ^ 1, 31);
                if (p.dbg) dbg_acc[0] += clock64() - t0;
                if (elect_one()) {
                    mbar_expect_tx(bar_afull(stage), A_STAGE_BYTES);
                    const bool second = (c == 1) && p.chunk1_src1;
                    tma_load_2d(s_a + stage * A_STAGE_BYTES, second ? &tmA1 : &tmA0, second ? 0 : c * 64, row0, bar_afull(stage));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || (warp == 2 && DUAL)) {
        // ================= MMA issuers =================
        // DUAL: warp 1 takes even, warp 2 odd tiles of this CTA so one warp's barrier waits overlap the other's
        // MMAs.  Only legal when every pipeline stage is always consumed by the SAME warp (STAGES a multiple of
        // 2*CHUNKS): an mbarrier parity wait is only meaningful one phase ahead, and a warp that waits on a stage
        // whose previous phase belongs to the other warp may observe the phase before that one as "complete".
        const int parity = (warp == 1) ? 0 : 1;
        constexpr int NISS = DUAL ? 2 : 1;
        mbar_wait(bar_w(), 0, 32);
        tc_fence_after();
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo_base = ((s_a & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo_base = ((s_w & 0x3FFFFu) >> 4) | (1u << 16);
        for (int seq = parity, tile = blockIdx.x + parity * gridDim.x; tile < p.num_tiles; seq += NISS, tile += NISS * gridDim.x) {
            const int buf = seq % ACC_BUFS;
            const uint32_t acc_phase = (uint32_t)(seq / ACC_BUFS) & 1u;
            long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(bar_accempty(buf), acc_phase ^ 1, 33 + parity * 10);
            if (p.dbg) dbg_acc[1] += clock64() - t0;
            const uint32_t d_tmem = tmem_base + buf * 192;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int step = seq * CHUNKS + c;
                const int stage = step % STAGES;
                const uint32_t phase = (uint32_t)(step / STAGES) & 1u;
                t0 = p.dbg ? clock64() : 0;
                mbar_wait(bar_afull(stage), phase, 34 + parity * 10);
                if (p.dbg) dbg_acc[2] += clock64() - t0;
                tc_fence_after();
                t0 = p.dbg ? clock64() : 0;
                if (elect_one()) {
                    const uint32_t a_lo = a_lo_base + stage * (A_STAGE_BYTES >> 4);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // window of kernel row dy = d-1: slab rows [d*Wp, d*Wp + 128)
                            umma_f16_lh(d_tmem, a_lo + ((d * WP * 128 + ks * 32) >> 4),
                                        b_lo_base + (((d * CHUNKS + c) * W_TILE_BYTES + ks * 32) >> 4), DESC_HI, IDESC,
                                        (c | d | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(bar_aempty(stage));
                    if (c == CHUNKS - 1) umma_commit(bar_accfull(buf));
                }
                __syncwarp();
                if (p.dbg) dbg_acc[3] += clock64() - t0;
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: 2 sets x 4 warps; set s owns channels [32s, 32s+32) of every row =================
        const int lane_grp = warp & 3;                        // TMEM lanes 32*lane_grp .. +31
        const int set = (warp - 4) >> 2;
        const int r = lane_grp * 32 + lane;                   // D row == TMEM lane
        const int cset = set * 32;                            // first channel of this set
        int buf = 0, obuf = 0, par = 0;
        uint32_t acc_phase = 0;
        TOut* out = reinterpret_cast<TOut*>(p.out);
        const bool is0 = lane == 0, is31 = lane == 31;
        const int wl = lane_grp > 0 ? lane_grp - 1 : 0, wr = lane_grp < 3 ? lane_grp + 1 : 3;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const long long pos = (long long)tile * C3_BOUT - 1 + r;
            int n_img, hh, ww;
            const bool interior = (r >= 1) && (r <= C3_BOUT);
            const bool valid = p.g.decode(pos, n_img, hh, ww) && interior;
            long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(bar_accfull(buf), acc_phase, 35);
            long long t1 = p.dbg ? clock64() : 0;
            if (p.dbg) dbg_acc[4] += t1 - t0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * 192 + cset;
            const uint32_t xb = smem_u32(s_xchg) + (uint32_t)par * (4 * 2 * 64 * 4);
            uint32_t stage_o = 0;
            if (TMAST) {
                stage_o = s_o + (uint32_t)obuf * (TC_BM * 128);
                if (threadIdx.x == EPI_T0) tma_store_wait_read<1>();
            }
            // ---- phase B: out[p] = D[p-1][blk0] + D[p][blk1] + D[p+1][blk2] for this set's 32 channels
            float dot = 0.f;
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld16(taddr + gi * 16, r0);
                tmem_ld16(taddr + 64 + gi * 16, r1);
                tmem_ld16(taddr + 128 + gi * 16, r2);
                tmem_ld_wait();
                if (gi == 1) {   // last TMEM read of this tile: hand the accumulator back to the MMA warps
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_accempty(buf));
                }
                const int cb = cset + gi * 16;
                // the rows at the warp boundaries publish the 16 values their neighbours in the adjacent warps need
                // (TMEM is read exactly once per tile: its 64 B/clk read port is the epilogue's bottleneck)
                if (is31) {
                    const uint32_t a = xb + (uint32_t)((lane_grp * 2 + 0) * 64 + cb) * 4;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        sts128(a + j * 4, __uint_as_float(r0[j]), __uint_as_float(r0[j + 1]), __uint_as_float(r0[j + 2]),
                               __uint_as_float(r0[j + 3]));
                }
                if (is0) {
                    const uint32_t a = xb + (uint32_t)((lane_grp * 2 + 1) * 64 + cb) * 4;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        sts128(a + j * 4, __uint_as_float(r2[j]), __uint_as_float(r2[j + 1]), __uint_as_float(r2[j + 2]),
                               __uint_as_float(r2[j + 3]));
                }
                named_bar_sync(1, EPI_THREADS);
                const uint32_t a_up = xb + (uint32_t)((wl * 2 + 0) * 64 + cb) * 4;
                const uint32_t a_dn = xb + (uint32_t)((wr * 2 + 1) * 64 + cb) * 4;
                const uint32_t a_sh = smem_u32(s_shift) + (uint32_t)cb * 4;
                float ub[16], db[16], sh[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 u4 = lds128(a_up + j * 4), d4 = lds128(a_dn + j * 4), s4 = lds128(a_sh + j * 4);
                    ub[j] = u4.x; ub[j + 1] = u4.y; ub[j + 2] = u4.z; ub[j + 3] = u4.w;
                    db[j] = d4.x; db[j + 1] = d4.y; db[j + 2] = d4.z; db[j + 3] = d4.w;
                    sh[j] = s4.x; sh[j + 1] = s4.y; sh[j + 2] = s4.z; sh[j + 3] = s4.w;
                }
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);
                    float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);
                    up = is0 ? ub[j] : up;
                    dn = is31 ? db[j] : dn;
                    const float sum = (up + __uint_as_float(r1[j])) + (dn + sh[j]);
                    v[j] = p.relu ? fmaxf(sum, 0.f) : sum;
                }
                if (EPI == 2) {
                    const uint32_t a_wf = smem_u32(s_wf) + (uint32_t)cb * 4;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 w4 = lds128(a_wf + j * 4);
                        dot = fmaf(v[j], w4.x, dot); dot = fmaf(v[j + 1], w4.y, dot);
                        dot = fmaf(v[j + 2], w4.z, dot); dot = fmaf(v[j + 3], w4.w, dot);
                    }
                } else if (TMAST) {
                    if (interior) {
                        if (!valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;
                        }
                        uint4 qa, qb;
                        pack16<TOut>(v, qa, qb);
                        const uint32_t rr = (uint32_t)(r - 1);                 // staging row of output row r
                        const uint32_t rbase = stage_o + rr * 128;
                        const uint32_t ch = (uint32_t)cb >> 3;
                        const uint32_t sw = rr & 7u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((ch ^ sw) << 4)), "r"(qa.x),
                                     "r"(qa.y), "r"(qa.z), "r"(qa.w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((ch + 1) ^ sw) << 4)), "r"(qb.x),
                                     "r"(qb.y), "r"(qb.z), "r"(qb.w) : "memory");
                    }
                } else if (valid) {
                    store16<TOut>(out + pos * p.out_cs + n_blk * 64 + cb, v);
                }
            }
            if (EPI == 2) {
                // eps_hat = <a10, wf> + bf: set 1 hands its 32-channel partial to set 0, which applies the update
                if (set == 1) s_dot[r] = dot;
                named_bar_sync(1, EPI_THREADS);
                if (set == 0 && valid) {
                    const long long pix = (long long)n_img * (p.g.H * p.g.W) + hh * p.g.W + ww;
                    const float e = dot + s_dot[r] + __ldg(p.bf);
                    const float xv = p.x[pix];
                    float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(p.sig, e)), p.sqa);
                    x0 = fminf(fmaxf(x0, -1.f), 1.f);
                    float xn = __fadd_rn(__fmul_rn(p.sqp, x0), __fmul_rn(p.sqv, __ldg(p.z + pix)));
                    if (p.final_clamp) xn = fminf(fmaxf(xn, -1.f), 1.f);
                    p.x[pix] = xn;
                }
            }
            if (TMAST) {
                fence_proxy_async();
                named_bar_sync(1, EPI_THREADS);
                if (threadIdx.x == EPI_T0) {
                    tma_store_2d(&tmO, stage_o, n_blk * 64, tile * C3_BOUT + p.g.guard);
                    tma_store_commit();
                }
            }
            if (p.dbg) { dbg_acc[5] += clock64() - t1; dbg_acc[6] += 1; }
            obuf ^= 1;
            par ^= 1;
            if (++buf == ACC_BUFS) { buf = 0; acc_phase ^= 1; }
        }
    }
    if (TMAST && threadIdx.x == EPI_T0) tma_store_wait_all();
    if (p.dbg && lane == 0 && (warp <= 1 || warp == 4)) {
        long long* d = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
        if (warp == 0) d[0] = dbg_acc[0];
        if (warp == 1) { d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; }
        if (warp == 4) { d[4] = dbg_acc[4]; d[5] = dbg_acc[5]; d[6] = dbg_acc[6]; d[7] = clock64() - t_kernel0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// 2-D output map with a 126-row box (the staging tile holds output rows 1..126 of the D tile)
template <typename T>
CUtensorMap make_map_out126(const void* base, uint64_t rows, uint64_t cols) {
    return make_map_2d<T>(base, rows, cols, C3_BOUT);
}

template <int CHUNKS, int WP, int TMAST>
constexpr int c3_stages() {
    constexpr int R = ((TC_BM + 2 * WP + 7) / 8) * 8;
    constexpr int budget = 227 * 1024 - 1024 - 3 * CHUNKS * 192 * 128 - (TMAST ? 2 * TC_BM * 128 : 0) - 7 * 1024;
    constexpr int s = budget / (R * 128);
    // CHUNKS == 1: an even stage count keeps each stage private to one of the two issuer warps
    return CHUNKS == 1 ? (s >= 4 ? 4 : 2) : (s > 6 ? 6 : s);
}

template <int CHUNKS, int WP, int EPI, int TMAST, typename TIn, typename TOut>
void launch_c3(cudaStream_t st, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o,
               const C3Params& p, int n_blocks_y) {
    constexpr int STAGES = c3_stages<CHUNKS, WP, TMAST>();
    static_assert(STAGES >= 2, "not enough shared memory");
    constexpr int R = ((TC_BM + 2 * WP + 7) / 8) * 8;
    constexpr size_t smem = 1024 + (size_t)3 * CHUNKS * 192 * 128 + (size_t)STAGES * R * 128 + (TMAST ? 2 * TC_BM * 128 : 0) +
                            7 * 1024;
    auto kern = conv3_tc_kernel<CHUNKS, WP, STAGES, EPI, TMAST, TIn, TOut>;
    static bool attr_set = false;
    if (!attr_set) {
        DDPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int ctas_x = state().num_sms / n_blocks_y;
    if (ctas_x > p.num_tiles) ctas_x = p.num_tiles;
    if (ctas_x < 1) ctas_x = 1;
    kern<<<dim3(ctas_x, n_blocks_y), C3_THREADS, smem, st>>>(a0, a1, w, o, p);
    DDPM_LAUNCH_CHECK();
}

// W3 layout (pack_conv3_rows_kernel): rows = (Cout/64 blocks) x (dx, co_local) = 192 per block, cols = (dy, ci) = 3*Cin
template <typename TIn, typename TOut>
bool conv3x3_v2(cudaStream_t st, const TIn* s0, int C0, const TIn* s1, int C1, const TIn* W3, int Cout, TOut* out, const Geo& g,
                const float* shift, int relu, const C3Params* fused = nullptr) {
    if (!available()) return false;
    if constexpr (sizeof(TIn) != 2 || sizeof(TOut) != 2) {
        return false;
    } else {
    const int Cin = C0 + C1;
    if ((Cin != 64 && Cin != 128) || (Cout != 64 && Cout != 128) || (g.Wp != 34 && g.Wp != 18)) return false;
    if (s1 && (C0 != 64 || C1 != 64)) return false;
    const uint64_t rows = (uint64_t)g.alloc_positions();
    C3Params p{};
    if (fused) p = *fused;
    p.out = out; p.out_cs = Cout; p.g = g; p.shift = shift; p.relu = relu;
    p.num_tiles = cdiv(g.npos, C3_BOUT);
    p.chunk1_src1 = (s1 != nullptr) ? 1 : 0;
    p.dbg = state().dbg;
    const int nblk = Cout / 64;
    const int R = ((TC_BM + 2 * g.Wp + 7) / 8) * 8;
    CUtensorMap a0 = make_map_2d<TIn>(s0 - (size_t)g.guard * C0, rows, C0, R);
    CUtensorMap a1 = s1 ? make_map_2d<TIn>(s1 - (size_t)g.guard * C1, rows, C1, R) : a0;
    CUtensorMap w = make_map_2d<TIn>(W3, (uint64_t)nblk * 192, (uint64_t)3 * Cin, 192);
    CUtensorMap o = fused ? a0 : make_map_out126<TOut>(out - (size_t)g.guard * Cout, rows, Cout);
    const bool wide = g.Wp == 34;
    if (fused) {
        if (!(wide && Cin == 64 && Cout == 64)) return false;
        launch_c3<1, 34, 2, 0, TIn, TOut>(st, a0, a1, w, o, p, 1);
    } else if (Cin == 64) {
        if (wide) launch_c3<1, 34, 0, 1, TIn, TOut>(st, a0, a1, w, o, p, nblk);
        else launch_c3<1, 18, 0, 1, TIn, TOut>(st, a0, a1, w, o, p, nblk);
    } else {
        if (wide) launch_c3<2, 34, 0, 0, TIn, TOut>(st, a0, a1, w, o, p, nblk);
        else launch_c3<2, 18, 0, 0, TIn, TOut>(st, a0, a1, w, o, p, nblk);
    }
    return true;
    }
}

}  // namespace tc
}  // namespace ddpm
