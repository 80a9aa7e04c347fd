// tcgen05 (5th-generation tensor core) implicit-GEMM convolution for sm_100a.
//
// Formulation ("shifted-window implicit GEMM on the padded position matrix"):
// activations are a 2-D matrix [position][channel] over the zero-padded, image-stacked position
// space of common.cuh, so the 3x3 tap (dy,dx) is the constant row shift dy*(W+2)+dx.  For one
// M-tile of 128 consecutive positions a CTA
//   1. TMA-loads ONE halo'ed slab of rows [m0-(Wp+1), m0+128+(Wp+1)) x 64 channels into shared memory
//      (128B-swizzled, K-major) -- 1.5x the tile instead of the 9x an im2col gather would move;
//   2. issues 9 taps x 4 K-steps of tcgen05.mma (kind::f16, FP32 accumulate in TMEM); tap (dy,dx) is just
//      the shared-memory matrix descriptor advanced by that many 128-byte rows -- no data is moved
//      between taps;
//   3. drains the accumulator with tcgen05.ld: y = acc + shift[c] (any scale is folded into the weights),
//      ReLU, convert, store (halo positions hold zeros).
// The 9*Cin x Cout weights stay resident in shared memory for the lifetime of the persistent CTA.
// By default two CTAs form a pair (cluster of 2, cta_group::2): M = 256 per instruction, each CTA stages
// its own slab and half of the weight rows (see the comment above conv_tc_kernel).
// Warp roles (352 threads): warp 0 = TMA producer, warps 1 and 6 = MMA issuers (alternate tiles), warps
// 2..5 and 7..10 = two epilogue sets (alternate tiles); smem full/empty ring (TMA <-> MMA), one barrier per
// weight tile, and a 4-deep TMEM full/empty ring (MMA <-> epilogue).
//
// Replaces NNlib.conv / ∇conv_data (im2col + SGEMM) behind Flux.Conv at
// /root/reference/src/train_brain.jl:113-140 (forward) and Zygote's pullback (:267-269).
#pragma once
#include <cuda.h>
#include <map>
#include <tuple>
#include <type_traits>
#include "common.cuh"

namespace ddpm {
namespace tc {

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Host-mapped diagnostics: a wait that never completes records where it was stuck before it traps.
__device__ int* g_trap_info = nullptr;     // [8] in pinned, mapped host memory (set by tc::init)

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site = 0) {
    uint32_t done = 0;
    unsigned long long spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1ull << 24)) {          // a lost arrival must fault, never hang the GPU
            if (g_trap_info) {
                g_trap_info[0] = 1 + site; g_trap_info[1] = blockIdx.x; g_trap_info[2] = blockIdx.y;
                g_trap_info[3] = threadIdx.x; g_trap_info[4] = (int)parity; g_trap_info[5] = (int)bar;
                g_trap_info[6] = gridDim.x;
                __threadfence_system();
            }
            __trap();
        }
    } while (!done);
}
// one lane of a fully converged warp (keeps the surrounding control flow warp-uniform, so descriptors
// and barrier addresses stay in uniform registers instead of per-MMA R2UR/ELECT loops)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// shared -> global tile store (bulk async group); the box is clipped at the tensor bounds
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream drains; everything before pdl_wait() must not touch global memory the predecessor
// (or anything before it) writes.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) plumbing: the two CTAs of a cluster run one M=256 MMA stream issued by rank 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    // default semantics (.release.cta) on purpose: a cluster-scope release would first drain this thread's outstanding
    // global stores of the previous tile (~1.5k cycles per tile measured); the TMEM hand-over it guards is ordered by
    // tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not by the memory model
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose byte count is credited to an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}

template <int COLS, int CG = 1>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
    uint32_t ncols = COLS;
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int COLS, int CG = 1>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    uint32_t ncols = COLS;
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], single-CTA, kind::f16 (FP16/BF16 inputs, FP32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors passed as (lo, hi) halves so that advancing the start address is one 32-bit add
// same for FP32 operands read as TF32 (kind::tf32: K = 8 elements = the same 32 bytes per instruction), single CTA
__device__ __forceinline__ void umma_tf32_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int CG = 1>
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        // M = 256 over the CTA pair: each CTA supplies its own 128 A rows and half of the B rows (same smem offsets)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (CG == 2: the arrival is multicast to the barrier at the same offset in both CTAs of the pair)
template <int CG = 1>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
    }
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 16-bit elements
// (128 B), 8-row core groups 1024 B apart (SBO); start may be advanced by whole rows and by
// 32-byte K-steps inside the swizzle atom (the XOR pattern is a function of the address bits).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t base_offset_mode) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                      // LBO (unused for swizzled K-major) = 16 B
    d |= (uint64_t)(1024 >> 4) << 32;            // SBO = 1024 B
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

// operand format field of the instruction descriptor: kind::f16 -> 0 = F16, 1 = BF16;  kind::tf32 -> 2 = TF32
template <typename T> struct IsBf16 { static constexpr uint32_t v = 0; };
template <> struct IsBf16<__nv_bfloat16> { static constexpr uint32_t v = 1; };
template <> struct IsBf16<float> { static constexpr uint32_t v = 2; };

// instruction descriptor for kind::f16: D=F32, A/B = F16 or BF16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------------------------ kernel
// warp 0 TMA, warps 1 and 6 MMA issuers (even / odd tiles), warps 2..5 and 7..10 epilogue sets (even / odd tiles)
constexpr int TC_THREADS = 352;
constexpr int TC_BM = 128;

template <typename TOut>
__device__ __forceinline__ void store16(TOut* p, const float v[16]);
template <>
__device__ __forceinline__ void store16<__half>(__half* p, const float v[16]) {
    uint4 a, b;
    __half2* ha = reinterpret_cast<__half2*>(&a);
    __half2* hb = reinterpret_cast<__half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { ha[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]); hb[i] = __floats2half2_rn(v[8 + 2 * i], v[9 + 2 * i]); }
    *reinterpret_cast<uint4*>(p) = a;
    *reinterpret_cast<uint4*>(p + 8) = b;
}
template <>
__device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* p, const float v[16]) {
    uint4 a, b;
    __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { ha[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); hb[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[9 + 2 * i]); }
    *reinterpret_cast<uint4*>(p) = a;
    *reinterpret_cast<uint4*>(p + 8) = b;
}

template <>
__device__ __forceinline__ void store16<float>(float* p, const float v[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

template <typename TOut>
__device__ __forceinline__ void pack16(const float v[16], uint4& a, uint4& b);
template <>
__device__ __forceinline__ void pack16<float>(const float*, uint4&, uint4&) {}      // (staged epilogues are 16-bit only)
template <>
__device__ __forceinline__ void pack16<__half>(const float v[16], uint4& a, uint4& b) {
    __half2* ha = reinterpret_cast<__half2*>(&a);
    __half2* hb = reinterpret_cast<__half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { ha[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]); hb[i] = __floats2half2_rn(v[8 + 2 * i], v[9 + 2 * i]); }
}
template <>
__device__ __forceinline__ void pack16<__nv_bfloat16>(const float v[16], uint4& a, uint4& b) {
    __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { ha[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); hb[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[9 + 2 * i]); }
}

struct TcParams {
    void* out;            // position 0 of the output tensor
    int out_cs;           // channel stride (elements) of the output tensor
    Geo g;                // geometry of the A operand rows (== output geometry for convs)
    Geo g_out;            // EPI==1 only: fine geometry of the ConvTranspose output
    const float* shift;   // per output channel, nullptr -> 0   (EPI==1: bias).  Any per-channel SCALE is
                          // folded into the packed weights by the caller (inference BatchNorm fold).
    int relu;
    int num_m_tiles;
    int chunk1_src1;      // 1: K-chunk 1 comes from the second tensor map (channel concat), 0: channels 64.. of the first
    int rev;              // 1: walk the tiles from the last to the first.  Consecutive layers of the sampler alternate the
                          // direction, so a layer starts with the part of its input the previous kernel wrote LAST -- the part
                          // that is still in the 126 MB L2 (an LRU-streamed 100-200 MB tensor read front to back misses all of it)
    long long* dbg;       // optional [gridDim.x*gridDim.y][8] cycle counters (role wait/busy breakdown), nullptr = off
    // EPI == 2: final Conv((1,1), 64=>1) + reverse-diffusion update fused into this conv's epilogue
    float* x;             // [N][H*W] current sample, updated in place
    const float* z;       // [N][H*W] noise of this step
    const float* wf;      // [64] final conv weight, bf[1] bias
    const float* bf;
    float sig, sqa, sqp, sqv;   // sigma_t, sqrt(a_t), sqrt(a_prev), sqrt(post_var)
    int final_clamp;
    // BNS != 0: per-channel reductions of the STORED (rounded) output tile, fused into the epilogue (train-mode
    // BatchNorm, train_brain.jl:112-140): Float64 sums[2][stats_C], one atomic per channel, quantity and CTA.
    //   BNS == 1 (forward):  sums[0][c] += sum out,  sums[1][c] += sum out^2      (batch statistics of y)
    //   BNS == 2 (backward, the output is da = dL/d(BN-ReLU output) of the layer whose pre-BN tensor is bn_y):
    //                        g = da * [bn_y*scale + shift > 0];  sums[0][c] += sum g,  sums[1][c] += sum g * xhat,
    //                        xhat = (bn_y - mean) * istd          (first pass of the BatchNorm backward)
    double* stats;
    int stats_C;                // channel count of the statistics vectors (second quantity starts at stats_C)
    int stats_nch;              // BNS == 2: only output channels [0, stats_nch) carry a BatchNorm (d(cat): first half)
    const void* bn_y;           // BNS == 2: position 0 of the pre-BatchNorm tensor y (same geometry, bn_cs channels)
    int bn_cs;
    const float *bn_scale, *bn_shift, *bn_mean, *bn_istd;
};

// TAPS: 9 (3x3 conv, halo'ed slab) or 1 (plain GEMM rows); CHUNKS: 64-channel K chunks (1|2);
// NOUT: output channels handled by this CTA (64|128); WP: padded row width (W+1, Geo::Wp) for TAPS==9
// EPI: 0 = conv store on the same geometry, 1 = ConvTranspose 2x2 pixel shuffle (blockIdx.y = q),
//      2 = no activation store: eps_hat = <relu(acc+shift), wf> + bf per pixel, then the reverse-diffusion update
//          x <- sqrt(a_prev)*clamp((x - sigma*eps_hat)/sqrt(a_t)) + sqrt(pv)*z   (generate_images.jl:196-208)
// TMAST: 1 = epilogue stages the tile in swizzled shared memory and writes it with a TMA store
// (halo rows are written as zeros, which is what they must hold), 0 = predicated 16-byte stores.
// ---- shared-memory plan (shared by the kernel and its launcher)
// slab stages that fit next to the resident weights and `extra` bytes of epilogue staging
template <int TAPS, int CHUNKS, int NOUT, int WP, int CG>
constexpr int stages_for(int extra) {
    constexpr int HALO = (TAPS == 9) ? (WP + 1) : 0;
    constexpr int R = ((TC_BM + 2 * HALO + 7) / 8) * 8;
    const int budget = 227 * 1024 - 1024 /*align*/ - 1280 /*barriers, shift*/ - TAPS * CHUNKS * (NOUT / CG) * 128 - extra;
    if (budget <= 0) return 0;
    int s = budget / (R * 128);
    if (s > 8) s = 8;
    // CTA pairs: prefer a ring the two issuer warps can split (see DUAL in the kernel) over one more stage
    if (CG == 2 && s >= 2 * CHUNKS) s = (s / (2 * CHUNKS)) * (2 * CHUNKS);
    return s;
}
// channels per pass of the per-warp store staging (direct-store epilogues): every epilogue warp transposes its
// 32 rows x STC channels through a private shared-memory patch so that a store instruction writes whole 128-byte
// (STC=64) or 64-byte (STC=32) row segments instead of 32 scattered 16-byte pieces.  0 = no room, direct stores.
template <int TAPS, int CHUNKS, int NOUT, int WP, int EPI, int TMAST, int CG, int ESZ = 2>
constexpr int stage_cols() {
    if (TMAST || EPI == 2 || ESZ != 2) return 0;       // FP32 outputs (TF32 variants): direct 16-byte stores
    const int full = stages_for<TAPS, CHUNKS, NOUT, WP, CG>(0);
    const int want = full < 4 ? full : 4;
    if (stages_for<TAPS, CHUNKS, NOUT, WP, CG>(8 * 32 * 64 * 2) >= want) return 64;
    if (stages_for<TAPS, CHUNKS, NOUT, WP, CG>(8 * 32 * 32 * 2) >= want) return 32;
    return 0;
}
template <int TAPS, int CHUNKS, int NOUT, int WP, int EPI, int TMAST, int CG, int ESZ = 2>
constexpr int epi_smem_bytes() {
    return TMAST ? 2 * TC_BM * NOUT * 2 : 8 * 32 * stage_cols<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, ESZ>() * 2;
}
template <int TAPS, int CHUNKS, int NOUT, int WP, int EPI, int TMAST, int CG = 1, int ESZ = 2>
constexpr int pick_stages() {
    return stages_for<TAPS, CHUNKS, NOUT, WP, CG>(epi_smem_bytes<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, ESZ>());
}

// CG: 1 = one CTA per 128-position tile; 2 = CTA pair (cluster of 2, tcgen05 cta_group::2): the pair works on 256
// consecutive positions, each CTA stages its own 128-position slab and HALF of the weight rows (NOUT/2 output
// channels); rank 0 issues M=256 MMAs that read both CTAs' shared memory, each CTA's TMEM receives all NOUT columns
// of its own 128 positions.  Per SM and MMA the shared-memory operand traffic drops from 4 KB + NOUT*32 B to
// 4 KB + NOUT*16 B, and 128 => 128 layers no longer need the Cout split that re-read every slab twice.
template <int TAPS, int CHUNKS, int NOUT, int WP, int STAGES, int EPI, int TMAST, int CG, int BNS, typename TIn, typename TOut>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const TcParams p) {
    constexpr int HALO = (TAPS == 9) ? (WP + 1) : 0;
    constexpr int R = ((TC_BM + 2 * HALO + 7) / 8) * 8;       // slab rows (multiple of 8 -> 1024 B multiple)
    constexpr uint32_t A_STAGE_BYTES = R * 128;
    constexpr int NB = NOUT / CG;                              // weight rows (output channels) staged by this CTA
    constexpr uint32_t W_TILE_BYTES = NB * 128;                // one (tap, chunk) weight tile [NB][64]
    constexpr uint32_t W_BYTES = TAPS * CHUNKS * W_TILE_BYTES;
    constexpr int ACC_BUFS = 4;                                // accumulator ring in TMEM (decouples MMA and epilogue)
    constexpr bool DUAL = (STAGES % (2 * CHUNKS)) == 0;        // two MMA issuer warps only if stages stay warp-private
    constexpr int TMEM_COLS = (ACC_BUFS * NOUT <= 256) ? 256 : 512;
    constexpr uint32_t IDESC = make_idesc(IsBf16<TIn>::v, TC_BM * CG, NOUT);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // epilogue staging: TMAST -> two tiles [128][NOUT] (one per epilogue set); else 8 per-warp patches [32][STC]
    constexpr int STC = stage_cols<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, (int)sizeof(TOut)>();
    constexpr uint32_t O_BYTES = (uint32_t)epi_smem_bytes<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, (int)sizeof(TOut)>();
    constexpr bool TF32 = std::is_same<TIn, float>::value;     // FP32 operands read as TF32 (kind::tf32), FP32 output
    constexpr int CPC = 128 / (int)sizeof(TIn);                // channels per 128-byte K chunk: 64 (16-bit) or 32 (FP32)
    static_assert(!TF32 || (CG == 1 && TMAST == 0 && BNS == 0 && EPI == 0), "TF32 variants: single CTA, plain conv epilogue");
    const uint32_t s_w = smem_u32(smem);
    const uint32_t s_a = s_w + W_BYTES;
    const uint32_t s_o = s_a + STAGES * A_STAGE_BYTES;
    const uint32_t s_bar = s_o + O_BYTES;
    // barrier slots (8 B each): [0..NWG) weight tiles (one per (chunk, tap), in the order the MMAs consume them, so the
    // first tile starts while most of the weights are still in flight), then S a_full, S a_empty, ACC_BUFS acc_full,
    // ACC_BUFS acc_empty
    constexpr int NWG = TAPS * CHUNKS;
    auto bar_w = [&](int g) { return s_bar + 8u * g; };
    auto bar_afull = [&](int s) { return s_bar + 8u * (NWG + s); };
    auto bar_aempty = [&](int s) { return s_bar + 8u * (NWG + STAGES + s); };
    auto bar_accfull = [&](int b) { return s_bar + 8u * (NWG + 2 * STAGES + b); };
    auto bar_accempty = [&](int b) { return s_bar + 8u * (NWG + 2 * STAGES + ACC_BUFS + b); };
    uint8_t* misc = smem + W_BYTES + STAGES * A_STAGE_BYTES + O_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (NWG + 2 * STAGES + 2 * ACC_BUFS));
    static_assert(8 * (NWG + 2 * STAGES + 2 * ACC_BUFS) + 4 <= 512, "barrier area overflows into the shift vector");
    float* s_shift = reinterpret_cast<float*>(misc + 512);               // [NOUT] per-channel shift of this CTA's N-slice
    float* s_wf = s_shift + 128;                                         // [64] final-conv weights (EPI == 2)
    static_assert(BNS == 0 || (EPI == 0 && (TMAST || stage_cols<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, (int)sizeof(TOut)>() > 0)),
                  "fused BatchNorm reductions read the staged output tile");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_kernel0 = p.dbg ? clock64() : 0;
    const int n_blk = blockIdx.y;                 // N-slice (conv: half of Cout; up2: sub-position q)
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;     // position half / weight-row half inside the CTA pair
    const bool leader = (rank == 0);
    auto ch_off_of = [](int nb) { return (EPI == 1) ? 0 : nb * NOUT; };   // first output channel of this CTA's N-slice
    const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;   // persistent work unit = CTA (CG 1) or CTA pair (CG 2)
    const int num_units_tiles = (p.num_m_tiles + CG - 1) / CG;    // tiles of CG*128 positions

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0); prefetch_tmap(&tmW);
        if (p.chunk1_src1) prefetch_tmap(&tmA1);
        for (int g = 0; g < NWG; ++g) mbar_init(bar_w(g), 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_afull(s), 1); mbar_init(bar_aempty(s), 1); }
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), 4 * CG); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS, CG>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0 && TMAST) prefetch_tmap(&tmO);
    // everything above is private to this CTA: with a programmatic dependent launch it overlaps the predecessor's tail
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x >= 64 && threadIdx.x < 64 + NOUT) {
        const int c = threadIdx.x - 64;
        const int co = (EPI == 1) ? c : (n_blk * NOUT + c);
        s_shift[c] = p.shift ? p.shift[co] : 0.f;
        if (EPI == 2) s_wf[c] = p.wf[c];
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();          // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // barriers that collect contributions of both CTAs live in the leader (rank 0)
    auto lead = [&](uint32_t bar) { return (CG == 2) ? map_to_cta(bar, 0) : bar; };

    if (warp == 0) {
        // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
        if (elect_one()) {
            for (int c = 0; c < CHUNKS; ++c)
                for (int t = 0; t < TAPS; ++t) {
                    if (leader) mbar_expect_tx(bar_w(c * TAPS + t), W_TILE_BYTES * CG);
                    const uint32_t bw = lead(bar_w(c * TAPS + t));
                    if (CG == 2)
                        tma_load_2d_pair(s_w + (t * CHUNKS + c) * W_TILE_BYTES, &tmW, (t * CHUNKS + c) * CPC,
                                         n_blk * NOUT + (int)rank * NB, bw);
                    else
                        tma_load_2d(s_w + (t * CHUNKS + c) * W_TILE_BYTES, &tmW, (t * CHUNKS + c) * CPC, n_blk * NOUT, bw);
                }
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int ut = unit; ut < num_units_tiles; ut += n_units) {
            const int tile = (p.rev ? num_units_tiles - 1 - ut : ut) * CG + (int)rank;
            const int row0 = tile * TC_BM - HALO + p.g.guard;   // row coordinate in the tensor map (base = allocation start)
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                long long t0 = p.dbg ? clock64() : 0;
                mbar_wait(bar_aempty(stage), phase ^ 1);
                if (p.dbg) dbg_acc[0] += clock64() - t0;
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bar_afull(stage), A_STAGE_BYTES * CG);
                    // channel concat: the second half of the K chunks comes from the second tensor map
                    const bool second = p.chunk1_src1 && (c >= CHUNKS / 2);
                    const int ccol = (second ? c - CHUNKS / 2 : c) * CPC;
                    if (CG == 2)
                        tma_load_2d_pair(s_a + stage * A_STAGE_BYTES, second ? &tmA1 : &tmA0, ccol, row0, lead(bar_afull(stage)));
                    else
                        tma_load_2d(s_a + stage * A_STAGE_BYTES, second ? &tmA1 : &tmA0, ccol, row0, bar_afull(stage));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if ((warp == 1 || (warp == 6 && DUAL)) && leader) {
        // ================= MMA issuers (warp-uniform loops, one elected lane issues) =================
        // Two issuer warps alternate tiles (warp 1: even, warp 6: odd tiles of this CTA).  The tensor pipe's
        // instruction queue is shallow, so a single issuer drains it during every mbarrier wait (~150 cycles
        // each); with two independent streams one warp's waits overlap the other's MMAs.  Tiles are
        // independent (own TMEM accumulator, own slab), tcgen05.commit tracks the issuing thread's MMAs only.
        // Dual issue is only legal when each pipeline stage is always consumed by the same warp.
        const int parity = (warp == 1) ? 0 : 1;
        constexpr int NISS = DUAL ? 2 : 1;
        bool w_pending = true;                   // first tile of this warp: wait for each weight tile right before its MMAs
        // matrix descriptor halves (see make_desc_sw128): lo = start>>4 | LBO<<16, hi = SBO | version | SWIZZLE_128B
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo_base = ((s_a & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo_base = ((s_w & 0x3FFFFu) >> 4) | (1u << 16);
        for (int seq = parity, ut = unit + parity * n_units; ut < num_units_tiles; seq += NISS, ut += NISS * n_units) {
            const int buf = seq % ACC_BUFS;
            const uint32_t acc_phase = (uint32_t)(seq / ACC_BUFS) & 1u;
            long long t0 = p.dbg ? clock64() : 0;
            mbar_wait(bar_accempty(buf), acc_phase ^ 1);
            if (p.dbg) dbg_acc[1] += clock64() - t0;
            const uint32_t d_tmem = tmem_base + buf * NOUT;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int step = seq * CHUNKS + c;
                const int stage = step % STAGES;
                const uint32_t phase = (uint32_t)(step / STAGES) & 1u;
                t0 = p.dbg ? clock64() : 0;
                mbar_wait(bar_afull(stage), phase);
                if (p.dbg) dbg_acc[2] += clock64() - t0;
                tc_fence_after();
                t0 = p.dbg ? clock64() : 0;
                if (elect_one()) {
                    const uint32_t a_lo = a_lo_base + stage * (A_STAGE_BYTES >> 4);
#pragma unroll
                    for (int t = 0; t < TAPS; ++t) {
                        const int shift = (TAPS == 9) ? (HALO + (t / 3 - 1) * WP + (t % 3 - 1)) : 0;
                        if (w_pending) {
                            mbar_wait(bar_w(c * TAPS + t), 0);
                            tc_fence_after();
                        }
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if constexpr (TF32)
                                umma_tf32_lh(d_tmem, a_lo + ((shift * 128 + ks * 32) >> 4),
                                             b_lo_base + (((t * CHUNKS + c) * W_TILE_BYTES + ks * 32) >> 4), DESC_HI, IDESC,
                                             (c | t | ks) ? 1u : 0u);
                            else
                                umma_f16_lh<CG>(d_tmem, a_lo + ((shift * 128 + ks * 32) >> 4),
                                                b_lo_base + (((t * CHUNKS + c) * W_TILE_BYTES + ks * 32) >> 4), DESC_HI, IDESC,
                                                (c | t | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit<CG>(bar_aempty(stage));                       // slab reusable once these MMAs retire
                    if (c == CHUNKS - 1) umma_commit<CG>(bar_accfull(buf));   // accumulator ready for the epilogue
                }
                __syncwarp();
                if (p.dbg) dbg_acc[3] += clock64() - t0;
            }
            w_pending = false;
            dbg_acc[7] += 1;
        }
    } else if (warp == 1 || warp == 6) {
        // issuer warps without work: second issuer in single-issuer mode, both issuers of the non-leader CTA of a pair
    } else {
        // ================= epilogue warps (TMEM -> registers -> global) =================
        // Two sets of four warps (2..5 and 7..10) take alternate tiles: one tile's epilogue is a ~2.5k-cycle latency
        // chain (accumulator wait, tcgen05.ld round trips, staging-tile hand-over to the TMA store), so a single set
        // caps the CTA at one tile per chain; two sets overlap the chains of consecutive tiles.
        const int eset = (warp >= 7) ? 1 : 0;
        const int lane_grp = warp & 3;                       // TMEM lanes 32*lane_grp .. +31 are visible to this warp
        const int row = lane_grp * 32 + lane;
        const bool store_thread = (threadIdx.x == (eset ? 224 : 64));     // first lane of the set: issues its TMA stores
        TOut* out = reinterpret_cast<TOut*>(p.out);
        const uint32_t accempty_base = lead(bar_accempty(0));
        const uint32_t stage_o = s_o + (uint32_t)eset * (TC_BM * NOUT * 2);     // this set's staging tile (TMAST)
        const uint32_t wst = s_o + (uint32_t)(eset * 4 + lane_grp) * (32 * STC * 2);   // this warp's store patch (STC > 0)
        const int npos = (int)p.g.npos;
        // fused BatchNorm reductions: every lane owns one channel PAIR of each staged fill (64 or 32 channels wide) and
        // walks the rows its own warp staged; partial sums stay in registers for the whole kernel
        constexpr int FW = TMAST ? 64 : (STC > 0 ? STC : 64);               // channels per staged fill
        constexpr int NF = NOUT >= FW ? NOUT / FW : 1;                       // fills per tile
        float bs1[NF][2], bs2[NF][2];
#pragma unroll
        for (int f = 0; f < NF; ++f) { bs1[f][0] = bs1[f][1] = bs2[f][0] = bs2[f][1] = 0.f; }
        // BNS == 2: the pre-BatchNorm values y the lane needs for the fills of one 64-column round (32 rows x 1 channel
        // pair when a fill is 64 channels wide, 2 fills x 16 rows x 1 pair when it is 32 wide) are fetched with 32
        // INDEPENDENT loads issued before the round's TMEM loads, so one memory latency is paid per round, not per row
        uint32_t ypre[(BNS == 2) ? 32 : 1];
        auto bn_prefetch = [&](int cq, int tile) {
            if constexpr (BNS == 2) {
                constexpr int LPR = FW / 2;
                const int cp = lane % LPR, rpar = lane / LPR;
                const TOut* ybase = reinterpret_cast<const TOut*>(p.bn_y) +
                                    (long long)(tile * TC_BM + lane_grp * 32) * p.bn_cs + ch_off_of(n_blk) + cq + 2 * cp;
                const bool live = ch_off_of(n_blk) + cq < p.stats_nch;          // uniform per round (nch is a multiple of 64)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    // FW == 64: row i.  FW == 32: i < 16 -> first fill of the round, rows 2i+rpar; else second fill (+32 channels)
                    const int r = (FW == 64) ? i : (2 * (i & 15) + rpar);
                    const int coff = (FW == 64) ? 0 : ((i >> 4) * 32);
                    ypre[i] = live ? __ldg(reinterpret_cast<const unsigned int*>(ybase + (long long)r * p.bn_cs + coff)) : 0u;
                }
            }
        };
        // column walk of one staged fill (fill_base = shared address of the first row this warp staged); yoff = index
        // of the fill's first prefetched y word
        auto bn_accumulate = [&](uint32_t fill_base, int f, int c_first, uint32_t vmask, int yoff) {
            constexpr int ROWB = FW * 2;                                     // bytes per staged row
            constexpr int LPR = FW / 2;                                      // lanes per row (one channel pair each)
            constexpr int NR = LPR;                                          // rows per lane: 32 (FW=64) or 16 (FW=32)
            const int cp = lane % LPR;                                       // channel pair inside the fill
            const int rpar = lane / LPR;                                     // FW=32: row parity handled by this half-warp
            const int cg = c_first + 2 * cp;                                  // first of the two output channels (CTA-local)
            auto lds = [&](int r) {
                const uint32_t sw = (FW == 64) ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w)
                             : "r"(fill_base + (uint32_t)r * ROWB + ((((uint32_t)cp >> 2) ^ sw) << 4) + ((uint32_t)cp & 3u) * 4u));
                return w;
            };
            if constexpr (BNS == 1) {
                // forward statistics in FP32 (the variance is a difference of two sums: no 16-bit accumulation here)
                float a1x = 0.f, a1y = 0.f, a2x = 0.f, a2y = 0.f;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int r = (FW == 64) ? i : (2 * i + rpar);          // warp-local row 0..31
                    uint32_t w = lds(r);
                    float2 v;
                    if constexpr (std::is_same<TOut, __half>::value) v = __half22float2(*reinterpret_cast<__half2*>(&w));
                    else v = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
                    const float m = ((vmask >> r) & 1u) ? 1.f : 0.f;        // halo / out-of-range rows add nothing
                    v.x *= m; v.y *= m;
                    a1x += v.x; a1y += v.y; a2x = fmaf(v.x, v.x, a2x); a2y = fmaf(v.y, v.y, a2y);
                }
                bs1[f][0] += a1x; bs1[f][1] += a1y; bs2[f][0] += a2x; bs2[f][1] += a2y;
            } else if constexpr (BNS == 2) {
                // first BatchNorm-backward pass in packed 16-bit arithmetic (6 instructions per row and channel pair; the
                // FP32 form made the epilogue the bottleneck of the kernel).  Per channel: the ReLU mask z > 0 becomes
                // a threshold test on y (sgn*y > thr, thr rounded toward -inf so that the 16-bit test picks exactly the
                // values the FP32 test picks), xhat = y*istd - mean*istd.  Row sums of one fill (<= 32 terms) are kept in
                // 16 bits and widened to FP32 once per fill.
                using T2 = typename std::conditional<std::is_same<TOut, __half>::value, __half2, __nv_bfloat162>::type;
                if (ch_off_of(n_blk) + cg >= p.stats_nch) return;
                const int c = ch_off_of(n_blk) + cg;
                float thr[2], sg[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float scl = p.bn_scale[c + k], shf = p.bn_shift[c + k];
                    sg[k] = scl < 0.f ? -1.f : 1.f;
                    // scale == 0: the mask is the sign of the shift for every y
                    thr[k] = scl != 0.f ? -shf / fabsf(scl) : (shf > 0.f ? -INFINITY : INFINITY);
                }
                T2 sgn2, thr2, is2, nm2;
                const float i0 = p.bn_istd[c], i1 = p.bn_istd[c + 1];
                if constexpr (std::is_same<TOut, __half>::value) {
                    sgn2 = __floats2half2_rn(sg[0], sg[1]);
                    thr2 = __halves2half2(__float2half_rd(thr[0]), __float2half_rd(thr[1]));
                    is2 = __floats2half2_rn(i0, i1);
                    nm2 = __floats2half2_rn(-p.bn_mean[c] * i0, -p.bn_mean[c + 1] * i1);
                } else {
                    sgn2 = __floats2bfloat162_rn(sg[0], sg[1]);
                    thr2 = __halves2bfloat162(__float2bfloat16_rd(thr[0]), __float2bfloat16_rd(thr[1]));
                    is2 = __floats2bfloat162_rn(i0, i1);
                    nm2 = __floats2bfloat162_rn(-p.bn_mean[c] * i0, -p.bn_mean[c + 1] * i1);
                }
                T2 s1, s2;
                if constexpr (std::is_same<TOut, __half>::value) { s1 = __float2half2_rn(0.f); s2 = s1; }
                else { s1 = __float2bfloat162_rn(0.f); s2 = s1; }
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int r = (FW == 64) ? i : (2 * i + rpar);
                    uint32_t w = lds(r);
                    if (!((vmask >> r) & 1u)) w = 0u;                         // halo / out-of-range rows add nothing
                    uint32_t yw = ypre[yoff + i];
                    const T2 v = *reinterpret_cast<T2*>(&w), y = *reinterpret_cast<T2*>(&yw);
                    const T2 mk = __hgt2(__hmul2(y, sgn2), thr2);            // 1.0 where relu'(z) = 1
                    const T2 g = __hmul2(v, mk);
                    s1 = __hadd2(s1, g);
                    s2 = __hfma2(g, __hfma2(y, is2, nm2), s2);
                }
                float2 f1, f2;
                if constexpr (std::is_same<TOut, __half>::value) { f1 = __half22float2(s1); f2 = __half22float2(s2); }
                else { f1 = __bfloat1622float2(s1); f2 = __bfloat1622float2(s2); }
                bs1[f][0] += f1.x; bs1[f][1] += f1.y; bs2[f][0] += f2.x; bs2[f][1] += f2.y;
            }
        };
        for (int seq = eset, ut = unit + eset * n_units; ut < num_units_tiles; seq += 2, ut += 2 * n_units) {
            const int buf = seq % ACC_BUFS;
            const uint32_t acc_phase = (uint32_t)(seq / ACC_BUFS) & 1u;
            const int tile = (p.rev ? num_units_tiles - 1 - ut : ut) * CG + (int)rank;
            // position -> (image, row, column) with compile-time divisors (square images: Hs = H + 1 = W + 1 = WP)
            const int pos = tile * TC_BM + row;
            const int pr = pos / WP, pc = pos - pr * WP;
            const int n_img = pr / WP, prr = pr - n_img * WP;
            const int hh = prr - 1, ww = pc - 1;
            const bool valid = pos < npos && prr != 0 && pc >= 1 && n_img < p.g.N;
            long long opos = pos;
            int ch_off = n_blk * NOUT;
            if (EPI == 1) {
                if (valid) opos = p.g_out.pos(n_img, 2 * hh + (n_blk >> 1), 2 * ww + (n_blk & 1));
                ch_off = 0;
            }
            const int oidx = valid ? (int)opos : -1;          // output row of this thread's position (staged stores)
            long long t0 = p.dbg ? clock64() : 0;
            bn_prefetch(0, tile);                      // in flight while this warp waits for the accumulator
            mbar_wait(bar_accfull(buf), acc_phase);
            long long t1 = p.dbg ? clock64() : 0;
            if (p.dbg) dbg_acc[4] += t1 - t0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * NOUT;
            float dot = 0.f;
            if (TMAST) {
                // the TMA store that last read this set's staging tile (two tiles ago) must have drained it
                if (store_thread) tma_store_wait_read<0>();
                named_bar_sync(1 + eset, 128);
            }
            // 64 accumulator columns of this row per round trip (4 loads in flight, one wait); the TMEM buffer is handed
            // back to the MMA warp right after the last load, BEFORE the arithmetic / stores of the epilogue
            constexpr int CQ = NOUT < 64 ? NOUT : 64;              // accumulator columns per round trip
#pragma unroll
            for (int cq = 0; cq < NOUT; cq += CQ) {
                uint32_t racc[4][16];
                if (cq > 0) bn_prefetch(cq, tile);
#pragma unroll
                for (int q = 0; q < CQ / 16; ++q) tmem_ld16(taddr + cq + q * 16, racc[q]);
                tmem_ld_wait();
                if (cq + CQ == NOUT) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(accempty_base + 8u * buf);
                        else mbar_arrive(bar_accempty(buf));
                    }
                }
#pragma unroll
                for (int q = 0; q < CQ / 16; ++q) {
                    const uint32_t* r = racc[q];
                    const int cb = cq + q * 16;
                    float v[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 sh = *reinterpret_cast<const float4*>(s_shift + cb + 4 * j4);
                        v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + sh.x;
                        v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + sh.y;
                        v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + sh.z;
                        v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + sh.w;
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (TF32 && p.relu) {
                        // inference: this activation is the operand of the next TF32 convolution -- store the nearest
                        // TF32 value instead of letting the tensor core truncate it
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = tf32_rn(v[j]);
                    }
                    if (EPI == 2) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(s_wf + cb + 4 * j4);
                            dot = fmaf(v[4 * j4 + 0], w4.x, dot); dot = fmaf(v[4 * j4 + 1], w4.y, dot);
                            dot = fmaf(v[4 * j4 + 2], w4.z, dot); dot = fmaf(v[4 * j4 + 3], w4.w, dot);
                        }
                    } else if (TMAST) {
                        if (!valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;
                        }
                        uint4 qa, qb;
                        pack16<TOut>(v, qa, qb);
                        // staging tile = [NOUT/64] x [128 rows][128 B], 128B-swizzled like the tensor map expects
                        const uint32_t half_tile = stage_o + (uint32_t)(cb >> 6) * (TC_BM * 128);
                        const uint32_t rbase = half_tile + (uint32_t)row * 128;
                        const uint32_t ch = (uint32_t)(cb & 63) >> 3;            // 16-byte chunk index 0..7 (even)
                        const uint32_t sw = (uint32_t)row & 7u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((ch ^ sw) << 4)), "r"(qa.x),
                                     "r"(qa.y), "r"(qa.z), "r"(qa.w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((ch + 1) ^ sw) << 4)), "r"(qb.x),
                                     "r"(qb.y), "r"(qb.z), "r"(qb.w) : "memory");
                    } else if (STC > 0) {
                        constexpr int STCX = STC > 0 ? STC : 64;               // (dead branch when STC == 0)
                        // transpose through the warp's patch: row = lane, 16-byte chunks XOR-swizzled by row
                        uint4 qa, qb;
                        pack16<TOut>(v, qa, qb);
                        constexpr int CPR = STCX / 8;                           // 16-byte chunks per patch row
                        const uint32_t ch = (uint32_t)(cb % STCX) >> 3;          // even chunk index of these 16 channels
                        const uint32_t sw = (STCX == 64) ? ((uint32_t)lane & 7u) : (((uint32_t)lane >> 1) & 3u);
                        const uint32_t rbase = wst + (uint32_t)lane * (STCX * 2);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((ch ^ sw) << 4)), "r"(qa.x),
                                     "r"(qa.y), "r"(qa.z), "r"(qa.w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((ch + 1) ^ sw) << 4)), "r"(qb.x),
                                     "r"(qb.y), "r"(qb.z), "r"(qb.w) : "memory");
                        if ((cb + 16) % STCX == 0) {
                            // patch complete: CPR lanes per row write one contiguous row segment each
                            __syncwarp();
                            const int c_lo = cb + 16 - STCX;                      // first channel held by the patch
                            if (BNS) bn_accumulate(wst, c_lo / STCX, c_lo, __ballot_sync(0xffffffffu, valid), (STCX == 32) ? ((c_lo & 32) >> 1) : 0);
                            const uint32_t cj = (uint32_t)lane % CPR;
#pragma unroll
                            for (int it = 0; it < CPR; ++it) {
                                const int srow = it * (32 / CPR) + lane / CPR;
                                const uint32_t ssw = (STCX == 64) ? ((uint32_t)srow & 7u) : (((uint32_t)srow >> 1) & 3u);
                                uint4 val;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                             : "r"(wst + (uint32_t)srow * (STCX * 2) + ((cj ^ ssw) << 4)));
                                const int o = __shfl_sync(0xffffffffu, oidx, srow);
                                if (o >= 0)
                                    *reinterpret_cast<uint4*>(out + (long long)o * p.out_cs + ch_off + c_lo + cj * 8) = val;
                            }
                            __syncwarp();
                        }
                    } else if (valid) {
                        store16<TOut>(out + opos * p.out_cs + ch_off + cb, v);
                    }
                }
            }
            if (EPI == 2 && valid) {
                const long long pix = (long long)n_img * (p.g.H * p.g.W) + hh * p.g.W + ww;
                const float e = dot + __ldg(p.bf);
                const float xv = p.x[pix];
                float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(p.sig, e)), p.sqa);
                x0 = fminf(fmaxf(x0, -1.f), 1.f);
                float xn = __fadd_rn(__fmul_rn(p.sqp, x0), __fmul_rn(p.sqv, __ldg(p.z + pix)));
                if (p.final_clamp) xn = fminf(fmaxf(xn, -1.f), 1.f);
                p.x[pix] = xn;
            }
            if (TMAST && BNS) {
                __syncwarp();                        // rows of this warp are staged by its own lanes
                const uint32_t vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    bn_accumulate(stage_o + (uint32_t)f * (TC_BM * 128) + (uint32_t)(lane_grp * 32) * 128, f, f * 64, vm, 0);
            }
            if (TMAST) {
                fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA engine
                named_bar_sync(1 + eset, 128);
                if (store_thread) {
#pragma unroll
                    for (int hsel = 0; hsel < NOUT / 64; ++hsel)
                        tma_store_2d(&tmO, stage_o + hsel * (TC_BM * 128), n_blk * NOUT + hsel * 64, tile * TC_BM + p.g.guard);
                    tma_store_commit();
                }
            }
            if (p.dbg) { dbg_acc[5] += clock64() - t1; dbg_acc[6] += 1; }
        }
        if (BNS) {
            // block-level reduction of the per-lane partials through the (now idle) staging area, then one Float64
            // atomic per (quantity, channel) and CTA
            if (TMAST && store_thread) tma_store_wait_all();          // the TMA engine has finished reading this set's tile
            named_bar_sync(3, 256);                                   // both epilogue sets are done with their tiles
            float* red = reinterpret_cast<float*>(smem + W_BYTES + STAGES * A_STAGE_BYTES);     // [8 warps][2][NOUT]
            const int ew = eset * 4 + lane_grp;
            constexpr int LPR = FW / 2;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float a = bs1[f][j], b = bs2[f][j];
                    if (FW == 32) {                                   // the two half-warps hold the two row parities
                        a += __shfl_xor_sync(0xffffffffu, a, 16);
                        b += __shfl_xor_sync(0xffffffffu, b, 16);
                    }
                    if (lane < LPR) {
                        const int c = f * FW + 2 * (lane % LPR) + j;
                        red[(ew * 2 + 0) * NOUT + c] = a;
                        red[(ew * 2 + 1) * NOUT + c] = b;
                    }
                }
            }
            named_bar_sync(3, 256);
            const int et = eset * 128 + lane_grp * 32 + lane;         // 0..255
            for (int i = et; i < 2 * NOUT; i += 256) {
                const int q = i / NOUT, c = i - q * NOUT;
                float acc = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) acc += red[(w8 * 2 + q) * NOUT + c];
                const int cgl = ch_off_of(n_blk) + c;
                if (BNS == 1 || cgl < p.stats_nch) atomicAdd(p.stats + (size_t)q * p.stats_C + cgl, (double)acc);
            }
        }
    }
    if (TMAST && (threadIdx.x == 64 || threadIdx.x == 224)) tma_store_wait_all();
    if (p.dbg && lane == 0 && warp <= 2) {
        long long* d = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
        if (warp == 1) { d[0] = dbg_acc[7]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; }
        if (warp == 2) { d[4] = dbg_acc[4]; d[5] = dbg_acc[5]; d[6] = dbg_acc[6]; d[7] = clock64() - t_kernel0; }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();          // neither CTA may retire (or free TMEM) while the pair's MMAs can still touch it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS, CG>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct State {
    int* trap_host = nullptr;   // pinned + mapped [8]: filled by a kernel right before it traps on a stuck barrier
    long long* dbg = nullptr;   // device buffer [512][8] when role profiling is on
    EncodeTiledFn encode = nullptr;
    bool ok = false;
    int num_sms = 148;
    int base_offset_mode = 0;
    bool enabled = true;
    bool tma_store = true;      // epilogue of the 64->64 @32x32 conv: TMA store via swizzled smem (1) or direct stores (0)
    // bit set -> that layer shape runs as CTA pairs (cta_group::2):
    // 1 = 128=>128 @16x16, 2 = 64=>64 @32x32 (incl. the fused sampler epilogue), 4 = 128=>64 @32x32 (concat),
    // 8 = 64=>128 @16x16, 16 = the two data-gradient-only shapes (128=>64 @16x16, 64=>128 @32x32)
    int pair_mask = 31;
    bool pdl = true;            // launch the tcgen05 conv kernels as programmatic dependents of their stream predecessor
    bool reverse = true;        // option tc_reverse: alternate the tile direction between consecutive layers of the sampler
    // per-DEVICE launch bookkeeping (function attributes and occupancy are properties of a (kernel, device) pair)
    std::map<const void*, int> smem_attr;     // kernel -> dynamic shared memory size already granted on this device
    std::map<const void*, int> max_pairs;     // kernel -> resident CTA pairs on this device
};
// One State per CUDA device: the ABI allows several handles on several devices in one process
// (ddpm_create(..., device)); everything here is keyed by the CURRENT device (callers cudaSetDevice first).
constexpr int MAX_DEVICES = 64;
inline State& state() {
    static State s[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return s[(dev >= 0 && dev < MAX_DEVICES) ? dev : 0];
}

// opt a kernel in to `smem` bytes of dynamic shared memory, once per (kernel, device)
template <typename K>
inline void ensure_smem_attr(K kern, size_t smem) {
    State& st = state();
    const void* key = reinterpret_cast<const void*>(kern);
    auto it = st.smem_attr.find(key);
    if (it != st.smem_attr.end() && it->second >= (int)smem) return;
    DDPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    st.smem_attr[key] = (int)smem;
}

inline void init() {
    State& s = state();
    if (s.encode) return;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        s.ok = false;
        return;
    }
    s.encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (cudaHostAlloc(&s.trap_host, 8 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
        for (int i = 0; i < 8; ++i) s.trap_host[i] = 0;
        int* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, s.trap_host, 0) == cudaSuccess)
            cudaMemcpyToSymbol(g_trap_info, &dptr, sizeof dptr);     // the symbol has one instance per device
    } else {
        cudaGetLastError();
        s.trap_host = nullptr;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, dev);
    s.ok = true;
}
inline bool available() { return state().ok && state().enabled; }

template <typename T> struct TmType;
template <> struct TmType<__half> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; };
template <> struct TmType<__nv_bfloat16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };
template <> struct TmType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };

// 2-D map over a row-major [rows][cols] matrix, box = 128 bytes of columns (64 16-bit / 32 FP32 elements) x box_rows, 128B swizzle
template <typename T>
CUtensorMap make_map_2d(const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(T)), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = state().encode(&m, TmType<T>::v, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        throw Error(buf);
    }
    return m;
}

template <int TAPS, int CHUNKS, int NOUT, int WP, int EPI, int TMAST, typename TIn, typename TOut, int CG = 1, int BNS = 0>
void launch(cudaStream_t st, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o,
            const TcParams& p, int n_blocks_y) {
    constexpr int STAGES = pick_stages<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, (int)sizeof(TOut)>();
    static_assert(STAGES >= 2, "not enough shared memory for a 2-stage pipeline");
    constexpr int HALO = (TAPS == 9) ? (WP + 1) : 0;
    constexpr int R = ((TC_BM + 2 * HALO + 7) / 8) * 8;
    constexpr size_t smem = 1024 + (size_t)TAPS * CHUNKS * (NOUT / CG) * 128 + (size_t)STAGES * R * 128 +
                            (size_t)epi_smem_bytes<TAPS, CHUNKS, NOUT, WP, EPI, TMAST, CG, (int)sizeof(TOut)>() + 1280;
    DDPM_CHECK(p.g.Wp == WP && p.g.Hs == WP && p.g.npos + 2 * TC_BM < (1ll << 31),
               "conv_tc: geometry does not match the kernel's compile-time row width");
    auto kern = conv_tc_kernel<TAPS, CHUNKS, NOUT, WP, STAGES, EPI, TMAST, CG, BNS, TIn, TOut>;
    ensure_smem_attr(kern, smem);
    int ctas_x = state().num_sms / n_blocks_y;
    if (ctas_x > p.num_m_tiles) ctas_x = p.num_m_tiles;
    if (ctas_x < 1) ctas_x = 1;
    if (CG == 2) {
        // CTA pairs: clusters of two along x (the hardware co-schedules them on the two SMs of a TPC)
        ctas_x = ((ctas_x + 1) / 2) * 2;
        if (ctas_x > state().num_sms / n_blocks_y) ctas_x -= 2;
        if (ctas_x < 2) ctas_x = 2;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas_x, n_blocks_y);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
        // the kernel is persistent: never launch more pairs than the device can keep resident at once (a part with
        // an unpaired SM would otherwise run the surplus pair as a second wave and double the kernel time)
        const void* kkey = reinterpret_cast<const void*>(kern);
        auto mp = state().max_pairs.find(kkey);
        if (mp == state().max_pairs.end()) {
            cfg.attrs = attr;
            cfg.numAttrs = na;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = state().num_sms / 2; }
            mp = state().max_pairs.emplace(kkey, nc).first;
        }
        const int max_pairs = mp->second;
        if (ctas_x * n_blocks_y > 2 * max_pairs) {
            ctas_x = (2 * max_pairs / n_blocks_y) & ~1;
            if (ctas_x < 2) ctas_x = 2;
            cfg.gridDim = dim3(ctas_x, n_blocks_y);
        }
    }
    if (state().pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    DDPM_CUDA(cudaLaunchKernelEx(&cfg, kern, a0, a1, w, o, p));
    DDPM_LAUNCH_CHECK();
}

// Optional BatchNorm reductions fused into the epilogue of conv3x3 (see TcParams::stats).
struct BnFuse {
    int mode = 0;                 // 0 none, 1 forward statistics (sum y, sum y^2), 2 backward pass 1 (sum g, sum g*xhat)
    double* sums = nullptr;       // [2][C]
    int C = 0;                    // channels of the statistics vectors
    int nch = 0;                  // mode 2: output channels [0, nch) carry the BatchNorm
    const void* y = nullptr;      // mode 2: pre-BatchNorm tensor (position 0), y_cs channels per position
    int y_cs = 0;
    const float *scale = nullptr, *shift = nullptr, *mean = nullptr, *istd = nullptr;
};

// Conv((3,3), C0+C1 => Cout, pad=1) on the padded layout.  src pointers are POSITION 0 pointers;
// the tensor maps are based at the allocation start (position -guard).
// bn: reductions to fuse into the epilogue; *bn_done tells the caller whether the launched variant did them (only the
// default CTA-pair variants carry the fused code; otherwise the caller runs the stand-alone reduction kernel).
template <typename TIn, typename TOut>
bool conv3x3(cudaStream_t st, const TIn* s0, int C0, const TIn* s1, int C1, const TIn* Wt, int Cout, TOut* out,
             const Geo& g, const float* shift, int relu, const BnFuse* bn = nullptr, bool* bn_done = nullptr, int rev = 0) {
    if (bn_done) *bn_done = false;
    if (!available()) return false;
    if constexpr (std::is_same<TIn, float>::value && std::is_same<TOut, float>::value) {
        // TF32 mode: FP32 activations and weights in HBM, tcgen05.mma.kind::tf32, FP32 accumulate and output.  A 128-byte
        // K chunk is 32 channels; the resident weights of one CTA are capped at 147 KB, so wide layers split Cout
        // over blockIdx.y (each N-slice re-reads the slab from L2 -- this is the parity mode, not the fast one).
        const int Cin = C0 + C1, WP = g.Wp;
        if (s1 && C0 != C1) return false;
        const uint64_t rows = (uint64_t)g.alloc_positions();
        TcParams p{};
        p.out = out; p.out_cs = Cout; p.g = g; p.g_out = g; p.shift = shift; p.relu = relu;
        p.num_m_tiles = cdiv(g.npos, TC_BM);
        p.chunk1_src1 = (s1 != nullptr) ? 1 : 0;
        p.dbg = nullptr;
        constexpr int R32 = ((TC_BM + 2 * (WP_32 + 1) + 7) / 8) * 8, R16 = ((TC_BM + 2 * (WP_16 + 1) + 7) / 8) * 8;
        CUtensorMap a0 = make_map_2d<float>(s0 - (size_t)g.guard * C0, rows, C0, WP == WP_32 ? R32 : R16);
        CUtensorMap a1 = a0;
        if (s1) a1 = make_map_2d<float>(s1 - (size_t)g.guard * C1, rows, C1, WP == WP_32 ? R32 : R16);
        if (WP == WP_32 && Cin == 64 && Cout == 64) {
            CUtensorMap w = make_map_2d<float>(Wt, 64, 9 * 64, 64);
            launch<9, 2, 64, WP_32, 0, 0, float, float>(st, a0, a1, w, a0, p, 1);
        } else if (WP == WP_32 && Cin == 128 && Cout == 64) {
            CUtensorMap w = make_map_2d<float>(Wt, 64, 9 * 128, 32);
            launch<9, 4, 32, WP_32, 0, 0, float, float>(st, a0, a1, w, a0, p, 2);
        } else if (WP == WP_32 && Cin == 64 && Cout == 128) {
            CUtensorMap w = make_map_2d<float>(Wt, 128, 9 * 64, 64);
            launch<9, 2, 64, WP_32, 0, 0, float, float>(st, a0, a1, w, a0, p, 2);
        } else if (WP == WP_16 && Cin == 64 && Cout == 128) {
            CUtensorMap w = make_map_2d<float>(Wt, 128, 9 * 64, 64);
            launch<9, 2, 64, WP_16, 0, 0, float, float>(st, a0, a1, w, a0, p, 2);
        } else if (WP == WP_16 && Cin == 128 && Cout == 128) {
            CUtensorMap w = make_map_2d<float>(Wt, 128, 9 * 128, 32);
            launch<9, 4, 32, WP_16, 0, 0, float, float>(st, a0, a1, w, a0, p, 4);
        } else if (WP == WP_16 && Cin == 128 && Cout == 64) {
            CUtensorMap w = make_map_2d<float>(Wt, 64, 9 * 128, 32);
            launch<9, 4, 32, WP_16, 0, 0, float, float>(st, a0, a1, w, a0, p, 2);
        } else {
            return false;
        }
        return true;
    } else if constexpr (sizeof(TIn) != 2 || sizeof(TOut) != 2) {
        return false;
    } else {
    const int Cin = C0 + C1;
    const uint64_t rows = (uint64_t)g.alloc_positions();
    const int WP = g.Wp;
    TcParams p{};
    p.out = out; p.out_cs = Cout; p.g = g; p.g_out = g; p.shift = shift; p.relu = relu;
    p.num_m_tiles = cdiv(g.npos, TC_BM);
    p.chunk1_src1 = (s1 != nullptr) ? 1 : 0;
    p.rev = (state().reverse && rev) ? 1 : 0;
    p.dbg = state().dbg;
    const int bm = (bn && bn->mode && !p.dbg) ? bn->mode : 0;
    if (bm) {
        p.stats = bn->sums; p.stats_C = bn->C; p.stats_nch = bn->nch; p.bn_y = bn->y; p.bn_cs = bn->y_cs;
        p.bn_scale = bn->scale; p.bn_shift = bn->shift; p.bn_mean = bn->mean; p.bn_istd = bn->istd;
    }
    auto done = [&]() { if (bn_done) *bn_done = true; };
    const TIn* base0 = s0 - (size_t)g.guard * C0;
    constexpr int R32 = ((TC_BM + 2 * (WP_32 + 1) + 7) / 8) * 8, R16 = ((TC_BM + 2 * (WP_16 + 1) + 7) / 8) * 8;
    CUtensorMap a0 = make_map_2d<TIn>(base0, rows, C0, WP == WP_32 ? R32 : R16);
    CUtensorMap a1 = a0;
    if (s1) a1 = make_map_2d<TIn>(s1 - (size_t)g.guard * C1, rows, C1, WP == WP_32 ? R32 : R16);
    if (WP == WP_32 && Cin == 64 && Cout == 64 && !s1) {
        const bool pair = (state().pair_mask & 2) != 0;
        CUtensorMap w = make_map_2d<TIn>(Wt, 64, 9 * 64, pair ? 32 : 64);
        CUtensorMap o = make_map_2d<TOut>(out - (size_t)g.guard * Cout, rows, Cout, TC_BM);
        if (pair && bm == 1) { launch<9, 1, 64, WP_32, 0, 1, TIn, TOut, 2, 1>(st, a0, a1, w, o, p, 1); done(); }
        else if (pair && bm == 2) { launch<9, 1, 64, WP_32, 0, 1, TIn, TOut, 2, 2>(st, a0, a1, w, o, p, 1); done(); }
        else if (pair) launch<9, 1, 64, WP_32, 0, 1, TIn, TOut, 2>(st, a0, a1, w, o, p, 1);
        else if (state().tma_store) launch<9, 1, 64, WP_32, 0, 1, TIn, TOut>(st, a0, a1, w, o, p, 1);
        else launch<9, 1, 64, WP_32, 0, 0, TIn, TOut>(st, a0, a1, w, o, p, 1);
    } else if (WP == WP_32 && Cin == 128 && Cout == 64) {
        const bool pair = (state().pair_mask & 4) != 0;
        CUtensorMap w = make_map_2d<TIn>(Wt, 64, 9 * 128, pair ? 32 : 64);
        if (pair && bm == 1) { launch<9, 2, 64, WP_32, 0, 0, TIn, TOut, 2, 1>(st, a0, a1, w, a0, p, 1); done(); }
        else if (pair) launch<9, 2, 64, WP_32, 0, 0, TIn, TOut, 2>(st, a0, a1, w, a0, p, 1);
        else launch<9, 2, 64, WP_32, 0, 0, TIn, TOut>(st, a0, a1, w, a0, p, 1);
    } else if (WP == WP_16 && Cin == 64 && Cout == 128 && !s1) {
        const bool pair = (state().pair_mask & 8) != 0;
        CUtensorMap w = make_map_2d<TIn>(Wt, 128, 9 * 64, pair ? 64 : 128);
        if (pair && bm == 1) { launch<9, 1, 128, WP_16, 0, 0, TIn, TOut, 2, 1>(st, a0, a1, w, a0, p, 1); done(); }
        else if (pair) launch<9, 1, 128, WP_16, 0, 0, TIn, TOut, 2>(st, a0, a1, w, a0, p, 1);
        else launch<9, 1, 128, WP_16, 0, 0, TIn, TOut>(st, a0, a1, w, a0, p, 1);
    } else if (WP == WP_16 && Cin == 128 && Cout == 128 && !s1) {
        CUtensorMap w = make_map_2d<TIn>(Wt, 128, 9 * 128, 64);
        const bool pair = (state().pair_mask & 1) != 0;
        if (pair && bm == 1) { launch<9, 2, 128, WP_16, 0, 0, TIn, TOut, 2, 1>(st, a0, a1, w, a0, p, 1); done(); }
        else if (pair && bm == 2) { launch<9, 2, 128, WP_16, 0, 0, TIn, TOut, 2, 2>(st, a0, a1, w, a0, p, 1); done(); }
        else if (pair)
            launch<9, 2, 128, WP_16, 0, 0, TIn, TOut, 2>(st, a0, a1, w, a0, p, 1);  // CTA pair: each CTA stages 64 of the 128 weight rows
        else
            launch<9, 2, 64, WP_16, 0, 0, TIn, TOut>(st, a0, a1, w, a0, p, 2);      // Cout split over blockIdx.y so the weights fit
    } else if (WP == WP_16 && Cin == 128 && Cout == 64 && !s1) {
        const bool pair = (state().pair_mask & 16) != 0;
        CUtensorMap w = make_map_2d<TIn>(Wt, 64, 9 * 128, pair ? 32 : 64);         // dgrad of down2.conv1
        if (pair) launch<9, 2, 64, WP_16, 0, 0, TIn, TOut, 2>(st, a0, a1, w, a0, p, 1);
        else launch<9, 2, 64, WP_16, 0, 0, TIn, TOut>(st, a0, a1, w, a0, p, 1);
    } else if (WP == WP_32 && Cin == 64 && Cout == 128 && !s1) {
        const bool pair = (state().pair_mask & 16) != 0;
        CUtensorMap w = make_map_2d<TIn>(Wt, 128, 9 * 64, pair ? 64 : 128);        // dgrad of up1.conv1 (d cat)
        if (pair && bm == 2) { launch<9, 1, 128, WP_32, 0, 0, TIn, TOut, 2, 2>(st, a0, a1, w, a0, p, 1); done(); }
        else if (pair) launch<9, 1, 128, WP_32, 0, 0, TIn, TOut, 2>(st, a0, a1, w, a0, p, 1);
        else launch<9, 1, 128, WP_32, 0, 0, TIn, TOut>(st, a0, a1, w, a0, p, 1);
    } else {
        return false;
    }
    return true;
    }
}

// Last 3x3 conv (64=>64 @32x32) of the sampler with the final 1x1 conv and the reverse update fused in its epilogue
template <typename TIn>
bool conv3x3_final(cudaStream_t st, const TIn* s0, const TIn* Wt, const Geo& g, const float* shift, float* x, const float* z,
                   const float* wf, const float* bf, const float scal[4], int final_clamp, int rev = 0) {
    if (!available()) return false;
    if constexpr (sizeof(TIn) != 2) {
        return false;
    } else {
    if (g.Wp != WP_32) return false;
    TcParams p{};
    p.out = nullptr; p.out_cs = 64; p.g = g; p.g_out = g; p.shift = shift; p.relu = 1;
    p.num_m_tiles = cdiv(g.npos, TC_BM);
    p.chunk1_src1 = 0;
    p.rev = (state().reverse && rev) ? 1 : 0;
    p.dbg = nullptr;
    p.x = x; p.z = z; p.wf = wf; p.bf = bf;
    p.sig = scal[0]; p.sqa = scal[1]; p.sqp = scal[2]; p.sqv = scal[3];
    p.final_clamp = final_clamp;
    constexpr int R32 = ((TC_BM + 2 * (WP_32 + 1) + 7) / 8) * 8;
    CUtensorMap a0 = make_map_2d<TIn>(s0 - (size_t)g.guard * 64, (uint64_t)g.alloc_positions(), 64, R32);
    const bool pair = (state().pair_mask & 2) != 0;
    CUtensorMap w = make_map_2d<TIn>(Wt, 64, 9 * 64, pair ? 32 : 64);
    if (pair) launch<9, 1, 64, WP_32, 2, 0, TIn, TIn, 2>(st, a0, a0, w, a0, p, 1);
    else launch<9, 1, 64, WP_32, 2, 0, TIn, TIn>(st, a0, a0, w, a0, p, 1);
    return true;
    }
}

// ConvTranspose((2,2), 128 => 64, stride=2): [pos16][128] x Wt[q*64+co][128]^T, pixel-shuffle epilogue + bias
template <typename TA>
bool up2(cudaStream_t st, const TA* a6, const TA* Wt, TA* u, const Geo& gi, const Geo& go, const float* bias, int rev = 0) {
    if (!available()) return false;
    if constexpr (sizeof(TA) != 2) {
        return false;
    } else {
    TcParams p{};
    p.out = u; p.out_cs = 64; p.g = gi; p.g_out = go; p.shift = bias; p.relu = 0;
    p.num_m_tiles = cdiv(gi.npos, TC_BM);
    p.chunk1_src1 = 0;
    p.rev = (state().reverse && rev) ? 1 : 0;
    CUtensorMap a0 = make_map_2d<TA>(a6 - (size_t)gi.guard * 128, (uint64_t)gi.alloc_positions(), 128, TC_BM);
    CUtensorMap w = make_map_2d<TA>(Wt, 256, 128, 64);
    launch<1, 2, 64, WP_16, 1, 0, TA, TA>(st, a0, a0, w, a0, p, 4);
    return true;
    }
}

// out[pos][N=128] = A[pos][K=256] * Wt[n][k]^T over the valid positions of g (ConvTranspose data gradient
// on the pixel-un-shuffled gradient): the conv kernel with one tap and four 64-channel K chunks.
template <typename T>
bool gemm_rows(cudaStream_t st, const T* a, int K, const T* Wt, int Nout, T* out, const Geo& g, const BnFuse* bn = nullptr,
               bool* bn_done = nullptr) {
    if (bn_done) *bn_done = false;
    if (!available()) return false;
    if constexpr (sizeof(T) != 2) {
        return false;
    } else {
    if (K != 256 || Nout != 128) return false;
    TcParams p{};
    p.out = out; p.out_cs = Nout; p.g = g; p.g_out = g; p.shift = nullptr; p.relu = 0;
    p.num_m_tiles = cdiv(g.npos, TC_BM);
    p.chunk1_src1 = 0;
    p.dbg = nullptr;
    CUtensorMap a0 = make_map_2d<T>(a - (size_t)g.guard * K, (uint64_t)g.alloc_positions(), K, TC_BM);
    CUtensorMap w = make_map_2d<T>(Wt, Nout, K, 128);
    if (bn && bn->mode == 2) {
        p.stats = bn->sums; p.stats_C = bn->C; p.stats_nch = bn->nch; p.bn_y = bn->y; p.bn_cs = bn->y_cs;
        p.bn_scale = bn->scale; p.bn_shift = bn->shift; p.bn_mean = bn->mean; p.bn_istd = bn->istd;
        launch<1, 4, 128, WP_16, 0, 0, T, T, 1, 2>(st, a0, a0, w, a0, p, 1);
        if (bn_done) *bn_done = true;
    } else {
        launch<1, 4, 128, WP_16, 0, 0, T, T>(st, a0, a0, w, a0, p, 1);
    }
    return true;
    }
}

// =====================================================================================
// Weight gradient of a 3x3 convolution on tensor cores.
//   dWcc[co][ty,tx][ci] = sum_pos dy[pos][co] * x[pos + ty*Wp + tx][ci]      (halo rows of dy are zero)
// is a GEMM over the POSITION dimension whose operands are both "MN-major" (channels contiguous):
//   A'[k][(ty,co)] = dy[k - ty*Wp][co]     M = 128 = two kernel rows ty at once: the second 64-row group of
//                                          the descriptor is the SAME shared-memory slab, LBO = Wp rows further
//   B'[k][(tx,ci)] = x [k + tx    ][ci]     N = 192 = the three kernel columns tx: groups LBO = 1 row apart
// so per 16 positions two tcgen05.mma (M=128,N=192,K=16) produce all nine 64x64 tap blocks:
//   MMA1 -> (ty=+1 | ty=0) x (tx=-1,0,+1),   MMA2 -> (ty=-1 | unused) x (tx=-1,0,+1).
// Each CTA streams its share of the positions through a TMA ring (dy slab [KB+2Wp][64], x slab [KB+2][64]),
// accumulates in TMEM for the whole launch (2 x 192 columns) and writes ONE partial [9][64][64] FP32 block;
// wgrad_reduce_kernel sums the partials into the gradient arena in Flux layout (deterministic, no atomics).
// Replaces NNlib.∇conv_filter (im2col + SGEMM) behind Zygote's pullback (train_brain.jl:267-269).
constexpr int WG_KB = 128;        // positions per pipeline stage
constexpr int WG_THREADS = 192;

__host__ __device__ constexpr uint32_t make_idesc_mn(uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (afmt << 7) | (bfmt << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct WgParams {
    float* partial;      // [gridDim.y][gridDim.x][9][64][64]
    int num_kblocks;     // ceil(npos / WG_KB)
    int guard;
    int dy_chunk[4], x_chunk[4];   // per blockIdx.y: 64-channel chunk index of dy (co) and x (ci)
};

template <int WP, int STAGES, typename TDy, typename TX>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgParams p) {
    constexpr int RDY = ((WG_KB + 2 * WP + 7) / 8) * 8;    // dy slab rows: positions k0-Wp .. k0+KB+Wp
    constexpr int RX = ((WG_KB + 2 + 7) / 8) * 8;          // x slab rows:  positions k0-1  .. k0+KB+1
    constexpr uint32_t DY_BYTES = RDY * 128, X_BYTES = RX * 128, STAGE_BYTES = DY_BYTES + X_BYTES;
    constexpr uint32_t IDESC = make_idesc_mn(IsBf16<TDy>::v, IsBf16<TX>::v, 128, 192);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + STAGES * STAGE_BYTES;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (STAGES + s); };
    const uint32_t bar_acc = s_bar + 8u * (2 * STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDy); prefetch_tmap(&tmX);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int dyc = p.dy_chunk[blockIdx.y] * 64, xc = p.x_chunk[blockIdx.y] * 64;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (elect_one()) {
                const int k0 = kb * WG_KB + p.guard;
                mbar_expect_tx(bar_full(stage), STAGE_BYTES);
                tma_load_2d(s_base + stage * STAGE_BYTES, &tmDy, dyc, k0 - WP, bar_full(stage));
                tma_load_2d(s_base + stage * STAGE_BYTES + DY_BYTES, &tmX, xc, k0 - 1, bar_full(stage));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // MN-major, 128B swizzle: hi = SBO (8 K-rows = 1024 B) | version | SWIZZLE_128B ; lo = start>>4 | LBO<<16
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t LBO_A = (uint32_t)(WP * 128) >> 4;     // second M group = Wp positions further (ty one lower)
        constexpr uint32_t LBO_B = 128u >> 4;                     // N groups = consecutive positions (tx = -1, 0, +1)
        int stage = 0;
        uint32_t phase = 0;
        uint32_t first = 1;
        for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = (((s_base + stage * STAGE_BYTES) & 0x3FFFFu) >> 4);
                const uint32_t b0 = (((s_base + stage * STAGE_BYTES + DY_BYTES) & 0x3FFFFu) >> 4) | (LBO_B << 16);
#pragma unroll
                for (int ks = 0; ks < WG_KB / 16; ++ks) {
                    const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                    // rows of the dy slab: ty=+1 at +0, ty=0 at +Wp, ty=-1 at +2Wp (slab starts at k0-Wp)
                    umma_f16_lh(tmem_base, (a0 + ((ks * 16 * 128) >> 4)) | (LBO_A << 16), b0 + ((ks * 16 * 128) >> 4), DESC_HI,
                                IDESC, acc);
                    umma_f16_lh(tmem_base + 192, (a0 + (((2 * WP + ks * 16) * 128) >> 4)) | (LBO_B << 16),
                                b0 + ((ks * 16 * 128) >> 4), DESC_HI, IDESC, acc);
                }
                umma_commit(bar_empty(stage));
            }
            __syncwarp();
            first = 0;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
    } else {
        // epilogue: one partial block per CTA.  TMEM lane l: D1 -> (ty = l<64 ? +1 : 0, co = l%64), D2 lanes 0..63 -> ty=-1
        const int lane_grp = warp & 3;
        const int l = lane_grp * 32 + lane;
        const int co = l & 63;
        float* dst = p.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (9 * 64 * 64);
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const bool has_work = blockIdx.x < p.num_kblocks;   // always true (grid <= kblocks), kept for safety
#pragma unroll 1
        for (int d = 0; d < 2; ++d) {
            const int ty = (d == 0) ? (l < 64 ? 1 : 0) : -1;
            const bool live = has_work && (d == 0 || l < 64);
#pragma unroll 1
            for (int txi = 0; txi < 3; ++txi) {
                const int tap = (ty + 1) * 3 + txi;
                float* row = dst + ((size_t)tap * 64 + co) * 64;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + d * 192 + txi * 64 + c0, r);
                    tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(row + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// arena[flux index of (tap, co_off+co, ci_off+ci)] = alpha * sum_cta partial[y][cta][tap][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int n_cta, int n_sub, const int* __restrict__ sub_co,
                                    const int* __restrict__ sub_ci, int Cin_total, int ci_base, float alpha,
                                    float* __restrict__ dW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over n_sub * 9*64*64
    if (i >= n_sub * 36864) return;
    const int y = i / 36864, e = i - y * 36864;
    const int ci = e & 63, co = (e >> 6) & 63, tap = e >> 12;
    const float* src = partial + (size_t)y * n_cta * 36864 + e;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += src[(size_t)c * 36864];
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const long long idx = (long long)(1 - dx) + 3 * (1 - dy) + 9LL * (ci_base + sub_ci[y] * 64 + ci) +
                          9LL * Cin_total * (sub_co[y] * 64 + co);
    dW[idx] = alpha * s;
}

// Weight gradient of ConvTranspose((2,2), 128=>64, stride 2) on the un-shuffled gradient:
//   dWt[q*64+co][ci] = sum_pos du4[pos][q*64+co] * a6[pos][ci]
// Plain MN-major GEMM over coarse positions: M = 128 = two q blocks per MMA (second 64-row group = the next
// 64-channel slab, LBO = slab size), N = 128 = both 64-channel slabs of a6; two MMAs cover q = 0..3.
constexpr int WU_KB = 64;
template <int STAGES, typename T>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_up2_tc_kernel(const __grid_constant__ CUtensorMap tmDu4, const __grid_constant__ CUtensorMap tmA6, float* partial,
                    int num_kblocks, int guard) {
    constexpr uint32_t SLAB = WU_KB * 128;                 // [64 positions][64 ch]
    constexpr uint32_t STAGE_BYTES = 6 * SLAB;             // du4 chunks 0..3, a6 chunks 0..1
    constexpr uint32_t IDESC = make_idesc_mn(IsBf16<T>::v, IsBf16<T>::v, 128, 128);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + STAGES * STAGE_BYTES;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (STAGES + s); };
    const uint32_t bar_acc = s_bar + 8u * (2 * STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDu4); prefetch_tmap(&tmA6);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<256>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = blockIdx.x; kb < num_kblocks; kb += gridDim.x) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (elect_one()) {
                const int k0 = kb * WU_KB + guard;
                const uint32_t dst = s_base + stage * STAGE_BYTES;
                mbar_expect_tx(bar_full(stage), STAGE_BYTES);
                for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * SLAB, &tmDu4, c * 64, k0, bar_full(stage));
                for (int c = 0; c < 2; ++c) tma_load_2d(dst + (4 + c) * SLAB, &tmA6, c * 64, k0, bar_full(stage));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t LBO = SLAB >> 4;
        int stage = 0;
        uint32_t phase = 0, first = 1;
        for (int kb = blockIdx.x; kb < num_kblocks; kb += gridDim.x) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = (((s_base + stage * STAGE_BYTES) & 0x3FFFFu) >> 4) | (LBO << 16);
                const uint32_t b0 = (((s_base + stage * STAGE_BYTES + 4 * SLAB) & 0x3FFFFu) >> 4) | (LBO << 16);
#pragma unroll
                for (int ks = 0; ks < WU_KB / 16; ++ks) {
                    const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                    umma_f16_lh(tmem_base, a0 + ((ks * 16 * 128) >> 4), b0 + ((ks * 16 * 128) >> 4), DESC_HI, IDESC, acc);
                    umma_f16_lh(tmem_base + 128, a0 + ((2 * SLAB + ks * 16 * 128) >> 4), b0 + ((ks * 16 * 128) >> 4), DESC_HI,
                                IDESC, acc);
                }
                umma_commit(bar_empty(stage));
            }
            __syncwarp();
            first = 0;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
    } else {
        const int lane_grp = warp & 3;
        const int l = lane_grp * 32 + lane;
        float* dst = partial + (size_t)blockIdx.x * (256 * 128);
        mbar_wait(bar_acc, 0);
        tc_fence_after();
#pragma unroll 1
        for (int d = 0; d < 2; ++d) {
            const int row = (2 * d + (l >> 6)) * 64 + (l & 63);      // q*64 + co
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + d * 128 + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + (size_t)row * 128 + c0 + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// arena (Flux ConvTranspose layout w[a,b,co,ci], a = 1-px, b = 1-py) = alpha * sum_cta partial[cta][q*64+co][ci]
__global__ void wgrad_up2_reduce_kernel(const float* __restrict__ partial, int n_cta, float alpha, float* __restrict__ dW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 256 * 128) return;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += partial[(size_t)c * (256 * 128) + i];
    const int ci = i & 127, row = i >> 7, q = row >> 6, co = row & 63;
    const int py = q >> 1, px = q & 1;
    dW[(1 - px) + 2 * (1 - py) + 4 * co + 256 * ci] = alpha * s;
}

// Per-ENGINE scratch of the weight-gradient kernels (partial blocks of every CTA, sub-block tables): two handles on
// one device (or on two devices) must not share it.
struct WgScratch {
    float* partial = nullptr;
    size_t cap = 0;
    int* sub = nullptr;     // device [4][8]: co chunk [0..3], ci chunk [4..7] per (nco, nci) configuration
    void release() {
        if (partial) cudaFree(partial);
        if (sub) cudaFree(sub);
        partial = nullptr; sub = nullptr; cap = 0;
    }
};

// allocate the scratch once, outside any stream capture (wgrad3x3 / wgrad_up2 never allocate afterwards)
inline void wg_reserve(WgScratch& sc) {
    const size_t need = (size_t)4 * state().num_sms * 36864 * sizeof(float);
    if (sc.cap < need) {
        if (sc.partial) cudaFree(sc.partial);
        sc.partial = nullptr; sc.cap = 0;
        DDPM_CUDA(cudaMalloc(&sc.partial, need));
        sc.cap = need;
    }
    if (!sc.sub) {
        // all (co chunk, ci chunk) enumerations for nco x nci in {1,2}x{1,2}, indexed by (nco-1)*2 + (nci-1)
        static const int h[4][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 1, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0}, {0, 0, 1, 1, 0, 1, 0, 1}};
        DDPM_CUDA(cudaMalloc(&sc.sub, sizeof h));
        DDPM_CUDA(cudaMemcpy(sc.sub, h, sizeof h, cudaMemcpyHostToDevice));   // blocking copy from static storage
        DDPM_CUDA(cudaDeviceSynchronize());
    }
}

template <int WP, typename TDy, typename TX>
void launch_wgrad(cudaStream_t st, const CUtensorMap& mdy, const CUtensorMap& mx, const WgParams& p, int ctas_x, int n_sub) {
    constexpr int RDY = ((WG_KB + 2 * WP + 7) / 8) * 8, RX = ((WG_KB + 2 + 7) / 8) * 8;
    constexpr int STAGE = (RDY + RX) * 128;
    constexpr int STAGES = (227 * 1024 - 2048) / STAGE > 6 ? 6 : (227 * 1024 - 2048) / STAGE;
    constexpr size_t smem = 1024 + (size_t)STAGES * STAGE + 512;
    auto kern = wgrad_tc_kernel<WP, STAGES, TDy, TX>;
    ensure_smem_attr(kern, smem);
    kern<<<dim3(ctas_x, n_sub), WG_THREADS, smem, st>>>(mdy, mx, p);
    DDPM_LAUNCH_CHECK();
}

// dW of Conv((3,3)) for dy [pos][Cout] and x [pos][Cx] (one source of the input; ci_off = its offset in the
// layer's Cin_total input channels).  Writes (not accumulates) the Flux-layout block of the gradient arena.
template <typename TG, typename TA>
bool wgrad3x3(cudaStream_t st, WgScratch& sc, const TG* dy, int Cout, const TA* x, int Cx, const Geo& g, float* dW, int Cin_total,
              int ci_off, float alpha) {
    if (!available()) return false;
    // kind::f16 requires A and B of the SAME 16-bit format (mixed bf16 x f16 raises an illegal-instruction
    // fault on B200, measured in round 1), hence gradients share the activation format.
    if constexpr (sizeof(TG) != 2 || sizeof(TA) != 2 || !std::is_same<TG, TA>::value) {
        return false;
    } else {
    if ((Cout != 64 && Cout != 128) || (Cx != 64 && Cx != 128) || (g.Wp != WP_32 && g.Wp != WP_16)) return false;
    const int nco = Cout / 64, nci = Cx / 64, n_sub = nco * nci;
    const int num_kblocks = cdiv(g.npos, WG_KB);
    int ctas_x = state().num_sms / n_sub;
    if (ctas_x > num_kblocks) ctas_x = num_kblocks;
    if (ctas_x < 1) ctas_x = 1;
    wg_reserve(sc);                     // no-op after the first call (the engine reserves it when it builds a training set)
    DDPM_CHECK((size_t)n_sub * ctas_x * 36864 * sizeof(float) <= sc.cap, "wgrad scratch too small");
    const int cfg = (nco - 1) * 2 + (nci - 1);
    static const int hsub[4][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 1, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0}, {0, 0, 1, 1, 0, 1, 0, 1}};
    WgParams p{};
    p.partial = sc.partial; p.num_kblocks = num_kblocks; p.guard = g.guard;
    for (int y = 0; y < 4; ++y) { p.dy_chunk[y] = hsub[cfg][y]; p.x_chunk[y] = hsub[cfg][4 + y]; }
    const uint64_t rows = (uint64_t)g.alloc_positions();
    constexpr int RX = ((WG_KB + 2 + 7) / 8) * 8;
    const int RDY = ((WG_KB + 2 * g.Wp + 7) / 8) * 8;
    CUtensorMap mdy = make_map_2d<TG>(dy - (size_t)g.guard * Cout, rows, Cout, RDY);
    CUtensorMap mx = make_map_2d<TA>(x - (size_t)g.guard * Cx, rows, Cx, RX);
    if (g.Wp == WP_32) launch_wgrad<WP_32, TG, TA>(st, mdy, mx, p, ctas_x, n_sub);
    else launch_wgrad<WP_16, TG, TA>(st, mdy, mx, p, ctas_x, n_sub);
    const int total = n_sub * 36864;
    wgrad_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(sc.partial, ctas_x, n_sub, sc.sub + cfg * 8, sc.sub + cfg * 8 + 4,
                                                          Cin_total, ci_off, alpha, dW);
    DDPM_LAUNCH_CHECK();
    return true;
    }
}

template <typename T>
bool wgrad_up2(cudaStream_t st, WgScratch& sc, const T* du4, const T* a6, const Geo& g, float* dW, float alpha) {
    if (!available()) return false;
    if constexpr (sizeof(T) != 2) {
        return false;
    } else {
    const int num_kblocks = cdiv(g.npos, WU_KB);
    int ctas = state().num_sms;
    if (ctas > num_kblocks) ctas = num_kblocks;
    wg_reserve(sc);                     // shared with wgrad3x3 (>= ctas*256*128 floats)
    constexpr int STAGES = 4;
    constexpr size_t smem = 1024 + (size_t)STAGES * 6 * WU_KB * 128 + 512;
    auto kern = wgrad_up2_tc_kernel<STAGES, T>;
    ensure_smem_attr(kern, smem);
    const uint64_t rows = (uint64_t)g.alloc_positions();
    CUtensorMap m4 = make_map_2d<T>(du4 - (size_t)g.guard * 256, rows, 256, WU_KB);
    CUtensorMap m6 = make_map_2d<T>(a6 - (size_t)g.guard * 128, rows, 128, WU_KB);
    kern<<<ctas, WG_THREADS, smem, st>>>(m4, m6, sc.partial, num_kblocks, g.guard);
    wgrad_up2_reduce_kernel<<<cdiv(256 * 128, 256), 256, 0, st>>>(sc.partial, ctas, alpha, dW);
    DDPM_LAUNCH_CHECK();
    return true;
    }
}

}  // namespace tc
}  // namespace ddpm
