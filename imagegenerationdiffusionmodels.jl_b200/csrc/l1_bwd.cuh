// Backward of the first convolution, 16-bit gradient tensors: dy of down1.conv1 (64 channels @32x32) ->
//   * Ccls[n][cls][co]  per-image sums of dy over the nine border classes (the folded embedding channels'
//                       gradient is S[n] (x) pe[ts[n]], kernels.cuh: l1_tap_sums_kernel + the embedding GEMM)
//   * wimg_acc[tap*64+co] += sum_{n,p} dy[n,p,co] * x[n, p+shift(tap)]     (image-channel weights, Float64)
// Replaces Zygote's pullback of Conv((3,3), 129=>64) on cat(x, tile(t_emb)) (/root/reference/src/train_brain.jl:111,
// 164-168,267-269) for the K = 9 image part and the constant-channel part.
//
// HBM-bound: the kernel must stream dy once (128 KB per image) while doing 72 FMAs per loaded 16-byte vector.  The
// first version issued its loads from the FMA loop (one block per SM, eight warps, one dependent global load per row):
// 331 us for 2048 images = 0.8 TB/s.  Here a block stages dy through shared memory with bulk asynchronous copies
// (cp.async.bulk, the 1-D TMA path): eight image rows = 264 consecutive padded positions = 33 KB per copy, double
// buffered on two mbarriers, so the copy of chunk k+1 is in flight while chunk k is consumed from shared memory.
#pragma once
#include "conv_tc.cuh"

namespace ddpm {

__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int L1B_ROWS = 8;                                  // image rows per staged chunk
constexpr int L1B_POS = L1B_ROWS * WP_32;                    // padded positions per chunk
constexpr int L1B_CHUNK_BYTES = L1B_POS * 64 * 2;            // 33,792
constexpr size_t L1B_SMEM = 1024 + 2 * (size_t)L1B_CHUNK_BYTES + 34 * 34 * 4 + 32 * 3 * 64 * 4 + 64;

template <typename T>
__global__ void __launch_bounds__(256, 1)
l1_bwd_bulk_kernel(const T* __restrict__ dy0 /*position 0, 64 channels per position*/, Geo g, const float* __restrict__ x,
                   float* __restrict__ Ccls, double* __restrict__ wimg_acc) {
    constexpr int H = 32, W = 32, TW = W + 2, PW = WP_32, CPI = H / L1B_ROWS;     // x window width, position row stride, chunks per image
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t s_buf = tc::smem_u32(smem);
    float* tile = reinterpret_cast<float*>(smem + 2 * L1B_CHUNK_BYTES);
    float* red = tile + 34 * 34;                                       // [32 w][3 row classes][64]; reused as [8][576]
    const uint32_t s_bar = tc::smem_u32(red + 32 * 3 * 64);
    const int t = threadIdx.x;
    const int c0 = (t & 7) * 8, w = t >> 3;
    if (t == 0) {
        tc::mbar_init(s_bar, 1);
        tc::mbar_init(s_bar + 8, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int n_mine = (g.N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // images of this block
    const int n_chunks = n_mine * CPI;
    auto issue = [&](int k) {                                          // chunk k of this block's sequence -> buffer k&1
        const int n = blockIdx.x + (k / CPI) * gridDim.x, c = k % CPI;
        const long long p0 = (long long)(n * (H + 1) + 1 + c * L1B_ROWS) * PW;
        tc::mbar_expect_tx(s_bar + 8 * (k & 1), L1B_CHUNK_BYTES);
        bulk_load_1d(s_buf + (k & 1) * L1B_CHUNK_BYTES, dy0 + p0 * 64, L1B_CHUNK_BYTES, s_bar + 8 * (k & 1));
    };
    if (t == 0 && n_chunks > 0) issue(0);
    float tw[9][8];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) tw[a][j] = 0.f;
    float top[8], mid[8], bot[8];
    for (int k = 0; k < n_chunks; ++k) {
        const int n = blockIdx.x + (k / CPI) * gridDim.x, c = k % CPI;
        if (c == 0) {
            // new image: zero-padded FP32 input window (taps become unconditional broadcast loads)
            const float* xi = x + (long long)n * H * W;
            for (int i = t; i < (H + 2) * TW; i += 256) {
                const int rr = i / TW, cc = i - rr * TW;
                const int hh = rr - 1, ww = cc - 1;
                tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { top[j] = 0.f; mid[j] = 0.f; bot[j] = 0.f; }
        }
        __syncthreads();                 // tile ready; every thread is done with buffer (k+1)&1 (chunk k-1)
        if (t == 0 && k + 1 < n_chunks) issue(k + 1);
        tc::mbar_wait(s_bar + 8 * (k & 1), (uint32_t)(k >> 1) & 1u, 40);
        const uint8_t* buf = smem + (k & 1) * L1B_CHUNK_BYTES;
#pragma unroll
        for (int hl = 0; hl < L1B_ROWS; ++hl) {
            const int h = c * L1B_ROWS + hl;
            float d[8];
            V8<T>::ld(reinterpret_cast<const T*>(buf + (size_t)(hl * PW + w + 1) * 128) + c0, d);
            const float* tp = tile + h * TW + w;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float xv = tp[(tap / 3) * TW + (tap % 3)];
#pragma unroll
                for (int j = 0; j < 8; ++j) tw[tap][j] = fmaf(d[j], xv, tw[tap][j]);
            }
            if (h == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) top[j] = d[j];
            } else if (h == H - 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bot[j] = d[j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) mid[j] += d[j];
            }
        }
        if (c == CPI - 1) {
            // image complete: nine border-class sums (column class of a thread is fixed: w == 0 | 1..30 | 31)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                red[(w * 3 + 0) * 64 + c0 + j] = top[j];
                red[(w * 3 + 1) * 64 + c0 + j] = mid[j];
                red[(w * 3 + 2) * 64 + c0 + j] = bot[j];
            }
            __syncthreads();
            for (int i = t; i < 576; i += 256) {
                const int co = i & 63, cls = i >> 6, rc = cls / 3, cc = cls - rc * 3;
                float acc;
                if (cc == 0) acc = red[(0 * 3 + rc) * 64 + co];
                else if (cc == 2) acc = red[((W - 1) * 3 + rc) * 64 + co];
                else {
                    acc = 0.f;
                    for (int ww = 1; ww < W - 1; ++ww) acc += red[(ww * 3 + rc) * 64 + co];
                }
                Ccls[(long long)n * 576 + i] = acc;
            }
        }
    }
    __syncthreads();
    // image-channel weight gradient: combine the 4 pixel columns of a warp, then the 8 warps
    const int warp = t >> 5;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = tw[tap][j];
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            if ((t & 31) < 8) red[warp * 576 + tap * 64 + c0 + j] = a;
        }
    __syncthreads();
    for (int i = t; i < 576; i += 256) {
        float acc = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) acc += red[w8 * 576 + i];
        atomicAdd(&wimg_acc[i], (double)acc);
    }
}

template <typename T>
void launch_l1_bwd_bulk(cudaStream_t st, const T* dy0, const Geo& g, const float* x, float* Ccls, double* wimg_acc) {
    auto kern = l1_bwd_bulk_kernel<T>;
    tc::ensure_smem_attr(kern, L1B_SMEM);
    int blocks = tc::state().num_sms;
    if (blocks > g.N) blocks = g.N;
    kern<<<blocks, 256, L1B_SMEM, st>>>(dy0, g, x, Ccls, wimg_acc);
    DDPM_LAUNCH_CHECK();
}

}  // namespace ddpm
