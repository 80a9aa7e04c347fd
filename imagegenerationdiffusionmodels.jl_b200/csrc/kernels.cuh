// HBM-bound kernels of the DDPM path: forward noising, the first (1+128 channel) convolution
// with the folded timestep embedding, BatchNorm statistics / apply / backward, pooling, the final
// 1x1 convolution fused with the reverse-diffusion update, MSE, Adam, weight packing, Philox.
// All multi-channel tensors use the padded [position][channel] layout of common.cuh and are
// accessed 8 channels (16 B in 16-bit modes) per thread, channel-fastest => fully coalesced.
#pragma once
#include "common.cuh"

namespace ddpm {

// ------------------------------------------------------------------------------------ 8-wide access
template <typename T> struct V8;
template <> struct V8<float> {
    static __device__ __forceinline__ void ld(const float* p, float o[8]) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* p, const float o[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
};
template <> struct V8<__half> {
    static __device__ __forceinline__ void ld(const __half* p, float o[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(__half* p, const float o[8]) {
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float o[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float o[8]) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};

// ------------------------------------------------------------------------------------ Philox-4x32-10
struct Philox {
    static __device__ __forceinline__ uint4 gen(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u;
            k.y += 0xBB67AE85u;
        }
        return c;
    }
    // four N(0,1) draws for pixel quad `quad` of global image `img` at `step` (oracle: device_normal)
    static __device__ __forceinline__ float4 normal4(unsigned long long seed, unsigned long long img,
                                                     uint32_t step, uint32_t quad) {
        uint4 r = gen(make_uint4(quad, (uint32_t)img, (uint32_t)(img >> 32), step),
                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float s = 2.3283064365386963e-10f;  // 2^-32
        float u0 = ((float)r.x + 0.5f) * s, u1 = ((float)r.y + 0.5f) * s;
        float u2 = ((float)r.z + 0.5f) * s, u3 = ((float)r.w + 0.5f) * s;
        // (float)r + 0.5f can round up to 2^32 -> u == 1 -> log(1) = 0: harmless; u > 0 always.
        float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
        float s0, c0, s1, c1;
        sincospif(2.f * u1, &s0, &c0);
        sincospif(2.f * u3, &s1, &c1);
        return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
    }
};

// x[n][p] = N(0,1), keyed by (seed, first_index + n, step)
__global__ void randn_kernel(float* __restrict__ x, long long n_img, int hw, unsigned long long seed,
                             long long first_index, uint32_t step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // quad index
    int qpi = hw / 4;
    if (i >= n_img * qpi) return;
    long long n = i / qpi;
    int q = (int)(i - n * qpi);
    float4 z = Philox::normal4(seed, (unsigned long long)(first_index + n), step, (uint32_t)q);
    *reinterpret_cast<float4*>(x + n * hw + 4 * q) = z;
}

// same with (seed, first_index) read from device memory, so one captured graph serves every chunk
__global__ void randn_dev_kernel(float* __restrict__ x, long long n_img, int hw, const unsigned long long* __restrict__ rng_dev,
                                 uint32_t step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qpi = hw / 4;
    if (i >= n_img * qpi) return;
    long long n = i / qpi;
    int q = (int)(i - n * qpi);
    float4 z = Philox::normal4(rng_dev[0], rng_dev[1] + (unsigned long long)n, step, (uint32_t)q);
    *reinterpret_cast<float4*>(x + n * hw + 4 * q) = z;
}

// ts[n] ~ U{1..T}: mulhi of a Philox word (step-keyed, component by image index)
__global__ void randint_ts_kernel(int* __restrict__ ts, int B, int T, unsigned long long seed,
                                  long long first_index, uint32_t step) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B) return;
    unsigned long long img = (unsigned long long)(first_index + n);
    uint4 r = Philox::gen(make_uint4(0xFFFFFFFFu, (uint32_t)img, (uint32_t)(img >> 32), step),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    ts[n] = 1 + (int)__umulhi(r.x, (uint32_t)T);
}

// training-step draws with (seed, first global image index, step) read from device memory (rng3), so that one
// captured training iteration serves every step: ts as randint_ts_kernel, eps as randn_kernel with the eps stream's
// seed (seed ^ 0x9E3779B97F4A7C15)
__global__ void randint_ts_dev_kernel(int* __restrict__ ts, int B, int T, const unsigned long long* __restrict__ rng3) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B) return;
    const unsigned long long seed = rng3[0], img = rng3[1] + (unsigned long long)n;
    uint4 r = Philox::gen(make_uint4(0xFFFFFFFFu, (uint32_t)img, (uint32_t)(img >> 32), (uint32_t)rng3[2]),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    ts[n] = 1 + (int)__umulhi(r.x, (uint32_t)T);
}
__global__ void randn_train_dev_kernel(float* __restrict__ x, long long n_img, int hw, const unsigned long long* __restrict__ rng3) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qpi = hw / 4;
    if (i >= n_img * qpi) return;
    long long n = i / qpi;
    int q = (int)(i - n * qpi);
    float4 z = Philox::normal4(rng3[0] ^ 0x9E3779B97F4A7C15ull, rng3[1] + (unsigned long long)n, (uint32_t)rng3[2], (uint32_t)q);
    *reinterpret_cast<float4*>(x + n * hw + 4 * q) = z;
}

// ------------------------------------------------------------------------------------ K1: q_sample
// x_t = a_n*x0 + b_n*eps with a = sqrt(acum[t]), b = sqrt(1-acum[t]) from host-built Float32 tables.
// Separately rounded multiplies and add (no FMA contraction) => bit-exact with the reference's
// broadcast `a .* x0 .+ b .* ϵ` (/root/reference/src/train_brain.jl:230-233).
// idx != nullptr gathers x0 rows from a resident dataset.
__global__ void qsample_kernel(const float* __restrict__ x0, const int* __restrict__ idx, const float* __restrict__ eps,
                               const int* __restrict__ ts, const float* __restrict__ sqrt_ac,
                               const float* __restrict__ sqrt_1mac, float* __restrict__ xt, long long B, int hw) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
    int v = hw / 4;
    if (i >= B * v) return;
    long long n = i / v;
    int q = (int)(i - n * v);
    int t = ts[n] - 1;
    float a = sqrt_ac[t], b = sqrt_1mac[t];
    long long src = idx ? (long long)idx[n] : n;
    float4 x = *reinterpret_cast<const float4*>(x0 + src * hw + 4 * q);
    float4 e = *reinterpret_cast<const float4*>(eps + n * hw + 4 * q);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(b, e.x));
    o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(b, e.y));
    o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(b, e.z));
    o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(b, e.w));
    *reinterpret_cast<float4*>(xt + n * hw + 4 * q) = o;
}

// final `clamp.(x_t, -1f0, 1f0)` of generate_image (/root/reference/src/generate_images.jl:242)
__global__ void clamp_kernel(float* __restrict__ x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = fminf(fmaxf(x[i], -1.f), 1.f);
}

// apply_noise (/root/reference/src/ImageGenerationDiffusionModels.jl:60-73): Float64 recurrence
__global__ void apply_noise_f64_kernel(const double* __restrict__ img, const double* __restrict__ eps, long long n,
                                       const double* __restrict__ sa, const double* __restrict__ sb, int nb,
                                       double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = img[i], e = eps[i];
    for (int k = 0; k < nb; ++k) x = __dadd_rn(__dmul_rn(sa[k], x), __dmul_rn(sb[k], e));
    out[i] = x;
}

// ------------------------------------------------------------------------------------ first convolution
// down1.conv1 = Conv((3,3), 1+128 => 64, pad=1) on cat(x, tile(t_emb)) (train_brain.jl:111,164-168).
// The 128 tiled embedding channels are constant over the image, so their contribution is a
// per-(timestep, border-class, cout) constant Ecls (border class = which taps fall inside the
// image); only the image channel is convolved (K = 9), in FP32 on CUDA cores.
//   y = (sum_tap x[h+dy,w+dx]*Wimg[tap][co] + Ecls[t][cls][co]) * scale[co] + shift[co]
// x is the unpadded boundary layout [N][H][W].
constexpr int CONV1_PIX_PER_BLOCK = 256;

// Thread layout: 16 lanes per pixel, 4 output channels per lane (16 pixels per 256-thread pass).  The 36
// image-channel weights, scale and shift of a lane's 4 channels live in registers for the whole block; keeping the
// per-thread state small (vs 8 channels = 146 registers, 1 block/SM) triples the resident warps of this
// latency-bound kernel.  H = W = 32 is a compile-time constant (the engine rejects other sizes).
template <typename T> __device__ __forceinline__ void st4(T* p, const float o[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float o[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void st4<__half>(__half* p, const float o[4]) {
    uint2 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
    h[0] = __floats2half2_rn(o[0], o[1]); h[1] = __floats2half2_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(p) = v;
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float o[4]) {
    uint2 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    h[0] = __floats2bfloat162_rn(o[0], o[1]); h[1] = __floats2bfloat162_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(p) = v;
}

// One block = 8 full rows (256 pixels) of one image.  The block first stages the zero-padded 10 x 34 input
// window in shared memory, so the nine taps are unconditional broadcast LDS (the kernel was issue-bound on
// per-tap bounds checks: 175 instructions per pixel-thread, profiles/README.md).  Thread layout: 16 lanes per
// pixel, 4 output channels per lane, weights/scale/shift of the lane's channels in registers.
template <typename TA>
__global__ void __launch_bounds__(256)
conv1_kernel(const float* __restrict__ x, const int* __restrict__ ts, int t_fixed, const float* __restrict__ Wimg,
             const float* __restrict__ Ecls, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
             View<TA> out, Geo g, double* __restrict__ stats) {
    constexpr int H = 32, W = 32, HW = H * W, ROWS = CONV1_PIX_PER_BLOCK / W, TW = W + 2;
    __shared__ float tile[(ROWS + 2) * TW];
    __shared__ float red[2][64];
    const int t = threadIdx.x;
    const long long pbeg = (long long)blockIdx.x * CONV1_PIX_PER_BLOCK;
    const int n = (int)(pbeg >> 10);
    const int h0 = (int)((pbeg & (HW - 1)) >> 5);
    if (n >= g.N) return;
    const float* xi = x + (long long)n * HW;
    for (int i = t; i < (ROWS + 2) * TW; i += 256) {
        const int rr = i / TW, cc = i - rr * TW;
        const int hh = h0 + rr - 1, ww = cc - 1;
        tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
    }
    if (t < 64) { red[0][t] = 0.f; red[1][t] = 0.f; }
    const int cg = (t & 15) * 4;
    float wr[9][4], sc[4], sh[4];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(Wimg + tap * 64 + cg));
        wr[tap][0] = a.x; wr[tap][1] = a.y; wr[tap][2] = a.z; wr[tap][3] = a.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = scale ? scale[cg + j] : 1.f;
        sh[j] = shift ? shift[cg + j] : 0.f;
    }
    const int trow = ts ? (ts[n] - 1) : (t_fixed - 1);
    const float* erow = Ecls + (long long)trow * 9 * 64 + cg;
    __syncthreads();
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    TA* obase = out.p + ((long long)(n * g.Hs + 1 + h0) * g.Wp + 1) * out.cs + cg;
#pragma unroll 2
    for (int lp = (t >> 4); lp < CONV1_PIX_PER_BLOCK; lp += 16) {
        const int hl = lp >> 5, w = lp & 31, h = h0 + hl;
        const float* tp = tile + hl * TW + w;
        float xv[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) xv[tap] = tp[(tap / 3) * TW + (tap % 3)];
        const int cls = (h == 0 ? 0 : (h == H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
        const float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + cls * 64));
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = ev[j];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) acc = fmaf(xv[tap], wr[tap][j], acc);
            const float v = acc * sc[j] + sh[j];
            s1[j] += v; s2[j] = fmaf(v, v, s2[j]);
            o[j] = relu ? fmaxf(v, 0.f) : v;
        }
        st4<TA>(obase + (long long)(hl * g.Wp + w) * out.cs, o);
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&red[0][cg + j], s1[j]); atomicAdd(&red[1][cg + j], s2[j]); }
        __syncthreads();
        if (t < 64) { atomicAdd(&stats[t], (double)red[0][t]); atomicAdd(&stats[64 + t], (double)red[1][t]); }
    }
}

// Ecls[t][cls][co] = sum over the taps that are inside the image for border class cls of P[t][tap][co]
__global__ void emb_class_sums_kernel(const float* __restrict__ P, float* __restrict__ Ecls, int T, int Cout) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 9 * Cout) return;
    int co = i % Cout, cls = (i / Cout) % 9, t = i / (9 * Cout);
    int rc = cls / 3, cc = cls % 3;  // 0 = first row/col, 1 = interior, 2 = last
    float s = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        bool ok = !(rc == 0 && dy < 0) && !(rc == 2 && dy > 0) && !(cc == 0 && dx < 0) && !(cc == 2 && dx > 0);
        if (ok) s += P[((long long)t * 9 + tap) * Cout + co];
    }
    Ecls[i] = s;
}

// ------------------------------------------------------------------------------------ one-shot all-reduce over peer memory
// SyncBN needs 20 all-reduces of 128..256 doubles per training step, each on the critical path between a convolution
// and the BatchNorm that follows it.  As NCCL calls they cost a launch plus a multi-hop protocol each; here the
// exchange is a few lines INSIDE the kernel that consumes the result (bn_finalize_kernel / bn_bwd_means_kernel):
// every rank stores its vector straight into every peer's mailbox over NVLink (peer pointers from cudaIpc handles),
// publishes a flag, waits for the world's flags in its own mailbox and sums the rows in rank order -- bit-identical
// on every rank.  Mailbox of one rank:  data[parity][slot][src rank][384] doubles, flag[parity][slot][src rank] u32.
// `slot` identifies the call site inside a step (forward layer l: l-1, backward layer l: 10+l-1), the epoch of a slot
// (how often it ran) comes from a per-slot device counter, so the kernel arguments never change (CUDA-graph safe);
// the epoch's parity double-buffers the mailbox.  A rank can only be one slot ahead of the slowest rank (it needs
// everybody's flag to leave a slot), so a mailbox entry is never overwritten before its reader is done.
constexpr int XR_SLOTS = 24;
constexpr int XR_MAX_WORLD = 16;
constexpr int XR_VEC = 384;
struct XReduce {
    double* peer[XR_MAX_WORLD];     // base of every rank's mailbox in THIS process's address space (peer[rank] = own)
    unsigned* epoch;                // [XR_SLOTS] per-slot run counters (device, private)
    int rank, world;
    __host__ __device__ static size_t data_doubles() { return (size_t)2 * XR_SLOTS * XR_MAX_WORLD * XR_VEC; }
    __host__ __device__ static size_t bytes() { return data_doubles() * sizeof(double) + (size_t)2 * XR_SLOTS * XR_MAX_WORLD * sizeof(unsigned); }
};
__device__ __forceinline__ void xr_store_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned xr_load_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Block-wide (one block, >= max(n, world) threads): out[i] = sum over ranks of local[i], i < n <= 384.
__device__ __forceinline__ void xr_allreduce_block(const XReduce& x, int slot, const double* __restrict__ local, int n,
                                                   double* __restrict__ out) {
    const int t = threadIdx.x;
    const unsigned e = x.epoch[slot];
    const int par = (int)(e & 1u);
    const size_t row = ((size_t)(par * XR_SLOTS + slot) * XR_MAX_WORLD);
    if (t < n) {
        const double v = local[t];
        for (int r = 0; r < x.world; ++r) x.peer[r][(row + x.rank) * XR_VEC + t] = v;       // NVLink stores
    }
    __threadfence_system();
    __syncthreads();
    if (t < x.world) {
        unsigned* fl = reinterpret_cast<unsigned*>(x.peer[t] + XReduce::data_doubles());
        xr_store_release_sys(fl + row + x.rank, e + 1u);
    }
    if (t < x.world) {
        const unsigned* mine = reinterpret_cast<const unsigned*>(x.peer[x.rank] + XReduce::data_doubles()) + row + t;
        unsigned long long spins = 0;
        while (xr_load_acquire_sys(mine) != e + 1u) {
            if (++spins > (1ull << 22)) __trap();       // a lost peer must fault (after a few seconds), never hang the GPU
        }
    }
    __syncthreads();
    if (t < n) {
        const volatile double* d = x.peer[x.rank] + row * XR_VEC + t;
        double s = 0.0;
        for (int r = 0; r < x.world; ++r) s += d[(size_t)r * XR_VEC];
        out[t] = s;
    }
    __syncthreads();
    if (t == 0) x.epoch[slot] = e + 1u;
}

// ------------------------------------------------------------------------------------ BatchNorm forward
// Flux BatchNorm(c, relu), eps=1e-5, momentum=0.1.
// sums = [sum y | sum y^2] (Float64, accumulated from the FP32 conv accumulators).
// xr_slot >= 0: `sums` are this rank's LOCAL sums; the kernel first all-reduces them over peer memory into gsums
// (SyncBN: the statistics of the GLOBAL batch, `count` = global element count) -- compute step and collective in one
// launch.  xr_slot < 0: `sums` are used as they are.  One block of 256 threads.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ run_mu, float* __restrict__ run_var,
                   float* __restrict__ mean_o, float* __restrict__ istd_o, float* __restrict__ scale_o,
                   float* __restrict__ shift_o, int C, float eps, float momentum, int update_running,
                   XReduce xr, int xr_slot, double* __restrict__ gsums) {
    if (xr_slot >= 0) {
        xr_allreduce_block(xr, xr_slot, sums, 2 * C, gsums);
        sums = gsums;
    }
    int c = threadIdx.x;
    if (c >= C) return;
    double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;
    if (var < 0) var = 0;
    float meanf = (float)mean, varf = (float)var;
    float istd = 1.f / sqrtf(varf + eps);
    float sc = gamma[c] * istd;
    mean_o[c] = meanf; istd_o[c] = istd; scale_o[c] = sc; shift_o[c] = beta[c] - sc * meanf;
    if (update_running) {
        float corr = (float)(count / (count - 1.0));
        run_mu[c] = (1.f - momentum) * run_mu[c] + momentum * meanf;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * (corr * varf);
    }
}

// inference: BN folded into an affine epilogue: scale = gamma/sqrt(var_run+eps), shift = (b-mu_run)*scale + beta
__global__ void bn_inference_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                           const float* __restrict__ mu, const float* __restrict__ var,
                                           const float* __restrict__ bias, float* __restrict__ scale_o,
                                           float* __restrict__ shift_o, int C, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sc = gamma[c] / sqrtf(var[c] + eps);
    scale_o[c] = sc;
    shift_o[c] = (bias[c] - mu[c]) * sc + beta[c];
}

// a = relu(y*scale + shift) over valid pixels
template <typename TA>
__global__ void __launch_bounds__(256)
bn_apply_kernel(View<const TA> y, View<TA> a, Geo g, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                int rev = 0) {
    // rev: blocks walk the tensor from its end.  The conv that produced y wrote it front to back, so its tail is what the
    // L2 still holds; this kernel then leaves the FRONT of a in L2, where the next (front-to-back) conv starts reading
    const int groups = C / 8;
    long long idx = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 256 + threadIdx.x;
    long long pix = idx / groups;
    int c0 = (int)(idx - pix * groups) * 8;
    if (pix >= (long long)g.N * g.H * g.W) return;
    int n = (int)(pix / (g.H * g.W));
    int rem = (int)(pix - (long long)n * g.H * g.W);
    long long p = g.pos(n, rem / g.W, rem % g.W);
    float v[8];
    V8<TA>::ld(y.p + p * y.cs + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], scale[c0 + j], shift[c0 + j]), 0.f);
    V8<TA>::st(a.p + p * a.cs + c0, v);
}

// a = relu(y*scale+shift) at the four pixels of a 2x2 window and pooled = max of them
// (BatchNorm(relu) followed by MaxPool((2,2)), train_brain.jl:114,117).  y == nullptr-scale variant:
// if scale == nullptr the input is already activated (inference: plain max-pool of h1).
template <typename TA>
__global__ void __launch_bounds__(256)
bn_apply_pool_kernel(View<const TA> y, View<TA> a, View<TA> pooled, Geo gf, Geo gc, int C,
                     const float* __restrict__ scale, const float* __restrict__ shift, int rev = 0) {
    const int groups = C / 8;
    long long idx = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 256 + threadIdx.x;
    long long pix = idx / groups;
    int c0 = (int)(idx - pix * groups) * 8;
    if (pix >= (long long)gc.N * gc.H * gc.W) return;
    int n = (int)(pix / (gc.H * gc.W));
    int rem = (int)(pix - (long long)n * gc.H * gc.W);
    int i = rem / gc.W, j = rem % gc.W;
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        long long p = gf.pos(n, 2 * i + (q >> 1), 2 * j + (q & 1));
        float v[8];
        V8<TA>::ld(y.p + p * y.cs + c0, v);
        if (scale) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], scale[c0 + k], shift[c0 + k]), 0.f);
            V8<TA>::st(a.p + p * a.cs + c0, v);
            // pooled value must equal the max of the STORED (rounded) activations
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = to_f<TA>(from_f<TA>(v[k]));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
    }
    V8<TA>::st(pooled.p + gc.pos(n, i, j) * pooled.cs + c0, m);
}

// ------------------------------------------------------------------------------------ final conv (+ reverse update)
// final = Conv((1,1), 64 => 1) (train_brain.jl:142): eps_hat[n,p] = sum_c a[p][c]*wf[c] + bf.
// 8 lanes per pixel, 8 channels each, 3-step shuffle reduction.
// mode 0: write eps_hat.  mode 1: fused reverse-diffusion update (generate_images.jl:196-208)
//   x <- sqrt(a_prev)*clamp((x - sigma_t*eps_hat)/sqrt(a_t), -1, 1) + sqrt(post_var)*z   (in place)
// with z read from zbuf (host-supplied noise) or drawn from Philox(seed, image, step).
template <typename TA>
__global__ void __launch_bounds__(256)
final_conv_kernel(View<const TA> a, Geo g, const float* __restrict__ wf, const float* __restrict__ bf,
                  float* __restrict__ eps_hat, int mode, float* __restrict__ x, const float* __restrict__ zbuf,
                  float4 scal /*sigma_t, sqrt_at, sqrt_aprev, sqrt_pv*/,
                  const unsigned long long* __restrict__ rng_dev /*[seed, first_index]*/, uint32_t step, int final_clamp) {
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    long long pix = idx >> 3;
    int c0 = (int)(idx & 7) * 8;
    const int HW = g.H * g.W;
    bool live = pix < (long long)g.N * HW;
    float part = 0.f;
    int n = 0, rem = 0;
    if (live) {
        n = (int)(pix / HW);
        rem = (int)(pix - (long long)n * HW);
        long long p = g.pos(n, rem / g.W, rem % g.W);
        float v[8];
        V8<TA>::ld(a.p + p * a.cs + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) part = fmaf(v[j], wf[c0 + j], part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (!live || c0 != 0) return;
    float e = part + bf[0];
    if (mode == 0) {
        eps_hat[pix] = e;
        return;
    }
    float z;
    if (zbuf) {
        z = zbuf[pix];
    } else {
        // seed and first image index live in device memory so one captured graph serves every chunk
        float4 zz = Philox::normal4(rng_dev[0], rng_dev[1] + (unsigned long long)n, step, (uint32_t)(rem >> 2));
        int k = rem & 3;
        z = k == 0 ? zz.x : (k == 1 ? zz.y : (k == 2 ? zz.z : zz.w));
    }
    float xv = x[pix];
    float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(scal.x, e)), scal.y);
    x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float xn = __fadd_rn(__fmul_rn(scal.z, x0), __fmul_rn(scal.w, z));
    if (final_clamp) xn = fminf(fmaxf(xn, -1.f), 1.f);
    x[pix] = xn;
}

// ------------------------------------------------------------------------------------ MSE
// Flux.Losses.mse = mean(abs2.(pred .- target)) (train_brain.jl:240).  Warp-shuffle + block
// reduction, one Float64 atomic per block; also emits d(loss)/d(pred) = 2(pred-target)*inv_count.
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n4, float dscale /*inv_count * loss scale*/,
           double* __restrict__ loss_sum, float* __restrict__ dpred) {
    __shared__ float wsum[8];
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    float s = 0.f;
    if (i < n4) {
        float4 p = *reinterpret_cast<const float4*>(pred + 4 * i);
        float4 t = *reinterpret_cast<const float4*>(target + 4 * i);
        float4 d = make_float4(p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w);
        s = d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (dpred) {
            float k = 2.f * dscale;
            *reinterpret_cast<float4*>(dpred + 4 * i) = make_float4(k * d.x, k * d.y, k * d.z, k * d.w);
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? wsum[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_sum, (double)v);
    }
}

constexpr int FINAL_BWD_PIX_PER_BLOCK = 512;

// ------------------------------------------------------------------------------------ backward: final conv
// da[p][c] = deps[p]*wf[c];  dwf[c] = sum_p deps[p]*a[p][c];  dbf = sum_p deps[p]   (sums -> Float64)
template <typename TA, typename TG>
__global__ void __launch_bounds__(256)
final_bwd_kernel(View<const TA> a, View<TG> da, Geo g, const float* __restrict__ wf, const float* __restrict__ deps,
                 double* __restrict__ sums /*[64 dwf | 1 dbf]*/) {
    __shared__ float red[32 * 65];                    // [pixel lane][64 dwf | dbf], reduced once per block
    const int t = threadIdx.x;
    const int c0 = (t & 7) * 8, pl = t >> 3;          // 8 lanes per pixel, 32 pixels per pass
    const int HW = g.H * g.W;
    const long long total = (long long)g.N * HW;
    float w8[8], acc[8], accb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { w8[j] = wf[c0 + j]; acc[j] = 0.f; }
    const long long nchunks = (total + FINAL_BWD_PIX_PER_BLOCK - 1) / FINAL_BWD_PIX_PER_BLOCK;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long pbeg = chunk * FINAL_BWD_PIX_PER_BLOCK;
        for (long long pix = pbeg + pl; pix < pbeg + FINAL_BWD_PIX_PER_BLOCK && pix < total; pix += 32) {
            const int n = (int)(pix / HW);
            const int rem = (int)(pix - (long long)n * HW);
            const long long p = g.pos(n, rem / g.W, rem % g.W);
            const float d = deps[pix];
            float v[8], o[8];
            V8<TA>::ld(a.p + p * a.cs + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j] = d * w8[j]; acc[j] = fmaf(d, v[j], acc[j]); }
            V8<TG>::st(da.p + p * da.cs + c0, o);
            if (c0 == 0) accb += d;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl * 65 + c0 + j] = acc[j];
    if (c0 == 0) red[pl * 65 + 64] = accb;
    __syncthreads();
    if (t < 65) {
        float s = 0.f;
        for (int l = 0; l < 32; ++l) s += red[l * 65 + t];
        atomicAdd(&sums[t], (double)s);
    }
}

// ------------------------------------------------------------------------------------ BatchNorm backward
// z = y*scale+shift; g = da * [z > 0]; xhat = (y-mean)*istd
// pass 1: sums[0:C] += sum g, sums[C:2C] += sum g*xhat
// pass 2: dy = scale*(g - mg - xhat*mgx), sums[2C:3C] += sum dy   (conv-bias gradient)
// Each thread owns 8 channels and walks PIX_PER_THREAD pixels, so the block-level reduction is
// one shared-memory atomic per channel per thread.
constexpr int BNB_PIX_PER_BLOCK = 512;
// grid of the chunk-striding backward kernels: at most `per_sm` resident blocks per SM
static inline int stride_blocks(long long items, int per_block, int num_sms, int per_sm) {
    const long long chunks = (items + per_block - 1) / per_block;
    const long long cap = (long long)per_sm * num_sms;
    return (int)(chunks < cap ? (chunks < 1 ? 1 : chunks) : cap);
}

template <typename TA, typename TG, int PASS>
__global__ void __launch_bounds__(256, PASS == 2 ? 4 : 2)
bn_bwd_kernel(View<const TA> y, View<const TG> da, View<TG> dy, Geo g, int C, const float* __restrict__ scale,
              const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ istd,
              const float* __restrict__ mg, const float* __restrict__ mgx, double* __restrict__ sums, int rev = 0) {
    // blocks stride over 512-pixel chunks (rev: from the last chunk down -- da was just written front to back); per-thread partial sums live in registers for the whole block and are reduced
    // ONCE through a [lanes][C] shared-memory table (per-chunk shared float atomics were 32-way contended)
    __shared__ float red[2][256 * 8];
    const int t = threadIdx.x;
    const int groups = C / 8;
    const int lanes = 256 / groups;
    const int c0 = (t % groups) * 8, pl = t / groups;
    const long long total = (long long)g.N * g.H * g.W;
    // PASS 2 keeps four blocks per SM resident (<= 64 registers): dy = scale*(g - mg - xhat*mgx) with xhat = (y - mean)*istd
    // is evaluated as  g*scale + (y*k1 + k0),  k1 = -scale*istd*mgx,  k0 = scale*(mean*istd*mgx - mg)
    float sc[8], sh[8], mu[8], is[8], r0[8], r1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j];
        r0[j] = 0.f; r1[j] = 0.f;
        if (PASS == 2) {
            const float m_ = mean[c0 + j], i_ = istd[c0 + j], a0 = mg[c0 + j], a1 = mgx[c0 + j];
            mu[j] = -sc[j] * i_ * a1;                        // k1
            is[j] = sc[j] * (m_ * i_ * a1 - a0);             // k0
        } else {
            mu[j] = mean[c0 + j]; is[j] = istd[c0 + j];
        }
    }
    const int HW = g.H * g.W;
    const long long nchunks = (total + BNB_PIX_PER_BLOCK - 1) / BNB_PIX_PER_BLOCK;
    for (long long ck = blockIdx.x; ck < nchunks; ck += gridDim.x) {
        const long long chunk = rev ? nchunks - 1 - ck : ck;
        const long long pbeg = chunk * BNB_PIX_PER_BLOCK;
        long long pend = pbeg + BNB_PIX_PER_BLOCK;
        if (pend > total) pend = total;
        for (long long pix = pbeg + pl; pix < pend; pix += lanes) {
            int n = (int)(pix / HW);
            int rem = (int)(pix - (long long)n * HW);
            long long p = g.pos(n, rem / g.W, rem % g.W);
            float yv[8], gv[8];
            V8<TA>::ld(y.p + p * y.cs + c0, yv);
            V8<TG>::ld(da.p + p * da.cs + c0, gv);
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float z = fmaf(yv[j], sc[j], sh[j]);
                float gg = z > 0.f ? gv[j] : 0.f;
                if (PASS == 1) {
                    float xh = (yv[j] - mu[j]) * is[j];
                    r0[j] += gg; r1[j] += gg * xh;
                } else {
                    float d = fmaf(gg, sc[j], fmaf(yv[j], mu[j], is[j]));
                    o[j] = d; r0[j] += d;
                }
            }
            if (PASS == 2) V8<TG>::st(dy.p + p * dy.cs + c0, o);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[0][pl * C + c0 + j] = r0[j];
        if (PASS == 1) red[1][pl * C + c0 + j] = r1[j];
    }
    __syncthreads();
    const int nq = (PASS == 1) ? 2 : 1;
    if (t < nq * C) {
        const int q = t / C, c = t - q * C;
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += red[q][l * C + c];
        atomicAdd(&sums[(PASS == 1 ? q : 2) * C + c], (double)acc);
    }
}

// Reduction-only variants that stream over ALL padded positions (halo rows hold zeros in y and in every
// gradient tensor, so they add nothing to any of the sums): no per-pixel index arithmetic, rows are
// contiguous, four independent 16-byte loads per operand in flight per thread.
//   MODE 0: sums[0:C] += sum y, sums[C:2C] += sum y^2                       (forward statistics)
//   MODE 1: sums[0:C] += sum g, sums[C:2C] += sum g*xhat,  g = da*[y*scale+shift > 0]   (backward pass 1)
constexpr int LIN_POS_PER_BLOCK = 1024;     // positions per block-chunk; blocks stride over chunks (persistent-style grid)

// launch geometry of bn_reduce_linear_kernel: at most two resident blocks per SM, each striding over 1024-position chunks
static inline int lin_reduce_blocks(long long npos, int num_sms) {
    const long long chunks = (npos + LIN_POS_PER_BLOCK - 1) / LIN_POS_PER_BLOCK;
    const long long cap = 2LL * num_sms;
    return (int)(chunks < cap ? chunks : cap);
}

template <typename TA, typename TG, int MODE>
__global__ void __launch_bounds__(256, 2)
bn_reduce_linear_kernel(View<const TA> y, View<const TG> da, long long npos, int C, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ istd,
                        double* __restrict__ sums) {
    // per-thread partials over all chunks of this block, ONE block-level reduction at the end: a [lanes][C] table in
    // shared memory summed column-wise (the earlier per-chunk shared-memory float atomics were 32-way contended)
    __shared__ float red[2][256 * 8];           // [2][lanes * C]: lanes * C == 256 threads * 8 channels == 2048
    const int t = threadIdx.x;
    const int groups = C / 8, lanes = 256 / groups;
    const int c0 = (t % groups) * 8, pl = t / groups;
    float sc[8], sh[8], mu[8], is[8], r0[8], r1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        r0[j] = 0.f; r1[j] = 0.f;
        if (MODE == 1) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; mu[j] = mean[c0 + j]; is[j] = istd[c0 + j]; }
    }
    constexpr int U = 4;
    const long long nchunks = (npos + LIN_POS_PER_BLOCK - 1) / LIN_POS_PER_BLOCK;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long pbeg = chunk * LIN_POS_PER_BLOCK;
        long long pend = pbeg + LIN_POS_PER_BLOCK;
        if (pend > npos) pend = npos;
        for (long long p0 = pbeg + pl; p0 < pend; p0 += (long long)U * lanes) {
            float yv[U][8], gv[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                long long pp = p0 + (long long)u * lanes;
                if (pp >= pend) pp = pend - 1;       // clamp (in-bounds re-read, masked below)
                V8<TA>::ld(y.p + pp * y.cs + c0, yv[u]);
                if (MODE == 1) V8<TG>::ld(da.p + pp * da.cs + c0, gv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float m = (p0 + (long long)u * lanes < pend) ? 1.f : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (MODE == 0) {
                        const float v = yv[u][j] * m;
                        r0[j] += v; r1[j] = fmaf(v, v, r1[j]);
                    } else {
                        const float z = fmaf(yv[u][j], sc[j], sh[j]);
                        const float gg = z > 0.f ? gv[u][j] * m : 0.f;
                        r0[j] += gg; r1[j] = fmaf(gg, (yv[u][j] - mu[j]) * is[j], r1[j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[0][pl * C + c0 + j] = r0[j]; red[1][pl * C + c0 + j] = r1[j]; }
    __syncthreads();
    if (t < 2 * C) {
        const int q = t / C, c = t - q * C;
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += red[q][l * C + c];
        atomicAdd(&sums[q * C + c], (double)acc);
    }
}

// after pass 1: local sums -> gradient arena (d beta, d gamma); global sums -> means for pass 2.
// xr_slot >= 0: the global sums are produced here by the peer-memory all-reduce of the local ones (see above).
__global__ void __launch_bounds__(256)
bn_bwd_means_kernel(const double* __restrict__ local_sums, const double* __restrict__ global_sums,
                    double count, int C, float* __restrict__ mg, float* __restrict__ mgx,
                    float* __restrict__ dbeta, float* __restrict__ dgamma, float alpha, XReduce xr, int xr_slot,
                    double* __restrict__ gsums) {
    if (xr_slot >= 0) {
        xr_allreduce_block(xr, xr_slot, local_sums, 2 * C, gsums);
        global_sums = gsums;
    }
    int c = threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)((double)alpha * local_sums[c]);        // alpha = 1/loss-scale
    dgamma[c] = (float)((double)alpha * local_sums[C + c]);
    mg[c] = (float)(global_sums[c] / count);
    mgx[c] = (float)(global_sums[C + c] / count);
}

// generic: out[i] = (float)(alpha * in[i])
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n, double alpha) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(alpha * in[i]);
}

// ------------------------------------------------------------------------------------ MaxPool backward + skip merge
// dh1[p] = dskip[p] + (p is the first arg-max of its 2x2 window ? dpool[window] : 0)
// (NNlib maxpool backward routes the gradient to the first maximal element)
template <typename TA, typename TG>
__global__ void __launch_bounds__(256)
pool_bwd_merge_kernel(View<const TA> a, View<const TG> dskip, View<const TG> dpool, View<TG> out, Geo gf, Geo gc, int C) {
    const int groups = C / 8;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    long long pix = idx / groups;
    int c0 = (int)(idx - pix * groups) * 8;
    if (pix >= (long long)gc.N * gc.H * gc.W) return;
    int n = (int)(pix / (gc.H * gc.W));
    int rem = (int)(pix - (long long)n * gc.H * gc.W);
    int i = rem / gc.W, j = rem % gc.W;
    float av[4][8], dp[8];
    long long p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        p[q] = gf.pos(n, 2 * i + (q >> 1), 2 * j + (q & 1));
        V8<TA>::ld(a.p + p[q] * a.cs + c0, av[q]);
    }
    V8<TG>::ld(dpool.p + gc.pos(n, i, j) * dpool.cs + c0, dp);
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int best = 0;
        float bv = av[0][k];
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (av[q][k] > bv) { bv = av[q][k]; best = q; }
        arg[k] = best;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float s[8];
        V8<TG>::ld(dskip.p + p[q] * dskip.cs + c0, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += (arg[k] == q) ? dp[k] : 0.f;
        V8<TG>::st(out.p + p[q] * out.cs + c0, s);
    }
}

// ------------------------------------------------------------------------------------ backward of the first conv
// One block per image.  For the image channel:  Tw[n][tap][co] = sum_p dy[p][co]*x[p+shift(tap)]
// For the folded embedding channels: class sums Ccls[n][cls][co] = sum_{p in border class} dy[p][co]
template <typename T> __device__ __forceinline__ void ld4(const T* p, float o[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float o[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<__half>(const __half* p, float o[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float o[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// 16 lanes x 4 channels per pixel, 16 pixels per pass (72 accumulators per thread instead of 144)
template <typename TG>
__global__ void __launch_bounds__(256)
l1_bwd_kernel(View<const TG> dy, Geo g, const float* __restrict__ x, float* __restrict__ Tw, float* __restrict__ Ccls) {
    // block reduction: the two pixel lanes of a warp are combined with one shuffle, the eight warps through a
    // [2][8][576] shared-memory table summed by all threads (72 contended shared-memory atomics per thread before)
    __shared__ float red[2][8][9 * 64];
    const int t = threadIdx.x, n = blockIdx.x;
    const int H = g.H, W = g.W, TW = W + 2;
    const int c0 = (t & 15) * 4, pl = t >> 4;
    const float* xi = x + (long long)n * H * W;
    // zero-padded image window in shared memory: taps become unconditional broadcast loads
    __shared__ float tile[34 * 34];
    for (int i = t; i < (H + 2) * TW; i += 256) {
        const int rr = i / TW, cc = i - rr * TW;
        const int hh = rr - 1, ww = cc - 1;
        tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? xi[hh * W + ww] : 0.f;
    }
    __syncthreads();
    float tw[9][4], cl[9][4];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tw[a][j] = 0.f; cl[a][j] = 0.f; }
    for (int pix = pl; pix < H * W; pix += 16) {
        int h = pix / W, w = pix - h * W;
        float d[4];
        ld4<TG>(dy.p + g.pos(n, h, w) * dy.cs + c0, d);
        int cls = (h == 0 ? 0 : (h == H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const float m = (a == cls) ? 1.f : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) cl[a][j] = fmaf(m, d[j], cl[a][j]);
        }
        const float* tp = tile + h * TW + w;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float xv = tp[(tap / 3) * TW + (tap % 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) tw[tap][j] = fmaf(d[j], xv, tw[tap][j]);
        }
    }
    const int warp = t >> 5;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = tw[tap][j] + __shfl_xor_sync(0xffffffffu, tw[tap][j], 16);
            const float b = cl[tap][j] + __shfl_xor_sync(0xffffffffu, cl[tap][j], 16);
            if ((t & 16) == 0) {
                red[0][warp][tap * 64 + c0 + j] = a;
                red[1][warp][tap * 64 + c0 + j] = b;
            }
        }
    __syncthreads();
    for (int i = t; i < 2 * 576; i += 256) {
        const int q = i / 576, k = i - q * 576;
        float acc = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) acc += red[q][w8][k];
        (q == 0 ? Tw : Ccls)[(long long)n * 576 + k] = acc;
    }
}

// Second formulation of the first conv's backward (the default): persistent blocks, one image per iteration.
//   thread t: channels 8*(t&7).., pixel column w = t>>3 (fixed), rows h = 0..31 in turn  => every load instruction of
//   a warp covers 4 whole 128-byte pixel rows; the border class of a thread's column never changes and the row class
//   only at h = 0 / 31, so the nine class sums need no masks: top = row 0, bottom = row 31, middle = rows 1..30.
//   The image-channel weight gradient is a sum over ALL images, so its 72 accumulators per thread live in registers
//   across the block's images and are reduced ONCE (shuffles, shared memory, one Float64 atomic per value and block).
// Outputs: Ccls[n][cls][co] per image (feeds the embedding-weight GEMM), wimg_acc[tap*64+co] += sum_n,p dy*x (Float64).
__global__ void __launch_bounds__(256)
l1_bwd_fused_kernel(View<const __half> dyh, View<const __nv_bfloat16> dyb, View<const float> dyf, int dtype /*0 f32, 1 f16, 2 bf16*/,
                    Geo g, const float* __restrict__ x, float* __restrict__ Ccls, double* __restrict__ wimg_acc) {
    constexpr int H = 32, W = 32, TW = W + 2;
    __shared__ float tile[(H + 2) * TW];
    __shared__ float red[32 * 3 * 64];                // class partials [w][row class][co]; reused as [8 warps][576] at the end
    const int t = threadIdx.x;
    const int c0 = (t & 7) * 8, w = t >> 3;
    float tw[9][8];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) tw[a][j] = 0.f;
    for (int n = blockIdx.x; n < g.N; n += gridDim.x) {
        const float* xi = x + (long long)n * H * W;
        for (int i = t; i < (H + 2) * TW; i += 256) {
            const int rr = i / TW, cc = i - rr * TW;
            const int hh = rr - 1, ww = cc - 1;
            tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
        }
        __syncthreads();
        float top[8], mid[8], bot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { top[j] = 0.f; mid[j] = 0.f; bot[j] = 0.f; }
#pragma unroll 2
        for (int h = 0; h < H; ++h) {
            float d[8];
            const long long pos = g.pos(n, h, w);
            if (dtype == 1) V8<__half>::ld(dyh.p + pos * dyh.cs + c0, d);
            else if (dtype == 2) V8<__nv_bfloat16>::ld(dyb.p + pos * dyb.cs + c0, d);
            else V8<float>::ld(dyf.p + pos * dyf.cs + c0, d);
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
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[(w * 3 + 0) * 64 + c0 + j] = top[j];
            red[(w * 3 + 1) * 64 + c0 + j] = mid[j];
            red[(w * 3 + 2) * 64 + c0 + j] = bot[j];
        }
        __syncthreads();
        // class (rc, cc): cc = 0 -> column 0, cc = 2 -> column 31, cc = 1 -> columns 1..30
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
        __syncthreads();                               // tile / red are rewritten by the next image
    }
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

// dWimg: arena[idx(tap,co)] = alpha * wimg_acc[tap*64+co]   (Flux index of input channel 0)
__global__ void l1_wimg_finish_kernel(const double* __restrict__ acc, float alpha, int Cin_total, float* __restrict__ dW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 576) return;
    const int tap = i / 64, co = i % 64;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    dW[(1 - dx) + 3 * (1 - dy) + 9LL * Cin_total * co] = (float)((double)alpha * acc[i]);
}

// S[n][tap][co] = sum of the class sums of the border classes for which tap stays inside the image
__global__ void l1_tap_sums_kernel(const float* __restrict__ Ccls, float* __restrict__ S, long long B) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 576) return;
    int co = (int)(i % 64), tap = (int)((i / 64) % 9);
    long long n = i / 576;
    int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float s = 0.f;
    for (int rc = 0; rc < 3; ++rc) {
        if ((rc == 0 && dy < 0) || (rc == 2 && dy > 0)) continue;
        for (int cc = 0; cc < 3; ++cc) {
            if ((cc == 0 && dx < 0) || (cc == 2 && dx > 0)) continue;
            s += Ccls[n * 576 + (rc * 3 + cc) * 64 + co];
        }
    }
    S[i] = s;
}

// dWimg: arena[idx(tap,co)] = alpha * sum_n Tw[n][tap][co]   (Flux index of input channel 0)
__global__ void l1_wimg_grad_kernel(const float* __restrict__ Tw, long long B, float alpha, int Cin_total,
                                    float* __restrict__ dW) {
    int i = blockIdx.x;  // tap*64 + co
    int tap = i / 64, co = i % 64;
    double s = 0;
    for (long long n = threadIdx.x; n < B; n += blockDim.x) s += Tw[n * 576 + i];
    __shared__ double ws[8];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += ws[k];
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        dW[(1 - dx) + 3 * (1 - dy) + 9LL * Cin_total * co] = (float)(alpha * tot);
    }
}

// per-channel sum over valid pixels of a tensor -> Float64 (ConvTranspose bias gradient)
template <typename TG>
__global__ void __launch_bounds__(256)
channel_sum_kernel(View<const TG> d, Geo g, int C, double* __restrict__ sums) {
    __shared__ float red[256 * 8];                    // [lanes][C], reduced once per block (blocks stride over chunks)
    const int t = threadIdx.x;
    const int groups = C / 8, lanes = 256 / groups;
    const int c0 = (t % groups) * 8, pl = t / groups;
    const long long total = (long long)g.N * g.H * g.W;
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = 0.f;
    const int HW = g.H * g.W;
    const long long nchunks = (total + BNB_PIX_PER_BLOCK - 1) / BNB_PIX_PER_BLOCK;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long pbeg = chunk * BNB_PIX_PER_BLOCK;
        long long pend = pbeg + BNB_PIX_PER_BLOCK;
        if (pend > total) pend = total;
        for (long long pix = pbeg + pl; pix < pend; pix += lanes) {
            int n = (int)(pix / HW);
            int rem = (int)(pix - (long long)n * HW);
            float v[8];
            V8<TG>::ld(d.p + g.pos(n, rem / g.W, rem % g.W) * d.cs + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl * C + c0 + j] = r[j];
    __syncthreads();
    if (t < C) {
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += red[l * C + t];
        atomicAdd(&sums[t], (double)acc);
    }
}

// per-channel sum and sum of squares over valid pixels (train-mode BatchNorm statistics when the
// producing convolution is the tensor-core kernel, whose epilogue does not reduce)
template <typename TA>
__global__ void __launch_bounds__(256)
bn_stats_kernel(View<const TA> y, Geo g, int C, double* __restrict__ sums) {
    __shared__ float red[2][128];
    const int t = threadIdx.x;
    if (t < 128) { red[0][t] = 0.f; red[1][t] = 0.f; }
    __syncthreads();
    const int groups = C / 8, lanes = 256 / groups;
    const int c0 = (t % groups) * 8, pl = t / groups;
    const long long total = (long long)g.N * g.H * g.W;
    const long long pbeg = (long long)blockIdx.x * BNB_PIX_PER_BLOCK;
    long long pend = pbeg + BNB_PIX_PER_BLOCK;
    if (pend > total) pend = total;
    float r[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { r[j] = 0.f; q[j] = 0.f; }
    const int HW = g.H * g.W;
    for (long long pix = pbeg + pl; pix < pend; pix += lanes) {
        int n = (int)(pix / HW);
        int rem = (int)(pix - (long long)n * HW);
        float v[8];
        V8<TA>::ld(y.p + g.pos(n, rem / g.W, rem % g.W) * y.cs + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&red[0][c0 + j], r[j]); atomicAdd(&red[1][c0 + j], q[j]); }
    __syncthreads();
    if (t < C) { atomicAdd(&sums[t], (double)red[0][t]); atomicAdd(&sums[C + t], (double)red[1][t]); }
}

// ConvTranspose((2,2), stride 2) backward operand: du4[coarse pos][q*C + c] = du[fine pos(2i+py, 2j+px)][c],
// q = py*2+px.  With it both the data gradient (K = 4C GEMM) and the weight gradient are plain GEMMs
// over coarse positions that the tensor-core kernels can TMA-load.
template <typename TG>
__global__ void __launch_bounds__(256)
unshuffle2_kernel(View<const TG> du, View<TG> du4, Geo gf, Geo gc, int C) {
    const int groups = C / 8;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    long long pix = idx / (4 * groups);
    int rem4 = (int)(idx - pix * 4 * groups);
    int q = rem4 / groups, c0 = (rem4 - q * groups) * 8;
    if (pix >= (long long)gc.N * gc.H * gc.W) return;
    int n = (int)(pix / (gc.H * gc.W));
    int rem = (int)(pix - (long long)n * gc.H * gc.W);
    int i = rem / gc.W, j = rem % gc.W;
    float v[8];
    V8<TG>::ld(du.p + gf.pos(n, 2 * i + (q >> 1), 2 * j + (q & 1)) * du.cs + c0, v);
    V8<TG>::st(du4.p + gc.pos(n, i, j) * du4.cs + q * C + c0, v);
}

// ------------------------------------------------------------------------------------ Adam (Optimisers.jl 0.4.6)
// Per-step optimiser state kept on the device so that a training iteration needs no host-side scalar:
//   bt1, bt2  = beta1^t, beta2^t of the NEXT update (Optimisers.jl keeps them in the rule state)
//   nonfinite = 1 if the gradient of the current step holds an inf/NaN (set by grad_check_kernel)
//   skipped / applied = number of updates skipped by the overflow guard / applied so far
struct TrainState {
    float bt1, bt2;
    int nonfinite;
    int skipped;
    long long applied;
};

// flags any non-finite element of the flat gradient arena (the 16-bit gradient tensors use a static loss scale;
// an overflow anywhere upstream reaches the FP32 weight gradients as inf or NaN)
__global__ void __launch_bounds__(256)
grad_check_kernel(const float* __restrict__ g, long long n, TrainState* __restrict__ st) {
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    bool bad = false;
    if (i0 + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(g + i0);
        bad = !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    } else {
        for (long long i = i0; i < n; ++i) bad |= !isfinite(g[i]);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(&st->nonfinite, 1);
}

// m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p -= m/(1-bt1) / (sqrt(v/(1-bt2)) + eps) * eta
// over the whole flat parameter arena (running statistics have g = m = v = 0 => unchanged).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float eta, float b1, float b2, float eps, const TrainState* __restrict__ st) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (st->nonfinite) return;            // overflow guard: skip the whole update
    const float bt1 = st->bt1, bt2 = st->bt2;
    float gi = g[i];
    float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(1.f - b1, gi));
    float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(1.f - b2, __fmul_rn(gi, gi)));
    m[i] = mi; v[i] = vi;
    float num = __fdiv_rn(mi, 1.f - bt1);
    float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, 1.f - bt2)), eps);
    p[i] = __fsub_rn(p[i], __fmul_rn(__fdiv_rn(num, den), eta));
}

// after adam_kernel: beta^t <- beta^t * beta (Optimisers.jl), counters, and re-arm the overflow flag
__global__ void adam_advance_kernel(TrainState* __restrict__ st, float b1, float b2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st->nonfinite) {
        st->skipped += 1;
        st->nonfinite = 0;
    } else {
        st->bt1 = __fmul_rn(st->bt1, b1);
        st->bt2 = __fmul_rn(st->bt2, b2);
        st->applied += 1;
    }
}

// (x+1)/2 clamped to [0,1] and quantised to 8 bits, round-half-even in Float64 like the host PNG writer
// (`(img .+ 1) ./ 2` then Gray -> N0f8, /root/reference/src/generate_images.jl:256-265)
__global__ void __launch_bounds__(256)
quantize_u8_kernel(const float* __restrict__ x, long long n4, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 v = *reinterpret_cast<const float4*>(x + 4 * i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float a = __fmul_rn(__fadd_rn(f[k], 1.f), 0.5f);
        a = fminf(fmaxf(a, 0.f), 1.f);
        const uint32_t q = (uint32_t)__double2int_rn((double)a * 255.0);
        w |= q << (8 * k);
    }
    out[i] = w;
}

// ------------------------------------------------------------------------------------ weight packing
// Flux Conv weight w[a,b,ci,co] (column-major; true convolution) -> cross-correlation, K-major:
//   fwd  : Wf[co][tap][ci]  = w[1-dx, 1-dy, ci+ci_off, co]              tap = (dy+1)*3 + (dx+1)
//   dgrad: Wd[ci][tap][co]  = w[1+dx, 1+dy, ci, co]                     (transposed, un-flipped)
// row_scale (forward packing only): per-output-channel factor folded into the weights in FP32 before
// rounding (inference BatchNorm fold: W' = W * gamma/sqrt(var_run+eps)).
template <typename TW>
__global__ void pack_conv3_kernel(const float* __restrict__ w, int Cin_total, int ci_off, int Cin, int Cout,
                                  int dgrad, const float* __restrict__ row_scale, TW* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 9LL * Cin * Cout;
    if (i >= total) return;
    if (!dgrad) {
        int ci = (int)(i % Cin), tap = (int)((i / Cin) % 9), co = (int)(i / (9LL * Cin));
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        float v = w[(1 - dx) + 3 * (1 - dy) + 9LL * (ci + ci_off) + 9LL * Cin_total * co];
        out[i] = from_f<TW>(row_scale ? v * row_scale[co] : v);
    } else {
        int co = (int)(i % Cout), tap = (int)((i / Cout) % 9), ci = (int)(i / (9LL * Cout));
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        out[i] = from_f<TW>(w[(1 + dx) + 3 * (1 + dy) + 9LL * (ci + ci_off) + 9LL * Cin_total * co]);
    }
}
// All 3x3 convolutions of the network in ONE launch (the training step re-packs after every Adam update; nine to
// eighteen 3-us launches were a visible share of a small-batch step): blockIdx.y = layer, blockIdx.z = 0 forward
// layout (type TA, optional per-output-channel scale), 1 data-gradient layout (type TG).
struct PackJobs {
    const float* w[9];
    int cin[9], cout[9];
    void* out_f[9];
    void* out_d[9];
    const float* row_scale[9];
    int tf32_round;      // FP32 outputs rounded to the nearest TF32 value (DDPM_PREC_TF32: operands of kind::tf32 MMAs)
};
template <typename TA, typename TG>
__global__ void pack_conv3_batch_kernel(const PackJobs J) {
    const int l = blockIdx.y, dgrad = blockIdx.z;
    const int Cin = J.cin[l], Cout = J.cout[l];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9LL * Cin * Cout) return;
    const float* __restrict__ w = J.w[l];
    if (!dgrad) {
        const int ci = (int)(i % Cin), tap = (int)((i / Cin) % 9), co = (int)(i / (9LL * Cin));
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const float v = w[(1 - dx) + 3 * (1 - dy) + 9LL * ci + 9LL * Cin * co];
        float vs = J.row_scale[l] ? v * J.row_scale[l][co] : v;
        if (J.tf32_round) vs = tf32_rn(vs);
        reinterpret_cast<TA*>(J.out_f[l])[i] = from_f<TA>(vs);
    } else {
        const int co = (int)(i % Cout), tap = (int)((i / Cout) % 9), ci = (int)(i / (9LL * Cout));
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        float vd = w[(1 + dx) + 3 * (1 + dy) + 9LL * ci + 9LL * Cin * co];
        if (J.tf32_round) vd = tf32_rn(vd);
        reinterpret_cast<TG*>(J.out_d[l])[i] = from_f<TG>(vd);
    }
}
// ConvTranspose weight w[a,b,co,ci]:  Wt[q*Cout+co][ci] (fwd, K = ci)  /  Wtd[ci][q*Cout+co] (dgrad, K = q,co)
template <typename TW>
__global__ void pack_up2_kernel(const float* __restrict__ w, int Cin, int Cout, int dgrad, TW* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 4LL * Cin * Cout;
    if (i >= total) return;
    int q, co, ci;
    if (!dgrad) { ci = (int)(i % Cin); co = (int)((i / Cin) % Cout); q = (int)(i / ((long long)Cin * Cout)); }
    else { co = (int)(i % Cout); q = (int)((i / Cout) % 4); ci = (int)(i / (4LL * Cout)); }
    int py = q >> 1, px = q & 1;
    out[i] = from_f<TW>(w[(1 - px) + 2 * (1 - py) + 4LL * co + 4LL * Cout * ci]);
}
// down1.conv1: Wimg[tap][co] (input channel 0) and Wemb[tap*Cout+co][c] (input channels 1..128), FP32
__global__ void pack_l1_kernel(const float* __restrict__ w, int D, int Cout, float* __restrict__ Wimg,
                               float* __restrict__ Wemb) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 9LL * Cout * (D + 1);
    if (i >= total) return;
    int c = (int)(i % (D + 1));
    int co = (int)((i / (D + 1)) % Cout), tap = (int)(i / ((long long)(D + 1) * Cout));
    int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float v = w[(1 - dx) + 3 * (1 - dy) + 9LL * c + 9LL * (D + 1) * co];
    if (c == 0) Wimg[tap * Cout + co] = v;
    else Wemb[((long long)tap * Cout + co) * D + (c - 1)] = v;
}

}  // namespace ddpm
