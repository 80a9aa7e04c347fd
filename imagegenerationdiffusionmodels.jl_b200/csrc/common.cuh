// Shared definitions of libddpm: error handling, the padded activation geometry and small
// device helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <cstdio>

namespace ddpm {

// ------------------------------------------------------------------------------------ errors
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define DDPM_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _buf[512];                                                                      \
            snprintf(_buf, sizeof _buf, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                     __FILE__, __LINE__);                                                        \
            throw ::ddpm::Error(_buf);                                                           \
        }                                                                                        \
    } while (0)

#define DDPM_CHECK(cond, msg)                                                                    \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            char _buf[512];                                                                      \
            snprintf(_buf, sizeof _buf, "%s (%s:%d)", msg, __FILE__, __LINE__);                  \
            throw ::ddpm::Error(_buf);                                                           \
        }                                                                                        \
    } while (0)

#define DDPM_LAUNCH_CHECK() DDPM_CUDA(cudaGetLastError())

// ------------------------------------------------------------------------------------ geometry
// Every multi-channel activation / gradient tensor lives in HBM as a 2-D matrix
//   [position][channel]   (channel fastest, i.e. NHWC)
// over a ZERO-PADDED, IMAGE-STACKED position space:
//   rows  r = 0 .. N*(H+1)      row n*(H+1) is an all-zero separator (bottom pad of image n-1 ==
//                               top pad of image n), rows n*(H+1)+1+h hold image row h
//   cols  c = 0 .. W            c = 0 is zero, c = 1+w holds pixel w; the right neighbour of pixel W-1 is position
//                               (r+1, 0): ONE zero column serves as the right pad of row r and the left pad of row r+1
//   pos = r*(W+1) + c
// so that a 3x3 tap (dy,dx) is the constant row shift dy*(W+1)+dx of that matrix and the zero halo
// implements pad=1 (tests/test_layout_host.py checks exactly this property on the host).  The implicit-GEMM kernels therefore never test bounds on the input side;
// producers only ever write valid positions, halos stay zero from allocation time.
// `guard` zero positions precede position 0 and follow the last one so shifted tile reads and the
// overhang of the last 128-row tile stay inside the allocation.
constexpr int WP_32 = 33, WP_16 = 17;      // Geo::Wp of the 32x32 and 16x16 levels (kernel template arguments)

struct Geo {
    int N, H, W;
    int Wp;        // W + 1  (row stride in positions: ONE zero column per row, see make())
    int Hs;        // H + 1  (image stride in rows)
    int L;         // N*(H+1) + 1 rows
    long long npos;  // L * Wp
    int guard;     // positions of zero guard on each side

    __host__ __device__ static Geo make(int N, int H, int W) {
        Geo g;
        g.N = N; g.H = H; g.W = W;
        // one zero column per row: position (r, 0) is the right pad of row r-1 and the left pad of row r, exactly like the
        // shared separator row between images (33x33 instead of 34x33 positions per 32x32 image: fewer MMA rows and bytes)
        g.Wp = W + 1; g.Hs = H + 1; g.L = N * (H + 1) + 1;
        g.npos = (long long)g.L * g.Wp;
        g.guard = 2 * g.Wp + 160;
        return g;
    }
    __host__ __device__ long long alloc_positions() const { return npos + 2LL * guard; }
    __host__ __device__ long long pos(int n, int h, int w) const {
        return (long long)(n * Hs + 1 + h) * Wp + (w + 1);
    }
    // decode a position; returns false for halo positions
    __host__ __device__ bool decode(long long p, int& n, int& h, int& w) const {
        if (p < 0 || p >= npos) return false;
        int r = (int)(p / Wp);
        int c = (int)(p - (long long)r * Wp);
        int rr = r % Hs;
        n = r / Hs; h = rr - 1; w = c - 1;
        return rr != 0 && c >= 1 && c <= W && n < N;
    }
    __host__ __device__ bool valid(long long p) const {
        int n, h, w;
        return decode(p, n, h, w);
    }
};

// A channel-slice view of such a tensor: p points at (position 0, first channel of the slice),
// cs = channel count of the underlying tensor (row stride in elements).
template <typename T>
struct View {
    T* p;
    int cs;
};

// ------------------------------------------------------------------------------------ numeric helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// FP32 -> nearest TF32 (10-bit mantissa), returned as an FP32 bit pattern.  tcgen05.mma.kind::tf32 ignores the low 13
// mantissa bits (truncation); rounding operands to nearest beforehand halves the operand error and removes its bias.
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ddpm
