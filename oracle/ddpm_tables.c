/* C restatement of the bit-exact pieces of the DDPM path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * (Only tests/ link this; the product never does.)  PARITY UNPINNED vs Julia, see ddpm_oracle.py.
 *
 *   oracle_schedule          src/train_brain.jl:20-24   (range -> Base twice-precision _linspace, Float32)
 *   oracle_embedding         src/train_brain.jl:54-62
 *   oracle_sampler_scalars   src/generate_images.jl:186-208
 *   oracle_q_sample          src/train_brain.jl:230-233
 *   oracle_apply_noise       src/ImageGenerationDiffusionModels.jl:60-73
 *   oracle_philox4x32_10     counter-based RNG of the device sampler (Salmon et al. 2011)
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, every op rounds once)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float truncbits(float x, int nb) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= (uint32_t)(0xFFFFFFFFu << nb);
    memcpy(&x, &u, 4);
    return x;
}
static void add12(float x, float y, float* hi, float* lo) {
    if (fabsf(y) > fabsf(x)) { float t = x; x = y; y = t; }
    volatile float h = x + y;
    volatile float d = x - h;
    *hi = h;
    *lo = d + y;
}

/* collect(range(Float32(bmin), Float32(bmax), length=T)); alpha = 1 .- beta; acum = accumulate(*, alpha) */
void oracle_schedule(int T, float bmin, float bmax, float* beta, float* alpha, float* acum) {
    volatile float start = bmin, stop = bmax;
    volatile float delta = stop - start;
    volatile float q = start / delta;
    volatile float tmin = -q;
    volatile float tt = tmin * (float)(T - 1);
    volatile float ti = tt + 1.0f;
    long imin = lroundf(ti);
    volatile float ref, step;
    if (imin > 1 && imin < T) {
        volatile float t = (float)(imin - 1) / (float)(T - 1);
        volatile float omt = 1.0f - t;
        volatile float p1 = omt * start, p2 = t * stop;
        ref = p1 + p2;
        if (imin - 1 < T - imin) { volatile float n = ref - start; step = n / (float)(imin - 1); }
        else { volatile float n = stop - ref; step = n / (float)(T - imin); }
    } else if (imin <= 1) {
        imin = 1; ref = start; step = delta / (float)(T - 1);
    } else {
        imin = T; ref = stop; step = delta / (float)(T - 1);
    }
    long big = (imin - 1 > T - imin) ? imin - 1 : T - imin;
    int nb = (int)ceil(log2((double)big));
    if (nb > 12) nb = 12;
    float step_hi = truncbits(step, nb);
    float x1_hi, x1_lo, x2_hi, x2_lo;
    volatile float m1 = (float)(1 - imin) * step_hi, m2 = (float)(T - imin) * step_hi;
    add12(m1, ref, &x1_hi, &x1_lo);
    add12(m2, ref, &x2_hi, &x2_lo);
    volatile float a0 = start - x1_hi, a = a0 - x1_lo;
    volatile float b0 = stop - x2_hi, b = b0 - x2_lo;
    volatile float ba = b - a;
    volatile float step_lo = ba / (float)(T - 1);
    volatile float rl0 = (float)(1 - imin) * step_lo;
    volatile float ref_lo = a - rl0;
    double ref64 = (double)ref + (double)ref_lo, step64 = (double)step_hi + (double)step_lo;
    float prod = 1.0f;
    for (int i = 1; i <= T; ++i) {
        beta[i - 1] = (float)(ref64 + (double)(i - imin) * step64);
        volatile float al = 1.0f - beta[i - 1];
        alpha[i - 1] = al;
        volatile float pr = (i == 1) ? al : prod * al;
        prod = pr;
        acum[i - 1] = prod;
    }
}

void oracle_embedding(int t, int D, float* pe) {
    const double neg_log = -(double)logf(1e4f);
    for (int i = 1; i <= D / 2; ++i) {
        double div = exp(neg_log * (2.0 * (double)(i - 1) / (double)(D - 1)));
        pe[2 * i - 2] = (float)sin((double)t * div);
        pe[2 * i - 1] = (float)cos((double)t * div);
    }
}

/* out = [sigma_t, sqrt(a_t), sqrt(a_prev), sqrt(post_var)], t is 1-based */
void oracle_sampler_scalars(const float* acum, int t, float* out) {
    volatile float a_t = acum[t - 1];
    volatile float a_prev = t > 1 ? acum[t - 2] : 1.0f;
    volatile float beta_t = 1.0f - a_t;
    volatile float beta_prev = 1.0f - a_prev;
    volatile float om = 1.0f - a_t;
    volatile float num = beta_prev * om;
    volatile float pv = num / om;
    out[0] = sqrtf(beta_t); out[1] = sqrtf(a_t); out[2] = sqrtf(a_prev); out[3] = sqrtf(pv);
}

void oracle_q_sample(const float* x0, const int* ts, const float* eps, const float* acum, int B, int hw, float* xt) {
    for (int n = 0; n < B; ++n) {
        volatile float ac = acum[ts[n] - 1];
        volatile float om = 1.0f - ac;
        float a = sqrtf(ac), b = sqrtf(om);
        for (int p = 0; p < hw; ++p) {
            volatile float u = a * x0[n * hw + p], v = b * eps[n * hw + p];
            xt[n * hw + p] = u + v;
        }
    }
}

void oracle_apply_noise(const double* img, const double* eps, long n, int steps, double bmin, double bmax, double* out) {
    double step = (bmax - bmin) / steps;
    for (long i = 0; i < n; ++i) out[i] = img[i];
    for (int k = 0; k <= steps; ++k) {
        double beta = bmin + k * step;
        double sa = sqrt(1 - beta), sb = sqrt(beta);
        for (long i = 0; i < n; ++i) {
            volatile double u = sa * out[i], v = sb * eps[i];
            out[i] = u + v;
        }
    }
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
