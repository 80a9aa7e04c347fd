// Host-only check of the padded position layout (csrc/common.cuh: Geo).  Compiled with nvcc, runs on the CPU.
// For every pixel and every 3x3 tap, position + dy*Wp + dx must be the position of the neighbouring pixel when that pixel
// exists and a halo position (never a pixel of this or any other image) otherwise; every tap and every 128-row tile
// overhang must stay inside [-guard, npos + guard); pos() and decode() must be inverse to each other.
#include <cstdio>
#include <vector>
#include "../../imagegenerationdiffusionmodels.jl_b200/csrc/common.cuh"

using ddpm::Geo;

static int check(int N, int H, int W) {
    const Geo g = Geo::make(N, H, W);
    int bad = 0;
    std::vector<int> owner((size_t)g.npos, -1);
    long long pixels = 0;
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const long long p = g.pos(n, h, w);
                if (p < 0 || p >= g.npos || owner[(size_t)p] != -1) { ++bad; continue; }
                owner[(size_t)p] = (n * H + h) * W + w;
                int n2, h2, w2;
                if (!g.decode(p, n2, h2, w2) || n2 != n || h2 != h || w2 != w) ++bad;
                ++pixels;
            }
    for (long long p = 0; p < g.npos; ++p)
        if ((owner[(size_t)p] >= 0) != g.valid(p)) ++bad;
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const long long q = g.pos(n, h, w) + (long long)dy * g.Wp + dx;
                        if (q < -(long long)g.guard || q >= g.npos + g.guard) { ++bad; continue; }
                        const bool inside = h + dy >= 0 && h + dy < H && w + dx >= 0 && w + dx < W;
                        const bool in_range = q >= 0 && q < g.npos;
                        if (inside) {
                            if (!in_range || q != g.pos(n, h + dy, w + dx)) ++bad;
                        } else if (in_range && owner[(size_t)q] >= 0) {
                            ++bad;                       // pad = 1 must read a zero, not somebody's pixel
                        }
                    }
    // the last 128-position tile and its halo'ed slab stay inside the guards
    const long long tiles = (g.npos + 127) / 128;
    if (tiles * 128 + (g.Wp + 1) > g.npos + g.guard || g.guard < g.Wp + 1) ++bad;
    std::printf("N=%d H=%d W=%d Wp=%d Hs=%d npos=%lld pixels=%lld fill=%.4f bad=%d\n", N, H, W, g.Wp, g.Hs, g.npos, pixels,
                (double)pixels / (double)g.npos, bad);
    return bad;
}

int main() {
    int bad = 0;
    bad += check(1, 32, 32);
    bad += check(3, 32, 32);
    bad += check(5, 16, 16);
    bad += check(2, 8, 8);
    if (Geo::make(1, 32, 32).Wp != ddpm::WP_32 || Geo::make(1, 16, 16).Wp != ddpm::WP_16) ++bad;
    return bad ? 1 : 0;
}
