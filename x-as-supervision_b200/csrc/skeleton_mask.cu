// K4/K5 — skeleton rasteriser, channel max and mask-reconstruction loss, forward and backward.
//
// Replaces modules/util.py:21-59 (draw_lines), modules/model.py:91-96 (max over the line
// heat-maps) and modules/base_losses/loss_func.py:4-16 (compute_mask_reconstruction_loss) of the
// reference.  The reference materialises ~10 temporaries of [B, L, S*S(,2)] fp32 (L = 25 lines,
// S = 256: 6.5 MB each per sample) plus the [B, L, S, S] heat-maps before the channel max.  Here:
//
//   skeleton_mask_fwd_kernel  one pass, writes only recon [B,1,S,S] + a byte per pixel (winning line);
//                             exp is monotone, so max_l exp(u_l) = exp(max_l u_l): one exp per pixel.
//                             Each warp owns a 16x8 pixel tile and first CULLS the lines that cannot
//                             win anywhere in the tile (distance is 1-Lipschitz: a line whose lower
//                             bound exceeds the smallest upper bound is out) - exact, typically 2-5 of
//                             25 lines survive.  Optionally fuses the loss sums against gt / weight.
//   skeleton_mask_bwd_kernel  one pass over recon / winner byte / upstream gradient: closed-form
//                             d recon / d joints of the winning segment (including the path through the
//                             projection parameter t), reduced per line without float atomics
//                             (warp shuffles -> per-warp shared rows -> per-CTA partials -> fixed-order
//                             scatter to joints), so gradients are bit-reproducible.
//   draw_lines_{fwd,bwd}      the un-maxed [B,L,S,S] heat-maps for API parity with util.draw_lines.
//   mask_loss_{fwd,bwd}       the loss on an arbitrary mask tensor (physique_recons_loss, model.py:176).
#include "xsup_internal.h"

namespace xsup {

constexpr int kSkelWarps = 8;
constexpr int kSkelThreads = kSkelWarps * 32;
constexpr int kTileW = 16, kTileH = 8;
constexpr int kTilesPerWarp = 8;
constexpr int kTilesPerCta = kSkelWarps * kTilesPerWarp;
#define kInf __int_as_float(0x7f800000)
constexpr float kZeroExp = 40.0f;

// segment `l` of sample `b`: A = (start.x, start.y, d.x, d.y), Bv = (1/(1e-8+|d|^2), end.x, end.y, c)
// start = child joint, end = parent joint (util.py:34-36); c = 2 for the arm lines when L >= 21 (util.py:50-53)
__device__ __forceinline__ void load_lines(const SkelParams& p, int b, float4* sA, float4* sB, float4* sC = nullptr) {
    const int l = threadIdx.x;
    if (l < p.L) {
        const float* s = p.kps + (size_t)b * p.kbs + (size_t)p.child[l] * p.kjs;
        const float* e = p.kps + (size_t)b * p.kbs + (size_t)p.parent[l] * p.kjs;
        const float sx = s[0], sy = s[1], ex = e[0], ey = e[1];
        const float dx = ex - sx, dy = ey - sy;
        const float den = 1e-8f + __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float c = (p.L >= 21 && (l == 11 || l == 12 || l == 14 || l == 15)) ? 2.0f : 1.0f;
        const float inv = 1.0f / den;
        sA[l] = make_float4(sx, sy, dx, dy);
        sB[l] = make_float4(inv, ex, ey, c);
        if (sC) sC[l] = make_float4(dx * inv, dy * inv, c, 0.0f);       // forward form: t = a . (d/den)
    }
}

// pixel-centre coordinate of make_coordinate_grid (util.py:8-12): 2*(i/(S-1)) - 1
__device__ __forceinline__ float grid_coord(int i, float fS1) { return fmaf(2.0f, __fdiv_rn((float)i, fS1), -1.0f); }

// the S coordinates once per CTA (true divisions), padded to a multiple of 4 so rows can be fetched as float4
__device__ __forceinline__ void load_coord_table(float* tab, int S) {
    const float fS1 = (float)(S - 1);
    for (int i = threadIdx.x; i < S; i += blockDim.x) tab[i] = grid_coord(i, fS1);
}

// q / bw with one Newton correction of the reciprocal product (within 1 ulp of the IEEE quotient, 3 instructions)
__device__ __forceinline__ float div_by(float q, float bw, float rbw) {
    const float y = q * rbw;
    return fmaf(fmaf(-y, bw, q), rbw, y);
}

// Squared distance from the pixel centre to the segment.  ax, ay = g - start; aydyi = ay * d.y/den; C = (d.x/den, d.y/den, c).
// t = a.d/den clamped to [0,1] (one FFMA.SAT) selects the three cases of util.py:44-46 at once: t <= 0 -> |g - start|^2,
// t >= 1 -> |a - d|^2 = |g - end|^2, else the foot of the perpendicular.  6 instructions per pixel and line.
__device__ __forceinline__ float seg_sqdist(float ax, float ay, float aydyi, const float4& A, const float4& C) {
    const float t = __saturatef(fmaf(ax, C.x, aydyi));
    const float rx = fmaf(-t, A.z, ax), ry = fmaf(-t, A.w, ay);
    return fmaf(rx, rx, ry * ry);
}

__device__ __forceinline__ float warp_min(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));   // CREDUX.MIN.F32 (sm_100a)
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// lines that can be the winner somewhere in the tile whose first pixel is (x0, y0); lane = line
// Returns 0 when every pixel of the tile is below exp(-kZeroExp) = 4.2e-18 (the reference's value there is a positive number below
// that; ours is exactly 0: invisible in every fp32 sum the loss and its gradient form, and 12 orders below the 1e-5 parity bound).
__device__ __forceinline__ unsigned cull_lines(int L, const float4* sA, const float4* sC, int x0, int y0, float two_over, float rho,
                                               float rbw, int lane) {
    float lo = kInf, up = kInf;
    if (lane < L) {
        const float cx = fmaf((float)x0 + 0.5f * (kTileW - 1), two_over, -1.0f);
        const float cy = fmaf((float)y0 + 0.5f * (kTileH - 1), two_over, -1.0f);
        const float4 A = sA[lane], C = sC[lane];
        const float ay = cy - A.y;
        const float dist = sqrt_approx(seg_sqdist(cx - A.x, ay, ay * C.y, A, C));
        const float wl = C.z == 2.0f ? 1.41421356f : 1.0f;      // sqrt(c): the quantity minimised is c*q = (sqrt(c)*dist)^2
        lo = wl * fmaxf(dist - rho, 0.0f);
        up = wl * (dist + rho);
    }
    const float U = warp_min(up);
    // slack: the fp32 operations above (incl. the approximate sqrt) err by ~1e-6 on values <= 4; 1e-4 only keeps a few more lines
    const unsigned keep = __ballot_sync(0xffffffffu, lo <= U + 1e-4f);
    // min over the lines of the lower bound of c*q/bw (beyond 105 every exp() in the tile would be exactly 0 in fp32, for the
    // reference as well; kZeroExp cuts earlier: with body_width 3e-3 that is 44 pixels from the nearest line instead of 72)
    const float lmin = warp_min(lo);
    return (lmin * lmin * rbw > kZeroExp) ? 0u : keep;
}

// ---------------------------------------------------------------------------------------------- fused forward
template <bool LOSS>
__global__ void __launch_bounds__(kSkelThreads) skeleton_mask_fwd_kernel(const SkelParams p, float* __restrict__ recon,
                                                                         uint8_t* __restrict__ line_idx,
                                                                         const float* __restrict__ gt,
                                                                         const float* __restrict__ weight, int use_clip,
                                                                         float* __restrict__ ws_loss) {
    __shared__ float4 sA[XSUP_MAX_LINES], sB[XSUP_MAX_LINES], sC[XSUP_MAX_LINES];
    __shared__ float red[kSkelWarps][3];
    extern __shared__ __align__(16) float tab[];                  // [S] pixel-centre coordinates
    const int b = blockIdx.y, chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_lines(p, b, sA, sB, sC);
    load_coord_table(tab, p.S);
    __syncthreads();
    const float two_over = 2.0f / (float)(p.S - 1);
    const float rho = 8.27648f * two_over * 1.0001f;             // half diagonal of the tile's pixel centres
    const float rbw = 1.0f / p.bw;
    const int S = p.S, L = p.L;
    const size_t sample = (size_t)b * S * S;
    recon += sample;
    line_idx += sample;
    if (LOSS) {
        gt += sample;
        if (weight) weight += sample;
    }
    float s_sq = 0.f, s_f = 0.f, s_w = 0.f;
    const int t_end = min(p.tiles, (chunk + 1) * kTilesPerCta);
    for (int tile = chunk * kTilesPerCta + warp; tile < t_end; tile += kSkelWarps) {
        int ty = (int)__umulhi((unsigned)tile, p.tiles_x_magic);   // tile / tiles_x without the integer-division sequence
        int tx = tile - ty * p.tiles_x;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        const int x0 = tx * kTileW, y0 = ty * kTileH;
        const int py = y0 + (lane >> 2), px = x0 + (lane & 3) * 4;
        const bool in = py < S && px < S;                         // ragged edge tiles: lanes outside the image idle
        const int o = py * S + px;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (LOSS && in) {                                         // issued before the line loop: their latency hides behind it
            g4 = *reinterpret_cast<const float4*>(gt + o);
            if (weight) w4 = *reinterpret_cast<const float4*>(weight + o);
        }
        const unsigned keep = cull_lines(L, sA, sC, x0, y0, two_over, rho, rbw, lane);
        if (in) {
            float h[4] = {0.f, 0.f, 0.f, 0.f};
            int bl[4] = {0, 0, 0, 0};
            if (keep) {
                const float gy = tab[py];
                const float4 gx4 = *reinterpret_cast<const float4*>(tab + px);
                const float gx[4] = {gx4.x, gx4.y, gx4.z, gx4.w};
                float best[4] = {kInf, kInf, kInf, kInf};
                for (unsigned m = keep; m; m &= m - 1) {
                    const int l = __ffs(m) - 1;
                    const float4 A = sA[l], C = sC[l];
                    const float ay = gy - A.y, aydyi = ay * C.y;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float cq = seg_sqdist(gx[i] - A.x, ay, aydyi, A, C) * C.z;
                        if (cq < best[i]) { best[i] = cq; bl[i] = l; }                     // ties: lowest line index
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // util.py:52-55: exp(-(q / bw) * c).  best = c*q with c in {1, 2}: the power-of-two factor commutes with every
                    // rounding of the division, so no look-up of c is needed; exp as one FMUL + MUFU.EX2 (absolute error < 3e-8)
                    const float u = div_by(best[i], p.bw, rbw);
                    h[i] = u > kZeroExp ? 0.0f : ex2(u * -kLog2e);                          // same cut per pixel as per tile
                }
            }
            *reinterpret_cast<float4*>(recon + o) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uchar4*>(line_idx + o) = make_uchar4((uint8_t)bl[0], (uint8_t)bl[1], (uint8_t)bl[2], (uint8_t)bl[3]);
            if (LOSS) {
                const float g[4] = {g4.x, g4.y, g4.z, g4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float d = h[i] - g[i], sq = d * d;
                    const float f = (!use_clip || h[i] > 0.1f) ? 1.0f : 0.0f;              // loss_func.py:9
                    s_sq += sq;
                    s_f += f;
                    s_w = fmaf(sq * f, w[i], s_w);
                }
            }
        }
    }
    if (LOSS) {
        s_sq = warp_sum(s_sq); s_f = warp_sum(s_f); s_w = warp_sum(s_w);
        if (lane == 0) { red[warp][0] = s_sq; red[warp][1] = s_f; red[warp][2] = s_w; }
        __syncthreads();
        if (threadIdx.x < 3) {
            float s = 0.f;
            for (int w = 0; w < kSkelWarps; ++w) s += red[w][threadIdx.x];
            ws_loss[((size_t)b * p.NC + chunk) * 4 + threadIdx.x] = s;
        }
    }
}

// sums[0..2] = fixed-order sums of the per-CTA partials (sum sq, sum filter, sum sq*filter*w), sums[3] = loss
__global__ void __launch_bounds__(256) mask_loss_finalize_kernel(const float* __restrict__ part, int n_part, double n, int mode,
                                                                 float* __restrict__ sums) {
    __shared__ double sh[256][3];
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n_part; i += 256)
        for (int j = 0; j < 3; ++j) a[j] += (double)part[(size_t)i * 4 + j];
    for (int j = 0; j < 3; ++j) sh[threadIdx.x][j] = a[j];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int j = 0; j < 3; ++j) sh[threadIdx.x][j] += sh[threadIdx.x + o][j];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double sq = sh[0][0], f = sh[0][1], w = sh[0][2];
        sums[0] = (float)sq; sums[1] = (float)f; sums[2] = (float)w;
        double loss = sq / n;                                        // XSUP_MASK_MSE: MSELoss(mean)
        if (mode == XSUP_MASK_CLIP_MEAN) loss = (sq / n) * (f / n);  // mean of (scalar mse * filter tensor)
        if (mode == XSUP_MASK_WEIGHTED) loss = w / n;                // mean of sq * filter * weight
        sums[3] = (float)loss;
    }
}

// d loss / d mask-pixel coefficient: returns s such that dL/dm = s * (m - gt) [* filter * weight in weighted mode]
__device__ __forceinline__ float loss_grad_scale(int mode, double n, const float* sums, const float* g_loss) {
    const double g = (double)*g_loss;
    if (mode == XSUP_MASK_CLIP_MEAN) return (float)(g * 2.0 * ((double)sums[1] / n) / n);
    return (float)(g * 2.0 / n);
}

// ---------------------------------------------------------------------------------------------- fused backward
// gradient of q (squared distance to the winning segment) w.r.t. its start and end joints
__device__ __forceinline__ void seg_sqdist_grad(float gx, float gy, const float4& A, const float4& Bv, float& dsx, float& dsy,
                                                float& dex, float& dey) {
    const float ax = gx - A.x, ay = gy - A.y;
    const float t = fmaf(ax, A.z, ay * A.w) * Bv.x;
    if (t <= 0.0f) {                       // q = |g - s|^2
        dsx = -2.0f * ax; dsy = -2.0f * ay; dex = 0.f; dey = 0.f;
    } else if (t >= 1.0f) {                // q = |g - e|^2
        dsx = 0.f; dsy = 0.f; dex = -2.0f * (gx - Bv.y); dey = -2.0f * (gy - Bv.z);
    } else {                               // q = |r|^2, r = a - t d, t = (a.d)/(1e-8+|d|^2) depends on both joints
        const float rx = fmaf(-t, A.z, ax), ry = fmaf(-t, A.w, ay);
        const float rho = fmaf(rx, A.z, ry * A.w) * Bv.x;          // (r.d)/den: ~1e-8*t/den analytically, kept for autograd parity
        const float mx = fmaf(-2.0f * t, A.z, ax), my = fmaf(-2.0f * t, A.w, ay);   // a - 2 t d
        dsx = 2.0f * (fmaf(-(1.0f - t), rx, rho * (A.z + mx)));
        dsy = 2.0f * (fmaf(-(1.0f - t), ry, rho * (A.w + my)));
        dex = 2.0f * (fmaf(-t, rx, -rho * mx));
        dey = 2.0f * (fmaf(-t, ry, -rho * my));
    }
}

template <bool LOSS>
__global__ void __launch_bounds__(kSkelThreads) skeleton_mask_bwd_kernel(const SkelParams p, const float* __restrict__ recon,
                                                                         const uint8_t* __restrict__ line_idx,
                                                                         const float* __restrict__ g_ext,
                                                                         const float* __restrict__ gt,
                                                                         const float* __restrict__ weight, int mode, int use_clip,
                                                                         double n, const float* __restrict__ sums,
                                                                         const float* __restrict__ g_loss,
                                                                         float* __restrict__ ws_grad) {
    __shared__ float4 sA[XSUP_MAX_LINES], sB[XSUP_MAX_LINES];
    __shared__ float4 acc[kSkelWarps][XSUP_MAX_LINES];
    extern __shared__ __align__(16) float tab[];                  // [S] pixel-centre coordinates
    const int b = blockIdx.y, chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_lines(p, b, sA, sB);
    load_coord_table(tab, p.S);
    acc[warp][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int S = p.S;
    const float lscale = LOSS ? loss_grad_scale(mode, n, sums, g_loss) : 0.f;
    const float nibw = -1.0f / p.bw;
    const size_t sample = (size_t)b * S * S;
    recon += sample;
    line_idx += sample;
    if (g_ext) g_ext += sample;
    if (LOSS) {
        gt += sample;
        if (weight) weight += sample;
    }
    struct TileIn {
        float4 m4, e4, g4, w4;
        uchar4 l4;
        int px, py;
        bool in;
    };
    auto fetch = [&](int tile, TileIn& t) {
        int ty = (int)__umulhi((unsigned)tile, p.tiles_x_magic);
        int tx = tile - ty * p.tiles_x;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        t.py = ty * kTileH + (lane >> 2);
        t.px = tx * kTileW + (lane & 3) * 4;
        t.in = t.py < S && t.px < S;
        t.e4 = make_float4(0.f, 0.f, 0.f, 0.f);
        t.w4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (t.in) {
            const int o = t.py * S + t.px;
            t.m4 = *reinterpret_cast<const float4*>(recon + o);
            t.l4 = *reinterpret_cast<const uchar4*>(line_idx + o);
            if (g_ext) t.e4 = *reinterpret_cast<const float4*>(g_ext + o);
            if (LOSS) {
                t.g4 = *reinterpret_cast<const float4*>(gt + o);
                if (weight) t.w4 = *reinterpret_cast<const float4*>(weight + o);
            }
        }
    };
    // (prefetching the next tile's inputs was measured: 80 registers, 3 CTAs/SM, 64 -> 76 us; not kept)
    const int t_end = min(p.tiles, (chunk + 1) * kTilesPerCta);
    for (int tile = chunk * kTilesPerCta + warp; tile < t_end; tile += kSkelWarps) {
        TileIn cur;
        fetch(tile, cur);
        float4 c[4];
        int li[4] = {-1, -1, -1, -1};
        unsigned bits = 0;
        if (cur.in) {
            float G[4] = {cur.e4.x, cur.e4.y, cur.e4.z, cur.e4.w};
            const float m[4] = {cur.m4.x, cur.m4.y, cur.m4.z, cur.m4.w};
            const int ls[4] = {cur.l4.x & 31, cur.l4.y & 31, cur.l4.z & 31, cur.l4.w & 31};
            if (LOSS) {
                const float g[4] = {cur.g4.x, cur.g4.y, cur.g4.z, cur.g4.w}, w[4] = {cur.w4.x, cur.w4.y, cur.w4.z, cur.w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s = lscale * (m[i] - g[i]);
                    if (mode == XSUP_MASK_WEIGHTED) s *= ((!use_clip || m[i] > 0.1f) ? w[i] : 0.0f);
                    G[i] += s;
                }
            }
            const float gy = tab[cur.py];
            const float4 gx4 = *reinterpret_cast<const float4*>(tab + cur.px);
            const float gx[4] = {gx4.x, gx4.y, gx4.z, gx4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 Bv = sB[ls[i]];
                const float wq = G[i] * m[i] * (Bv.w * nibw);            // dL/dq = G * heat * (-c / body_width)
                if (wq != 0.0f) {
                    float dsx, dsy, dex, dey;
                    seg_sqdist_grad(gx[i], gy, sA[ls[i]], Bv, dsx, dsy, dex, dey);
                    c[i] = make_float4(wq * dsx, wq * dsy, wq * dex, wq * dey);
                    li[i] = ls[i];
                    bits |= 1u << ls[i];
                }
            }
        }
        // per-line warp reduction of the contributions present in this tile (usually 1-3 lines)
        for (unsigned present = __reduce_or_sync(0xffffffffu, bits); present; present &= present - 1) {
            const int l = __ffs(present) - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (li[i] == l) { v.x += c[i].x; v.y += c[i].y; v.z += c[i].z; v.w += c[i].w; }
            // the four components in 6 shuffles instead of 4 x 5: the first exchange leaves (x, y) on the lower and (z, w) on the upper
            // half-warp, the second one component per quarter, three more finish it; lanes 0 / 8 / 16 / 24 then hold x / y / z / w.
            // A fixed tree, so the sums stay bit-reproducible (65.1 -> 63.2 us at B = 256, 256 x 256)
            const bool up = lane & 16;
            float k0 = up ? v.z : v.x, k1 = up ? v.w : v.y;
            k0 += __shfl_xor_sync(0xffffffffu, up ? v.x : v.z, 16);
            k1 += __shfl_xor_sync(0xffffffffu, up ? v.y : v.w, 16);
            const bool odd = lane & 8;
            float k = odd ? k1 : k0;
            k += __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 8);
            k += __shfl_xor_sync(0xffffffffu, k, 4);
            k += __shfl_xor_sync(0xffffffffu, k, 2);
            k += __shfl_xor_sync(0xffffffffu, k, 1);
            if ((lane & 7) == 0) reinterpret_cast<float*>(&acc[warp][l])[lane >> 3] += k;
        }
    }
    __syncthreads();
    if (threadIdx.x < XSUP_MAX_LINES) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < kSkelWarps; ++w) {
            const float4 a = acc[w][threadIdx.x];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        reinterpret_cast<float4*>(ws_grad)[((size_t)b * p.NC + chunk) * XSUP_MAX_LINES + threadIdx.x] = s;
    }
}

// per-line partials [B][NC][32] float4 (d/d start.xy, d/d end.xy) -> g_kps [B,K,2], fixed summation order
__global__ void __launch_bounds__(128) skeleton_scatter_kernel(const SkelParams p, int NC, const float* __restrict__ ws_grad,
                                                               float* __restrict__ g_kps) {
    __shared__ float ls[XSUP_MAX_LINES][4];
    const int b = blockIdx.x, t = threadIdx.x;
    {
        const int l = t >> 2, j = t & 3;
        float s = 0.f;
        if (l < p.L)
            for (int c = 0; c < NC; ++c) s += ws_grad[(((size_t)b * NC + c) * XSUP_MAX_LINES + l) * 4 + j];
        ls[l][j] = s;
    }
    __syncthreads();
    for (int k = t; k < p.K; k += 128) {
        float gx = 0.f, gy = 0.f;
        for (int l = 0; l < p.L; ++l) {
            if (p.child[l] == k) { gx += ls[l][0]; gy += ls[l][1]; }
            if (p.parent[l] == k) { gx += ls[l][2]; gy += ls[l][3]; }
        }
        g_kps[((size_t)b * p.K + k) * 2 + 0] = gx;
        g_kps[((size_t)b * p.K + k) * 2 + 1] = gy;
    }
}

// ---------------------------------------------------------------------------------------------- un-maxed heat-maps (util.draw_lines)
__global__ void __launch_bounds__(256) draw_lines_fwd_kernel(const SkelParams p, float* __restrict__ heat) {
    __shared__ float4 sA[XSUP_MAX_LINES], sB[XSUP_MAX_LINES], sC[XSUP_MAX_LINES];
    const int b = blockIdx.z, l = blockIdx.y;
    load_lines(p, b, sA, sB, sC);
    __syncthreads();
    const int S = p.S, quad = blockIdx.x * 256 + threadIdx.x;          // 4 consecutive pixels of one row
    if (quad * 4 >= S * S) return;
    const int py = (quad * 4) / S, px = quad * 4 - py * S;
    const float fS1 = (float)(S - 1);
    const float4 A = sA[l], C = sC[l];
    const float gy = grid_coord(py, fS1), ay = gy - A.y, aydyi = ay * C.y;
    const float rbw = 1.0f / p.bw;
    float h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)          // same operation sequence as the fused kernel: max over l of these == its output, bit for bit
    {
        const float u = div_by(seg_sqdist(grid_coord(px + i, fS1) - A.x, ay, aydyi, A, C) * C.z, p.bw, rbw);
        h[i] = u > kZeroExp ? 0.0f : ex2(u * -kLog2e);
    }
    *reinterpret_cast<float4*>(heat + (((size_t)b * p.L + l) * S + py) * S + px) = make_float4(h[0], h[1], h[2], h[3]);
}

__global__ void __launch_bounds__(256) draw_lines_bwd_kernel(const SkelParams p, const float* __restrict__ heat,
                                                             const float* __restrict__ g_heat, float* __restrict__ ws_grad) {
    __shared__ float4 sA[XSUP_MAX_LINES], sB[XSUP_MAX_LINES];
    __shared__ float4 red[8];
    const int b = blockIdx.z, l = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_lines(p, b, sA, sB);
    __syncthreads();
    const int S = p.S, quad = blockIdx.x * 256 + threadIdx.x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (quad * 4 < S * S) {
        const int py = (quad * 4) / S, px = quad * 4 - py * S;
        const float fS1 = (float)(S - 1);
        const float4 A = sA[l], Bv = sB[l];
        const size_t o = (((size_t)b * p.L + l) * S + py) * S + px;
        const float4 h4 = *reinterpret_cast<const float4*>(heat + o), g4 = *reinterpret_cast<const float4*>(g_heat + o);
        const float h[4] = {h4.x, h4.y, h4.z, h4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        const float gy = grid_coord(py, fS1), k = -Bv.w / p.bw;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float wq = g[i] * h[i] * k;
            if (wq != 0.0f) {
                float dsx, dsy, dex, dey;
                seg_sqdist_grad(grid_coord(px + i, fS1), gy, A, Bv, dsx, dsy, dex, dey);
                v.x = fmaf(wq, dsx, v.x); v.y = fmaf(wq, dsy, v.y); v.z = fmaf(wq, dex, v.z); v.w = fmaf(wq, dey, v.w);
            }
        }
    }
    v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 s = red[0];
        for (int w = 1; w < 8; ++w) { s.x += red[w].x; s.y += red[w].y; s.z += red[w].z; s.w += red[w].w; }
        reinterpret_cast<float4*>(ws_grad)[((size_t)b * gridDim.x + blockIdx.x) * XSUP_MAX_LINES + l] = s;
    }
}

// ---------------------------------------------------------------------------------------------- loss on an arbitrary mask tensor
__global__ void __launch_bounds__(256) mask_loss_fwd_kernel(const float* __restrict__ mask, const float* __restrict__ gt,
                                                            const float* __restrict__ weight, float* __restrict__ filter_out,
                                                            long long n, int use_clip, float* __restrict__ part) {
    __shared__ float red[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s_sq = 0.f, s_f = 0.f, s_w = 0.f;
    const long long nq = n >> 2;
    for (long long qd = (long long)blockIdx.x * 256 + threadIdx.x; qd < nq; qd += (long long)gridDim.x * 256) {
        const float4 m4 = reinterpret_cast<const float4*>(mask)[qd], g4 = reinterpret_cast<const float4*>(gt)[qd];
        float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (weight) w4 = reinterpret_cast<const float4*>(weight)[qd];
        const float m[4] = {m4.x, m4.y, m4.z, m4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
        float f[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = m[i] - g[i], sq = d * d;
            f[i] = (!use_clip || m[i] > 0.1f) ? 1.0f : 0.0f;
            s_sq += sq; s_f += f[i]; s_w = fmaf(sq * f[i], w[i], s_w);
        }
        if (filter_out) reinterpret_cast<float4*>(filter_out)[qd] = make_float4(f[0], f[1], f[2], f[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {                // ragged tail (n not a multiple of 4)
        const long long i = (nq << 2) + threadIdx.x;
        const float d = mask[i] - gt[i], sq = d * d, f = (!use_clip || mask[i] > 0.1f) ? 1.0f : 0.0f;
        s_sq += sq; s_f += f; s_w = fmaf(sq * f, weight ? weight[i] : 1.0f, s_w);
        if (filter_out) filter_out[i] = f;
    }
    s_sq = warp_sum(s_sq); s_f = warp_sum(s_f); s_w = warp_sum(s_w);
    if (lane == 0) { red[warp][0] = s_sq; red[warp][1] = s_f; red[warp][2] = s_w; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        part[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ gt,
                                                            const float* __restrict__ weight, long long n, int mode, int use_clip,
                                                            const float* __restrict__ sums, const float* __restrict__ g_loss,
                                                            float* __restrict__ g_mask) {
    const float lscale = loss_grad_scale(mode, (double)n, sums, g_loss);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float m = mask[i];
        float s = lscale * (m - gt[i]);
        if (mode == XSUP_MASK_WEIGHTED) s *= ((!use_clip || m > 0.1f) ? (weight ? weight[i] : 1.0f) : 0.0f);
        g_mask[i] = s;
    }
}

// ---------------------------------------------------------------------------------------------- launchers
int skel_chunks(int S) {
    const int tiles = ((S + kTileW - 1) / kTileW) * ((S + kTileH - 1) / kTileH);
    return (tiles + kTilesPerCta - 1) / kTilesPerCta;
}
int draw_lines_chunks(int S) { return (S * S / 4 + 255) / 256; }
int mask_loss_ctas(long long n) {
    const long long want = ((n >> 2) + 255) / 256;
    return (int)(want < 1 ? 1 : (want > 1184 ? 1184 : want));          // fixed by n alone: sums are reproducible on any device
}

static void fill_tiles(SkelParams& p) {
    p.tiles_x = (p.S + kTileW - 1) / kTileW;
    p.tiles_x_magic = p.tiles_x == 1 ? 0xffffffffu : (unsigned)(0x100000000ULL / (unsigned)p.tiles_x);   // floor: __umulhi may be one short, fixed up in the kernel
    p.tiles = p.tiles_x * ((p.S + kTileH - 1) / kTileH);
    p.NC = skel_chunks(p.S);
}

cudaError_t launch_skeleton_mask_fwd(SkelParams p, float* recon, uint8_t* line_idx, const float* gt, const float* weight,
                                     const xsup_mask_loss_t* loss, float* loss_sums, float* ws, cudaStream_t st) {
    fill_tiles(p);
    const dim3 grid(p.NC, p.B);
    const size_t tab_bytes = (size_t)p.S * sizeof(float);
    float* ws_loss = ws + (size_t)p.B * p.NC * XSUP_MAX_LINES * 4;
    if (loss) {
        skeleton_mask_fwd_kernel<true><<<grid, kSkelThreads, tab_bytes, st>>>(p, recon, line_idx, gt, weight,
                                                                      loss->use_clip || loss->mode == XSUP_MASK_CLIP_MEAN, ws_loss);
        mask_loss_finalize_kernel<<<1, 256, 0, st>>>(ws_loss, p.B * p.NC, (double)loss->n, loss->mode, loss_sums);
    } else {
        skeleton_mask_fwd_kernel<false><<<grid, kSkelThreads, tab_bytes, st>>>(p, recon, line_idx, nullptr, nullptr, 0, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t launch_skeleton_mask_bwd(SkelParams p, const float* recon, const uint8_t* line_idx, const float* g_recon,
                                     const float* gt, const float* weight, const xsup_mask_loss_t* loss, const float* loss_sums,
                                     const float* g_loss, float* g_kps, float* ws, cudaStream_t st) {
    fill_tiles(p);
    const dim3 grid(p.NC, p.B);
    const size_t tab_bytes = (size_t)p.S * sizeof(float);
    if (loss)
        skeleton_mask_bwd_kernel<true><<<grid, kSkelThreads, tab_bytes, st>>>(p, recon, line_idx, g_recon, gt, weight, loss->mode,
                                                                     loss->use_clip || loss->mode == XSUP_MASK_CLIP_MEAN,
                                                                     (double)loss->n, loss_sums, g_loss, ws);
    else
        skeleton_mask_bwd_kernel<false><<<grid, kSkelThreads, tab_bytes, st>>>(p, recon, line_idx, g_recon, nullptr, nullptr, 0, 0, 1.0,
                                                                      nullptr, nullptr, ws);
    skeleton_scatter_kernel<<<p.B, 128, 0, st>>>(p, p.NC, ws, g_kps);
    return cudaGetLastError();
}

cudaError_t launch_draw_lines_fwd(SkelParams p, float* heat, cudaStream_t st) {
    draw_lines_fwd_kernel<<<dim3(draw_lines_chunks(p.S), p.L, p.B), 256, 0, st>>>(p, heat);
    return cudaGetLastError();
}

cudaError_t launch_draw_lines_bwd(SkelParams p, const float* heat, const float* g_heat, float* g_kps, float* ws, cudaStream_t st) {
    const int nc = draw_lines_chunks(p.S);
    draw_lines_bwd_kernel<<<dim3(nc, p.L, p.B), 256, 0, st>>>(p, heat, g_heat, ws);
    skeleton_scatter_kernel<<<p.B, 128, 0, st>>>(p, nc, ws, g_kps);
    return cudaGetLastError();
}

cudaError_t launch_mask_loss_fwd(const float* mask, const float* gt, const float* weight, float* filter_out,
                                 const xsup_mask_loss_t& c, float* loss_sums, float* ws, cudaStream_t st) {
    const int ctas = mask_loss_ctas(c.n);
    mask_loss_fwd_kernel<<<ctas, 256, 0, st>>>(mask, gt, weight, filter_out, c.n, c.use_clip || c.mode == XSUP_MASK_CLIP_MEAN, ws);
    mask_loss_finalize_kernel<<<1, 256, 0, st>>>(ws, ctas, (double)c.n, c.mode, loss_sums);
    return cudaGetLastError();
}

cudaError_t launch_mask_loss_bwd(const float* mask, const float* gt, const float* weight, const xsup_mask_loss_t& c,
                                 const float* loss_sums, const float* g_loss, float* g_mask, int num_sms, cudaStream_t st) {
    const long long want = (c.n + 255) / 256;
    const int ctas = (int)(want < 1 ? 1 : (want > (long long)num_sms * 8 ? (long long)num_sms * 8 : want));
    mask_loss_bwd_kernel<<<ctas, 256, 0, st>>>(mask, gt, weight, c.n, c.mode, c.use_clip || c.mode == XSUP_MASK_CLIP_MEAN,
                                               loss_sums, g_loss, g_mask);
    return cudaGetLastError();
}

}  // namespace xsup
