// K2 — per-sample camera geometry, loss terms and min-over-hypotheses selection.
//
// Replaces, for one camera,
//   modules/util.py:61-95,128-152     convert_patch_to_world (inverse crop affine, px->mm depth,
//                                     pinhole back-projection, inverse extrinsics) and :98-125,155-168
//                                     its inverse (forward perspective projection),
//   modules/base_losses/loss_func.py:18-52   bone / keypoint symmetry and pseudo-GT MSE,
//   modules/model.py:71-79,105-114,158-162   the per-hypothesis Python loops and torch.min(torch.stack()),
//   eval.py:138-145 / loss_func.py:59        per-joint argmin / per-sample min.
// The reference issues ~10^2 tiny kernels and two batched LU inversions per camera for this; here it is
// three launches of a few microseconds.  All reductions are fixed-order (no float atomics), so the
// selected hypothesis is reproducible run to run.
#include "xsup_internal.h"

namespace xsup {

__constant__ int c_bone_child[8] = {16, 15, 13, 12, 3, 2, 6, 5};    // loss_func.py:20
__constant__ int c_bone_parent[8] = {15, 14, 12, 11, 2, 1, 5, 4};
__constant__ int c_mid_a[2] = {11, 1};                              // loss_func.py:28
__constant__ int c_mid_b[2] = {14, 4};

enum { F_NORM = 1, F_MONO = 2, F_PATCH = 4 };

struct Cam {
    float ai00, ai01, ai10, ai11;   // inverse of trans_image[:, :, :2]
    float a00, a01, a10, a11, t0, t1;
    float pz;                       // pelvis depth
    float fx, fy, cx, cy;
    float ri[9];                    // inverse of rot_world (general inverse, util.py:93)
    float r[9], tw[3];
};

__device__ __forceinline__ Cam load_cam(const xsup_cam_t& c, int b) {
    Cam m;
    const float* A = c.trans_image + (size_t)b * 6;
    m.a00 = A[0]; m.a01 = A[1]; m.t0 = A[2]; m.a10 = A[3]; m.a11 = A[4]; m.t1 = A[5];
    const float idet = 1.0f / (m.a00 * m.a11 - m.a01 * m.a10);
    m.ai00 = m.a11 * idet; m.ai01 = -m.a01 * idet; m.ai10 = -m.a10 * idet; m.ai11 = m.a00 * idet;
    m.pz = c.pelvis[(size_t)b * 3 + 2];
    const float* K = c.k_mat + (size_t)b * 9;
    m.fx = K[0]; m.fy = K[4]; m.cx = K[2]; m.cy = K[5];
    const float* R = c.rot_world + (size_t)b * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.r[i] = R[i];
    const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
    const float rdet = 1.0f / (R[0] * c00 + R[1] * c01 + R[2] * c02);
    m.ri[0] = c00 * rdet; m.ri[1] = (R[2] * R[7] - R[1] * R[8]) * rdet; m.ri[2] = (R[1] * R[5] - R[2] * R[4]) * rdet;
    m.ri[3] = c01 * rdet; m.ri[4] = (R[0] * R[8] - R[2] * R[6]) * rdet; m.ri[5] = (R[2] * R[3] - R[0] * R[5]) * rdet;
    m.ri[6] = c02 * rdet; m.ri[7] = (R[1] * R[6] - R[0] * R[7]) * rdet; m.ri[8] = (R[0] * R[4] - R[1] * R[3]) * rdet;
    const float* T = c.trans_world + (size_t)b * 3;
    m.tw[0] = T[0]; m.tw[1] = T[1]; m.tw[2] = T[2];
    return m;
}

struct Img { float wm1, hm1, dm1, ds; };
__device__ __forceinline__ Img make_img(int img_h, int img_w, float rect_width) {
    // util.py:137-138: depth extent = image width, depth_scale = RECT_WIDTH / width
    return Img{(float)(img_w - 1), (float)(img_h - 1), (float)(img_w - 1), 1.0f / (float)img_w * rect_width};
}

// (x,y,z) patch -> world; also returns the intermediates the VJP needs
__device__ __forceinline__ void patch_to_world(const Cam& m, const Img& im, int flags, float x, float y, float z,
                                               float (&w)[3], float& u, float& v, float& Z) {
    if (flags & F_PATCH) {
        if (flags & F_NORM) {
            x = (x + 1.0f) / 2.0f * im.wm1;                       // util.py:70-72
            y = (y + 1.0f) / 2.0f * im.hm1;
            z = z * im.dm1;
        }
        const float du = x - m.t0, dv = y - m.t1;                 // util.py:65,76
        u = m.ai00 * du + m.ai01 * dv;
        v = m.ai10 * du + m.ai11 * dv;
        Z = z * im.ds + m.pz;                                     // util.py:79-80
    } else {
        u = x; v = y; Z = z;
    }
    if (flags & F_MONO) {                                         // util.py:145-150
        w[0] = -u; w[1] = -(Z + 128.0f); w[2] = -v;
        return;
    }
    const float X = (u - m.cx) / m.fx * Z - m.tw[0];              // util.py:89-90,93
    const float Y = (v - m.cy) / m.fy * Z - m.tw[1];
    const float Zc = Z - m.tw[2];
    w[0] = m.ri[0] * X + m.ri[1] * Y + m.ri[2] * Zc;
    w[1] = m.ri[3] * X + m.ri[4] * Y + m.ri[5] * Zc;
    w[2] = m.ri[6] * X + m.ri[7] * Y + m.ri[8] * Zc;
}

// vector-Jacobian product of patch_to_world: gw (d loss / d world) -> (gx, gy, gz)
__device__ __forceinline__ void patch_to_world_vjp(const Cam& m, const Img& im, int flags, float u, float v, float Z,
                                                   const float (&gw)[3], float (&g)[3]) {
    float gu, gv, gZ;
    if (flags & F_MONO) {
        gu = -gw[0]; gZ = -gw[1]; gv = -gw[2];
    } else {
        const float gX = m.ri[0] * gw[0] + m.ri[3] * gw[1] + m.ri[6] * gw[2];
        const float gY = m.ri[1] * gw[0] + m.ri[4] * gw[1] + m.ri[7] * gw[2];
        const float gC = m.ri[2] * gw[0] + m.ri[5] * gw[1] + m.ri[8] * gw[2];
        gu = gX * Z / m.fx;
        gv = gY * Z / m.fy;
        gZ = gC + gX * (u - m.cx) / m.fx + gY * (v - m.cy) / m.fy;
    }
    if (flags & F_PATCH) {
        float gx = m.ai00 * gu + m.ai10 * gv, gy = m.ai01 * gu + m.ai11 * gv, gz = gZ * im.ds;
        if (flags & F_NORM) { gx *= im.wm1 * 0.5f; gy *= im.hm1 * 0.5f; gz *= im.dm1; }
        g[0] = gx; g[1] = gy; g[2] = gz;
    } else {
        g[0] = gu; g[1] = gv; g[2] = gZ;
    }
}

// ---------------------------------------------------------------------------------------------- standalone geometry
// The four stage functions of util.py:61-125 and their two composites (:128-168) are one pair of kernels driven by
// stage flags: XSUP_GEOM_PATCH_STAGE = patch<->image (crop affine, px<->mm depth), XSUP_GEOM_CAMERA_STAGE =
// image<->world (pinhole, extrinsics).  Tensors a stage does not need may be NULL.
__device__ __forceinline__ Cam load_cam_g(const xsup_geom_t& g, int b) {
    Cam m;
    m.a00 = m.a11 = m.ai00 = m.ai11 = 1.f; m.a01 = m.a10 = m.ai01 = m.ai10 = 0.f; m.t0 = m.t1 = 0.f; m.pz = 0.f;
    m.fx = m.fy = 1.f; m.cx = m.cy = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.r[i] = m.ri[i] = (i % 4 == 0) ? 1.f : 0.f;
    m.tw[0] = m.tw[1] = m.tw[2] = 0.f;
    if (g.flags & XSUP_GEOM_PATCH_STAGE) {
        const float* A = g.trans_image + (size_t)b * 6;
        m.a00 = A[0]; m.a01 = A[1]; m.t0 = A[2]; m.a10 = A[3]; m.a11 = A[4]; m.t1 = A[5];
        const float idet = 1.0f / (m.a00 * m.a11 - m.a01 * m.a10);
        m.ai00 = m.a11 * idet; m.ai01 = -m.a01 * idet; m.ai10 = -m.a10 * idet; m.ai11 = m.a00 * idet;
        m.pz = g.pelvis[(size_t)b * 3 + 2];
    }
    if ((g.flags & XSUP_GEOM_CAMERA_STAGE) && !(g.flags & XSUP_GEOM_MONO)) {
        const size_t is = (size_t)b * g.intr_stride;
        m.fx = g.fx[is]; m.fy = g.fy[is]; m.cx = g.cx[is]; m.cy = g.cy[is];
        const float* R = g.rot_world + (size_t)b * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) m.r[i] = R[i];
        const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
        const float rdet = 1.0f / (R[0] * c00 + R[1] * c01 + R[2] * c02);
        m.ri[0] = c00 * rdet; m.ri[1] = (R[2] * R[7] - R[1] * R[8]) * rdet; m.ri[2] = (R[1] * R[5] - R[2] * R[4]) * rdet;
        m.ri[3] = c01 * rdet; m.ri[4] = (R[0] * R[8] - R[2] * R[6]) * rdet; m.ri[5] = (R[2] * R[3] - R[0] * R[5]) * rdet;
        m.ri[6] = c02 * rdet; m.ri[7] = (R[1] * R[6] - R[0] * R[7]) * rdet; m.ri[8] = (R[0] * R[4] - R[1] * R[3]) * rdet;
        const float* T = g.trans_world + (size_t)b * 3;
        m.tw[0] = T[0]; m.tw[1] = T[1]; m.tw[2] = T[2];
    }
    return m;
}
__device__ __forceinline__ Img make_img_g(const xsup_geom_t& g) {
    return Img{(float)(g.img_w - 1), (float)(g.img_h - 1), (float)(g.img_d - 1), g.depth_scale};
}
__device__ __forceinline__ int internal_flags(const xsup_geom_t& g) {
    return ((g.flags & XSUP_GEOM_NORM) ? F_NORM : 0) | ((g.flags & XSUP_GEOM_MONO) ? F_MONO : 0) |
           ((g.flags & XSUP_GEOM_PATCH_STAGE) ? F_PATCH : 0);
}

// world/image -> image/patch (util.py:116-125 then :98-113) and the intermediates its VJP needs
__device__ __forceinline__ void world_to_patch(const Cam& m, const Img& im, int gflags, float wx, float wy, float wz, float (&o)[3],
                                               float& X, float& Y, float& Z) {
    float u = wx, v = wy, zc = wz;
    X = wx; Y = wy; Z = wz;
    if (gflags & XSUP_GEOM_CAMERA_STAGE) {
        X = m.r[0] * wx + m.r[1] * wy + m.r[2] * wz + m.tw[0];               // util.py:120
        Y = m.r[3] * wx + m.r[4] * wy + m.r[5] * wz + m.tw[1];
        Z = m.r[6] * wx + m.r[7] * wy + m.r[8] * wz + m.tw[2];
        u = X / Z * m.fx + m.cx;                                             // util.py:122-123
        v = Y / Z * m.fy + m.cy;
        zc = Z;
    }
    if (gflags & XSUP_GEOM_PATCH_STAGE) {
        float z = (zc - m.pz) / im.ds;                                       // util.py:102-103
        float x = m.a00 * u + m.a01 * v + m.t0, y = m.a10 * u + m.a11 * v + m.t1;
        if (gflags & XSUP_GEOM_NORM) {                                       // util.py:108-111
            x = x / im.wm1 * 2.0f - 1.0f;
            y = y / im.hm1 * 2.0f - 1.0f;
            z = z / im.dm1;
        }
        o[0] = x; o[1] = y; o[2] = z;
    } else {
        o[0] = u; o[1] = v; o[2] = zc;
    }
}
__device__ __forceinline__ void world_to_patch_vjp(const Cam& m, const Img& im, int gflags, float X, float Y, float Z,
                                                   const float (&go)[3], float (&gi)[3]) {
    float gu = go[0], gv = go[1], gz = go[2];
    if (gflags & XSUP_GEOM_PATCH_STAGE) {
        float gx = go[0], gy = go[1], gzz = go[2];
        if (gflags & XSUP_GEOM_NORM) { gx = gx / im.wm1 * 2.0f; gy = gy / im.hm1 * 2.0f; gzz = gzz / im.dm1; }
        gu = m.a00 * gx + m.a10 * gy;
        gv = m.a01 * gx + m.a11 * gy;
        gz = gzz / im.ds;
    }
    if (gflags & XSUP_GEOM_CAMERA_STAGE) {
        const float iz = 1.0f / Z;
        const float gX = gu * m.fx * iz, gY = gv * m.fy * iz;
        const float gZ = gz - (gX * X + gY * Y) * iz;
        gi[0] = m.r[0] * gX + m.r[3] * gY + m.r[6] * gZ;
        gi[1] = m.r[1] * gX + m.r[4] * gY + m.r[7] * gZ;
        gi[2] = m.r[2] * gX + m.r[5] * gY + m.r[8] * gZ;
    } else {
        gi[0] = gu; gi[1] = gv; gi[2] = gz;
    }
}

// dir 0: patch -> world direction (convert_patch_to_image / convert_image_to_world / convert_patch_to_world)
// dir 1: world -> patch direction (convert_world_to_image / convert_image_to_patch / convert_world_to_patch)
// g_out == nullptr: forward (out = f(in)); else out = VJP of g_out at `in`.
template <int DIR>
__global__ void __launch_bounds__(128) geom_kernel(const float* __restrict__ in, const float* __restrict__ g_out, float* __restrict__ out,
                                                   const xsup_geom_t g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.B * g.J) return;
    const Cam m = load_cam_g(g, i / g.J);
    const Img im = make_img_g(g);
    const float a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
    float o[3];
    if (DIR == 0) {
        const int fl = internal_flags(g);
        float w[3], u, v, Z;
        if (g.flags & XSUP_GEOM_CAMERA_STAGE) {
            patch_to_world(m, im, fl, a, b, c, w, u, v, Z);
            if (g_out) {
                const float gw[3] = {g_out[3 * i], g_out[3 * i + 1], g_out[3 * i + 2]};
                patch_to_world_vjp(m, im, fl, u, v, Z, gw, o);
            } else { o[0] = w[0]; o[1] = w[1]; o[2] = w[2]; }
        } else {                                                  // convert_patch_to_image alone: the affine part only
            patch_to_world(m, im, fl | F_MONO, a, b, c, w, u, v, Z);   // (u, v, Z) are the image-space point
            if (g_out) {
                float gx = m.ai00 * g_out[3 * i] + m.ai10 * g_out[3 * i + 1], gy = m.ai01 * g_out[3 * i] + m.ai11 * g_out[3 * i + 1],
                      gz = g_out[3 * i + 2] * im.ds;
                if (g.flags & XSUP_GEOM_NORM) { gx *= im.wm1 * 0.5f; gy *= im.hm1 * 0.5f; gz *= im.dm1; }
                if (!(g.flags & XSUP_GEOM_PATCH_STAGE)) { gx = g_out[3 * i]; gy = g_out[3 * i + 1]; gz = g_out[3 * i + 2]; }
                o[0] = gx; o[1] = gy; o[2] = gz;
            } else { o[0] = u; o[1] = v; o[2] = Z; }
        }
    } else {
        float X, Y, Z;
        world_to_patch(m, im, g.flags, a, b, c, o, X, Y, Z);
        if (g_out) {
            const float go[3] = {g_out[3 * i], g_out[3 * i + 1], g_out[3 * i + 2]};
            float gi[3];
            world_to_patch_vjp(m, im, g.flags, X, Y, Z, go, gi);
            o[0] = gi[0]; o[1] = gi[1]; o[2] = gi[2];
        }
    }
    out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

cudaError_t launch_geom(int dir, const float* in, const float* g_out, float* out, const xsup_geom_t& g, cudaStream_t st) {
    const int n = g.B * g.J;
    if (dir == 0) geom_kernel<0><<<(n + 127) / 128, 128, 0, st>>>(in, g_out, out, g);
    else geom_kernel<1><<<(n + 127) / 128, 128, 0, st>>>(in, g_out, out, g);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- loss forward
// one warp per (sample, hypothesis); lane = joint (K <= 32)
__device__ __forceinline__ void loss_fwd_warp(const float* __restrict__ kps, const float* __restrict__ target, const xsup_cam_t& cam,
                                              float* __restrict__ world, float* __restrict__ sample_terms, const xsup_loss_cfg_t& c,
                                              int wid, int lane) {
    const int b = wid / c.NH, h = wid - b * c.NH, K = c.K;
    const Cam m = load_cam(cam, b);
    const Img im = make_img(c.img_h, c.img_w, c.rect_width);
    float x = 0.f, y = 0.f, z = 0.f, w[3] = {0.f, 0.f, 0.f}, se = 0.f;
    if (lane < K) {
        const size_t o = (((size_t)b * c.NH + h) * K + lane) * 3;
        x = kps[o]; y = kps[o + 1]; z = kps[o + 2];
        const float* tg = target + ((size_t)b * K + lane) * 3;
        const float dx = x - tg[0], dy = y - tg[1], dz = z - tg[2];
        se = dx * dx + dy * dy + dz * dz;
        float u, v, Z;
        patch_to_world(m, im, F_NORM | F_PATCH, x, y, z, w, u, v, Z);
        world[o] = w[0]; world[o + 1] = w[1]; world[o + 2] = w[2];
    }
    se = warp_sum(se);
    float bone = 0.f, kp3 = 0.f, kp2 = 0.f;
    if (c.use_sym) {
        // bones: lane i < 8 owns bone i (loss_func.py:20-21)
        const int ci = c_bone_child[lane & 7], pi = c_bone_parent[lane & 7];
        const float vx = __shfl_sync(0xffffffffu, w[0], ci) - __shfl_sync(0xffffffffu, w[0], pi);
        const float vy = __shfl_sync(0xffffffffu, w[1], ci) - __shfl_sync(0xffffffffu, w[1], pi);
        const float vz = __shfl_sync(0xffffffffu, w[2], ci) - __shfl_sync(0xffffffffu, w[2], pi);
        const float n = sqrtf(vx * vx + vy * vy + vz * vz) * 1e-3f;
        const float nn = __shfl_down_sync(0xffffffffu, n, 1);
        const float df = n - nn;
        bone = warp_sum((lane < 8 && !(lane & 1)) ? df * df : 0.f);           // pairs (0,1),(2,3),(4,5),(6,7)
        // midpoints: lane s < 2 owns pair s (loss_func.py:28-35)
        const int ia = c_mid_a[lane & 1], ib = c_mid_b[lane & 1], ir = (lane & 1) ? 0 : K - 1;
        float e3 = 0.f, e2 = 0.f;
        {
            const float a0 = __shfl_sync(0xffffffffu, w[0], ia), b0 = __shfl_sync(0xffffffffu, w[0], ib), r0 = __shfl_sync(0xffffffffu, w[0], ir);
            const float a1 = __shfl_sync(0xffffffffu, w[1], ia), b1 = __shfl_sync(0xffffffffu, w[1], ib), r1 = __shfl_sync(0xffffffffu, w[1], ir);
            const float a2 = __shfl_sync(0xffffffffu, w[2], ia), b2 = __shfl_sync(0xffffffffu, w[2], ib), r2 = __shfl_sync(0xffffffffu, w[2], ir);
            const float d0 = (a0 + b0) / 2.0f * 1e-3f - r0 * 1e-3f, d1 = (a1 + b1) / 2.0f * 1e-3f - r1 * 1e-3f,
                        d2 = (a2 + b2) / 2.0f * 1e-3f - r2 * 1e-3f;
            e3 = d0 * d0 + d1 * d1 + d2 * d2;
            const float xa = __shfl_sync(0xffffffffu, x, ia), xb = __shfl_sync(0xffffffffu, x, ib), xr = __shfl_sync(0xffffffffu, x, ir);
            const float ya = __shfl_sync(0xffffffffu, y, ia), yb = __shfl_sync(0xffffffffu, y, ib), yr = __shfl_sync(0xffffffffu, y, ir);
            const float q0 = (xa + xb) / 2.0f - xr, q1 = (ya + yb) / 2.0f - yr;
            e2 = q0 * q0 + q1 * q1;
        }
        kp3 = warp_sum(lane < 2 ? e3 : 0.f);
        kp2 = warp_sum(lane < 2 ? e2 : 0.f);
    }
    if (lane == 0) {
        float* o = sample_terms + (size_t)b * XSUP_LOSS_TERMS * c.NH + h;
        o[0] = se; o[c.NH] = bone; o[2 * c.NH] = kp3; o[3 * c.NH] = kp2;
    }
}

__global__ void __launch_bounds__(128) reproj_loss_fwd_kernel(const float* __restrict__ kps, const float* __restrict__ target,
                                                              const xsup_cam_t cam, float* __restrict__ world,
                                                              float* __restrict__ sample_terms, const xsup_loss_cfg_t c) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= c.B * c.NH) return;
    loss_fwd_warp(kps, target, cam, world, sample_terms, c, wid, lane);
}

// partial[col] = sum_b sample_terms[b,col], col = (term,h), by one block of 256 threads, in a fixed order, so the sums - and with
// them the selected slot - are bit-reproducible.  A row is 4*NH floats = NH float4s: thread t owns float4-column t % NHP (NHP = NH
// rounded up to a power of two) of the rows t / NHP, t / NHP + 256 / NHP, ...: neighbouring threads read neighbouring 16 bytes and
// every load is independent of the others (B = 4 096, NH = 3: 64 LDG.128 per thread, where one warp per column took 2 x 128 dependent
// round trips).  Then the xor tree over the lanes that share a column, the 8 warps (or the 256 / NHP row groups) in order.
// L2 loads: in the fused kernel the terms were written by other CTAs of the same grid.  `sh`: 256 float4.
static_assert(XSUP_LOSS_TERMS == 4, "a row of sample_terms is read as NH float4s");
__device__ void partial_block(const float* sample_terms, float* partial, int B, int NH, float4* sh) {
    const int t = threadIdx.x;
    int nhp_log2 = 0;
    while ((1 << nhp_log2) < NH) ++nhp_log2;
    const int nhp = 1 << nhp_log2, c4 = t & (nhp - 1), rg = t >> nhp_log2, nrg = 256 >> nhp_log2;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < NH) {
        const float4* src = reinterpret_cast<const float4*>(sample_terms) + c4;
#pragma unroll 8
        for (int b = rg; b < B; b += nrg) {
            const float4 v = __ldcg(src + (size_t)b * NH);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    }
    for (int o = 16; o >= nhp; o >>= 1) {                                     // lanes of a warp that share a column (NHP < 32)
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
        a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    sh[t] = a;
    __syncthreads();
    const float* shf = reinterpret_cast<const float*>(sh);
    const int step = nhp < 32 ? 32 : nhp, n = 256 / step;                     // contributors of a column: threads c4, c4 + step, ...
    for (int col = t; col < XSUP_LOSS_TERMS * NH; col += 256) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r += shf[(i * step + (col >> 2)) * 4 + (col & 3)];
        partial[col] = r;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256) reproj_partial_kernel(const float* sample_terms, float* __restrict__ partial, int B, int NH) {
    __shared__ float4 sh[256];
    partial_block(sample_terms, partial, B, NH, sh);
}

cudaError_t launch_reproj_loss_fwd(const float* kps, const float* target, const xsup_cam_t& cam, float* world,
                                   float* sample_terms, float* partial, const xsup_loss_cfg_t& c, cudaStream_t st) {
    const int warps = c.B * c.NH;
    reproj_loss_fwd_kernel<<<(warps + 3) / 4, 128, 0, st>>>(kps, target, cam, world, sample_terms, c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reproj_partial_kernel<<<1, 256, 0, st>>>(sample_terms, partial, c.B, c.NH);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- cross-GPU exchange
// All-reduce(SUM) of the [4,NH] partial sums in one kernel over NVLink peer memory (no NCCL, no stream
// hand-off): publish into every peer's mailbox, acquire-spin on the own mailbox, sum in rank order.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Block-level body (any blockDim >= world, all threads of the block call it; ends with every thread seeing buf[]).
// A peer that never arrives (~10 s) turns the sums into NaN and raises the sticky word *err, which the host checks
// (dist.PeerExchange.check): after a timeout the mailboxes are out of step and the exchange must be rebuilt.
__device__ void xchg_allreduce_block(float* buf, int n, void* const* __restrict__ peers, int rank, int world, uint32_t step_arg,
                                     uint32_t* seq, uint32_t* err) {
    __shared__ int timed_out;
    __shared__ uint32_t s_step;
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) {
        timed_out = 0;
        s_step = seq ? (*seq += 1u) : step_arg;          // device-side sequence number: replayable from a CUDA graph
    }
    __syncthreads();
    const uint32_t step = s_step;
    const size_t par = step & 1u;
    for (int d = 0; d < world; ++d) {                  // publish: coalesced P2P stores into slot [par][rank] of every mailbox
        float* dst = static_cast<float*>(peers[d]) + (par * world + rank) * XSUP_XCHG_SLOT;
        for (int i = t; i < n; i += nt) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (t < world)                                     // one thread per destination GPU raises that mailbox's flag
        st_release_sys(reinterpret_cast<uint32_t*>(static_cast<float*>(peers[t]) + (par * world + rank) * XSUP_XCHG_SLOT + XSUP_XCHG_SLOT - 1), step);
    float* mine = static_cast<float*>(peers[rank]) + par * world * XSUP_XCHG_SLOT;
    if (t < world) {                                   // one thread per source GPU
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + (size_t)t * XSUP_XCHG_SLOT + XSUP_XCHG_SLOT - 1);
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != step) {
            __nanosleep(64);
            if (clock64() - t0 > 20000000000LL) { timed_out = 1; break; }     // ~10 s at 2 GHz: a peer died
        }
    }
    __syncthreads();
    for (int i = t; i < n; i += nt) {                  // fixed rank order: bit-identical sums on every rank
        float a = 0.f;
        for (int r = 0; r < world; ++r) a += *reinterpret_cast<volatile float*>(mine + (size_t)r * XSUP_XCHG_SLOT + i);
        buf[i] = timed_out ? __int_as_float(0x7fc00000) : a;
    }
    if (t == 0 && timed_out && err) *err = 1u;
    __syncthreads();
}

__global__ void __launch_bounds__(64) partial_allreduce_kernel(float* partial, int n, const xsup_xchg_t x) {
    xchg_allreduce_block(partial, n, x.peer_bufs, x.rank, x.world, x.step, x.seq, x.err);
}

cudaError_t launch_partial_allreduce(float* partial, int n, const xsup_xchg_t& x, cudaStream_t st) {
    partial_allreduce_kernel<<<1, 64, 0, st>>>(partial, n, x);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- selection
__device__ __forceinline__ float sym_value(const xsup_loss_cfg_t& c, float bone, float kp3, float kp2, float n) {
    // model.py:108-112: bone MSE over B*4, kp MSE over B*2*3, kp_2d MSE over B*2*2 (x 1e2)
    return c.w_bone * bone / (n * 4.0f) + c.w_kp * kp3 / (n * 6.0f) + c.w_kp2d * 1e2f * kp2 / (n * 4.0f);
}

__device__ float block_sum_256(float v, float* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += scratch[i];
    return r;
}

// 256 threads.  `sample_terms` / `partial` may have been written earlier in the same kernel (fused forward): no
// __restrict__ / read-only path on them.
__device__ void select_block(const float* __restrict__ kps, const float* __restrict__ target, const float* sample_terms,
                             const float* partial, float* loss, int64_t* __restrict__ sel, const xsup_loss_cfg_t& c) {
    __shared__ float scratch[8];
    const int NH = c.NH, K = c.K, B = c.B;
    const float n = (float)c.batch_total;
    if (c.reduction == XSUP_REDUCE_BATCH) {
        if (threadIdx.x == 0) {
            int sm = 0, ss = -1;
            float vm = partial[0] / (n * (float)K * 3.0f), vs = 0.f;
            for (int h = 1; h < NH; ++h) {
                const float v = partial[h] / (n * (float)K * 3.0f);
                if (v < vm) { vm = v; sm = h; }
            }
            if (c.use_sym) {
                ss = 0;
                vs = sym_value(c, partial[NH], partial[2 * NH], partial[3 * NH], n);
                for (int h = 1; h < NH; ++h) {
                    const float v = sym_value(c, partial[NH + h], partial[2 * NH + h], partial[3 * NH + h], n);
                    if (v < vs) { vs = v; ss = h; }
                }
            }
            loss[0] = c.w_mse * vm; loss[1] = vs;
            sel[0] = sm; sel[1] = ss;
        }
        return;
    }
    if (c.reduction == XSUP_REDUCE_SAMPLE) {
        float am = 0.f, as = 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float* t = sample_terms + (size_t)b * XSUP_LOSS_TERMS * NH;
            int sm = 0, ss = -1;
            float vm = c.w_mse * t[0] / ((float)K * 3.0f), vs = 0.f;
            for (int h = 1; h < NH; ++h) {
                const float v = c.w_mse * t[h] / ((float)K * 3.0f);
                if (v < vm) { vm = v; sm = h; }
            }
            if (c.use_sym) {
                ss = 0;
                vs = sym_value(c, t[NH], t[2 * NH], t[3 * NH], 1.0f);
                for (int h = 1; h < NH; ++h) {
                    const float v = sym_value(c, t[NH + h], t[2 * NH + h], t[3 * NH + h], 1.0f);
                    if (v < vs) { vs = v; ss = h; }
                }
            }
            sel[b] = sm; sel[B + b] = ss;
            am += vm; as += vs;
        }
        am = block_sum_256(am, scratch);
        as = block_sum_256(as, scratch);
        if (threadIdx.x == 0) { loss[0] = am / n; loss[1] = as / n; }
        return;
    }
    // joint: eval.py:138-145
    float acc = 0.f;
    for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
        const int b = i / K, k = i - b * K;
        const float* tg = target + (size_t)i * 3;
        int s = 0;
        float best = 0.f;
        for (int h = 0; h < NH; ++h) {
            const float* p = kps + (((size_t)b * NH + h) * K + k) * 3;
            const float dx = p[0] - tg[0], dy = p[1] - tg[1], dz = p[2] - tg[2];
            const float e = dx * dx + dy * dy + dz * dz;
            if (h == 0 || e < best) { best = e; s = h; }
        }
        sel[i] = s;
        acc += best;
    }
    acc = block_sum_256(acc, scratch);
    if (threadIdx.x == 0) { loss[0] = c.w_mse * acc / (n * (float)K * 3.0f); loss[1] = 0.f; }
}

__global__ void __launch_bounds__(256) reproj_select_kernel(const float* __restrict__ kps, const float* __restrict__ target,
                                                            const float* sample_terms, const float* partial, float* loss,
                                                            int64_t* __restrict__ sel, const xsup_loss_cfg_t c) {
    select_block(kps, target, sample_terms, partial, loss, sel, c);
}

// ---------------------------------------------------------------------------------------------- fused forward
// loss terms per (sample, hypothesis) -> last CTA done: fixed-order batch sums -> [NVLink exchange] -> selection.
// One launch instead of loss_fwd + partial + exchange + select; the exchange rides inside the compute kernel.
__global__ void __launch_bounds__(256) reproj_fused_fwd_kernel(const float* __restrict__ kps, const float* __restrict__ target,
                                                               const xsup_cam_t cam, float* __restrict__ world, float* sample_terms,
                                                               float* partial, float* loss, int64_t* __restrict__ sel,
                                                               const xsup_loss_cfg_t c, const xsup_xchg_t x, unsigned int* ticket) {
    __shared__ int s_last;
    __shared__ float4 sh_part[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wid = blockIdx.x * 8 + warp;
    if (wid < c.B * c.NH) loss_fwd_warp(kps, target, cam, world, sample_terms, c, wid, lane);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int cols = XSUP_LOSS_TERMS * c.NH;
    partial_block(sample_terms, partial, c.B, c.NH, sh_part);
    const bool xch = x.peer_bufs != nullptr && x.world > 1;
    if (xch && c.reduction == XSUP_REDUCE_BATCH) xchg_allreduce_block(partial, cols, x.peer_bufs, x.rank, x.world, x.step, x.seq, x.err);
    select_block(kps, target, sample_terms, partial, loss, sel, c);
    if (xch && c.reduction != XSUP_REDUCE_BATCH) {       // reporting only: the gradient needs just batch_total
        __syncthreads();
        xchg_allreduce_block(loss, 2, x.peer_bufs, x.rank, x.world, x.step, x.seq, x.err);
    }
    if (threadIdx.x == 0) *ticket = 0u;                  // re-armed for the next call on the same buffers
}

cudaError_t launch_reproj_fused_fwd(const float* kps, const float* target, const xsup_cam_t& cam, float* world, float* sample_terms,
                                    float* partial, float* loss, int64_t* sel, const xsup_loss_cfg_t& c, const xsup_xchg_t& x,
                                    unsigned int* ticket, cudaStream_t st) {
    const int warps = c.B * c.NH;
    reproj_fused_fwd_kernel<<<(warps + 7) / 8, 256, 0, st>>>(kps, target, cam, world, sample_terms, partial, loss, sel, c, x, ticket);
    return cudaGetLastError();
}

cudaError_t launch_reproj_select(const float* kps, const float* target, const float* sample_terms, const float* partial,
                                 float* loss, int64_t* sel, const xsup_loss_cfg_t& c, cudaStream_t st) {
    reproj_select_kernel<<<1, 256, 0, st>>>(kps, target, sample_terms, partial, loss, sel, c);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- loss backward
// d (gl0 * pseudo + gl1 * symmetry) / d kps[b,h,lane,:] (+ the vector-Jacobian product of an upstream gradient on
// kps_world, + an upstream gradient on kps itself) for one (sample, hypothesis); lane = joint.  Warp-uniform control flow.
__device__ __forceinline__ void loss_bwd_warp(const float* __restrict__ kps, const float* __restrict__ target, const xsup_cam_t& cam,
                                              const int64_t* __restrict__ sel, float gl0, float gl1,
                                              const float* __restrict__ g_kps_in, const float* __restrict__ g_world,
                                              const xsup_loss_cfg_t& c, int b, int h, int lane, float (&g)[3]) {
    const int K = c.K;
    const float n = (float)c.batch_total;
    bool on_sym = false;
    if (c.use_sym) {
        if (c.reduction == XSUP_REDUCE_BATCH) on_sym = (sel[1] == h);
        else if (c.reduction == XSUP_REDUCE_SAMPLE) on_sym = (sel[c.B + b] == h);
    }
    float x = 0.f, y = 0.f, z = 0.f;
    g[0] = g[1] = g[2] = 0.f;
    const size_t o = (((size_t)b * c.NH + h) * K + (lane < K ? lane : 0)) * 3;
    if (lane < K) {
        x = kps[o]; y = kps[o + 1]; z = kps[o + 2];
        bool on_mse;
        if (c.reduction == XSUP_REDUCE_BATCH) on_mse = (sel[0] == h);
        else if (c.reduction == XSUP_REDUCE_SAMPLE) on_mse = (sel[b] == h);
        else on_mse = (sel[(size_t)b * K + lane] == h);
        if (on_mse) {
            const float* tg = target + ((size_t)b * K + lane) * 3;
            const float s = gl0 * c.w_mse * 2.0f / (n * (float)K * 3.0f);
            g[0] = s * (x - tg[0]); g[1] = s * (y - tg[1]); g[2] = s * (z - tg[2]);
        }
        if (g_kps_in) { g[0] += g_kps_in[o]; g[1] += g_kps_in[o + 1]; g[2] += g_kps_in[o + 2]; }
    }
    if (!on_sym && !g_world) return;                                         // warp-uniform
    const Cam m = load_cam(cam, b);
    const Img im = make_img(c.img_h, c.img_w, c.rect_width);
    float w[3] = {0.f, 0.f, 0.f}, u = 0.f, v = 0.f, Z = 1.f;
    if (lane < K) patch_to_world(m, im, F_NORM | F_PATCH, x, y, z, w, u, v, Z);
    float gw[3] = {0.f, 0.f, 0.f};
    if (g_world && lane < K) { gw[0] = g_world[o]; gw[1] = g_world[o + 1]; gw[2] = g_world[o + 2]; }
    if (on_sym) {
        // ---- bones
        const int ci = c_bone_child[lane & 7], pi = c_bone_parent[lane & 7];
        const float vx = __shfl_sync(0xffffffffu, w[0], ci) - __shfl_sync(0xffffffffu, w[0], pi);
        const float vy = __shfl_sync(0xffffffffu, w[1], ci) - __shfl_sync(0xffffffffu, w[1], pi);
        const float vz = __shfl_sync(0xffffffffu, w[2], ci) - __shfl_sync(0xffffffffu, w[2], pi);
        const float len = sqrtf(vx * vx + vy * vy + vz * vz);
        const float nrm = len * 1e-3f;
        const float other = __shfl_xor_sync(0xffffffffu, nrm, 1);             // partner bone of the pair (2p, 2p+1)
        // d bone_sum / d n_i = +2(n_i - n_partner) for even i, and the same expression for odd i
        const float gn = gl1 * c.w_bone / (n * 4.0f) * 2.0f * (nrm - other) * 1e-3f;
        const float il = len > 0.f ? 1.0f / len : 0.f;                        // torch.norm's backward is 0 at 0
        const float bx = gn * vx * il, by = gn * vy * il, bz = gn * vz * il;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float fx_ = __shfl_sync(0xffffffffu, bx, i), fy_ = __shfl_sync(0xffffffffu, by, i), fz_ = __shfl_sync(0xffffffffu, bz, i);
            if (lane == c_bone_child[i]) { gw[0] += fx_; gw[1] += fy_; gw[2] += fz_; }
            if (lane == c_bone_parent[i]) { gw[0] -= fx_; gw[1] -= fy_; gw[2] -= fz_; }
        }
        // ---- midpoints (3-D on world, 2-D on the patch x,y)
        const int ia = c_mid_a[lane & 1], ib = c_mid_b[lane & 1], ir = (lane & 1) ? 0 : K - 1;
        float e[3], q[2];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            e[d] = (__shfl_sync(0xffffffffu, w[d], ia) + __shfl_sync(0xffffffffu, w[d], ib)) / 2.0f * 1e-3f - __shfl_sync(0xffffffffu, w[d], ir) * 1e-3f;
        q[0] = (__shfl_sync(0xffffffffu, x, ia) + __shfl_sync(0xffffffffu, x, ib)) / 2.0f - __shfl_sync(0xffffffffu, x, ir);
        q[1] = (__shfl_sync(0xffffffffu, y, ia) + __shfl_sync(0xffffffffu, y, ib)) / 2.0f - __shfl_sync(0xffffffffu, y, ir);
        const float s3 = gl1 * c.w_kp / (n * 6.0f) * 2.0f * 1e-3f;            // d/d(mid or ref), before the 1/2 of the midpoint
        const float s2 = gl1 * c.w_kp2d * 1e2f / (n * 4.0f) * 2.0f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int ja = c_mid_a[s], jb = c_mid_b[s], jr = s ? 0 : K - 1;
            float e0 = __shfl_sync(0xffffffffu, e[0], s), e1 = __shfl_sync(0xffffffffu, e[1], s), e2 = __shfl_sync(0xffffffffu, e[2], s);
            float q0 = __shfl_sync(0xffffffffu, q[0], s), q1 = __shfl_sync(0xffffffffu, q[1], s);
            if (lane == ja || lane == jb) {
                gw[0] += 0.5f * s3 * e0; gw[1] += 0.5f * s3 * e1; gw[2] += 0.5f * s3 * e2;
                g[0] += 0.5f * s2 * q0; g[1] += 0.5f * s2 * q1;
            }
            if (lane == jr) {
                gw[0] -= s3 * e0; gw[1] -= s3 * e1; gw[2] -= s3 * e2;
                g[0] -= s2 * q0; g[1] -= s2 * q1;
            }
        }
    }
    if (lane < K) {
        float gk[3];
        patch_to_world_vjp(m, im, F_NORM | F_PATCH, u, v, Z, gw, gk);
        g[0] += gk[0]; g[1] += gk[1]; g[2] += gk[2];
    }
}

__global__ void __launch_bounds__(128) reproj_loss_bwd_kernel(const float* __restrict__ kps, const float* __restrict__ target,
                                                              const xsup_cam_t cam, const int64_t* __restrict__ sel,
                                                              const float* __restrict__ g_loss, float* __restrict__ g_kps,
                                                              const xsup_loss_cfg_t c) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= c.B * c.NH) return;
    const int b = wid / c.NH, h = wid - b * c.NH;
    float g[3];
    loss_bwd_warp(kps, target, cam, sel, g_loss[0], g_loss[1], nullptr, nullptr, c, b, h, lane, g);
    if (lane < c.K) {
        const size_t o = (((size_t)b * c.NH + h) * c.K + lane) * 3;
        g_kps[o] = g[0]; g_kps[o + 1] = g[1]; g_kps[o + 2] = g[2];
    }
}

// ---------------------------------------------------------------------------------------------- fused backward
// One CTA per sample: (1) the loss VJP of every hypothesis of the sample (warp per hypothesis, lane = joint), summed
// with the upstream gradients on kps / kps_world, kept in shared memory; (2) the coefficient blocks of the sample's K
// units for the streaming head backward (what integral_coef_kernel derives from grad_kps in global memory).  Replaces
// loss_bwd + the ATen zeros/stack/add glue + integral_coef: the gradient w.r.t. kps never round-trips HBM.
// kLossBwdWarps warps per sample: the hypotheses' vector-Jacobian products on warps 0..NH-1, then one (b,k) unit per warp and round
// (see launch_reproj_fused_bwd for the two configurations).
template <int kLossBwdWarps, int MINB>
__global__ void __launch_bounds__(kLossBwdWarps * 32, MINB) reproj_fused_bwd_kernel(const float* __restrict__ kps, const float* __restrict__ target,
                                                               const xsup_cam_t cam, const int64_t* __restrict__ sel,
                                                               const float* __restrict__ g_lp, const float* __restrict__ g_ls,
                                                               const float* __restrict__ g_kps_in, const float* __restrict__ g_world,
                                                               float* __restrict__ g_kps_out, const xsup_loss_cfg_t c, const CoefParams p) {
    extern __shared__ float sm[];
    const int NH = c.NH, K = c.K;
    float* gz = sm;                       // [NH][32]
    float* gxp = sm + (size_t)NH * 32;    // [kLossBwdWarps][32] per-warp partial sums over this warp's hypotheses
    float* gyp = gxp + kLossBwdWarps * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
    if (b == 0 && threadIdx.x == 0) *p.counter = 0;                          // work-claim counter of the streaming kernel that follows
    const int D = p.D;
    // Second-phase inputs that do not depend on the first phase, fetched before it so that the two round trips to L2 overlap:
    // lane h holds (peak bin, window sum, window mean) of hypothesis h of this warp's first unit.  (Read inside the window test
    // of the loop over d they were 2 * NH dependent round trips per unit: the whole kernel is latency.)
    float t_idx = 0.f, t_sw = 1.f, t_zb = 0.f, pz0 = 0.f, pz1 = 0.f;
    if (warp < K) {
        const float* __restrict__ st0 = p.stats + (size_t)(b * K + warp) * p.stats_stride;
        if (lane < NH) {
            const float* sh = st0 + 4 + D + 3 * lane;
            t_idx = sh[0]; t_sw = sh[1]; t_zb = sh[2];
        }
        if (lane < D) pz0 = st0[4 + lane];                   // the depth marginal of the first two 32-bin chunks (all of it at D <= 64)
        if (32 + lane < D) pz1 = st0[4 + 32 + lane];
    }
    const float gl0 = g_lp ? *g_lp : 0.f, gl1 = g_ls ? *g_ls : 0.f;
    float ax = 0.f, ay = 0.f;
    for (int h = warp; h < NH; h += kLossBwdWarps) {
        float g[3];
        loss_bwd_warp(kps, target, cam, sel, gl0, gl1, g_kps_in, g_world, c, b, h, lane, g);
        gz[h * 32 + lane] = g[2];
        ax += g[0];
        ay += g[1];
        if (g_kps_out && lane < K) {
            const size_t o = (((size_t)b * NH + h) * K + lane) * 3;
            g_kps_out[o] = g[0]; g_kps_out[o + 1] = g[1]; g_kps_out[o + 2] = g[2];
        }
    }
    gxp[warp * 32 + lane] = ax;
    gyp[warp * 32 + lane] = ay;
    __syncthreads();
    const float zs = 2.0f / (float)D;
    const int half = p.NS >> 1;
    const int nh_reg = NH < 32 ? NH : 32;
    for (int k = warp; k < K; k += kLossBwdWarps) {
        const int unit = b * K + k;
        const float* __restrict__ st = p.stats + (size_t)unit * p.stats_stride;
        float* __restrict__ cf = p.coef + (size_t)unit * p.coef_stride;
        if (k != warp) {                                     // more joints than warps: this round's inputs
            if (lane < NH) {
                const float* sh = st + 4 + D + 3 * lane;
                t_idx = sh[0]; t_sw = sh[1]; t_zb = sh[2];
            }
            pz0 = lane < D ? st[4 + lane] : 0.f;
            pz1 = 32 + lane < D ? st[4 + 32 + lane] : 0.f;
        }
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int w = 0; w < kLossBwdWarps; ++w) { gx += gxp[w * 32 + k]; gy += gyp[w * 32 + k]; }
        const float a = gx * (2.0f / (float)p.H);            // x was normalised by H (…_multi.py:78)
        const float bb = gy * (2.0f / (float)p.W);           // y by W (…:79)
        float dot = 0.f;
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            const float pzd = d0 == 0 ? pz0 : d0 == 32 ? pz1 : d < D ? st[4 + d] : 0.f;
            float cd = 0.f;
            for (int h = 0; h < nh_reg; ++h) {               // same terms, same order as integral_coef_kernel
                const int idx = (int)__shfl_sync(0xffffffffu, t_idx, h);
                const float sw = __shfl_sync(0xffffffffu, t_sw, h), zb = __shfl_sync(0xffffffffu, t_zb, h);
                if (d >= idx - half && d <= idx + half) cd += gz[h * 32 + k] * zs * ((float)d - zb) / sw;
            }
            for (int h = 32; h < NH; ++h) {
                const float* sh = st + 4 + D + 3 * h;
                const int idx = (int)sh[0];
                if (d >= idx - half && d <= idx + half) cd += gz[h * 32 + k] * zs * ((float)d - sh[2]) / sh[1];
            }
            if (d < D) {
                cf[8 + d] = cd;
                dot = fmaf(cd, pzd, dot);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) {
            const float wc = rintf(st[1]), hc = rintf(st[2]);
            cf[0] = st[0];
            cf[1] = a;
            cf[2] = bb;
            cf[3] = -fmaf(a, st[1] - wc, fmaf(bb, st[2] - hc, dot));
            cf[4] = wc;
            cf[5] = hc;
            cf[6] = 0.f;
            cf[7] = 0.f;
        }
    }
}

cudaError_t launch_reproj_fused_bwd(const float* kps, const float* target, const xsup_cam_t& cam, const int64_t* sel, const float* g_lp,
                                    const float* g_ls, const float* g_kps_in, const float* g_world, float* g_kps_out,
                                    const xsup_loss_cfg_t& c, const CoefParams& p, cudaStream_t st) {
    auto smem_for = [&](int warps) { return ((size_t)c.NH * 32 + 2 * warps * 32) * sizeof(float); };    // <= 37 KB (NH <= 254)
    // The kernel is a chain of two round trips to L2 per CTA, so its duration is CTAs / (CTAs in flight) x that latency.  Up to two
    // samples per SM: 18 warps (every joint of the reference's skeletons in one round), 56 registers so that two CTAs fit on an SM
    // (with 87 it was one: B = 256 took two waves).  Beyond that: 9 warps, four CTAs per SM, the joints in two rounds.
    // Measured, 32^3 NH = 3, us at B = 64 / 256 / 1 024 / 4 096: 18 warps x 1 per SM 10.7 / 14.4 / 30.9 / 100; 18 x 2: 10.6 / 11.8 / 22.6 /
    // 65.6; 9 x 4: 12.3 / 12.4 / 18.3 / 45.0; 6 x 6: 13.4 / 14.2 / 18.7 / 41.3.
    if (c.B <= 2 * 148)
        reproj_fused_bwd_kernel<18, 2><<<c.B, 18 * 32, smem_for(18), st>>>(kps, target, cam, sel, g_lp, g_ls, g_kps_in, g_world, g_kps_out, c, p);
    else
        reproj_fused_bwd_kernel<9, 4><<<c.B, 9 * 32, smem_for(9), st>>>(kps, target, cam, sel, g_lp, g_ls, g_kps_in, g_world, g_kps_out, c, p);
    return cudaGetLastError();
}

cudaError_t launch_reproj_loss_bwd(const float* kps, const float* target, const xsup_cam_t& cam, const int64_t* sel,
                                   const float* g_loss, float* g_kps, const xsup_loss_cfg_t& c, cudaStream_t st) {
    const int warps = c.B * c.NH;
    reproj_loss_bwd_kernel<<<(warps + 3) / 4, 128, 0, st>>>(kps, target, cam, sel, g_loss, g_kps, c);
    return cudaGetLastError();
}

}  // namespace xsup
