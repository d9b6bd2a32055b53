// K6 — eval-side hypothesis selection + multi-view triangulation, and the discriminator-side glue.
//
// Replaces, in the reference,
//   eval.py:122-148 + eval_utils.py:7-41   ground-truth normalisation, left/right `switch_points` per hypothesis,
//                                          per-joint best-hypothesis argmin + gather, `per_act_mse`
//                                          (~40 tiny launches per camera)  ->  eval_select_kernel, 1 launch
//   modules/util.py:171-230                `triangulation` / `batch_triangulate`: patch -> image, P = K [R|T], the DLT
//                                          rows and a batched LAPACK/cuSOLVER SVD of [B,K,2V,4]
//                                          ->  triangulate_kernel: one thread per (sample, joint), one-sided Jacobi
//                                          SVD in fp64 registers/local memory (the matrices are 2V x 4)
//   modules/model.py:123-124               root-centring / 1000 of the world joints fed to the discriminator
//   modules/base_losses/loss_func.py:54-76 compute_disc_loss: per-sample min over hypotheses of (logit - target)^2
// Everything here is latency-sized (KBs of data); the point is launch count and a GPU-resident eval loop.
#include "xsup_internal.h"

namespace xsup {

// ---------------------------------------------------------------------------------------------- eval selection
__global__ void __launch_bounds__(128) eval_select_kernel(const EvalParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= p.B) return;
    const int K = p.K, NH = p.NH, k = lane;
    const bool on = k < K;
    const float s1 = p.img_size - 1.0f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (on) {
        const float* g = p.joints_px + ((size_t)b * K + k) * 3;
        if (p.img_size > 0.0f) {
            gx = fmaf(2.0f, __fdiv_rn(g[0], s1), -1.0f);           // eval.py:126-127
            gy = fmaf(2.0f, __fdiv_rn(g[1], s1), -1.0f);
            gz = __fdiv_rn(g[2], s1);
        } else {                                                   // img_size == 0: the ground truth is already normalised
            gx = g[0]; gy = g[1]; gz = g[2];
        }
        if (p.gt_norm) {
            float* o = p.gt_norm + ((size_t)b * K + k) * 3;
            o[0] = gx; o[1] = gy; o[2] = gz;
        }
    }
    float b3 = 0.f, b2 = 0.f, best3[3] = {0.f, 0.f, 0.f}, best2[2] = {0.f, 0.f};
    int i3 = 0, i2 = 0;
    bool last_trans = false;
    if (on) {
        const int ks = p.perm[k];
        for (int h = 0; h < NH; ++h) {
            const float* q = p.kps + (((size_t)b * NH + h) * K + k) * 3;
            const float* qs = p.kps + (((size_t)b * NH + h) * K + ks) * 3;
            // switch_points (eval_utils.py:17-27): L1 error in (x,y) of the swapped joint strictly smaller
            const float e = __fadd_rn(fabsf(q[0] - gx), fabsf(q[1] - gy));
            const float es = __fadd_rn(fabsf(qs[0] - gx), fabsf(qs[1] - gy));
            const bool tr = es < e;
            const float x = tr ? qs[0] : q[0], y = tr ? qs[1] : q[1], z = tr ? qs[2] : q[2];
            last_trans = tr;
            const float dx = x - gx, dy = y - gy, dz = z - gz;
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));       // eval.py:142: 2-D squared error
            const float d3 = __fadd_rn(d2, __fmul_rn(dz, dz));                      // eval.py:139: 3-D squared error
            const bool take3 = h == 0 || (p.best && d3 < b3);                       // argmin: first minimum wins
            const bool take2 = h == 0 || (p.best && d2 < b2);
            if (take3) { b3 = d3; i3 = h; best3[0] = x; best3[1] = y; best3[2] = z; }
            if (take2) { b2 = d2; i2 = h; best2[0] = x; best2[1] = y; }
        }
        const size_t o = (size_t)b * K + k;
        p.kp3d[o * 3] = best3[0]; p.kp3d[o * 3 + 1] = best3[1]; p.kp3d[o * 3 + 2] = best3[2];
        p.kp2d[o * 2] = best2[0]; p.kp2d[o * 2 + 1] = best2[1];
        if (p.is_trans) p.is_trans[o] = last_trans ? 1 : 0;
        if (p.best_idx) p.best_idx[o] = i3;
        if (p.best_2d_idx) p.best_2d_idx[o] = i2;
    }
    // per_act_mse (eval_utils.py:31-41) of the selected 2-D points
    float e = 0.f;
    if (on) {
        const float ex = (best2[0] + 1.0f) / 2.0f - (gx + 1.0f) / 2.0f, ey = (best2[1] + 1.0f) / 2.0f - (gy + 1.0f) / 2.0f;
        e = sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    }
    e = warp_sum(e);
    if (lane == 0 && p.err2d) p.err2d[b] = e / (float)K;
}

cudaError_t launch_eval_select(const EvalParams& p, cudaStream_t st) {
    eval_select_kernel<<<(p.B + 3) / 4, 128, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- triangulation
// One thread per (sample, joint).  A is [2V,4] in fp64; one-sided (Hestenes) Jacobi rotations orthogonalise its
// columns (A <- A J, W <- W J), which is accurate for the badly graded columns of a DLT matrix (the translation
// column is ~10^3 times larger) where forming A^T A would square the condition number.  The right singular
// vector of the smallest singular value is the column of W whose image has the smallest norm.
__global__ void __launch_bounds__(128) triangulate_kernel(const TriParams p, float* __restrict__ world) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.K) return;
    const int b = i / p.K;
    double A[2 * XSUP_MAX_VIEWS][4];
    const int V = p.V, R = 2 * V;
    const double wm1 = p.img_w - 1, hm1 = p.img_h - 1, ds = 1.0 / (double)p.img_w * (double)p.rect_width;
    for (int v = 0; v < V; ++v) {
        const xsup_cam_t& c = p.cam[v];
        const float* kp = p.kps[v] + (size_t)i * 3;
        double x = kp[0], y = kp[1], z = kp[2];
        if (p.is_norm) {                                          // util.py:70-72 (image_depth = image width, :183-184)
            x = (x + 1.0) / 2.0 * wm1;
            y = (y + 1.0) / 2.0 * hm1;
            z = z * wm1;
        }
        const float* T = c.trans_image + (size_t)b * 6;
        const double a00 = T[0], a01 = T[1], t0 = T[2], a10 = T[3], a11 = T[4], t1 = T[5];
        const double idet = 1.0 / (a00 * a11 - a01 * a10);
        const double du = x - t0, dv = y - t1;
        const double u = (a11 * du - a01 * dv) * idet, w = (-a10 * du + a00 * dv) * idet;   // util.py:64,76
        const double conf = z * ds + (double)c.pelvis[(size_t)b * 3 + 2];                   // util.py:79-80; the weight at :215-217
        const float* Km = c.k_mat + (size_t)b * 9;
        const float* Rm = c.rot_world + (size_t)b * 9;
        const float* Tw = c.trans_world + (size_t)b * 3;
        double P[3][4];                                           // K [R | T] (util.py:189)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s += (double)Km[r * 3 + j] * (cc < 3 ? (double)Rm[j * 3 + cc] : (double)Tw[j]);
                P[r][cc] = s;
            }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            A[v][cc] = conf * (u * P[2][cc] - P[0][cc]);          // util.py:215-216
            A[V + v][cc] = conf * (w * P[2][cc] - P[1][cc]);
        }
    }
    double W[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pc = 0; pc < 3; ++pc)
#pragma unroll
            for (int qc = pc + 1; qc < 4; ++qc) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = 0; r < R; ++r) {
                    al += A[r][pc] * A[r][pc];
                    be += A[r][qc] * A[r][qc];
                    ga += A[r][pc] * A[r][qc];
                }
                if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < R; ++r) {
                    const double ap = A[r][pc], aq = A[r][qc];
                    A[r][pc] = cs * ap - sn * aq;
                    A[r][qc] = sn * ap + cs * aq;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double wp = W[r][pc], wq = W[r][qc];
                    W[r][pc] = cs * wp - sn * wq;
                    W[r][qc] = sn * wp + cs * wq;
                }
            }
        if (!rotated) break;
    }
    int jmin = 0;
    double nmin = 0.0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += A[r][cc] * A[r][cc];
        if (cc == 0 || s < nmin) { nmin = s; jmin = cc; }
    }
    double X[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) X[r] = jmin == 0 ? W[r][0] : jmin == 1 ? W[r][1] : jmin == 2 ? W[r][2] : W[r][3];
    float* o = world + (size_t)i * 3;
    o[0] = (float)(X[0] / X[3]);                                  // util.py:221
    o[1] = (float)(X[1] / X[3]);
    o[2] = (float)(X[2] / X[3]);
}

cudaError_t launch_triangulate(const TriParams& p, float* world, cudaStream_t st) {
    triangulate_kernel<<<(p.B * p.K + 127) / 128, 128, 0, st>>>(p, world);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- root-centring (model.py:123-124)
// world [N, M, R] (R = 3 * joints-per-item): out[n, m, :] = (world[n, m, :] - world[n, 0, :]) / 1000, keeping the first
// `dim` of every coordinate triple.  For the reference's stacked [B, NH, K, 3] input N = B, M = NH, R = 3K (relative to
// HYPOTHESIS 0 - the literal meaning of `[:, [0], :]` there); for a [B, K, 3] input N = B, M = K, R = 3 (root joint).
__global__ void __launch_bounds__(128) root_centre_fwd_kernel(const float* __restrict__ world, float* __restrict__ out, int N, int M,
                                                              int R, int dim) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* w = world + (size_t)n * M * R;
    const int Ro = R / 3 * dim;
    for (int e = lane; e < M * R; e += 32) {
        const int m = e / R, r = e - m * R, c = r % 3;
        if (c < dim) out[((size_t)n * M + m) * Ro + (r / 3) * dim + c] = (w[e] - w[r]) / 1000.0f;
    }
}
__global__ void __launch_bounds__(128) root_centre_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_world, int N,
                                                              int M, int R, int dim) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const int Ro = R / 3 * dim;
    const float* g = g_out + (size_t)n * M * Ro;
    float* gw = g_world + (size_t)n * M * R;
    // item 0 receives minus the sum over all items (its own +1/1000 cancels); fixed order over m per lane
    for (int r = lane; r < R; r += 32) {
        const int c = r % 3, ro = (r / 3) * dim + c;
        float s = 0.f;
        if (c < dim)
            for (int m = 0; m < M; ++m) s += g[m * Ro + ro];
        for (int m = 0; m < M; ++m) gw[m * R + r] = c < dim ? (g[m * Ro + ro] - (m == 0 ? s : 0.f)) / 1000.0f : 0.f;
    }
}
cudaError_t launch_root_centre_fwd(const float* world, float* out, int N, int M, int R, int dim, cudaStream_t st) {
    root_centre_fwd_kernel<<<(N + 3) / 4, 128, 0, st>>>(world, out, N, M, R, dim);
    return cudaGetLastError();
}
cudaError_t launch_root_centre_bwd(const float* g_out, float* g_world, int N, int M, int R, int dim, cudaStream_t st) {
    root_centre_bwd_kernel<<<(N + 3) / 4, 128, 0, st>>>(g_out, g_world, N, M, R, dim);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- compute_disc_loss term
// logits [B,NH,C]: loss = mean over (b,c) of min_h (x - target)^2; sel[b,c] = argmin (first minimum).  One CTA:
// the batch is a few hundred scalars.
__global__ void __launch_bounds__(256) disc_min_loss_fwd_kernel(const float* __restrict__ logits, int B, int NH, int C, float target,
                                                                float* __restrict__ loss, int64_t* __restrict__ sel) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int r = threadIdx.x; r < B * C; r += 256) {
        const int b = r / C, c = r - b * C;
        float best = 0.f;
        int bi = 0;
        for (int h = 0; h < NH; ++h) {
            const float d = logits[((size_t)b * NH + h) * C + c] - target, e = d * d;
            if (h == 0 || e < best) { best = e; bi = h; }
        }
        sel[r] = bi;
        a += (double)best;
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(sh[0] / (double)(B * C));
}
__global__ void __launch_bounds__(256) disc_min_loss_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ sel,
                                                                const float* __restrict__ g_loss, int B, int NH, int C, float target,
                                                                float* __restrict__ g_logits) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * NH * C) return;
    const int c = i % C, h = (i / C) % NH, b = i / (C * NH);
    const float g = *g_loss * 2.0f / (float)(B * C);
    g_logits[i] = sel[b * C + c] == h ? g * (logits[i] - target) : 0.0f;
}
cudaError_t launch_disc_min_loss_fwd(const float* logits, int B, int NH, int C, float target, float* loss, int64_t* sel,
                                     cudaStream_t st) {
    disc_min_loss_fwd_kernel<<<1, 256, 0, st>>>(logits, B, NH, C, target, loss, sel);
    return cudaGetLastError();
}
cudaError_t launch_disc_min_loss_bwd(const float* logits, const int64_t* sel, const float* g_loss, int B, int NH, int C, float target,
                                     float* g_logits, cudaStream_t st) {
    disc_min_loss_bwd_kernel<<<(B * NH * C + 255) / 256, 256, 0, st>>>(logits, sel, g_loss, B, NH, C, target, g_logits);
    return cudaGetLastError();
}

}  // namespace xsup

// ---------------------------------------------------------------------------------------------- stand-alone pose loss terms
// compute_supervision / compute_bone_sym_loss / compute_kp_sym_loss (modules/base_losses/loss_func.py:18-52) on their own,
// for callers that use them outside the fused per-camera op (which evaluates the same terms per hypothesis in
// reproj_loss_fwd_kernel).  x [B,K,C]; one thread per sample forms the sample's un-normalised sum, one CTA adds them in
// fixed order.  The backward is the analytic VJP, one thread per sample.
namespace xsup {

__constant__ int c_pt_bone_child[8] = {16, 15, 13, 12, 3, 2, 6, 5};     // loss_func.py:20
__constant__ int c_pt_bone_parent[8] = {15, 14, 12, 11, 2, 1, 5, 4};
__constant__ int c_pt_mid_a[2] = {11, 1};                               // loss_func.py:28
__constant__ int c_pt_mid_b[2] = {14, 4};

__device__ __forceinline__ float pt_scaled(const PoseTermParams& p, float v, int c) {
    // compute_supervision's feature_shape branch (loss_func.py:39-45): x,y -> (v+1)/2*(fs-1), z -> v*(fs-1)
    if (!p.use_fs) return v;
    return c < 2 ? (v + 1.0f) / 2.0f * (p.fs[c] - 1.0f) : v * (p.fs[2] - 1.0f);
}

__global__ void __launch_bounds__(128) pose_term_fwd_kernel(const PoseTermParams p, float* __restrict__ sample_sums) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= p.B) return;
    const float* x = p.x + (size_t)b * p.K * p.C;
    float a = 0.f;
    if (p.term == XSUP_TERM_MSE) {
        const float* g = p.gt + (size_t)b * p.K * p.C;
        for (int i = 0; i < p.K * p.C; ++i) {
            const float d = pt_scaled(p, x[i], i % p.C) - g[i];
            a = fmaf(d, d, a);
        }
    } else if (p.term == XSUP_TERM_BONE) {
        float n[8];
        for (int i = 0; i < 8; ++i) {
            const float* c = x + c_pt_bone_child[i] * p.C;
            const float* q = x + c_pt_bone_parent[i] * p.C;
            float s = 0.f;
            for (int d = 0; d < p.C; ++d) s = fmaf(c[d] - q[d], c[d] - q[d], s);
            n[i] = sqrtf(s) * 1e-3f;
        }
        for (int i = 0; i < 8; i += 2) a = fmaf(n[i] - n[i + 1], n[i] - n[i + 1], a);
    } else {                                                           // XSUP_TERM_KP: is_3D scales by 1e-3
        const float sc = p.is_3d ? 1e-3f : 1.0f;
        for (int s = 0; s < 2; ++s) {
            const int ir = s ? 0 : p.K - 1;
            for (int d = 0; d < p.C; ++d) {
                const float mid = (x[c_pt_mid_a[s] * p.C + d] + x[c_pt_mid_b[s] * p.C + d]) / 2.0f;
                const float e = mid * sc - x[ir * p.C + d] * sc;
                a = fmaf(e, e, a);
            }
        }
    }
    sample_sums[b] = a;
}

__global__ void __launch_bounds__(256) pose_term_reduce_kernel(const float* __restrict__ sample_sums, int B, double denom, float* __restrict__ loss) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) a += (double)sample_sums[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(sh[0] / denom);
}

__global__ void __launch_bounds__(128) pose_term_bwd_kernel(const PoseTermParams p, const float* __restrict__ g_loss, float inv_denom,
                                                            float* __restrict__ g_x) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= p.B) return;
    const float* x = p.x + (size_t)b * p.K * p.C;
    float* g = g_x + (size_t)b * p.K * p.C;
    const float s0 = *g_loss * inv_denom;
    if (p.term == XSUP_TERM_MSE) {
        const float* t = p.gt + (size_t)b * p.K * p.C;
        for (int i = 0; i < p.K * p.C; ++i) {
            const int c = i % p.C;
            const float jac = !p.use_fs ? 1.0f : (c < 2 ? 0.5f * (p.fs[c] - 1.0f) : p.fs[2] - 1.0f);
            g[i] = s0 * 2.0f * (pt_scaled(p, x[i], c) - t[i]) * jac;
        }
        return;
    }
    for (int i = 0; i < p.K * p.C; ++i) g[i] = 0.f;
    if (p.term == XSUP_TERM_BONE) {
        float n[8], len[8];
        for (int i = 0; i < 8; ++i) {
            float s = 0.f;
            for (int d = 0; d < p.C; ++d) {
                const float v = x[c_pt_bone_child[i] * p.C + d] - x[c_pt_bone_parent[i] * p.C + d];
                s = fmaf(v, v, s);
            }
            len[i] = sqrtf(s);
            n[i] = len[i] * 1e-3f;
        }
        for (int i = 0; i < 8; ++i) {
            const float dn = s0 * 2.0f * (n[i] - n[i ^ 1]) * 1e-3f;   // d/d n_i of (n_2p - n_2p+1)^2, times d n / d |v|
            const float il = len[i] > 0.f ? 1.0f / len[i] : 0.f;     // torch.norm's backward is 0 at 0
            for (int d = 0; d < p.C; ++d) {
                const float v = x[c_pt_bone_child[i] * p.C + d] - x[c_pt_bone_parent[i] * p.C + d];
                g[c_pt_bone_child[i] * p.C + d] += dn * v * il;
                g[c_pt_bone_parent[i] * p.C + d] -= dn * v * il;
            }
        }
    } else {
        const float sc = p.is_3d ? 1e-3f : 1.0f;
        for (int s = 0; s < 2; ++s) {
            const int ia = c_pt_mid_a[s], ib = c_pt_mid_b[s], ir = s ? 0 : p.K - 1;
            for (int d = 0; d < p.C; ++d) {
                const float e = (x[ia * p.C + d] + x[ib * p.C + d]) / 2.0f * sc - x[ir * p.C + d] * sc;
                const float ge = s0 * 2.0f * e * sc;
                g[ia * p.C + d] += 0.5f * ge;
                g[ib * p.C + d] += 0.5f * ge;
                g[ir * p.C + d] -= ge;
            }
        }
    }
}

cudaError_t launch_pose_term_fwd(const PoseTermParams& p, double denom, float* sample_sums, float* loss, cudaStream_t st) {
    pose_term_fwd_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p, sample_sums);
    pose_term_reduce_kernel<<<1, 256, 0, st>>>(sample_sums, p.B, denom, loss);
    return cudaGetLastError();
}
cudaError_t launch_pose_term_bwd(const PoseTermParams& p, double denom, const float* g_loss, float* g_x, cudaStream_t st) {
    pose_term_bwd_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p, g_loss, (float)(1.0 / denom), g_x);
    return cudaGetLastError();
}

// compute_supervision(mode='none') (loss_func.py:46-47, nn.MSELoss(reduction='none')): the element-wise squared error
// tensor [B,K,C], and its VJP.
__global__ void __launch_bounds__(256) pose_sqerr_kernel(const PoseTermParams p, const float* __restrict__ g_out, float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= p.B * p.K * p.C) return;
    const int c = i % p.C;
    const float d = pt_scaled(p, p.x[i], c) - p.gt[i];
    if (!g_out) { out[i] = d * d; return; }
    const float jac = !p.use_fs ? 1.0f : (c < 2 ? 0.5f * (p.fs[c] - 1.0f) : p.fs[2] - 1.0f);
    out[i] = g_out[i] * 2.0f * d * jac;
}
cudaError_t launch_pose_sqerr(const PoseTermParams& p, const float* g_out, float* out, cudaStream_t st) {
    const int n = p.B * p.K * p.C;
    pose_sqerr_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, g_out, out);
    return cudaGetLastError();
}

}  // namespace xsup
