// K3 — backward of the integral multi-hypothesis head as ONE pass over the volume.
//
// The reference back-propagates through softmax + three expanded marginal sums + gathers
// (autograd of keypoint_detector_integral_multi.py:69-88): ~4 reads and ~3 writes of the volume.
// In closed form the gradient of every logit is
//     dL/dl[d,h,w] = p[d,h,w] * (a*w + b*h + c[d] - gbar),   p = 2^(l*log2e - lse2)
// evaluated about the integer bin (wc, hc) nearest the expectation, a*(w-wc) + b*(h-hc) + (c[d] + base0),
// so that no term of size a*W cancels against gbar in fp32.  (lse2, a, b, base0, wc, hc, c[0..D)) per
// (b,k) unit is an (8+D)-float coefficient block that the tiny
// `integral_coef_kernel` derives from grad_kps and the statistics saved by the forward.  The
// streaming kernel then reads each logit once and writes each gradient once.
//
// Streaming kernel structure: persistent CTA per SM; warp 16 = TMA producer (bulk copy of the
// stage and of its unit's coefficient block into the same ring slot), warps 0..15 = stateless
// consumers (LDS.128 -> registers -> one MUFU.EX2 + 3 FP ops per element -> coalesced STG.128).
#include "xsup_internal.h"

namespace xsup {

// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) integral_coef_kernel(const CoefParams p) {
    // one warp per (b,k) unit
    const int unit = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.counter = 0;       // work-claim counter of the streaming kernel that follows
    if (unit >= p.n_units) return;
    const int b = unit / p.K, k = unit - b * p.K;
    const int D = p.D, NH = p.NH;
    const float* st = p.stats + (size_t)unit * p.stats_stride;
    float* cf = p.coef + (size_t)unit * p.coef_stride;
    float gx = 0.f, gy = 0.f;
    for (int h = 0; h < NH; ++h) {
        const float* g = p.g_kps + (((size_t)b * NH + h) * p.K + k) * 3;
        gx += g[0];
        gy += g[1];
    }
    const float a = gx * (2.0f / (float)p.H);                // x was normalised by H (…_multi.py:78)
    const float bb = gy * (2.0f / (float)p.W);               // y by W (…:79)
    const float zs = 2.0f / (float)D;
    const int half = p.NS >> 1;
    const float zc = p.head == XSUP_HEAD_SINGLE ? rintf(st[4 + D]) : 0.f;   // single head: centre d as well
    // lane h: (peak bin, window sum, window mean, upstream z gradient) of hypothesis h, one round trip instead of a dependent one per
    // hypothesis inside the window test below
    float t_idx = 0.f, t_sw = 1.f, t_zb = 0.f, t_gz = 0.f;
    if (p.head != XSUP_HEAD_SINGLE && lane < NH) {
        const float* sh = st + 4 + D + 3 * lane;
        t_idx = sh[0]; t_sw = sh[1]; t_zb = sh[2];
        t_gz = p.g_kps[(((size_t)b * NH + lane) * p.K + k) * 3 + 2];
    }
    const int nh_reg = NH < 32 ? NH : 32;
    float dot = 0.f;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        float c = 0.f;
        if (p.head == XSUP_HEAD_SINGLE) {
            c = p.g_kps[((size_t)b * p.K + k) * 3 + 2] * zs * ((float)d - zc);
        } else {
            for (int h = 0; h < nh_reg; ++h) {
                const int idx = (int)__shfl_sync(0xffffffffu, t_idx, h);
                const float sw = __shfl_sync(0xffffffffu, t_sw, h), zb = __shfl_sync(0xffffffffu, t_zb, h);
                const float gz = __shfl_sync(0xffffffffu, t_gz, h);
                if (d >= idx - half && d <= idx + half) c += gz * zs * ((float)d - zb) / sw;
            }
            for (int h = 32; h < NH; ++h) {
                const float* sh = st + 4 + D + 3 * h;
                const int idx = (int)sh[0];
                if (d >= idx - half && d <= idx + half) {
                    const float gz = p.g_kps[(((size_t)b * NH + h) * p.K + k) * 3 + 2];
                    c += gz * zs * ((float)d - sh[2]) / sh[1];
                }
            }
        }
        if (d < D) {
            cf[8 + d] = c;
            dot = fmaf(c, st[4 + d], dot);
        }
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        const float wc = rintf(st[1]), hc = rintf(st[2]);
        cf[0] = st[0];
        cf[1] = a;
        cf[2] = bb;
        // gbar = a*xbar + b*ybar + sum_d c[d]*pz[d];  base0 = a*wc + b*hc - gbar, formed from small differences
        cf[3] = -fmaf(a, st[1] - wc, fmaf(bb, st[2] - hc, dot));
        cf[4] = wc;
        cf[5] = hc;
        cf[6] = 0.f;
        cf[7] = 0.f;
    }
}

cudaError_t launch_integral_coef(const CoefParams& p, cudaStream_t st) {
    integral_coef_kernel<<<(p.n_units + 3) / 4, 128, 0, st>>>(p);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------
template <typename T, int U>
__global__ void __launch_bounds__(kBwdThreads, 1) integral_bwd_kernel(const BwdParams p) {
    constexpr int VEC = Vec<T>::N;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Tiling& t = p.t;
    const int nst = p.nst, TU = t.tasks_per_unit, SPU = t.stages_per_unit;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nst * p.slot_bytes);
    volatile int2* hdr = reinterpret_cast<volatile int2*>(bars + 2 * kMaxStages);       // [nst] (unit, stage in unit)
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * nst;
    const uint32_t ring0 = smem_u32(smem);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) {
            mbar_init(full0 + 8u * i, 1);
            mbar_init(empty0 + 8u * i, kTasksPerStage);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t coef_bytes = (uint32_t)p.coef_stride * 4u;

    if (warp == kConsumerWarps) {
        // ------------------------------------------------------------------ producer + scheduler
        // the consumers are stateless, so work is claimed in chunks of `p.chunk` ring stages of the
        // global stage sequence (unit-major); chunks may straddle units
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const int total = p.n_units * SPU;
            const int n_chunks = (total + p.chunk - 1) / p.chunk;
            int slot = 0;
            uint32_t eph = 1;                                        // parity to wait for on `empty`: the first pass over the ring does not wait
            int cur = atomicAdd(p.counter, 1);
            while (cur < n_chunks) {
                const int nxt = atomicAdd(p.counter, 1);            // claim ahead
                const int gs0 = cur * p.chunk, gs1 = min(gs0 + p.chunk, total);
                int unit = gs0 / SPU, j = gs0 - unit * SPU;
                for (int gs = gs0; gs < gs1; ++gs) {
                    mbar_wait(empty0 + 8u * slot, eph);
                    hdr[slot].x = unit;
                    hdr[slot].y = j;
                    const uint8_t* src = static_cast<const uint8_t*>(p.logits) + (size_t)unit * (size_t)t.unit_bytes;
                    const long long off = (long long)j * t.stage_bytes;
                    const uint32_t bytes = (uint32_t)min((long long)t.stage_bytes, t.unit_bytes - off);
                    const uint32_t dst = ring0 + (uint32_t)slot * p.slot_bytes;
                    mbar_arrive_expect_tx(full0 + 8u * slot, bytes + coef_bytes);
                    bulk_g2s_hint(dst, src + off, bytes, full0 + 8u * slot, pol);
                    bulk_g2s(dst + t.stage_bytes, p.coef + (size_t)unit * p.coef_stride, coef_bytes, full0 + 8u * slot);
                    if (++j == SPU) { j = 0; ++unit; }
                    if (++slot == nst) { slot = 0; eph ^= 1; }
                }
                cur = nxt;
            }
            for (int g = 0; g < kGroups; ++g) {                     // end-of-stream sentinel per consumer group
                mbar_wait(empty0 + 8u * slot, eph);
                hdr[slot].x = -1;
                hdr[slot].y = 0;
                mbar_arrive(full0 + 8u * slot);
                if (++slot == nst) { slot = 0; eph ^= 1; }
            }
        }
    } else {
        const int g = warp / kTasksPerStage, q = warp % kTasksPerStage;
        const int lr = lane >> t.lpr_log2;
        const int w0 = (lane & (t.lpr - 1)) * VEC;
        const float rpi = (float)(32 >> t.lpr_log2);
        const uint32_t hdr0 = smem_u32(const_cast<int2*>(hdr));
        int slot = g;                                                // this group's slots: g, g + kGroups, ... (nst is a multiple of kGroups)
        uint32_t fph = 0;
        for (;; slot += kGroups) {
            if (slot >= nst) { slot -= nst; fph ^= 1; }
            mbar_wait(full0 + 8u * slot, fph);
            const int2 hd = lds_int2(hdr0 + 8u * slot);
            const int unit = hd.x, j = hd.y;
            if (unit < 0) break;
            for (int r = 0; r < t.rounds; ++r) {
            const int task = (j * t.rounds + r) * kTasksPerStage + q;
            const bool last_round = r == t.rounds - 1;
            if (task < TU) {
                const uint32_t sbase = ring0 + (uint32_t)slot * p.slot_bytes;
                const uint32_t addr = sbase + (uint32_t)(r * kTasksPerStage + q) * t.task_bytes + lane * 16u;
                int d, part;
                if (t.parts_log2 >= 0) { d = task >> t.parts_log2; part = task & (t.parts - 1); }
                else { d = task / t.parts; part = task - d * t.parts; }
                uint4 raw[U];
#pragma unroll
                for (int i = 0; i < U; ++i) raw[i] = lds128(addr + i * 512u);
                const uint4 c4 = lds128(sbase + t.stage_bytes);
                const uint4 c5 = lds128(sbase + t.stage_bytes + 16u);
                float cd;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cd) : "r"(sbase + t.stage_bytes + 32u + 4u * d));
                if (last_round) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8u * slot);
                }

                constexpr int P = Vec<T>::P;
                const float nlse = -__uint_as_float(c4.x), a = __uint_as_float(c4.y), bb = __uint_as_float(c4.z);
                const float base = cd + __uint_as_float(c4.w);
                const float wrel = (float)w0 - __uint_as_float(c5.x);
                f32x2 aw[P];                                       // a*(w - wc) for this lane's columns, as pairs
#pragma unroll
                for (int v = 0; v < P; ++v) aw[v] = pk2(a * (wrel + (float)(2 * v)), a * (wrel + (float)(2 * v + 1)));
                const f32x2 l2e2 = pk2(kLog2e, kLog2e), nlse2 = pk2(nlse, nlse);
                float hf = (float)(part * t.rows_per_task + lr) - __uint_as_float(c5.y);
                uint8_t* out = static_cast<uint8_t*>(p.g_logits) + (size_t)unit * (size_t)t.unit_bytes + (size_t)task * t.task_bytes + lane * 16u;
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    f32x2 x[P];
                    Vec<T>::unpack2(raw[i], x);
                    const float rowv = fmaf(bb, hf, base);
                    const f32x2 rowv2 = pk2(rowv, rowv);
#pragma unroll
                    for (int v = 0; v < P; ++v) x[v] = fmul2(ex2_2(ffma2(x[v], l2e2, nlse2)), fadd2(aw[v], rowv2));
                    *reinterpret_cast<uint4*>(out + i * 512u) = Vec<T>::pack2(x);
                    hf += rpi;
                }
            } else if (last_round) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * slot);
            }
            }
        }
    }
}

// Generic path (any shape): one CTA per unit, scalar accesses.
template <typename T>
__device__ __forceinline__ void store_elem(T* p, size_t i, float v);
template <>
__device__ __forceinline__ void store_elem<float>(float* p, size_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void store_elem<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }
template <typename T>
__device__ __forceinline__ float load_elem_b(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_elem_b<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_elem_b<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

template <typename T>
__global__ void __launch_bounds__(256) integral_bwd_generic_kernel(const BwdParams p) {
    const int D = p.t.D, H = p.t.H, W = p.t.W, HW = H * W;
    const size_t unit = blockIdx.x;
    const T* src = static_cast<const T*>(p.logits) + unit * D * HW;
    T* dst = static_cast<T*>(p.g_logits) + unit * D * HW;
    const float* cf = p.coef + unit * (size_t)p.coef_stride;
    const float nlse = -cf[0], a = cf[1], bb = cf[2], base0 = cf[3], wc = cf[4], hc = cf[5];
    for (int i = threadIdx.x; i < D * HW; i += blockDim.x) {
        const int d = i / HW, r = i - d * HW, h = r / W, w = r - h * W;
        const float pr = ex2(fmaf(load_elem_b(src, i), kLog2e, nlse));
        store_elem(dst, i, pr * (fmaf(a, (float)w - wc, fmaf(bb, (float)h - hc, cf[8 + d] + base0))));
    }
}

template <typename T, int U>
static cudaError_t launch_fast(const BwdParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = integral_bwd_kernel<T, U>;
    static unsigned long long attr_done = 0;             // per instantiation; one bit per device
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    kern<<<grid, kBwdThreads, smem, st>>>(p);
    return cudaGetLastError();
}
template <typename T>
static cudaError_t launch_fast_u(const BwdParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (p.t.U) {
        case 8: return launch_fast<T, 8>(p, grid, smem, st);
        case 4: return launch_fast<T, 4>(p, grid, smem, st);
        case 2: return launch_fast<T, 2>(p, grid, smem, st);
        default: return launch_fast<T, 1>(p, grid, smem, st);
    }
}

cudaError_t launch_integral_bwd(BwdParams p, bool fast, int dtype, int num_sms, cudaStream_t st) {
    if (!fast) {
        if (dtype == XSUP_F32) integral_bwd_generic_kernel<float><<<p.n_units, 256, 0, st>>>(p);
        else integral_bwd_generic_kernel<__nv_bfloat16><<<p.n_units, 256, 0, st>>>(p);
        return cudaGetLastError();
    }
    const int coef_pad = ((p.coef_stride * 4 + 127) / 128) * 128;
    p.slot_bytes = p.t.stage_bytes + coef_pad;
    const size_t fixed = (size_t)(2 * kMaxStages) * 8 + (size_t)kMaxStages * sizeof(int2);
    // Ring stages per work claim.  Measured at B=256, 64^3 fp32 (ms per launch): 16 -> 1.360, 8 -> 1.347, 4 -> 1.338, 2 -> 1.327,
    // 1 -> 1.377 (bf16: 16 -> 0.692, 2 -> 0.673): fine-grained claims keep the 148 SMs streaming through neighbouring
    // addresses and shorten the tail; below two stages the claim round trip is no longer hidden behind the copies.
    p.chunk = 2;
    int nst = (int)((kSmemBudget - fixed) / p.slot_bytes);
    nst = nst > kMaxStages ? kMaxStages : nst;
    nst = nst / kGroups * kGroups;        // one warp group per slot, see launch_integral_fwd
    if (nst < kGroups) return cudaErrorInvalidConfiguration;
    p.nst = nst;
    const size_t smem = (size_t)nst * p.slot_bytes + fixed;
    const int grid = p.n_units < num_sms ? p.n_units : num_sms;
    return dtype == XSUP_F32 ? launch_fast_u<float>(p, grid, smem, st) : launch_fast_u<__nv_bfloat16>(p, grid, smem, st);
}

}  // namespace xsup
