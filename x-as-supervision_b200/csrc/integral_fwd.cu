// K1 — fused integral (soft-argmax) multi-hypothesis head, forward.
//
// Replaces keypoint_detector_integral_multi.py:69-88 / keypoint_detector_integral.py:45-65 of the
// reference: softmax over each joint's D*H*W volume, the three marginals, x/y expectations, depth
// local-maxima + top-NH, windowed depth expectation, normalisation.  The reference makes ~6 passes
// over the volume and materialises the probabilities; this kernel reads every logit from HBM
// exactly once and writes O(D) floats per (b,k) unit.
//
// Structure (one persistent CTA per SM, 19 warps):
//   warp 16      producer : 1-D bulk async copies (TMA, UBLKCP) global -> smem ring, mbarrier tx-count
//   warps 0..15  consumers: 4 warps per ring stage, one 512*U-byte "task" each; LDS.128 into
//                registers, early release of the slot, warp-uniform running max (CREDUX.MAX.F32),
//                one MUFU.EX2 per element, per-lane column accumulators (x), row-weighted sum (y),
//                per-task depth-slice sum (pz)
//   warps 17..   finalisers (2, or 3 for units of at most 256 KB; unit i of the CTA goes to finaliser i mod n): log-sum-exp
//                combine of the per-task/per-warp partials, peaks, top-NH, window depth, outputs + saved-for-backward
//                stats; overlaps the stream of the next n-1 units
#ifdef XSUP_TRACE
#include <stdlib.h>
#endif

#include "xsup_internal.h"
#include "xsup_finalise.cuh"

namespace xsup {

// Diagnostics, compiled in only with -DXSUP_TRACE: clock64 timeline of CTA 0 (units 4..19 of that CTA) into the device buffer whose
// address is in XSUP_K1_TRACE; [unit - 4][16] slots: 0 producer claims, 1 first stage issued, 2 last stage issued, 4 consumer warp 0
// sees the first stage, 5 consumer warp 0 flushes, 8 finaliser starts, 9 partials merged, 10 unit finalised.
#ifdef XSUP_TRACE
__device__ long long* g_k1_trace = nullptr;
#define K1TRACE(slot, ui) do { if (g_k1_trace && blockIdx.x == 0 && (ui) >= 4 && (ui) < 20) g_k1_trace[((ui) - 4) * 16 + (slot)] = clock64(); } while (0)
#else
#define K1TRACE(slot, ui) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------------
// Fast path: TMA-fed ring, one pass over the volume.
// ----------------------------------------------------------------------------------------------
template <typename T, int U>
__global__ void __launch_bounds__(kFwdThreads, 1) integral_fwd_kernel(const FwdParams p) {
    constexpr int VEC = Vec<T>::N;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Tiling& t = p.t;
    const int nst = p.nst, TU = t.tasks_per_unit, SPU = t.stages_per_unit;
    // every unit occupies a multiple of kGroups ring stages (the tail ones carry no data), so that every
    // consumer group sees every unit and group g always handles the stages j = g (mod kGroups)
    const int SPUP = (SPU + kGroups - 1) / kGroups * kGroups;

    uint8_t* ring = smem;
    const int NF = p.nfin;                                                              // finaliser warps = partial buffers
    const int DP = (t.D + 1) & ~1;                                                      // keeps the tables 8-byte aligned
    float4* lane_part = reinterpret_cast<float4*>(smem + (size_t)nst * t.stage_bytes);  // [NF][kConsumerWarps][32] per-lane (sx, sa, sy, sr)
    float2* pz_table = reinterpret_cast<float2*>(lane_part + NF * kConsumerWarps * 32); // [NF][TU] (m, sum)
    float2* warp_hdr = pz_table + NF * TU;                                              // [NF][kConsumerWarps] (m_ref, unit)
    float* pz_final = reinterpret_cast<float*>(warp_hdr + NF * kConsumerWarps);         // [NF][DP]
    int* peak_bins = reinterpret_cast<int*>(pz_final + NF * DP);                        // [NF][DP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(peak_bins + NF * DP);
    volatile int2* hdr = reinterpret_cast<volatile int2*>(bars + 2 * kMaxStages + 2 * kMaxFinalisers);   // [nst] (unit, stage in unit)
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * nst;
    const uint32_t pfull0 = empty0 + 8u * nst, pempty0 = pfull0 + 8u * kMaxFinalisers;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) {
            mbar_init(full0 + 8u * i, 1);
            mbar_init(empty0 + 8u * i, kTasksPerStage);
        }
        for (int i = 0; i < NF; ++i) {
            mbar_init(pfull0 + 8u * i, kConsumerWarps);
            mbar_init(pempty0 + 8u * i, 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ------------------------------------------------------------------ producer + scheduler
        // SMs see different HBM bandwidth (die / L2 distance), so units are claimed from a global counter
        // instead of being dealt round-robin; the claim is published to the consumers in the slot header.
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t ring0 = smem_u32(ring);
            int slot = 0;
            uint32_t eph = 1;                                        // parity to wait for on `empty`: the first pass over the ring does not wait
            int cur = atomicAdd(p.counter, 1);
            int lu = 0;                                              // units this CTA has started (trace only)
            while (cur < p.n_units) {
                K1TRACE(0, lu);
                const int nxt = atomicAdd(p.counter, 1);            // claim ahead: the round trip overlaps this unit's copies
                const uint8_t* src = static_cast<const uint8_t*>(p.logits) + (size_t)cur * (size_t)t.unit_bytes;
                for (int j = 0; j < SPUP; ++j) {
                    mbar_wait(empty0 + 8u * slot, eph);
                    hdr[slot].x = cur;
                    hdr[slot].y = j;
                    if (j < SPU) {
                        const long long off = (long long)j * t.stage_bytes;
                        const uint32_t bytes = (uint32_t)min((long long)t.stage_bytes, t.unit_bytes - off);
                        mbar_arrive_expect_tx(full0 + 8u * slot, bytes);
                        bulk_g2s_hint(ring0 + (uint32_t)slot * t.stage_bytes, src + off, bytes, full0 + 8u * slot, pol);
                    } else {
                        mbar_arrive(full0 + 8u * slot);             // padding stage: header only
                    }
                    if (j == 0) K1TRACE(1, lu);
                    if (j == SPUP - 1) K1TRACE(2, lu);
                    if (++slot == nst) { slot = 0; eph ^= 1; }
                }
                cur = nxt;
                ++lu;
            }
            for (int g = 0; g < kGroups; ++g) {                     // one end-of-stream sentinel per consumer group
                mbar_wait(empty0 + 8u * slot, eph);
                hdr[slot].x = -1;
                hdr[slot].y = 0;
                mbar_arrive(full0 + 8u * slot);
                if (++slot == nst) { slot = 0; eph ^= 1; }
            }
        }
    } else if (warp > kConsumerWarps) {
        // ------------------------------------------------------------------ finalisers
        // NF warps, one per partial buffer: warp f takes this CTA's units f, f + NF, ..., so a unit's epilogue may last NF
        // unit-streaming times before it stalls the ring (64 KB bf16 units stream in 1.5 us, their epilogue takes 4 us on a
        // warp that shares its scheduler with four issue-bound consumers: with two buffers the consumers waited for `pempty`)
        const int buf = warp - kConsumerWarps - 1;
        if (buf >= NF) return;
        float* pz_mine = pz_final + buf * DP;
        int* bins_mine = peak_bins + buf * DP;
        uint32_t fpar = 0;                                           // this buffer's use number, mod 2
        for (int it = buf;; it += NF, fpar ^= 1) {
            mbar_wait(pfull0 + 8u * buf, fpar);
            if (lane == 0) K1TRACE(8, it);
            // the consumers leave their per-LANE partial sums (no shuffles on their side: they are the issue-bound warps, this
            // one has two unit-times per unit); merge the 16 warps with their log-sum-exp weights, then one reduction per quantity
            const float2 hd = lane < kConsumerWarps ? warp_hdr[buf * kConsumerWarps + lane] : make_float2(kNegHuge, 0.f);
            const int unit = __shfl_sync(0xffffffffu, __float_as_int(hd.y), 0);
            if (unit < 0) break;                                     // consumers reached the sentinel
            const float M = warp_max(hd.x);
            const float wsc = ex2(hd.x - M);
            // merged in fp64: 2 048 products per unit on a warp with time to spare, and the expectations keep all the bits the
            // per-lane fp32 partial sums carry (the gradient of a 128^3 unit is sensitive to the last ones)
            double ax = 0.0, as = 0.0, ay = 0.0, ar = 0.0;
#pragma unroll
            for (int w = 0; w < kConsumerWarps; ++w) {
                const double sc = (double)__shfl_sync(0xffffffffu, wsc, w);
                const float4 v = lane_part[(buf * kConsumerWarps + w) * 32 + lane];
                ax = fma((double)v.x, sc, ax);
                as = fma((double)v.y, sc, as);
                ay = fma((double)v.z, sc, ay);
                ar = fma((double)v.w, sc, ar);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                as += __shfl_xor_sync(0xffffffffu, as, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
                ar += __shfl_xor_sync(0xffffffffu, ar, o);
            }
            // (w-weighted sum) / (sum through the same column accumulators), likewise for rows
            const float xbar = (float)(ax / as);
            const float ybar = (float)(ay / ar);
            const float2* tab = pz_table + buf * TU;
            for (int d = lane; d < t.D; d += 32) {
                float a = 0.f;
                for (int q = 0; q < t.parts; ++q) {
                    const float2 e = tab[d * t.parts + q];
                    a = fmaf(e.y, ex2(e.x - M), a);
                }
                pz_mine[d] = a;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pempty0 + 8u * buf);      // partial buffers may be refilled
            if (lane == 0) K1TRACE(9, it);
            finalise_unit(p, unit, pz_mine, bins_mine, M, xbar, ybar, lane);
            __syncwarp();
            if (lane == 0) K1TRACE(10, it);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int g = warp / kTasksPerStage, q = warp % kTasksPerStage;
        const int lr = lane >> t.lpr_log2;
        const int w0 = (lane & (t.lpr - 1)) * VEC;
        const float rpi = (float)(32 >> t.lpr_log2);
        const uint32_t ring0 = smem_u32(ring);
        int it = 0;                                                  // units this warp has flushed
        int buf = 0;                                                 // it mod NF
        uint32_t bpar = 1;                                           // parity of the previous use of partial buffer `buf` (no wait in the first round)
        bool fresh = true;                                           // first stage of a unit: claim the partial buffer
        constexpr int P = Vec<T>::P;
        float m_ref = kNegHuge, sy = 0.f, sr = 0.f;
        f32x2 acc[P];                                                // column accumulators, as fp32 pairs
#pragma unroll
        for (int v = 0; v < P; ++v) acc[v] = pk2(0.f, 0.f);
        const f32x2 l2e2 = pk2(kLog2e, kLog2e);

        const uint32_t hdr0 = smem_u32(const_cast<int2*>(hdr));
        int slot = g;                                                // this group's slots: g, g + kGroups, ... (nst is a multiple of kGroups)
        uint32_t fph = 0;
        for (;; slot += kGroups) {
            if (slot >= nst) { slot -= nst; fph ^= 1; }
            mbar_wait(full0 + 8u * slot, fph);
            const int2 hd = lds_int2(hdr0 + 8u * slot);
            const int unit = hd.x, j = hd.y;
            if (unit < 0) break;
            if (fresh && warp == 0 && lane == 0) K1TRACE(4, it);
            if (fresh) {
                if (it >= NF) mbar_wait(pempty0 + 8u * buf, bpar);
                fresh = false;
            }
            for (int r = 0; r < t.rounds; ++r) {
            const int task = (j * t.rounds + r) * kTasksPerStage + q;
            const bool last_round = r == t.rounds - 1;
            if (task < TU) {
                const uint32_t addr = ring0 + (uint32_t)slot * t.stage_bytes + (uint32_t)(r * kTasksPerStage + q) * t.task_bytes + lane * 16u;
                uint4 raw[U];
#pragma unroll
                for (int i = 0; i < U; ++i) raw[i] = lds128(addr + i * 512u);
                if (last_round) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8u * slot);   // data is in registers: free the slot early
                }
                const float mt = warp_max(Vec<T>::template vmax<U>(raw)) * kLog2e;
                if (mt > m_ref) {                                 // warp-uniform, rare after the first tasks
                    const float sc = ex2(m_ref - mt);
                    const f32x2 sc2 = pk2(sc, sc);
#pragma unroll
                    for (int v = 0; v < P; ++v) acc[v] = fmul2(acc[v], sc2);
                    sy *= sc;
                    sr *= sc;
                    m_ref = mt;
                }
                int d, part;
                if (t.parts_log2 >= 0) { d = task >> t.parts_log2; part = task & (t.parts - 1); }
                else { d = task / t.parts; part = task - d * t.parts; }
                float hf = (float)(part * t.rows_per_task + lr);
                float tsum = 0.f;
                const f32x2 nm2 = pk2(-m_ref, -m_ref);
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    f32x2 x[P];
                    Vec<T>::unpack2(raw[i], x);
#pragma unroll
                    for (int v = 0; v < P; ++v) {
                        x[v] = ex2_2(ffma2(x[v], l2e2, nm2));     // e = 2^(l*log2e - m_ref)
                        acc[v] = fadd2(acc[v], x[v]);
                    }
                    f32x2 rs = fadd2(x[0], x[1]);
#pragma unroll
                    for (int v = 2; v < P; ++v) rs = fadd2(rs, x[v]);
                    float r0, r1;
                    upk2(rs, r0, r1);
                    const float r = r0 + r1;                      // sum of this row fragment
                    tsum += r;
                    sy = fmaf(hf, r, sy);
                    hf += rpi;
                }
                sr += tsum;
                tsum = warp_sum(tsum);
                if (lane == 0) pz_table[buf * TU + task] = make_float2(m_ref, tsum);
            } else if (last_round) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * slot);
            }
            }
            if (j + kGroups >= SPUP) {
                // this warp's last stage of the unit: hand its partials to the finaliser
                float sx = 0.f, sa = 0.f;
#pragma unroll
                for (int v = 0; v < P; ++v) {
                    float a0, a1;
                    upk2(acc[v], a0, a1);
                    sx = fmaf((float)(w0 + 2 * v), a0, sx);
                    sx = fmaf((float)(w0 + 2 * v + 1), a1, sx);
                    sa += a0 + a1;
                    acc[v] = pk2(0.f, 0.f);
                }
                if (warp == 0 && lane == 0) K1TRACE(5, it);
                lane_part[(buf * kConsumerWarps + warp) * 32 + lane] = make_float4(sx, sa, sy, sr);
                if (lane == 0) warp_hdr[buf * kConsumerWarps + warp] = make_float2(m_ref, __int_as_float(unit));
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull0 + 8u * buf);
                m_ref = kNegHuge;
                sy = 0.f;
                sr = 0.f;
                ++it;
                if (++buf == NF) { buf = 0; bpar ^= 1; }
                fresh = true;
            }
        }
        // end of stream: pass the sentinel on to every finaliser (the next use of each buffer)
        for (int e = 0; e < NF; ++e) {
            if (it + e >= NF) mbar_wait(pempty0 + 8u * buf, bpar);
            if (lane == 0) {
                warp_hdr[buf * kConsumerWarps + warp] = make_float2(kNegHuge, __int_as_float(-1));
                mbar_arrive(pfull0 + 8u * buf);
            }
            if (++buf == NF) { buf = 0; bpar ^= 1; }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Generic path for shapes the tiling cannot cut (tiny or odd volumes): one CTA per unit, plain
// loads, three passes.  Correctness net for edge cases, not a performance path.
// ----------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_elem(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_elem<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_elem<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

__device__ float block_reduce(float v, float* scratch, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = is_max ? kNegHuge : 0.f;
    for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, scratch[i]) : r + scratch[i];
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) integral_fwd_generic_kernel(const FwdParams p) {
    __shared__ float pz[kMaxD];
    __shared__ int peak_bins[kMaxD];
    __shared__ float scratch[8];
    const int D = p.t.D, H = p.t.H, W = p.t.W, HW = H * W;
    const int unit = blockIdx.x;
    const T* src = static_cast<const T*>(p.logits) + (size_t)unit * D * HW;
    float m = kNegHuge;
    for (int i = threadIdx.x; i < D * HW; i += blockDim.x) m = fmaxf(m, load_elem(src, i));
    const float M = block_reduce(m, scratch, true) * kLog2e;
    float sx = 0.f, sy = 0.f, sa = 0.f;
    for (int d = 0; d < D; ++d) {
        float sd = 0.f;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
            const float e = ex2(fmaf(load_elem(src, (size_t)d * HW + i), kLog2e, -M));
            const int h = i / W, w = i - h * W;
            sd += e;
            sx = fmaf((float)w, e, sx);
            sy = fmaf((float)h, e, sy);
        }
        sa += sd;
        sd = block_reduce(sd, scratch, false);
        if (threadIdx.x == 0) pz[d] = sd;
    }
    sx = block_reduce(sx, scratch, false);
    sy = block_reduce(sy, scratch, false);
    sa = block_reduce(sa, scratch, false);
    __syncthreads();
    if (threadIdx.x < 32) finalise_unit(p, unit, pz, peak_bins, M, sx / sa, sy / sa, threadIdx.x);
}

// ----------------------------------------------------------------------------------------------
// Stand-alone find_peak on rows of a [rows, D] matrix (API parity with KPDetector3DMulti.find_peak).
__global__ void __launch_bounds__(128) find_peak_kernel(const float* __restrict__ pz, int64_t* __restrict__ idx, int rows, int D, int NH) {
    __shared__ float row[4][kMaxD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 4 + w;
    if (r >= rows) return;
    for (int d = lane; d < D; d += 32) row[w][d] = pz[(size_t)r * D + d];
    __syncwarp();
    float cv[kMaxD / 32];
    peak_candidates(row[w], D, lane, cv);
    for (int h = 0; h < NH; ++h) {
        const int bd = take_best_peak(cv, lane);
        if (lane == 0) idx[(size_t)r * NH + h] = bd;
    }
}
cudaError_t launch_find_peak(const float* pz, int64_t* idx, int rows, int D, int NH, cudaStream_t st) {
    find_peak_kernel<<<(rows + 3) / 4, 128, 0, st>>>(pz, idx, rows, D, NH);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------
template <typename T, int U>
static cudaError_t launch_fast(const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
#ifdef XSUP_TRACE
    if (const char* d = getenv("XSUP_K1_TRACE")) {
        long long* ptr = reinterpret_cast<long long*>(strtoull(d, nullptr, 0));
        cudaMemcpyToSymbolAsync(g_k1_trace, &ptr, sizeof(ptr), 0, cudaMemcpyHostToDevice, st);
    }
#endif
    auto kern = integral_fwd_kernel<T, U>;
    static unsigned long long attr_done = 0;             // per instantiation; one bit per device
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    kern<<<grid, kFwdThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_fast_u(const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (p.t.U) {
        case 8: return launch_fast<T, 8>(p, grid, smem, st);
        case 4: return launch_fast<T, 4>(p, grid, smem, st);
        case 2: return launch_fast<T, 2>(p, grid, smem, st);
        default: return launch_fast<T, 1>(p, grid, smem, st);
    }
}

cudaError_t launch_integral_fwd(FwdParams p, bool fast, int dtype, int num_sms, cudaStream_t st) {
    if (!fast) {
        if (dtype == XSUP_F32) integral_fwd_generic_kernel<float><<<p.n_units, 256, 0, st>>>(p);
        else integral_fwd_generic_kernel<__nv_bfloat16><<<p.n_units, 256, 0, st>>>(p);
        return cudaGetLastError();
    }
    // Finaliser warps (= partial buffers): two; for short units (<= 256 KB, whose epilogue outlasts two unit-streaming times) as many
    // as kMaxFinalisers, as long as their tables do not cost a ring stage.
    const int dp = (p.t.D + 1) & ~1;
    auto fixed_for = [&](int nf) {
        return (size_t)nf * (kConsumerWarps * 32 * sizeof(float4) + p.t.tasks_per_unit * sizeof(float2) + kConsumerWarps * sizeof(float2) +
                             2 * dp * sizeof(float)) +
               (size_t)(2 * kMaxStages + 2 * kMaxFinalisers) * 8 + (size_t)kMaxStages * sizeof(int2);
    };
    auto stages_for = [&](int nf) {
        int n = (int)((kSmemBudget - fixed_for(nf)) / p.t.stage_bytes);
        n = n > kMaxStages ? kMaxStages : n;
        return n / kGroups * kGroups;
    };
    p.nfin = 2;
    if (p.t.unit_bytes <= 256 * 1024)
        for (int nf = kMaxFinalisers; nf > 2; --nf)
            if (stages_for(nf) == stages_for(2)) { p.nfin = nf; break; }
    const size_t fixed = fixed_for(p.nfin);
    int nst = (int)((kSmemBudget - fixed) / p.t.stage_bytes);
    nst = nst > kMaxStages ? kMaxStages : nst;
    // A slot must always be consumed by the same warp group (slot = s % nst, group = s % kGroups): a waiter
    // may only wait on phase k of an mbarrier if it observed phase k-1 itself, because bulk copies complete
    // out of order and try_wait.parity cannot tell "phase k done" from "phase k-1 still pending".
    nst = nst / kGroups * kGroups;
    if (nst < kGroups) return cudaErrorInvalidConfiguration;
    p.nst = nst;
    const size_t smem = (size_t)nst * p.t.stage_bytes + fixed;
    const int grid = p.n_units < num_sms ? p.n_units : num_sms;
    return dtype == XSUP_F32 ? launch_fast_u<float>(p, grid, smem, st) : launch_fast_u<__nv_bfloat16>(p, grid, smem, st);
}

}  // namespace xsup
