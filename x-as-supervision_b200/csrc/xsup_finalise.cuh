// Per-unit epilogue of the integral head (find_peak, top-NH, window depth, outputs, saved statistics), shared by the
// streaming forward (integral_fwd.cu) and the conv-fused forward (conv_head_fwd.cu).
#pragma once
#include "xsup_internal.h"

namespace xsup {

// ----------------------------------------------------------------------------------------------
// find_peak (…_multi.py:24-34) on one depth row held in shared memory, by one warp.
// cv[j] is the candidate value of bin lane+32j: pz if it is a non-strict interior local maximum,
// 0 if it is an interior non-peak (the reference's masked value), -1 if it cannot be chosen.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void peak_candidates(const float* pz, int D, int lane, float (&cv)[kMaxD / 32]) {
#pragma unroll
    for (int j = 0; j < kMaxD / 32; ++j) {
        const int d = lane + 32 * j;
        float v = -1.0f;
        if (d >= 1 && d <= D - 2) {
            const float c = pz[d];
            v = (c >= pz[d - 1] && c >= pz[d + 1]) ? c : 0.0f;
        }
        cv[j] = v;
    }
}
// next entry of topk: largest candidate, lowest bin on ties; marks it taken.  Warp-uniform result.
__device__ __forceinline__ int take_best_peak(float (&cv)[kMaxD / 32], int lane) {
    float bv = -2.0f;
    int bd = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kMaxD / 32; ++j)
        if (cv[j] > bv) { bv = cv[j]; bd = lane + 32 * j; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int od = __shfl_xor_sync(0xffffffffu, bd, o);
        if (ov > bv || (ov == bv && od < bd)) { bv = ov; bd = od; }
    }
#pragma unroll
    for (int j = 0; j < kMaxD / 32; ++j)
        if (lane + 32 * j == bd) cv[j] = -1.0f;
    return bd;
}

// All NH entries of topk at once: the position of every candidate in the order (value desc, bin asc) is the number of
// bins that precede it, counted against all D bins with one shuffle per bin.  ~D*(2 + 3*ceil(D/32)) issue slots whatever NH
// is, against ~250 per hypothesis for the serial selection above; same total order, so the same bins.
// NJ = ceil(D/32) is a template parameter: predicated-off compares of unused register slots still take issue slots, and
// this warp shares its scheduler with four streaming warps.
template <int NJ>
__device__ __forceinline__ void rank_peaks(const float (&cv)[kMaxD / 32], int NH, int lane, int* bins) {
    int rank[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) rank[j] = 0;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const float ov = __shfl_sync(0xffffffffu, cv[jj], l);
            const int od = l + 32 * jj;
#pragma unroll
            for (int j = 0; j < NJ; ++j) rank[j] += (ov > cv[j] || (ov == cv[j] && od < lane + 32 * j)) ? 1 : 0;
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        if (cv[j] >= 0.0f && rank[j] < NH) bins[rank[j]] = lane + 32 * j;      // bins >= D carry -1 and are never chosen
}

// ----------------------------------------------------------------------------------------------
// Unit epilogue, executed by one full warp.  On entry pz[0..D) in shared memory holds the raw
// (un-normalised) depth marginal relative to the log2-domain reference `M`; xbar, ybar are the
// w- and h-expectations in bin units (warp-uniform).  The callers form them as ratios of sums that
// went through the SAME accumulators, so the rounding of the accumulation cancels to first order and
// the residual error scales with the spread of the distribution, not with its position.
// ----------------------------------------------------------------------------------------------
__device__ inline void finalise_unit(const FwdParams& p, int unit, float* pz, int* bins, float M, float xbar, float ybar, int lane) {
    const int D = p.t.D, H = p.t.H, W = p.t.W, NH = p.NH;
    const int b = unit / p.K, k = unit - b * p.K;
    float ssum = 0.f;
    for (int d = lane; d < D; d += 32) ssum += pz[d];
    const float S = warp_sum(ssum);
    const float invS = 1.0f / S;
    float* st = p.stats + (size_t)unit * p.stats_stride;
    for (int d = lane; d < D; d += 32) {
        const float v = pz[d] * invS;
        pz[d] = v;
        st[4 + d] = v;
        if (b == 0) p.dmap[k * D + d] = v;                      // depth_prob_map = accu_z[0] (…_multi.py:48)
    }
    __syncwarp();
    if (lane == 0) {
        st[0] = M + log2f(S);                                   // log2-domain log-sum-exp: p = 2^(l*log2e - st[0])
        st[1] = xbar;
        st[2] = ybar;
        st[3] = M;
    }
    // the reference normalises x by H and y by W (…_multi.py:78-79); kept literally
    const float x = xbar / (float)H * 2.0f - 1.0f;
    const float y = ybar / (float)W * 2.0f - 1.0f;

    if (p.head == XSUP_HEAD_SINGLE) {                            // keypoint_detector_integral.py:37,41,59
        float zs = 0.f;
        for (int d = lane; d < D; d += 32) zs = fmaf((float)d, pz[d], zs);
        zs = warp_sum(zs);
        if (lane == 0) {
            float* o = p.kps + ((size_t)b * p.K + k) * 3;
            o[0] = x;
            o[1] = y;
            o[2] = zs / (float)D * 2.0f - 1.0f;
            st[4 + D] = zs;
        }
        return;
    }

    // find_peak (…_multi.py:24-34): non-strict interior local maxima, value-descending top-NH.
    // Candidates with value 0 (non-peaks) fill the remaining slots by ascending bin.
    // Phase 1: the NH peak bins into shared memory — serially (arg-max, mark, repeat: a dependency chain of ~250 issue
    // slots per hypothesis) or, when that would cost more than twice the all-at-once ranking, by rank.  At 32^3 with
    // NH = 16 the serial chain (~4 000 cycles) was as long as streaming the 128 KB unit.
    float cv[kMaxD / 32];
    peak_candidates(pz, D, lane, cv);
    const int nj = (D + 31) >> 5;
    if (125 * NH > D * (2 + 3 * nj)) {
        if (nj == 1) rank_peaks<1>(cv, NH, lane, bins);
        else if (nj == 2) rank_peaks<2>(cv, NH, lane, bins);
        else if (nj <= 4) rank_peaks<4>(cv, NH, lane, bins);
        else rank_peaks<kMaxD / 32>(cv, NH, lane, bins);
    } else {
        for (int h = 0; h < NH; ++h) {
            const int bd = take_best_peak(cv, lane);
            if (lane == 0) bins[h] = bd;
        }
    }
    __syncwarp();
    // Phase 2 (one lane per hypothesis, no shuffles): windowed depth expectation (…_multi.py:57-62): zero-padded,
    // count_include_pad average pools of d*pz and pz, gathered at the peak bin.  With the per-hypothesis warp
    // reductions this used to be the bottleneck of small volumes / many hypotheses (32^3, NH = 16: 9 us per unit).
    const int half = p.NS >> 1;
    const float fNS = (float)p.NS;
    for (int h = lane; h < NH; h += 32) {
        const int bd = bins[h];
        const int lo = max(0, bd - half), hi = min(D - 1, bd + half);
        float sw = 0.f, nw = 0.f;
        for (int d = lo; d <= hi; ++d) {
            const float v = pz[d];
            sw += v;
            nw = fmaf((float)d, v, nw);
        }
        const float zbar = (nw / fNS) / (sw / fNS);
        float* o = p.kps + (((size_t)b * NH + h) * p.K + k) * 3;
        o[0] = x;
        o[1] = y;
        o[2] = zbar / (float)D * 2.0f - 1.0f;
        if (p.peak_idx) p.peak_idx[((size_t)b * p.K + k) * NH + h] = bd;
        st[4 + D + 3 * h + 0] = (float)bd;
        st[4 + D + 3 * h + 1] = sw;
        st[4 + D + 3 * h + 2] = nw / sw;
    }
}


}  // namespace xsup
