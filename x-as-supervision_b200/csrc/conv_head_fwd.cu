// K7 — the head's final 1x1 convolution FUSED into the integral head, forward.
//
// Replaces modules/integral_base_modules/deconv_head.py:33-35 (`Conv2d(C, K*D, 1)`, the last layer of
// `self.net`) followed by modules/keypoint_detector_integral_multi.py:69-88 for the forward / eval path
// (eval.py:120): the logits `[B, K*D, H, W]` (4.56 GB at B=256) are never written to or read from HBM.
// This is the one dense contraction next to the path, so it runs on the 5th-generation tensor cores:
//
//   logits[b, k*D+d, p] = sum_c Wt[k*D+d, c] * X[b, p, c] + bias[k*D+d]        p = h*W+w, C = 256
//
// Persistent CTAs (one per SM) walk the work items (sample b, group of 128 output rows = 128/D joints) in b-major
// order, so the ~9 row groups of a sample run at the same time on neighbouring CTAs and share its activations in L2:
//   warp 16      TMA producer: the item's 128 x C weight slab (bf16, K-major, SWIZZLE_128B) as soon as the previous item's
//                MMAs have retired, then the sample's activations as 128-pixel tiles (channels-last bf16, K-major)
//                through a 2-stage ring; barrier phases run on across items
//   warp 17      MMA issuer: one elected thread; per item the weight slab goes from its landing buffer into tensor memory
//                (tcgen05.cp, 128 columns) and the buffer is released for the next slab; per tile 16 tcgen05.mma
//                cta_group::1.kind::f16 M=128 N=128 K=16 with A from TMEM and precomputed shared-memory descriptors for B
//                (the issue loop is ~5 instructions per MMA); accumulators in TMEM (3 x 128 columns: the epilogue of tile t
//                overlaps the MMAs of t+1, t+2); tcgen05.commit releases the smem stage and publishes the accumulator
//   warps 0..15  epilogue: thread = TMEM lane = output row (joint, d); warps w, w+4, w+8, w+12 share a lane quarter and
//                take 32 columns each; tcgen05.ld; the softmax statistics of the row (running max, sum e, sum w*e,
//                sum h*e) stay in FOUR registers per thread - rows are depth bins, so the depth marginal pz[d] is simply
//                the row sum; no shuffles per tile
//   warps 18,19  finalisers: merge the column parts and the rows of a joint (log-sum-exp through double-buffered shared
//                memory) and run the same finaliser as the streaming kernel (find_peak, top-NH, window depth, outputs,
//                saved statistics) while the other warps are already on the next item
// Measured on the way here (B=256, ablations): a non-persistent CTA per item spent 0.25 ms of 0.62 ms in start-up /
// tear-down, and computing the descriptors inside the issue loop cost ~65 cycles per MMA.
// Roofline: tensor (2*K*D*C flops per pixel = 584 GFLOP at B=256) with MUFU.EX2 of the epilogue at the same
// order (one exp per logit); HBM traffic is the activations only (0.54 GB).
#include "xsup_finalise.cuh"
#include "xsup_umma.cuh"

namespace xsup {

constexpr int kCvEpiWarps = 16;         // four per TMEM lane quarter: each takes 32 of a tile's 128 columns
constexpr int kCvParts = kCvEpiWarps / 4;
constexpr int kCvFinWarps = 2;
constexpr int kCvThreads = (kCvEpiWarps + 2 + kCvFinWarps) * 32;
constexpr int kCvRows = 128;           // UMMA M: output rows per CTA
constexpr int kCvPix = 128;            // UMMA N: pixels per tile
constexpr int kCvStages = 2;
constexpr int kCvAcc = 3;              // TMEM accumulator buffers (columns [128,512)); columns [0,128) hold the weight slab
constexpr int kCvAccCol = 128;
constexpr int kCvKBBytes = kCvRows * kCvKB * 2;   // 16 KB: one [128 x 64] bf16 k-block (same for W and X tiles)

struct ConvHeadParams {
    FwdParams f;                // kps, dmap, peak_idx, stats, K, NH, NS, head, stats_stride, t.{D,H,W}
    const float* bias;          // [K*D] or nullptr
    float* logits_out;          // optional [B, K*D, H*W] fp32 (validation); nullptr in production
    // backward mode (MODE = 1): d loss / d logits, recomputed from the same GEMM
    const float* coef;          // [B*K][coef_stride]: lse2, a, b, base0, wc, hc, -, -, c[0..D)  (integral_coef_kernel)
    int coef_stride;
    __nv_bfloat16* g_out;       // [B, K*D, H*W] bf16
    float* gbias_part;          // [B, kCvParts, K*D] fp32: per-item partial sums of g over the pixels
    int C, HW, rows_total;      // channels, pixels per sample, K*D
    int groups;                 // CTAs per sample = ceil(K*D / 128)
    int n_tiles;                // HW / 128
    int kblocks;                // C / 64
    int items;                  // B * groups
};

constexpr uint32_t kCvIdesc = umma_idesc_bf16(kCvRows, kCvPix);

// ------------------------------------------------------------------ kernel
// TF32 = true (forward only): fp32 operands, `kind::tf32` MMAs - the precision of the reference's own conv on this GPU.  A k-block is
// then 32 channels (still 128 bytes per row), KBN = C/32; the fp32 weight slab (128 KB at C = 256) passes through the 64 KB landing
// buffer in two halves into 256 TMEM columns; an activation ring stage holds 64 pixels (64 KB), and two stages fill one 128-column
// accumulator (two accumulators instead of three), so the epilogue and the finalisers are the very same code.
template <int KBN, int MODE, bool TF32 = false>   // MODE 0: forward statistics + finaliser; MODE 1: backward, emits d loss / d logits in bf16
__global__ void __launch_bounds__(kCvThreads, 1) conv_head_fwd_kernel(const __grid_constant__ CUtensorMap map_w,
                                                                      const __grid_constant__ CUtensorMap map_x,
                                                                      const ConvHeadParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the dynamic shared window is only 16-byte aligned by contract: round up to the 1024 B the swizzle needs
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWKB = TF32 ? KBN / 2 : KBN;                       // k-blocks the landing buffer holds (TF32: half the slab)
    constexpr uint32_t kXKBBytes = TF32 ? kCvKBBytes / 2 : kCvKBBytes;   // one k-block of a ring stage: [64 | 128 pixels x 128 B]
    constexpr int kAcc = TF32 ? 2 : kCvAcc;                          // accumulators of 128 columns
    constexpr int kAccCol0 = TF32 ? 256 : kCvAccCol;                 // TMEM columns [0, kAccCol0) hold the weight slab
    static_assert(!TF32 || (MODE == 0 && KBN % 2 == 0), "the tf32 variant is forward-only and needs C % 64 == 0");
    uint8_t* sW = smem;                                             // [kWKB][128 x 128 B]
    uint8_t* sX = sW + (size_t)kWKB * kCvKBBytes;                   // [stages][KBN][128 | 64 pixels x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + (size_t)kCvStages * KBN * kXKBBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    // MODE 0: row statistics + finaliser scratch (20 KB); MODE 1: the same region (32 KB) stages the gradient tiles
    float4* row_stat = reinterpret_cast<float4*>(bars + 32);                                   // [2 items][kCvParts][128] (m, s, sx, sy)
    float* pz_s = reinterpret_cast<float*>(row_stat + 2 * kCvParts * kCvRows);                 // [kCvFinWarps][kMaxD]
    int* bins_s = reinterpret_cast<int*>(pz_s + kCvFinWarps * kMaxD);                          // [kCvFinWarps][kMaxD]
    uint8_t* g_stage = reinterpret_cast<uint8_t*>(row_stat);                                   // [kCvEpiWarps][32 rows][64 B]
    const uint32_t b_wfull = smem_u32(bars), b_wempty = b_wfull + 8, b_xfull = b_wempty + 8, b_xempty = b_xfull + 8 * kCvStages,
                   b_afull = b_xempty + 8 * kCvStages, b_aempty = b_afull + 8 * kCvAcc, b_rsfull = b_aempty + 8 * kCvAcc,
                   b_rsfree = b_rsfull + 16;

    if (threadIdx.x == 0) {
        mbar_init(b_wfull, 1);
        mbar_init(b_wempty, 1);
        for (int i = 0; i < kCvStages; ++i) {
            mbar_init(b_xfull + 8 * i, 1);
            mbar_init(b_xempty + 8 * i, 1);
        }
        for (int i = 0; i < kCvAcc; ++i) {
            mbar_init(b_afull + 8 * i, 1);
            mbar_init(b_aempty + 8 * i, kCvEpiWarps);                // one elected lane per epilogue warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_rsfull + 8 * i, kCvEpiWarps);
            mbar_init(b_rsfree + 8 * i, kCvFinWarps);
        }
        mbar_fence_init();
    }
    if (warp == kCvEpiWarps + 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int T = p.n_tiles;

    if (warp == kCvEpiWarps) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int g = 0, n = 0;                                        // ring stages / items this CTA has started
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                const int b = item / p.groups, row0 = (item - b * p.groups) * kCvRows;
                if constexpr (!TF32) {
                    mbar_wait(b_wempty, (n & 1) ^ 1);                // the previous slab has been copied into tensor memory
                    mbar_arrive_expect_tx(b_wfull, (uint32_t)KBN * kCvKBBytes);
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) tma_load_2d(smem_u32(sW) + kb * kCvKBBytes, &map_w, kb * kCvKB, row0, b_wfull);
                    for (int t = 0; t < T; ++t, ++g) {
                        const int s = g % kCvStages, it = g / kCvStages;
                        mbar_wait(b_xempty + 8 * s, (it & 1) ^ 1);
                        mbar_arrive_expect_tx(b_xfull + 8 * s, (uint32_t)KBN * kCvKBBytes);
                        const uint32_t dst = smem_u32(sX) + (uint32_t)s * KBN * kCvKBBytes;
#pragma unroll
                        for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kCvKBBytes, &map_x, kb * kCvKB, b * p.HW + t * kCvPix, b_xfull + 8 * s);
                    }
                } else {
                    // slab half 0, the first ring stages, slab half 1 (its buffer is free once half 0 sits in TMEM), the rest
                    for (int ht = -1; ht < 2 * T; ++ht) {
                        if (ht == -1 || ht == kCvStages - 1) {
                            const int half = ht == -1 ? 0 : 1, wn = 2 * n + half;
                            mbar_wait(b_wempty, (wn & 1) ^ 1);
                            mbar_arrive_expect_tx(b_wfull, (uint32_t)kWKB * kCvKBBytes);
#pragma unroll
                            for (int kb = 0; kb < kWKB; ++kb)
                                tma_load_2d(smem_u32(sW) + kb * kCvKBBytes, &map_w, (half * kWKB + kb) * 32, row0, b_wfull);
                        }
                        if (ht < 0) continue;
                        const int s = g % kCvStages, it = g / kCvStages;
                        mbar_wait(b_xempty + 8 * s, (it & 1) ^ 1);
                        mbar_arrive_expect_tx(b_xfull + 8 * s, (uint32_t)KBN * kXKBBytes);
                        const uint32_t dst = smem_u32(sX) + (uint32_t)s * KBN * kXKBBytes;
#pragma unroll
                        for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kXKBBytes, &map_x, kb * 32, b * p.HW + ht * (kCvPix / 2), b_xfull + 8 * s);
                        ++g;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kCvEpiWarps + 1) {
        // ------------------------------------------------------------ MMA issuer
        // The item's weight slab goes from its shared-memory landing buffer into TENSOR memory (tcgen05.cp, 8 columns per 16
        // channels; in order behind the previous item's MMAs) and the buffer is released at once, so the producer prefetches the
        // next slab a whole item ahead and the MMAs read A from TMEM: with both operands in shared memory an M = N = 128 MMA
        // needs 128 B/clk of operand reads, which is all the shared memory delivers.
        uint64_t da[kWKB], db0[KBN];                                 // descriptors of the landing buffer and of ring stage 0, per k-block
#pragma unroll
        for (int kb = 0; kb < kWKB; ++kb) da[kb] = umma_desc_sw128(smem_u32(sW) + kb * kCvKBBytes);
#pragma unroll
        for (int kb = 0; kb < KBN; ++kb) db0[kb] = umma_desc_sw128(smem_u32(sX) + kb * kXKBBytes);
        constexpr uint64_t kStageStep = (uint64_t)((KBN * kXKBBytes) >> 4);    // start-address field units
        constexpr uint32_t kIdesc = TF32 ? umma_idesc_tf32(kCvRows, kCvPix / 2) : kCvIdesc;
        int s = 0, a = 0, n = 0;
        uint32_t xph = 0, aeph = 1;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
#pragma unroll
            for (int half = 0; half < (TF32 ? 2 : 1); ++half) {
                mbar_wait(b_wfull, (TF32 ? 2 * n + half : n) & 1);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int kb = 0; kb < kWKB; ++kb)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(((half * kWKB + kb) * 4 + q) * 8), da[kb] + 2 * q);
                    umma_commit(b_wempty);                           // the landing buffer may take the next (half) slab
                }
                __syncwarp();
            }
            for (int t = 0; t < T; ++t) {
                mbar_wait(b_aempty + 8 * a, aeph);                   // the epilogue has drained this accumulator
#pragma unroll
                for (int hf = 0; hf < (TF32 ? 2 : 1); ++hf) {        // TF32: two 64-pixel stages fill one 128-column accumulator
                    mbar_wait(b_xfull + 8 * s, xph);                 // the tile has landed
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t acc = tmem_base + (uint32_t)(kAccCol0 + a * kCvPix + hf * (kCvPix / 2));
                        const uint64_t soff = (uint64_t)s * kStageStep;
#pragma unroll
                        for (int kb = 0; kb < KBN; ++kb) {
                            const uint64_t b0 = db0[kb] + soff;
                            // +32 B per K step (16 bf16 / 8 tf32) inside the 128-byte swizzle atom: +2 in the start-address field
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if constexpr (TF32) umma_tf32_ts(acc, tmem_base + (uint32_t)((kb * 4 + q) * 8), b0 + 2 * q, kIdesc, (kb | q) ? 1u : 0u);
                                else umma_f16_ts(acc, tmem_base + (uint32_t)((kb * 4 + q) * 8), b0 + 2 * q, kIdesc, (kb | q) ? 1u : 0u);
                            }
                        }
                        umma_commit(b_xempty + 8 * s);               // smem stage reusable once these MMAs have read it
                        if (hf == (TF32 ? 1 : 0)) umma_commit(b_afull + 8 * a);   // accumulator complete
                    }
                    __syncwarp();
                    if (++s == kCvStages) { s = 0; xph ^= 1; }
                }
                if (++a == kAcc) { a = 0; aeph ^= 1; }
            }
        }
    } else if (warp < kCvEpiWarps) {
        // ------------------------------------------------------------ epilogue: thread = output row, 32 columns per warp and tile
        const int quarter = warp & 3, part = warp >> 2;              // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int row = quarter * 32 + lane;
        const int Wd = p.f.t.W, c0 = part * 32;
        int g = 0, n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
            const int b = item / p.groups, row0 = (item - b * p.groups) * kCvRows, grow = row0 + row;
            const bool live = grow < p.rows_total;
            const float bias = (live && p.bias) ? p.bias[grow] : 0.f, bl = bias * kLog2e;
            if (MODE == 1) {
                // ---- backward: g = p * (a (w - wc) + b (h - hc) + c[d] + base0), p = 2^(l log2e - lse2)  
                const int D = p.f.t.D, k = grow / D, d = grow - k * D;
                float nlse = 0.f, ca = 0.f, cb = 0.f, cbase = 0.f, wc = 0.f, hc = 0.f;
                if (live) {
                    const float* cf = p.coef + ((size_t)b * p.f.K + k) * p.coef_stride;
                    nlse = bl - cf[0]; ca = cf[1]; cb = cf[2]; cbase = cf[3] + cf[8 + d]; wc = cf[4]; hc = cf[5];
                }
                float gsum = 0.f;
                for (int t = 0; t < T; ++t, ++g) {
                    const int a = g % kAcc;
                    const uint32_t aph = (uint32_t)(g / kAcc) & 1u;
                    mbar_wait(b_afull + 8 * a, aph);
                    tc_fence_after();
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kAccCol0 + a * kCvPix + c0), r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_aempty + 8 * a);
                    if (row0 + quarter * 32 >= p.rows_total) continue;                       // warp-uniform: all 32 rows are padding
                    const int pix = t * kCvPix + c0;
                    const int hh = pix / Wd, w0 = pix - hh * Wd;
                    const float rowterm = fmaf(ca, (float)w0 - wc, fmaf(cb, (float)hh - hc, cbase));
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float g0 = ex2(fmaf(__uint_as_float(r[i]), kLog2e, nlse)) * fmaf(ca, (float)i, rowterm);
                        const float g1 = ex2(fmaf(__uint_as_float(r[i + 1]), kLog2e, nlse)) * fmaf(ca, (float)(i + 1), rowterm);
                        gsum += g0 + g1;
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i >> 1]) : "f"(g1), "f"(g0));
                    }
                    // A lane holds 64 B of ITS row; rows are H*W*2 bytes apart, so storing from here would write 32 half
                    // sectors per instruction (measured: +0.56 ms).  Transpose through 2 KB of shared memory per warp
                    // (16-byte chunks XOR-swizzled by the row pair: conflict-free both ways) so that each store
                    // instruction covers 8 rows x 64 contiguous bytes.
                    uint8_t* stg = g_stage + warp * 2048;
                    const int swz = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ swz) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = 8 * j + (lane >> 2), q = lane & 3;                     // row of this warp's 32, 16-byte chunk
                        const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((q ^ ((rr >> 1) & 3)) << 4));
                        const int gr = row0 + quarter * 32 + rr;
                        if (gr < p.rows_total)
                            *reinterpret_cast<uint4*>(p.g_out + ((size_t)b * p.rows_total + gr) * p.HW + pix + 8 * q) = v;
                    }
                    __syncwarp();
                }
                if (live && p.gbias_part) p.gbias_part[((size_t)b * kCvParts + part) * p.rows_total + grow] = gsum;
                continue;
            }
            float m = kNegHuge, s = 0.f, sx = 0.f, sy = 0.f;
            float* lrow = p.logits_out ? p.logits_out + ((size_t)b * p.rows_total + grow) * p.HW : nullptr;
            for (int t = 0; t < T; ++t, ++g) {
                const int a = g % kAcc;
                    const uint32_t aph = (uint32_t)(g / kAcc) & 1u;
                    mbar_wait(b_afull + 8 * a, aph);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kAccCol0 + a * kCvPix + c0), r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_aempty + 8 * a);        // values are in registers: release the accumulator early
                const int pix = t * kCvPix + c0;                     // 32 consecutive pixels of one image row (W >= 32)
                const int hh = pix / Wd, w0 = pix - hh * Wd;
                if (lrow && live) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(lrow + pix + i) = make_float4(__uint_as_float(r[i]) + bias, __uint_as_float(r[i + 1]) + bias,
                                                                                 __uint_as_float(r[i + 2]) + bias, __uint_as_float(r[i + 3]) + bias);
                }
                float cm4[4];                                        // four independent chains instead of one of length 32
#pragma unroll
                for (int j = 0; j < 4; ++j) cm4[j] = __uint_as_float(r[j]);
#pragma unroll
                for (int i = 4; i < 32; ++i) cm4[i & 3] = fmaxf(cm4[i & 3], __uint_as_float(r[i]));
                const float cm = fmaf(fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3])), kLog2e, bl);   // log2 domain, bias included
                if (cm > m) {                                        // per-row running max: rescale three registers
                    const float sc = ex2(m - cm);
                    s *= sc; sx *= sc; sy *= sc;
                    m = cm;
                }
                const float sh = bl - m;
                float cs4[4] = {0.f, 0.f, 0.f, 0.f}, cx4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float e = ex2(fmaf(__uint_as_float(r[i]), kLog2e, sh));
                    cs4[i & 3] += e;
                    cx4[i & 3] = fmaf((float)i, e, cx4[i & 3]);
                }
                const float cs = (cs4[0] + cs4[1]) + (cs4[2] + cs4[3]), cx = (cx4[0] + cx4[1]) + (cx4[2] + cx4[3]);
                s += cs;
                sx += fmaf((float)w0, cs, cx);
                sy = fmaf((float)hh, cs, sy);
            }
            // hand the row statistics of this item to the finalisers (double-buffered: they lag by up to one item)
            const int buf = n & 1, u = n >> 1;
            mbar_wait(b_rsfree + 8 * buf, (u & 1) ^ 1);
            row_stat[(buf * kCvParts + part) * kCvRows + row] = make_float4(m, s, sx, sy);
            __syncwarp();
            if (lane == 0) mbar_arrive(b_rsfull + 8 * buf);
        }
    } else if (MODE == 0) {
        // ------------------------------------------------------------ finalisers: one joint per warp and round
        const int fw = warp - (kCvEpiWarps + 2);
        const int D = p.f.t.D, jpc = kCvRows / D;                    // joints per item
        float* pz = pz_s + fw * kMaxD;
        int n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
            const int b = item / p.groups, row0 = (item - b * p.groups) * kCvRows;
            const int buf = n & 1, u = n >> 1;
            mbar_wait(b_rsfull + 8 * buf, u & 1);
            const float4* rs = row_stat + (size_t)buf * kCvParts * kCvRows;
            for (int jl = fw; jl < jpc; jl += kCvFinWarps) {
                const int k = row0 / D + jl;
                if (k >= p.f.K) continue;                            // padding rows of the last group
                float M = kNegHuge;
                for (int d = lane; d < D; d += 32)
#pragma unroll
                    for (int q = 0; q < kCvParts; ++q) M = fmaxf(M, rs[q * kCvRows + jl * D + d].x);
                M = warp_max(M);
                float ax = 0.f, ay = 0.f, as = 0.f;
                for (int d = lane; d < D; d += 32) {
                    float pd = 0.f;
#pragma unroll
                    for (int q = 0; q < kCvParts; ++q) {             // log-sum-exp merge of the column parts, fixed order
                        const float4 v = rs[q * kCvRows + jl * D + d];
                        const float sc = ex2(v.x - M);
                        pd = fmaf(v.y, sc, pd);
                        ax = fmaf(v.z, sc, ax);
                        ay = fmaf(v.w, sc, ay);
                    }
                    pz[d] = pd;
                    as += pd;
                }
                as = warp_sum(as); ax = warp_sum(ax); ay = warp_sum(ay);
                __syncwarp();
                finalise_unit(p.f, b * p.f.K + k, pz, bins_s + fw * kMaxD, M, ax / as, ay / as, lane);
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_rsfree + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kCvEpiWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------ activation packing
// [B, C, HW] fp32 (what a plain NCHW backbone hands over) -> [B, HW, C] bf16 (channels-last, the K-major operand the
// tensor maps above describe): a 64 x 64 tile transpose through shared memory with the cast fused in; reads are
// 128-bit along the pixels, writes 128-bit along the channels.  1.07 GB in + 0.54 GB out at B=256.
__global__ void __launch_bounds__(256) pack_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int HW) {
    __shared__ float tile[64][65];
    const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64, b = blockIdx.z;
    const float* src = x + ((size_t)b * C + c0) * HW + p0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = threadIdx.x + 256 * i;                       // 1024 float4 = 64 channels x 16 pixel quads
        const int c = q >> 4, pq = (q & 15) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)c * HW + pq);
        tile[c][pq] = v.x; tile[c][pq + 1] = v.y; tile[c][pq + 2] = v.z; tile[c][pq + 3] = v.w;
    }
    __syncthreads();
    __nv_bfloat16* dst = y + ((size_t)b * HW + p0) * C + c0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int q = threadIdx.x + 256 * i;                       // 512 chunks = 64 pixels x 8 groups of 8 channels
        const int pp = q >> 3, cg = (q & 7) * 8;
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(tile[cg + 2 * j][pp], tile[cg + 2 * j + 1][pp]);
            ow[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)pp * C + cg) = o;
    }
}

cudaError_t launch_pack_nhwc_bf16(const float* x, void* y, int B, int C, int HW, cudaStream_t st) {
    pack_nhwc_bf16_kernel<<<dim3(HW / 64, C / 64, B), 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(y), C, HW);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ host side
template <int KBN, int MODE>
static cudaError_t launch_kbn(const CUtensorMap& map_w, const CUtensorMap& map_x, const ConvHeadParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = conv_head_fwd_kernel<KBN, MODE>;
    static unsigned long long attr_done = 0;             // per instantiation; one bit per device
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    kern<<<grid, kCvThreads, smem, st>>>(map_w, map_x, p);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_conv_head(const void* x_nhwc, const void* w, ConvHeadParams p, int B, int num_sms, cudaStream_t st) {
    p.HW = p.f.t.H * p.f.t.W;
    p.rows_total = p.f.K * p.f.t.D;
    p.groups = (p.rows_total + kCvRows - 1) / kCvRows;
    p.n_tiles = p.HW / kCvPix;
    p.kblocks = p.C / kCvKB;
    p.items = B * p.groups;
    CUtensorMap map_w, map_x;
    if (!make_map(&map_w, w, p.rows_total, p.C) || !make_map(&map_x, x_nhwc, (long long)B * p.HW, p.C)) return cudaErrorNotSupported;
    const size_t scratch = MODE == 0 ? 2 * kCvParts * kCvRows * sizeof(float4) + 2 * kCvFinWarps * kMaxD * 4 : (size_t)kCvEpiWarps * 2048;
    const size_t smem = 1024 + (size_t)(1 + kCvStages) * p.kblocks * kCvKBBytes + 32 * 8 + scratch;
    const int grid = p.items < num_sms ? p.items : num_sms;          // persistent: one CTA per SM
    switch (p.kblocks) {
        case 1: return launch_kbn<1, MODE>(map_w, map_x, p, grid, smem, st);
        case 2: return launch_kbn<2, MODE>(map_w, map_x, p, grid, smem, st);
        case 3: return launch_kbn<3, MODE>(map_w, map_x, p, grid, smem, st);
        default: return launch_kbn<4, MODE>(map_w, map_x, p, grid, smem, st);
    }
}

// fp32 operands, tf32 tensor-core arithmetic (forward only): x [B, H*W, C] fp32 channels-last, w [K*D, C] fp32
template <int KBN>
static cudaError_t launch_kbn_tf32(const CUtensorMap& map_w, const CUtensorMap& map_x, const ConvHeadParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = conv_head_fwd_kernel<KBN, 0, true>;
    static unsigned long long attr_done = 0;
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    kern<<<grid, kCvThreads, smem, st>>>(map_w, map_x, p);
    return cudaGetLastError();
}

cudaError_t launch_conv_head_fwd_tf32(const void* x_nhwc_f32, const void* w_f32, const float* bias, float* logits_out, FwdParams f, int B, int C,
                                      int num_sms, cudaStream_t st) {
    ConvHeadParams p{};
    p.f = f;
    p.bias = bias;
    p.logits_out = logits_out;
    p.C = C;
    p.HW = p.f.t.H * p.f.t.W;
    p.rows_total = p.f.K * p.f.t.D;
    p.groups = (p.rows_total + kCvRows - 1) / kCvRows;
    p.n_tiles = p.HW / kCvPix;
    p.kblocks = C / 32;
    p.items = B * p.groups;
    CUtensorMap map_w, map_x;
    if (!make_map_f32(&map_w, w_f32, p.rows_total, C, kCvRows) || !make_map_f32(&map_x, x_nhwc_f32, (long long)B * p.HW, C, kCvPix / 2))
        return cudaErrorNotSupported;
    const size_t scratch = 2 * kCvParts * kCvRows * sizeof(float4) + 2 * kCvFinWarps * kMaxD * 4;
    const size_t smem = 1024 + (size_t)(p.kblocks / 2) * kCvKBBytes + (size_t)kCvStages * p.kblocks * (kCvKBBytes / 2) + 32 * 8 + scratch;
    const int grid = p.items < num_sms ? p.items : num_sms;
    switch (p.kblocks) {
        case 2: return launch_kbn_tf32<2>(map_w, map_x, p, grid, smem, st);
        case 4: return launch_kbn_tf32<4>(map_w, map_x, p, grid, smem, st);
        case 6: return launch_kbn_tf32<6>(map_w, map_x, p, grid, smem, st);
        default: return launch_kbn_tf32<8>(map_w, map_x, p, grid, smem, st);
    }
}

cudaError_t launch_conv_head_fwd(const void* x_nhwc, const void* w, const float* bias, float* logits_out, FwdParams f, int B, int C,
                                 int num_sms, cudaStream_t st) {
    ConvHeadParams p{};
    p.f = f;
    p.bias = bias;
    p.logits_out = logits_out;
    p.C = C;
    return launch_conv_head<0>(x_nhwc, w, p, B, num_sms, st);
}

// d loss / d logits of the fused head, recomputed (never read from memory): the same GEMM, then per element one exp and
// the closed form of SURVEY App. A.2 with the coefficient blocks of integral_coef_kernel; written as bf16 [B, K*D, H*W]
cudaError_t launch_conv_head_bwd_g(const void* x_nhwc, const void* w, const float* bias, const float* coef, int coef_stride, void* g_out,
                                   float* gbias_part, FwdParams f, int B, int C, int num_sms, cudaStream_t st) {
    ConvHeadParams p{};
    p.f = f;
    p.bias = bias;
    p.C = C;
    p.coef = coef;
    p.coef_stride = coef_stride;
    p.g_out = static_cast<__nv_bfloat16*>(g_out);
    p.gbias_part = gbias_part;
    return launch_conv_head<1>(x_nhwc, w, p, B, num_sms, st);
}

}  // namespace xsup
