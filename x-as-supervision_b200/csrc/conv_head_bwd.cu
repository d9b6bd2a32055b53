// K8 — backward of the conv-fused integral head, entirely on the tensor cores.
//
// Replaces autograd through `Conv2d(C, K*D, 1)` (modules/integral_base_modules/deconv_head.py:33-35) followed by
// modules/keypoint_detector_integral_multi.py:69-88.  With L[b,r,p] = sum_c Wt[r,c] X[b,p,c] + bias[r] the logits of sample b
// (r = k*D+d, p = h*W+w) and G = d loss / d L (G = softmax(L) * (a (w-wc) + b (h-hc) + c[d] + base0)),
//
//   d X[b,p,c] = sum_r G[b,r,p] Wt[r,c]          d Wt[r,c] = sum_{b,p} G[b,r,p] X[b,p,c]          d bias[r] = sum_{b,p} G[b,r,p]
//
// Neither L nor G ever exists in HBM: each of the two launches recomputes the logit tile it needs on the tensor cores, forms G
// in the epilogue (one MUFU.EX2 + 3 FP ops per element from the row coefficients), rounds it to bf16 into SHARED memory in the
// swizzled operand layout and feeds it straight back into a second tcgen05.mma chain whose accumulator stays in TMEM:
//
//   MODE_DW (W-stationary): item = (sample b, 128 weight rows).  The 128 x C weight slab is the stationary operand, the sample's
//            activations stream through as 128-pixel tiles:   S[128 rows x 128 pix] = Wt_slab * X_tile^T ;  G = f(S) ;
//            D2[128 rows x C] += G[128 x 128] * X_tile[128 x C].  D2 leaves once per item as a TMA fp32 reduce-add into d Wt.
//   MODE_DX (X-stationary): item = (sample b, 128 pixels).  The 128 x C activation tile is the stationary operand, the weights
//            stream through as 128-row tiles:   S[128 pix x 128 rows] = X_tile * Wt_tile^T ;  G^T = f(S) ;
//            D2[128 pix x C] += G^T[128 x 128] * Wt_tile[128 x C].  D2 is the finished d X tile (all K*D rows were summed in
//            TMEM) and is written once by the TMA, bf16 or fp32, channels-last.
//
// Both modes are ONE kernel.  Every operand tile is [128 x C] bf16 and travels through ONE 3-stage TMA ring (64 KB per stage
// at C = 256): per item first the stationary tile, then the streamed ones.
//   * the stationary tile is copied from its ring stage into TENSOR memory (tcgen05.cp, 8 columns per 16 channels) and the
//     stage is released at once: the first GEMM reads A from TMEM and only B from shared memory;
//   * a streamed tile is B of the first GEMM (K-major) and, from the very same bytes, B of the second GEMM (MN-major);
//   * G [128 x 128] bf16 goes through one 32 KB shared-memory buffer in the swizzled K-major operand layout (A of the second GEMM).
// Why these shapes (measured on the way here, tools/convhead_bwd_probe.py + the clock64 trace hook below): a tcgen05.mma
// with N = 64 occupies the pipe as long as one with N = 128 (64 cycles at M = 128), so 64-row streamed tiles ran the first
// GEMM at half rate (1500-1600 cycles per 64 rows whatever else was tuned); the tensor-pipe queue is shallow, an issuing
// thread is held while its MMAs execute, so each GEMM chain has its own issuing thread; descriptor arithmetic and integer
// divisions in the issue loop idled the pipe for ~1000 cycles per tile.
// Warp roles (19 warps, one persistent CTA per SM):
//   warp 16      TMA producer (ring entries in order: stationary, tile 0 .. T-1, next stationary, ...)
//   warp 17      first-GEMM issuer (one thread): per item 4*C/64 tcgen05.cp + release of that stage; per tile 4*C/64 MMAs
//                M=128 N=128 K=16 (A from TMEM) into the single S accumulator as soon as the epilogue has read the previous one
//   warp 18      second-GEMM issuer (one thread): per tile 8 MMAs M=128 N=C K=16 (A = G from shared memory) once the epilogue
//                has published G; its tcgen05.commit releases the ring stage and the G buffer
//   warps 0..15  epilogue: tcgen05.ld (thread = TMEM lane, 32 columns per warp), G = f(S) with packed fp32x2 arithmetic,
//                st.shared (XOR-swizzled 16-byte chunks), fence.proxy.async, mbarrier arrive; at the end of an item they drain D2
//                through the idle G buffer (16 KB chunks, swizzled, ping-pong) to the TMA: tensor store (d x) / reduce-add (d W)
// TMEM: columns [0,128) = stationary operand, [128,256) = S[128 x 128] fp32, [256, 256+C) = D2.
// Flops: 2 GEMMs of 2*K*D*C*H*W per sample and launch (1.17 TFLOP per launch at B=256, K=17, D=64, C=256).
#include <stdlib.h>

#include "xsup_internal.h"
#include "xsup_umma.cuh"

namespace xsup {

constexpr int kBwEpiWarps = 16;
constexpr int kBwParts = kBwEpiWarps / 4;                  // warps sharing a TMEM lane quarter: 32 of a tile's 128 columns each
constexpr int kBwThreads = (kBwEpiWarps + 3) * 32;         // + TMA producer, first-GEMM issuer, second-GEMM issuer
constexpr int kBwM = 128;                                  // rows of every operand tile = UMMA M, N of the first GEMM, K of the second
constexpr int kBwStages = 3;
constexpr int kBwKB = kBwM * kCvKB * 2;                    // 16 KB: one [128 x 64] bf16 k-block
constexpr int kBwGBytes = kBwM * kBwM * 2;                 // 32 KB: the G tile, two k-blocks
constexpr int kBwACol = 0;                                 // TMEM columns [0,128): the stationary operand, bf16, 8 columns per 16 channels
constexpr int kBwSCol = 128;                               // S accumulator [128 x 128] fp32
constexpr int kBwD2Col = 256;                              // first TMEM column of D2
enum { MODE_DW = 0, MODE_DX = 1 };

struct ConvBwdParams {
    const float* rowcoef;       // [B][rows_pad/2][8] per row pair {nlse, nlse', a, a', b, b', e, e'} (conv_rowcoef_kernel); zero for padding rows
    void* dx;                   // MODE_DX: [B, HW, C] bf16 or fp32
    int dx_f32;
    float* dw;                  // MODE_DW: [K*D, C] fp32, zeroed by the launcher
    float* dbias;               // MODE_DW: [K*D] fp32, zeroed by the launcher; may be NULL
    int C, HW, W, rows_total, rows_pad;
    int per_b;                  // items per sample: row groups (MODE_DW) or 128-pixel tiles (MODE_DX)
    int T;                      // streamed tiles per item
    int last_rows;              // valid rows of an item's last streamed tile (MODE_DX: K*D need not be a multiple of 128), else 128
    int items;
    long long* trace;           // diagnostics (-DXSUP_TRACE builds only): clock64 timeline of CTA 0, see TRACE below
};

// ------------------------------------------------------------------ row coefficients
// Everything the epilogue needs to turn a logit of row r into its gradient, G = 2^(L*log2e + nlse) * (a*w + b*h + e), laid
// out per PAIR of rows (2P, 2P+1) as 8 floats {nlse0, nlse1, a0, a1, b0, b1, e0, e1} so that the activation-stationary
// epilogue (columns = rows) reads packed fp32x2 operands.  out: [B][rows_pad / 2][8].
__global__ void __launch_bounds__(256) conv_rowcoef_kernel(const float* __restrict__ coef, int coef_stride, const float* __restrict__ bias,
                                                           float* __restrict__ out, int B, int K, int D, int rows_pad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * rows_pad) return;
    const int b = (int)(i / rows_pad), r = (int)(i - (long long)b * rows_pad);
    float nl = 0.f, a = 0.f, bb = 0.f, e = 0.f;
    if (r < K * D) {
        const int k = r / D, d = r - k * D;
        const float* cf = coef + ((size_t)b * K + k) * coef_stride;
        a = cf[1];
        bb = cf[2];
        nl = (bias ? bias[r] : 0.f) * kLog2e - cf[0];
        e = (cf[3] + cf[8 + d]) - fmaf(a, cf[4], bb * cf[5]);
    }
    float* o = out + ((size_t)b * rows_pad + (r & ~1)) * 4 + (r & 1);
    o[0] = nl;
    o[2] = a;
    o[4] = bb;
    o[6] = e;
}


// Diagnostics, compiled in only with -DXSUP_TRACE (XSUP_NVCC_EXTRA="-DXSUP_TRACE" python -c "import __graft_entry__ as g; g.build(force=True)"):
// a clock64 timeline of CTA 0, tiles / items 8..23, written to the device buffer whose address is in XSUP_CONVBWD_TRACE
// (tools/convhead_bwd_probe.py with TRACE=1 allocates and prints it).  This is what found the stalls listed in DESIGN.md (K8).
#ifdef XSUP_TRACE
#define TRACE(slot, gi) do { if (p.trace && blockIdx.x == 0 && (gi) >= 8 && (gi) < 24) p.trace[((gi) - 8) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot, gi) do { } while (0)
#endif
// ------------------------------------------------------------------ kernel
template <int KBN, int MODE>
__global__ void __launch_bounds__(kBwThreads, 1) conv_head_bwd_kernel(const __grid_constant__ CUtensorMap map_stat,
                                                                      const __grid_constant__ CUtensorMap map_str,
                                                                      const __grid_constant__ CUtensorMap map_out,
                                                                      const ConvBwdParams p) {
    constexpr int C = KBN * kCvKB;
    constexpr uint32_t kStageBytes = KBN * kBwKB;
    constexpr uint32_t kIdescS = umma_idesc_bf16(kBwM, kBwM);
    constexpr uint32_t kIdesc2 = umma_idesc_bf16(kBwM, C, false, true);       // B read MN-major
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sRing = smem;                                          // [stages][KBN][128 x 64] bf16
    uint8_t* sG = sRing + kBwStages * kStageBytes;                  // [2][128 x 64] bf16
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + kBwGBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    // ring: full[s] (TMA bytes landed) / empty[s] (the stationary copy, or the second GEMM of the tile, has retired)
    const uint32_t b_full = smem_u32(bars), b_empty = b_full + 8 * kBwStages, b_afull = b_empty + 8 * kBwStages, b_aempty = b_afull + 8,
                   b_gfull = b_aempty + 8, b_gempty = b_gfull + 8, b_dfull = b_gempty + 8, b_dempty = b_dfull + 8;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kBwStages; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        mbar_init(b_afull, 1);
        mbar_init(b_aempty, kBwEpiWarps);
        mbar_init(b_gfull, kBwEpiWarps);
        mbar_init(b_gempty, 1);
        mbar_init(b_dfull, 1);
        mbar_init(b_dempty, kBwEpiWarps);
        mbar_fence_init();
    }
    if (warp == kBwEpiWarps + 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int T = p.T;
    int n_items = 0;
    if ((int)blockIdx.x < p.items) n_items = (p.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

    if (warp == kBwEpiWarps) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t eph = 1;                                        // "empty" barriers: the first pass over the ring does not wait
            const uint32_t ring0 = smem_u32(sRing);
            for (int n = 0; n < n_items; ++n) {
                const int item = blockIdx.x + n * gridDim.x;
                const int b = item / p.per_b, j = item - b * p.per_b;
                // stationary: weight rows [128 j, +128) (MODE_DW) or pixels [b*HW + 128 j, +128) (MODE_DX)
                const int stat_row = MODE == MODE_DW ? j * kBwM : b * p.HW + j * kBwM;
                const int str_row0 = MODE == MODE_DW ? b * p.HW : 0;
                for (int t = -1; t < T; ++t) {
                    mbar_wait(b_empty + 8 * s, eph);
                    mbar_arrive_expect_tx(b_full + 8 * s, kStageBytes);
                    const uint32_t dst = ring0 + (uint32_t)s * kStageBytes;
                    if (t < 0) {
#pragma unroll
                        for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kBwKB, &map_stat, kb * kCvKB, stat_row, b_full + 8 * s);
                    } else {
#pragma unroll
                        for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kBwKB, &map_str, kb * kCvKB, str_row0 + t * kBwM, b_full + 8 * s);
                    }
                    if (++s == kBwStages) { s = 0; eph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kBwEpiWarps + 1) {
        // ------------------------------------------------------------ first-GEMM issuer: one thread, nothing but waits and issues
        if (lane == 0) {
            // descriptors: high word constant, low word = (address >> 4); K steps / stages / k-blocks add constants
            const uint64_t hiK = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t ring_lo0 = (smem_u32(sRing) & 0x3ffffu) >> 4;
            constexpr uint32_t kStageStep = kStageBytes >> 4, kKbStep = kBwKB >> 4;
            int s = 0, g = 0;
            uint32_t fph = 0, aeph = 1;
            for (int n = 0; n < n_items; ++n) {
                // new item: stationary operand from its ring stage into TMEM (in order behind the previous item's MMAs), stage released
                mbar_wait(b_full + 8 * s, fph);
                tc_fence_after();
                {
                    const uint32_t lo = ring_lo0 + (uint32_t)s * kStageStep;
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(kBwACol + (kb * 4 + q) * 8), hiK | (uint64_t)(lo + kb * kKbStep + 2 * q));
                    umma_commit(b_empty + 8 * s);
                    if (++s == kBwStages) { s = 0; fph ^= 1; }
                }
                for (int t = 0; t < T; ++t, ++g) {
                    TRACE(0, g);
                    mbar_wait(b_aempty, aeph);                       // the epilogue has read the previous S
                    aeph ^= 1;
                    TRACE(1, g);
                    mbar_wait(b_full + 8 * s, fph);
                    TRACE(2, g);
                    tc_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)kBwSCol;
                    const uint32_t lo = ring_lo0 + (uint32_t)s * kStageStep;
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            umma_f16_ts(acc, tmem_base + (uint32_t)(kBwACol + (kb * 4 + q) * 8), hiK | (uint64_t)(lo + kb * kKbStep + 2 * q), kIdescS,
                                        (kb | q) ? 1u : 0u);
                    }
                    umma_commit(b_afull);
                    TRACE(3, g);
                    if (++s == kBwStages) { s = 0; fph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kBwEpiWarps + 2) {
        // ------------------------------------------------------------ second-GEMM issuer: D2 += G (smem, K-major) * streamed (MN-major), K = 128 rows
        if (lane == 0) {
            const uint64_t hiK = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t mn_lo0 = ((smem_u32(sRing) & 0x3ffffu) >> 4) | ((uint32_t)(kBwKB >> 4) << 16);   // + leading byte offset = next 64 channels
            const uint32_t g_lo = (smem_u32(sG) & 0x3ffffu) >> 4;
            constexpr uint32_t kStageStep = kStageBytes >> 4;
            int s = 0, h = 0;
            uint32_t gph = 0, dph = 0;
            for (int n = 0; n < n_items; ++n) {
                if (++s == kBwStages) s = 0;                                      // the item's stationary entry is not ours
                if (n >= 1) { mbar_wait(b_dempty, dph); dph ^= 1; }               // the epilogue has drained the previous item's D2
                for (int t = 0; t < T; ++t, ++h) {
                    mbar_wait(b_gfull, gph);
                    gph ^= 1;
                    TRACE(4, h);
                    tc_fence_after();
                    const uint32_t d2 = tmem_base + (uint32_t)kBwD2Col;
                    const uint64_t db = hiK | (uint64_t)(mn_lo0 + (uint32_t)s * kStageStep);
                    const int nq = t == T - 1 ? (p.last_rows + 15) >> 4 : 8;       // K steps of 16 streamed rows; padding rows are skipped
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q < nq)
                            umma_f16_rt(d2, hiK | (uint64_t)(g_lo + (q >> 2) * (kBwKB >> 4) + 2 * (q & 3)), db + (uint64_t)(128 * q), kIdesc2,
                                        (t | q) ? 1u : 0u);
                    umma_commit(b_empty + 8 * s);
                    umma_commit(b_gempty);
                    TRACE(5, h);
                    if (++s == kBwStages) s = 0;
                }
                umma_commit(b_dfull);
            }
        }
        __syncwarp();
    } else if (warp < kBwEpiWarps) {
        // ------------------------------------------------------------ epilogue: thread = TMEM lane, 32 columns per warp and tile
        const int quarter = warp & 3, part = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // G tile: row = 128 bytes per k-block; this warp's 32 K-elements = chunks (part & 1) * 4 .. + 3 of k-block part >> 1
        const uint32_t g_row = smem_u32(sG) + (uint32_t)(part >> 1) * kBwKB + (uint32_t)row * 128u;
        const uint32_t sw = (uint32_t)(row & 7), ch0 = (uint32_t)(part & 1) * 4u;
        const int Wd = p.W;
        uint32_t aph = 0, geph = 1, dph = 0;
        int gtile = 0;
        for (int n = 0; n < n_items; ++n) {
            const int item = blockIdx.x + n * gridDim.x;
            const int b = item / p.per_b, j = item - b * p.per_b;
            // MODE_DW: (nlse, a, b, e) of this thread's logit row, fixed for the item.  MODE_DX: thread = pixel, columns = logit rows:
            // lane l fetches the coefficients of column l of the NEXT tile while this one is processed (one tile ahead), and the
            // columns' (nlse, e) reach all lanes by shuffle; (a, b) belong to the joint, which is the same for a warp's 32 columns.
            float4 rc = make_float4(0.f, 0.f, 0.f, 0.f);
            f32x2 gsum2 = pk2(0.f, 0.f);
            float fw = 0.f, fh = 0.f;
            const float* tabg = nullptr;
            if (MODE == MODE_DW) {
                const int r = j * kBwM + row;
                const float* q = p.rowcoef + ((size_t)b * p.rows_pad + (r & ~1)) * 4 + (r & 1);
                rc = make_float4(q[0], q[2], q[4], q[6]);
            } else {
                const int pix = j * kBwM + row;
                const int hh = pix / Wd;
                fh = (float)hh;
                fw = (float)(pix - hh * Wd);
                const int r = part * 32 + lane;                       // this lane's column of tile 0
                tabg = p.rowcoef + ((size_t)b * p.rows_pad + (r & ~1)) * 4 + (r & 1);
                rc = make_float4(tabg[0], tabg[2], tabg[4], tabg[6]);
            }
            const f32x2 l2e2 = pk2(kLog2e, kLog2e);
            for (int t = 0; t < T; ++t) {
                float4 rc_next = rc;
                if (MODE == MODE_DX && t + 1 < T) {
                    const float* q = tabg + (size_t)(t + 1) * kBwM * 4;
                    rc_next = make_float4(__ldg(q), __ldg(q + 2), __ldg(q + 4), __ldg(q + 6));
                }
                mbar_wait(b_afull, aph);
                aph ^= 1;
                if (threadIdx.x == 0) TRACE(8, gtile);
                tc_fence_after();
                if (MODE == MODE_DX && t == T - 1 && part * 32 >= p.last_rows) {
                    // this warp's 32 columns of the last weight tile are all padding: the second GEMM does not read them
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_aempty);
                    mbar_wait(b_gempty, geph);
                    geph ^= 1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_gfull);
                    ++gtile;
                    continue;
                }
                uint32_t r[32];
                tmem_ld32(lane_addr + (uint32_t)(kBwSCol + part * 32), r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_aempty);                // values are in registers: the next S may overwrite the accumulator
                if (threadIdx.x == 0) TRACE(9, gtile);
                uint32_t o[16];
                if (MODE == MODE_DW) {
                    // columns = 32 consecutive pixels of one image row (W % 32 == 0); packed fp32x2 arithmetic, two columns per op
                    const int pix = t * kBwM + part * 32;
                    const int hh = pix / Wd;
                    const float rowterm = fmaf(rc.y, (float)(pix - hh * Wd), fmaf(rc.z, (float)hh, rc.w));
                    f32x2 lin = pk2(rowterm, rowterm + rc.y);
                    const f32x2 step = pk2(2.f * rc.y, 2.f * rc.y), nl2 = pk2(rc.x, rc.x);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const f32x2 v = fmul2(ex2_2(ffma2(pk2u(r[2 * i], r[2 * i + 1]), l2e2, nl2)), lin);
                        gsum2 = fadd2(gsum2, v);
                        lin = fadd2(lin, step);
                        float v0, v1;
                        upk2(v, v0, v1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i]) : "f"(v1), "f"(v0));
                    }
                } else {
                    const float linw = fmaf(rc.y, fw, rc.z * fh);    // a_k w + b_k h of this pixel (one joint per 32 columns: 32 | D)
                    const f32x2 lin2 = pk2(linw, linw);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const f32x2 nl2 = pk2(__shfl_sync(0xffffffffu, rc.x, 2 * i), __shfl_sync(0xffffffffu, rc.x, 2 * i + 1));
                        const f32x2 e2 = pk2(__shfl_sync(0xffffffffu, rc.w, 2 * i), __shfl_sync(0xffffffffu, rc.w, 2 * i + 1));
                        const f32x2 v = fmul2(ex2_2(ffma2(pk2u(r[2 * i], r[2 * i + 1]), l2e2, nl2)), fadd2(e2, lin2));
                        float v0, v1;
                        upk2(v, v0, v1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i]) : "f"(v1), "f"(v0));
                    }
                    rc = rc_next;
                }
                // publish this warp's [32 lanes x 32 K-elements] of the G tile: 16-byte chunk c of a row sits at (c ^ (row & 7))
                if (threadIdx.x == 0) TRACE(10, gtile);
                mbar_wait(b_gempty, geph);                            // the second GEMM of the previous tile has read G
                geph ^= 1;
                if (threadIdx.x == 0) TRACE(11, gtile);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(g_row + (((ch0 + c) ^ sw) << 4)), "r"(o[4 * c]), "r"(o[4 * c + 1]),
                                 "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_gfull);
                if (threadIdx.x == 0) TRACE(12, gtile);
                ++gtile;
            }
            // ---- end of the item: drain D2 (this warp: its lane quarter, C/4 columns)
            if (threadIdx.x == 0) TRACE(13, 8 + n);
            mbar_wait(b_dfull, dph);
            dph ^= 1;
            if (threadIdx.x == 0) TRACE(14, 8 + n);
            tc_fence_after();
            // D2 [128 x C] fp32 leaves through the (now idle) G buffer: 32 rows at a time are laid out as [32 x 128-byte] blocks with
            // the 128-byte swizzle and handed to the TMA (tensor store for d x, fp32 reduce-add for d W), so that every row is
            // written as whole 128-byte lines.  (Storing straight from the TMEM lanes - one row per thread - cost ~7000 cycles per
            // item: 32 different lines per store instruction.)
            constexpr int CW = C / kBwParts;                           // columns per warp: 16, 32, 48 or 64
            constexpr int NB = C / 64;                                 // [32 rows x 128 B] blocks per 16 KB chunk
            const bool f32o = MODE == MODE_DW || p.dx_f32;
            // chunks of <= 16 KB ping-pong between the two halves of the buffer, so a chunk is staged while the TMA still reads
            // the previous one: bf16 - the 32 rows of one lane quarter, all columns; fp32 - 32 rows x half the columns
            const int n_chunks = f32o ? 8 : 4;
            const int out_row0 = (MODE == MODE_DW ? j * kBwM : b * p.HW + j * kBwM) + quarter * 32;
            const uint32_t stg = smem_u32(sG);
            for (int ch = 0; ch < n_chunks; ++ch) {
                const uint32_t half = stg + (uint32_t)(ch & 1) * (uint32_t)(kBwGBytes / 2);
                if (ch >= 2) {
                    if (threadIdx.x == 0) bulk_wait_read1();             // the chunk that used this half two steps ago has been read
                    asm volatile("bar.sync 1, %0;" ::"n"(kBwEpiWarps * 32) : "memory");
                }
                const bool active = f32o ? (quarter == (ch >> 1) && (part >> 1) == (ch & 1)) : (quarter == ch);
                if (active) {
                    const uint32_t rbase = half + (uint32_t)lane * 128u, rsw = (uint32_t)lane & 7u;
#pragma unroll
                    for (int q = 0; q < CW / 16; ++q) {
                        uint32_t r[16];
                        tmem_ld16(lane_addr + (uint32_t)(kBwD2Col + part * CW + q * 16), r);
                        if (f32o) {
                            const uint32_t off = (uint32_t)((part & 1) * CW + q * 16) * 4u, blk = off >> 7, c0 = (off & 127u) >> 4;
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rbase + blk * 4096u + (((c0 + c) ^ rsw) << 4)), "r"(r[4 * c]),
                                             "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3]) : "memory");
                        } else {
                            uint32_t ob[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 2)
                                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ob[i >> 1]) : "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i])));
                            const uint32_t off = (uint32_t)(part * CW + q * 16) * 2u, blk = off >> 7, c0 = (off & 127u) >> 4;
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rbase + blk * 4096u + (((c0 + c) ^ rsw) << 4)), "r"(ob[4 * c]),
                                             "r"(ob[4 * c + 1]), "r"(ob[4 * c + 2]), "r"(ob[4 * c + 3]) : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kBwEpiWarps * 32) : "memory");
                if (threadIdx.x == 0) {
                    // thread 0 sits in quarter 0: the chunk's rows / columns follow from ch, not from this thread's own quarter
                    const int crow = out_row0 - quarter * 32 + (f32o ? (ch >> 1) : ch) * 32;
                    const int ccol0 = f32o ? (ch & 1) * (C / 2) : 0, bcols = f32o ? 32 : 64;
                    for (int blk = 0; blk < NB; ++blk) {
                        if (MODE == MODE_DW) tma_reduce_add_2d(&map_out, half + (uint32_t)blk * 4096u, ccol0 + blk * bcols, crow);
                        else tma_store_2d(&map_out, half + (uint32_t)blk * 4096u, ccol0 + blk * bcols, crow);
                    }
                    bulk_commit();
                }
            }
            if (threadIdx.x == 0) bulk_wait_read0();                   // the buffer returns to the G tiles of the next item
            asm volatile("bar.sync 1, %0;" ::"n"(kBwEpiWarps * 32) : "memory");
            if (MODE == MODE_DW) {
                const int grow_i = j * kBwM + row;
                if (p.dbias && grow_i < p.rows_total) {
                    float g0, g1;
                    upk2(gsum2, g0, g1);
                    atomicAdd(p.dbias + grow_i, g0 + g1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_dempty);
            if (threadIdx.x == 0) TRACE(15, 8 + n);
        }
        if (threadIdx.x == 0) bulk_wait0();                            // all tensor stores / reductions of this CTA have been performed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBwEpiWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------ host side
template <int KBN, int MODE>
static cudaError_t launch_bwd_kbn(const CUtensorMap& map_stat, const CUtensorMap& map_str, const CUtensorMap& map_out, const ConvBwdParams& p, int grid,
                                  cudaStream_t st) {
    auto kern = conv_head_bwd_kernel<KBN, MODE>;
    static unsigned long long attr_done = 0;             // per instantiation; one bit per device
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    const size_t smem = 1024 + (size_t)kBwStages * KBN * kBwKB + kBwGBytes + 24 * 8;
    kern<<<grid, kBwThreads, smem, st>>>(map_stat, map_str, map_out, p);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_bwd_mode(const CUtensorMap& map_stat, const CUtensorMap& map_str, const CUtensorMap& map_out, const ConvBwdParams& p, int num_sms,
                                   cudaStream_t st) {
    const int grid = p.items < num_sms ? p.items : num_sms;
    switch (p.C / kCvKB) {
        case 1: return launch_bwd_kbn<1, MODE>(map_stat, map_str, map_out, p, grid, st);
        case 2: return launch_bwd_kbn<2, MODE>(map_stat, map_str, map_out, p, grid, st);
        case 3: return launch_bwd_kbn<3, MODE>(map_stat, map_str, map_out, p, grid, st);
        default: return launch_bwd_kbn<4, MODE>(map_stat, map_str, map_out, p, grid, st);
    }
}

int conv_bwd_rows_pad(int K, int D) { return (K * D + kBwM - 1) / kBwM * kBwM; }

// d x (bf16 or fp32, channels-last), d W (fp32), d bias (fp32) of the conv-fused head from the coefficient blocks of
// integral_coef_kernel; `rowcoef_ws` holds B * conv_bwd_rows_pad(K, D) float4.  Any of dx / dw may be NULL (that launch is skipped).
cudaError_t launch_conv_head_bwd(const void* x_nhwc, const void* w, const float* bias, const float* coef, int coef_stride, float* rowcoef_ws,
                                 void* dx, int dx_f32, float* dw, float* dbias, int B, int K, int D, int H, int W, int C, int num_sms,
                                 cudaStream_t st) {
    ConvBwdParams p{};
#ifdef XSUP_TRACE
    if (const char* d = getenv("XSUP_CONVBWD_TRACE")) p.trace = reinterpret_cast<long long*>(strtoull(d, nullptr, 0));
#endif
    p.rowcoef = rowcoef_ws;
    p.C = C; p.HW = H * W; p.W = W; p.rows_total = K * D; p.rows_pad = conv_bwd_rows_pad(K, D);
    const long long n = (long long)B * p.rows_pad;
    conv_rowcoef_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(coef, coef_stride, bias, rowcoef_ws, B, K, D, p.rows_pad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    CUtensorMap map_w, map_x, map_o;
    if (!make_map(&map_w, w, p.rows_total, C, kBwM) || !make_map(&map_x, x_nhwc, (long long)B * p.HW, C, kBwM)) return cudaErrorNotSupported;
    if (dw) {
        e = cudaMemsetAsync(dw, 0, (size_t)p.rows_total * C * sizeof(float), st);
        if (e != cudaSuccess) return e;
        if (dbias) {
            e = cudaMemsetAsync(dbias, 0, (size_t)p.rows_total * sizeof(float), st);
            if (e != cudaSuccess) return e;
        }
        ConvBwdParams q = p;
        q.dw = dw; q.dbias = dbias;
        q.per_b = p.rows_pad / kBwM;
        q.T = p.HW / kBwM;
        q.last_rows = kBwM;
        q.items = B * q.per_b;
        if (!make_map_out(&map_o, dw, p.rows_total, C, true, 32)) return cudaErrorNotSupported;
        e = launch_bwd_mode<MODE_DW>(map_w, map_x, map_o, q, num_sms, st);
        if (e != cudaSuccess) return e;
    }
    if (dx) {
        ConvBwdParams q = p;
        q.dx = dx; q.dx_f32 = dx_f32;
        q.per_b = p.HW / kBwM;
        q.T = p.rows_pad / kBwM;
        q.last_rows = p.rows_total - (q.T - 1) * kBwM;
        q.items = B * q.per_b;
        if (!make_map_out(&map_o, dx, (long long)B * p.HW, C, dx_f32 != 0, 32)) return cudaErrorNotSupported;
        e = launch_bwd_mode<MODE_DX>(map_x, map_w, map_o, q, num_sms, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace xsup
