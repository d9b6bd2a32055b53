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
//   MODE_DW (W-stationary): item = (sample b, 128 weight rows).  The 128 x C weight slab stays in shared memory, the sample's
//            activations stream through as 64-pixel tiles:   S[128 rows x 64 pix] = Wt_slab * X_tile^T ;  G = f(S) ;
//            D2[128 rows x C] += G[128 x 64] * X_tile[64 x C].  D2 is flushed once per item with fp32 RED.ADD into d Wt.
//   MODE_DX (X-stationary): item = (sample b, 128 pixels).  The 128 x C activation tile stays, the weights stream through as
//            64-row tiles:   S[128 pix x 64 rows] = X_tile * Wt_tile^T ;  G^T = f(S) ;  D2[128 pix x C] += G^T[128 x 64] * Wt_tile[64 x C].
//            D2 is the finished d X tile (all K*D rows were summed in TMEM) and is written once, bf16 or fp32, channels-last.
//
// Both modes are ONE kernel: "stationary" operand [128 x C] (A of the first GEMM, K-major), "streamed" tiles [64 x C] (B of
// the first GEMM, K-major; B of the second GEMM read MN-major from the very same bytes), G [128 x 64] bf16 = A of the second
// GEMM, which never leaves the tensor memory: the epilogue writes it with tcgen05.st next to the accumulators and the
// second GEMM reads its A operand from TMEM (the way attention kernels feed P into P*V).  Warp roles (18 warps, one
// persistent CTA per SM):
//   warp 16      TMA producer: stationary operand per item; streamed tiles through a 4-stage ring (+ in MODE_DX the 64 row
//                coefficients of the tile, one 1 KB bulk copy on the same barrier)
//   warp 17      first-GEMM issuer (ONE thread; descriptors precomputed, counters instead of divisions): per item 4*C/64
//                tcgen05.cp (stationary operand -> TMEM), per tile S = 4*C/64 MMAs M=128 N=64 K=16, A from TMEM
//   warp 18      second-GEMM issuer (one thread): 4 MMAs M=128 N=C K=16 per tile, A = G from TMEM, as soon as the epilogue has
//                published the tile's G.  Two issuers because the tensor-pipe queue is shallow - an issuing thread is held
//                while its MMAs execute, and with a single issuer every barrier wait idled the pipe (measured: 2070 -> 1700
//                cycles per tile); tcgen05.commit releases ring stage (and with it the G buffer) / accumulator
//   warps 0..15  epilogue: tcgen05.ld (thread = TMEM lane, 16 columns per warp), G, tcgen05.st (8 packed columns), mbarrier
//                arrive; at the end of an item they drain D2
// TMEM: columns [0,128) = the stationary operand (copied from its shared-memory landing buffer with tcgen05.cp once per item,
// so the first GEMM reads only its B operand from shared memory and the landing buffer is free to prefetch the next item),
// [128,256) = 2 x S[128 x 64] fp32 - the epilogue writes G[128 x 64] bf16 (32 columns) over the head of the accumulator it
// has just read, and the buffer returns to the first GEMM when the second GEMM that read G has retired -, [256, 256+C) = D2.
// Flops: 2 GEMMs of 2*K*D*C*H*W per sample and launch (1.17 TFLOP per launch at B=256, K=17, D=64, C=256).
#include <stdlib.h>

#include "xsup_internal.h"
#include "xsup_umma.cuh"

namespace xsup {

constexpr int kBwEpiWarps = 16;
constexpr int kBwParts = kBwEpiWarps / 4;                  // warps sharing a TMEM lane quarter: 16 of a tile's 64 columns each
constexpr int kBwThreads = (kBwEpiWarps + 3) * 32;             // + TMA producer, first-GEMM issuer, second-GEMM issuer
constexpr int kBwM = 128;                                  // stationary rows = UMMA M
constexpr int kBwN = 64;                                   // streamed rows per tile = N of the first GEMM, K of the second
constexpr int kBwStages = 4;
constexpr int kBwSAcc = 2;
constexpr int kBwStatKB = kBwM * kCvKB * 2;                // 16 KB: one [128 x 64] bf16 k-block of the stationary operand
constexpr int kBwStrKB = kBwN * kCvKB * 2;                 //  8 KB: one [64 x 64] bf16 k-block of a streamed tile
constexpr int kBwTabBytes = kBwN * 16;                     //  1 KB: row coefficients of a streamed weight tile (MODE_DX)
constexpr int kBwACol = 0;                                 // TMEM columns [0,128): the stationary operand, bf16, 8 columns per 16 channels
constexpr int kBwSCol = 128;                               // two S accumulators [128 x 64] fp32; G (bf16, 32 columns) overwrites the head of its own S
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
    int items;
    long long* trace;           // diagnostics: clock64 timeline of CTA 0, tiles 8..23 (XSUP_CONVBWD_TRACE = device pointer)
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

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (bf16, K-major, 2 elements per 32-bit column) comes from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// shared memory -> tensor memory, 128 lanes x 256 bits: one [128 rows x 16 bf16] K-slab of a K-major operand (same descriptor
// as the MMA would use for it); executes in issue order with the MMAs of the issuing thread
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

#define TRACE(slot, gi) do { if (p.trace && blockIdx.x == 0 && (gi) >= 8 && (gi) < 24) p.trace[((gi) - 8) * 16 + (slot)] = clock64(); } while (0)
// ------------------------------------------------------------------ kernel
template <int KBN, int MODE>
__global__ void __launch_bounds__(kBwThreads, 1) conv_head_bwd_kernel(const __grid_constant__ CUtensorMap map_stat,
                                                                      const __grid_constant__ CUtensorMap map_str,
                                                                      const ConvBwdParams p) {
    constexpr int C = KBN * kCvKB;
    constexpr uint32_t kStatBytes = KBN * kBwStatKB, kStageBytes = KBN * kBwStrKB;
    constexpr uint32_t kIdescS = umma_idesc_bf16(kBwM, kBwN);
    constexpr uint32_t kIdesc2 = umma_idesc_bf16(kBwM, C, false, true);       // B read MN-major
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sStat = smem;                                          // [KBN][128 x 64] bf16
    uint8_t* sStr = sStat + kStatBytes;                             // [stages][KBN][64 x 64] bf16
    uint8_t* sTab = sStr + kBwStages * kStageBytes;                 // [stages][64] float4
    uint64_t* bars = reinterpret_cast<uint64_t*>(sTab + kBwStages * kBwTabBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    // x_empty[s] completes when the second GEMM of the tile in stage s has retired: it frees the stage AND that tile's G buffer
    const uint32_t b_sfull = smem_u32(bars), b_sempty = b_sfull + 8, b_xfull = b_sempty + 8, b_xempty = b_xfull + 8 * kBwStages,
                   b_afull = b_xempty + 8 * kBwStages, b_gfull = b_afull + 8 * kBwSAcc, b_dfull = b_gfull + 16, b_dempty = b_dfull + 8;

    if (threadIdx.x == 0) {
        mbar_init(b_sfull, 1);
        mbar_init(b_sempty, 1);
        for (int i = 0; i < kBwStages; ++i) {
            mbar_init(b_xfull + 8 * i, 1);
            mbar_init(b_xempty + 8 * i, 1);
        }
        for (int i = 0; i < kBwSAcc; ++i) mbar_init(b_afull + 8 * i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(b_gfull + 8 * i, kBwEpiWarps);
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
            uint32_t xph = 1, sph = 1;                               // "empty" barriers: the first pass over the ring does not wait
            const uint32_t str0 = smem_u32(sStr), stat0 = smem_u32(sStat), tab0 = smem_u32(sTab);
            for (int n = 0; n < n_items; ++n) {
                const int item = blockIdx.x + n * gridDim.x;
                const int b = item / p.per_b, j = item - b * p.per_b;
                // stationary: weight rows [128 j, +128) (MODE_DW) or pixels [b*HW + 128 j, +128) (MODE_DX)
                const int stat_row = MODE == MODE_DW ? j * kBwM : b * p.HW + j * kBwM;
                const int str_row0 = MODE == MODE_DW ? b * p.HW : 0;
                const float* tab_src = p.rowcoef + (size_t)b * p.rows_pad * 4;
                mbar_wait(b_sempty, sph);                            // the previous item's first-GEMM MMAs have retired
                sph ^= 1;
                mbar_arrive_expect_tx(b_sfull, kStatBytes);
#pragma unroll
                for (int kb = 0; kb < KBN; ++kb) tma_load_2d(stat0 + kb * kBwStatKB, &map_stat, kb * kCvKB, stat_row, b_sfull);
                for (int t = 0; t < T; ++t) {
                    mbar_wait(b_xempty + 8 * s, xph);                // the second GEMM of the tile that used this stage has retired
                    TRACE(6, n * T + t);
                    mbar_arrive_expect_tx(b_xfull + 8 * s, kStageBytes + (MODE == MODE_DX ? (uint32_t)kBwTabBytes : 0u));
                    const uint32_t dst = str0 + (uint32_t)s * kStageBytes;
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kBwStrKB, &map_str, kb * kCvKB, str_row0 + t * kBwN, b_xfull + 8 * s);
                    if (MODE == MODE_DX) bulk_g2s(tab0 + (uint32_t)s * kBwTabBytes, tab_src + (size_t)t * kBwN * 4, kBwTabBytes, b_xfull + 8 * s);
                    if (++s == kBwStages) { s = 0; xph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kBwEpiWarps + 1) {
        // ------------------------------------------------------------ first-GEMM issuer: one thread, nothing but waits and issues
        // (the tensor-pipe queue is shallow: an issuing thread is held while its MMAs execute, so the two GEMM chains are issued
        //  by two threads - while one sits in its barrier waits the other one's MMAs keep the pipe busy)
        if (lane == 0) {
            const int total = n_items * T;
            // descriptors: high words are constant, low words = (address >> 4); K steps / stages / k-blocks add constants
            const uint64_t hiK = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint32_t stat_lo[KBN];
#pragma unroll
            for (int kb = 0; kb < KBN; ++kb) stat_lo[kb] = ((smem_u32(sStat) + kb * kBwStatKB) & 0x3ffffu) >> 4;
            const uint32_t str_lo0 = (smem_u32(sStr) & 0x3ffffu) >> 4;
            constexpr uint32_t kStageStep = kStageBytes >> 4, kKbStep = kBwStrKB >> 4;
            int t1 = 0, s1 = 0, a1 = 0;
            uint32_t xph1 = 0, sph1 = 0;
            // accumulator a1 held S and G of tile g-2: free once that tile's second GEMM has retired = x_empty of stage (g-2) & 3
            int s_prev2 = kBwStages - 2;
            uint32_t eph_prev2 = 1;
            for (int g = 0; g < total; ++g) {
                // ---- S = stationary * streamed^T, K = C
                TRACE(0, g);
                if (t1 == 0) {
                    // new item: stationary operand from its landing buffer into TMEM (in order behind the previous item's MMAs);
                    // the landing buffer is then free for the producer to prefetch the next item
                    mbar_wait(b_sfull, sph1);
                    sph1 ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(kBwACol + (kb * 4 + q) * 8), hiK | (uint64_t)(stat_lo[kb] + 2 * q));
                    umma_commit(b_sempty);
                }
                mbar_wait(b_xempty + 8 * s_prev2, eph_prev2);
                TRACE(1, g);
                mbar_wait(b_xfull + 8 * s1, xph1);
                TRACE(2, g);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(kBwSCol + a1 * kBwN);
                const uint32_t b_lo = str_lo0 + (uint32_t)s1 * kStageStep;
#pragma unroll
                for (int kb = 0; kb < KBN; ++kb) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        umma_f16_ts(acc, tmem_base + (uint32_t)(kBwACol + (kb * 4 + q) * 8), hiK | (uint64_t)(b_lo + kb * kKbStep + 2 * q), kIdescS,
                                    (kb | q) ? 1u : 0u);
                }
                umma_commit(b_afull + 8 * a1);
                TRACE(3, g);
                if (++t1 == T) t1 = 0;
                if (++s1 == kBwStages) { s1 = 0; xph1 ^= 1; }
                a1 ^= 1;
                if (++s_prev2 == kBwStages) { s_prev2 = 0; eph_prev2 ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == kBwEpiWarps + 2) {
        // ------------------------------------------------------------ second-GEMM issuer: D2 += G (TMEM) * streamed (MN-major), K = 64 rows
        if (lane == 0) {
            const int total = n_items * T;
            const uint64_t hiK = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t mn_lo0 = ((smem_u32(sStr) & 0x3ffffu) >> 4) | ((uint32_t)(kBwStrKB >> 4) << 16);   // + leading byte offset = next 64 channels
            constexpr uint32_t kStageStep = kStageBytes >> 4;
            int t2 = 0, s2 = 0, n2 = 0;
            uint32_t gb2 = 0, gph2 = 0, dph2 = 0;
            for (int h = 0; h < total; ++h) {
                if (t2 == 0 && n2 >= 1) { mbar_wait(b_dempty, dph2); dph2 ^= 1; }   // the epilogue has drained the previous item's D2
                mbar_wait(b_gfull + 8 * gb2, gph2);
                TRACE(4, h);
                tc_fence_after();
                const uint32_t d2 = tmem_base + (uint32_t)kBwD2Col;
                const uint32_t ga = tmem_base + (uint32_t)kBwSCol + gb2 * (uint32_t)kBwN;
                const uint64_t db = hiK | (uint64_t)(mn_lo0 + (uint32_t)s2 * kStageStep);
#pragma unroll
                for (int q = 0; q < 4; ++q) umma_f16_ts(d2, ga + 8u * q, db + (uint64_t)(128 * q), kIdesc2, (t2 | q) ? 1u : 0u);
                umma_commit(b_xempty + 8 * s2);
                TRACE(5, h);
                if (++t2 == T) { t2 = 0; ++n2; umma_commit(b_dfull); }
                if (++s2 == kBwStages) s2 = 0;
                if (gb2) gph2 ^= 1;
                gb2 ^= 1;
            }
        }
        __syncwarp();
    } else if (warp < kBwEpiWarps) {
        // ------------------------------------------------------------ epilogue: thread = TMEM lane, 16 columns per warp and tile
        const int quarter = warp & 3, part = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t tab0 = smem_u32(sTab) + (uint32_t)part * 256u;
        const int Wd = p.W;
        int a = 0, s = 0;
        uint32_t aph = 0, xph = 0;
        int gtile = 0;
        uint32_t dph = 0;
        for (int n = 0; n < n_items; ++n) {
            const int item = blockIdx.x + n * gridDim.x;
            const int b = item / p.per_b, j = item - b * p.per_b;
            float4 rc = make_float4(0.f, 0.f, 0.f, 0.f);            // (nlse, a, b, e) of this thread's row (MODE_DW)
            f32x2 fw2 = pk2(0.f, 0.f), fh2 = fw2, gsum2 = fw2;
            if (MODE == MODE_DW) {
                const int r = j * kBwM + row;
                const float* q = p.rowcoef + ((size_t)b * p.rows_pad + (r & ~1)) * 4 + (r & 1);
                rc = make_float4(q[0], q[2], q[4], q[6]);
            } else {
                const int pix = j * kBwM + row;
                const int hh = pix / Wd;
                fh2 = pk2((float)hh, (float)hh);
                fw2 = pk2((float)(pix - hh * Wd), (float)(pix - hh * Wd));
            }
            const f32x2 l2e2 = pk2(kLog2e, kLog2e);
            for (int t = 0; t < T; ++t) {
                mbar_wait(b_afull + 8 * a, aph);
                if (threadIdx.x == 0) TRACE(8, gtile);
                tc_fence_after();
                uint32_t r[16];
                tmem_ld16(lane_addr + (uint32_t)(kBwSCol + a * kBwN + part * 16), r);
                if (threadIdx.x == 0) TRACE(9, gtile);
                uint32_t o[8];
                if (MODE == MODE_DW) {
                    // columns = 16 consecutive pixels of one image row (W % 16 == 0); packed fp32x2 arithmetic, two columns per op
                    const int pix = t * kBwN + part * 16;
                    const int hh = pix / Wd;
                    const float rowterm = fmaf(rc.y, (float)(pix - hh * Wd), fmaf(rc.z, (float)hh, rc.w));
                    f32x2 lin = pk2(rowterm, rowterm + rc.y);
                    const f32x2 step = pk2(2.f * rc.y, 2.f * rc.y), nl2 = pk2(rc.x, rc.x);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const f32x2 v = fmul2(ex2_2(ffma2(pk2u(r[2 * i], r[2 * i + 1]), l2e2, nl2)), lin);
                        gsum2 = fadd2(gsum2, v);
                        lin = fadd2(lin, step);
                        float v0, v1;
                        upk2(v, v0, v1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i]) : "f"(v1), "f"(v0));
                    }
                } else {
                    // columns = 16 consecutive logit rows of the streamed weight tile: coefficients from the stage's table
                    mbar_wait(b_xfull + 8 * s, xph);                  // (complete already: the MMAs read this stage) acquires the table
                    const uint32_t tab = tab0 + (uint32_t)s * kBwTabBytes;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 q0 = lds_f4(tab + 32u * i), q1 = lds_f4(tab + 32u * i + 16u);   // {nlse, nlse', a, a'}, {b, b', e, e'}
                        const f32x2 lin = ffma2(pk2(q0.z, q0.w), fw2, ffma2(pk2(q1.x, q1.y), fh2, pk2(q1.z, q1.w)));
                        const f32x2 v = fmul2(ex2_2(ffma2(pk2u(r[2 * i], r[2 * i + 1]), l2e2, pk2(q0.x, q0.y))), lin);
                        float v0, v1;
                        upk2(v, v0, v1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i]) : "f"(v1), "f"(v0));
                    }
                }
                // publish this warp's [32 lanes x 16 K-elements] of the G tile: 8 packed columns over the head of the accumulator,
                // once the four warps of this lane quarter have all read their columns of it
                if (threadIdx.x == 0) TRACE(10, gtile);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
                if (threadIdx.x == 0) TRACE(11, gtile);
                tmem_st8(lane_addr + (uint32_t)(kBwSCol + a * kBwN + part * 8), o);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_gfull + 8 * a);
                if (threadIdx.x == 0) TRACE(12, gtile);
                ++gtile;
                if (++a == kBwSAcc) { a = 0; aph ^= 1; }
                if (++s == kBwStages) { s = 0; xph ^= 1; }
            }
            // ---- end of the item: drain D2 (this warp: its lane quarter, C/4 columns)
            mbar_wait(b_dfull, dph);
            dph ^= 1;
            tc_fence_after();
            constexpr int CW = C / kBwParts;                           // columns per warp: 16, 32, 48 or 64
            if (MODE == MODE_DW) {
                const int grow_i = j * kBwM + row;
                float* dst = p.dw + (size_t)grow_i * C + part * CW;
#pragma unroll
                for (int q = 0; q < CW / 16; ++q) {
                    uint32_t r[16];
                    tmem_ld16(lane_addr + (uint32_t)(kBwD2Col + part * CW + q * 16), r);
                    if (grow_i < p.rows_total) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            red_add_v4(dst + q * 16 + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                       __uint_as_float(r[i + 3]));
                    }
                }
                if (p.dbias && grow_i < p.rows_total) {
                    float g0, g1;
                    upk2(gsum2, g0, g1);
                    atomicAdd(p.dbias + grow_i, g0 + g1);
                }
            } else {
                const size_t pix = (size_t)b * p.HW + (size_t)j * kBwM + row;
#pragma unroll
                for (int q = 0; q < CW / 16; ++q) {
                    uint32_t r[16];
                    tmem_ld16(lane_addr + (uint32_t)(kBwD2Col + part * CW + q * 16), r);
                    if (p.dx_f32) {
                        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.dx) + pix * C + part * CW + q * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                                 __uint_as_float(r[4 * i + 3]));
                    } else {
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 16; i += 2)
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i >> 1]) : "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i])));
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.dx) + pix * C + part * CW + q * 16);
                        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_dempty);
        }
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
static cudaError_t launch_bwd_kbn(const CUtensorMap& map_stat, const CUtensorMap& map_str, const ConvBwdParams& p, int grid, cudaStream_t st) {
    auto kern = conv_head_bwd_kernel<KBN, MODE>;
    static unsigned long long attr_done = 0;             // per instantiation; one bit per device
    cudaError_t e = ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return e;
    const size_t smem = 1024 + (size_t)KBN * kBwStatKB + (size_t)kBwStages * KBN * kBwStrKB + kBwStages * kBwTabBytes + 32 * 8;
    kern<<<grid, kBwThreads, smem, st>>>(map_stat, map_str, p);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_bwd_mode(const CUtensorMap& map_stat, const CUtensorMap& map_str, const ConvBwdParams& p, int num_sms, cudaStream_t st) {
    const int grid = p.items < num_sms ? p.items : num_sms;
    switch (p.C / kCvKB) {
        case 1: return launch_bwd_kbn<1, MODE>(map_stat, map_str, p, grid, st);
        case 2: return launch_bwd_kbn<2, MODE>(map_stat, map_str, p, grid, st);
        case 3: return launch_bwd_kbn<3, MODE>(map_stat, map_str, p, grid, st);
        default: return launch_bwd_kbn<4, MODE>(map_stat, map_str, p, grid, st);
    }
}

int conv_bwd_rows_pad(int K, int D) { return (K * D + kBwM - 1) / kBwM * kBwM; }

// d x (bf16 or fp32, channels-last), d W (fp32), d bias (fp32) of the conv-fused head from the coefficient blocks of
// integral_coef_kernel; `rowcoef_ws` holds B * conv_bwd_rows_pad(K, D) float4.  Any of dx / dw may be NULL (that launch is skipped).
cudaError_t launch_conv_head_bwd(const void* x_nhwc, const void* w, const float* bias, const float* coef, int coef_stride, float* rowcoef_ws,
                                 void* dx, int dx_f32, float* dw, float* dbias, int B, int K, int D, int H, int W, int C, int num_sms,
                                 cudaStream_t st) {
    ConvBwdParams p{};
    if (const char* d = getenv("XSUP_CONVBWD_TRACE")) p.trace = reinterpret_cast<long long*>(strtoull(d, nullptr, 0));
    p.rowcoef = rowcoef_ws;
    p.C = C; p.HW = H * W; p.W = W; p.rows_total = K * D; p.rows_pad = conv_bwd_rows_pad(K, D);
    const long long n = (long long)B * p.rows_pad;
    conv_rowcoef_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(coef, coef_stride, bias, rowcoef_ws, B, K, D, p.rows_pad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    CUtensorMap map_w128, map_w64, map_x128, map_x64;
    if (!make_map(&map_w128, w, p.rows_total, C, kBwM) || !make_map(&map_w64, w, p.rows_total, C, kBwN) ||
        !make_map(&map_x128, x_nhwc, (long long)B * p.HW, C, kBwM) || !make_map(&map_x64, x_nhwc, (long long)B * p.HW, C, kBwN))
        return cudaErrorNotSupported;
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
        q.T = p.HW / kBwN;
        q.items = B * q.per_b;
        e = launch_bwd_mode<MODE_DW>(map_w128, map_x64, q, num_sms, st);
        if (e != cudaSuccess) return e;
    }
    if (dx) {
        ConvBwdParams q = p;
        q.dx = dx; q.dx_f32 = dx_f32;
        q.per_b = p.HW / kBwM;
        q.T = (p.rows_total + kBwN - 1) / kBwN;
        q.items = B * q.per_b;
        e = launch_bwd_mode<MODE_DX>(map_x128, map_w64, q, num_sms, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace xsup
