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
// the first GEMM, K-major; B of the second GEMM read MN-major from the very same bytes), G [128 x 64] bf16 (A of the second
// GEMM, K-major).  Warp roles (18 warps, one persistent CTA per SM):
//   warp 16      TMA producer: stationary operand per item; streamed tiles through a 3-stage ring (+ in MODE_DX the 64 row
//                coefficients of the tile, one 1 KB bulk copy on the same barrier)
//   warp 17      MMA issuer (one elected thread): S(g) = 4*C/64 MMAs M=128 N=64 K=16 into one of 4 TMEM accumulators, then
//                the second GEMM of the PREVIOUS tile (4 MMAs M=128 N=C K=16) as soon as its G tile is published, so the
//                epilogue of tile g-1 overlaps S(g); tcgen05.commit releases ring stage / G buffer / accumulators
//   warps 0..15  epilogue: tcgen05.ld (thread = TMEM lane, 16 columns per warp), G, st.shared (XOR-swizzled 16-byte chunks),
//                fence.proxy.async, mbarrier arrive; at the end of an item they drain D2
// TMEM: columns [0,256) = 4 x S[128 x 64] fp32, [256, 256+C) = D2.
// Flops: 2 GEMMs of 2*K*D*C*H*W per sample and launch (1.17 TFLOP per launch at B=256, K=17, D=64, C=256).
#include "xsup_internal.h"
#include "xsup_umma.cuh"

namespace xsup {

constexpr int kBwEpiWarps = 16;
constexpr int kBwParts = kBwEpiWarps / 4;                  // warps sharing a TMEM lane quarter: 16 of a tile's 64 columns each
constexpr int kBwThreads = (kBwEpiWarps + 2) * 32;
constexpr int kBwM = 128;                                  // stationary rows = UMMA M
constexpr int kBwN = 64;                                   // streamed rows per tile = N of the first GEMM, K of the second
constexpr int kBwStages = 3;
constexpr int kBwSAcc = 4;
constexpr int kBwStatKB = kBwM * kCvKB * 2;                // 16 KB: one [128 x 64] bf16 k-block of the stationary operand
constexpr int kBwStrKB = kBwN * kCvKB * 2;                 //  8 KB: one [64 x 64] bf16 k-block of a streamed tile
constexpr int kBwGBytes = kBwM * kBwN * 2;                 // 16 KB: one G tile
constexpr int kBwTabBytes = kBwN * 16;                     //  1 KB: row coefficients of a streamed weight tile (MODE_DX)
constexpr int kBwD2Col = kBwSAcc * kBwN;                   // first TMEM column of D2
enum { MODE_DW = 0, MODE_DX = 1 };

struct ConvBwdParams {
    const float4* rowcoef;      // [B][rows_pad] (bias*log2e - lse2, a, b, base0 + c[d] - a*wc - b*hc); zero for padding rows
    void* dx;                   // MODE_DX: [B, HW, C] bf16 or fp32
    int dx_f32;
    float* dw;                  // MODE_DW: [K*D, C] fp32, zeroed by the launcher
    float* dbias;               // MODE_DW: [K*D] fp32, zeroed by the launcher; may be NULL
    int C, HW, W, rows_total, rows_pad;
    int per_b;                  // items per sample: row groups (MODE_DW) or 128-pixel tiles (MODE_DX)
    int T;                      // streamed tiles per item
    int items;
};

// ------------------------------------------------------------------ row coefficients
// One float4 per (sample, logit row): everything the epilogue needs to turn a logit into its gradient,
//   G = 2^(L*log2e + q.x) * (q.y*w + q.z*h + q.w)
__global__ void __launch_bounds__(256) conv_rowcoef_kernel(const float* __restrict__ coef, int coef_stride, const float* __restrict__ bias,
                                                           float4* __restrict__ out, int B, int K, int D, int rows_pad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * rows_pad) return;
    const int b = (int)(i / rows_pad), r = (int)(i - (long long)b * rows_pad);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < K * D) {
        const int k = r / D, d = r - k * D;
        const float* cf = coef + ((size_t)b * K + k) * coef_stride;
        const float a = cf[1], bb = cf[2];
        q.x = (bias ? bias[r] : 0.f) * kLog2e - cf[0];
        q.y = a;
        q.z = bb;
        q.w = (cf[3] + cf[8 + d]) - fmaf(a, cf[4], bb * cf[5]);
    }
    out[i] = q;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

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
    uint8_t* sG = sStr + kBwStages * kStageBytes;                   // [2][128 x 64] bf16
    uint8_t* sTab = sG + 2 * kBwGBytes;                             // [stages][64] float4
    uint64_t* bars = reinterpret_cast<uint64_t*>(sTab + kBwStages * kBwTabBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    const uint32_t b_sfull = smem_u32(bars), b_sempty = b_sfull + 8, b_xfull = b_sempty + 8, b_xempty = b_xfull + 8 * kBwStages,
                   b_afull = b_xempty + 8 * kBwStages, b_aempty = b_afull + 8 * kBwSAcc, b_gfull = b_aempty + 8 * kBwSAcc,
                   b_gempty = b_gfull + 16, b_dfull = b_gempty + 16, b_dempty = b_dfull + 8;

    if (threadIdx.x == 0) {
        mbar_init(b_sfull, 1);
        mbar_init(b_sempty, 1);
        for (int i = 0; i < kBwStages; ++i) {
            mbar_init(b_xfull + 8 * i, 1);
            mbar_init(b_xempty + 8 * i, 1);
        }
        for (int i = 0; i < kBwSAcc; ++i) {
            mbar_init(b_afull + 8 * i, 1);
            mbar_init(b_aempty + 8 * i, kBwEpiWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_gfull + 8 * i, kBwEpiWarps);
            mbar_init(b_gempty + 8 * i, 1);
        }
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
            int g = 0;
            for (int n = 0; n < n_items; ++n) {
                const int item = blockIdx.x + n * gridDim.x;
                const int b = item / p.per_b, j = item - b * p.per_b;
                // stationary: weight rows [128 j, +128) (MODE_DW) or pixels [b*HW + 128 j, +128) (MODE_DX)
                const int stat_row = MODE == MODE_DW ? j * kBwM : b * p.HW + j * kBwM;
                const int str_row0 = MODE == MODE_DW ? b * p.HW : 0;
                mbar_wait(b_sempty, (n & 1) ^ 1);                    // the previous item's first-GEMM MMAs have retired
                mbar_arrive_expect_tx(b_sfull, kStatBytes);
#pragma unroll
                for (int kb = 0; kb < KBN; ++kb) tma_load_2d(smem_u32(sStat) + kb * kBwStatKB, &map_stat, kb * kCvKB, stat_row, b_sfull);
                for (int t = 0; t < T; ++t, ++g) {
                    const int s = g % kBwStages, it = g / kBwStages;
                    mbar_wait(b_xempty + 8 * s, (it & 1) ^ 1);       // the second GEMM of the tile that used this stage has retired
                    mbar_arrive_expect_tx(b_xfull + 8 * s, kStageBytes + (MODE == MODE_DX ? (uint32_t)kBwTabBytes : 0u));
                    const uint32_t dst = smem_u32(sStr) + (uint32_t)s * kStageBytes;
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) tma_load_2d(dst + kb * kBwStrKB, &map_str, kb * kCvKB, str_row0 + t * kBwN, b_xfull + 8 * s);
                    if (MODE == MODE_DX)
                        bulk_g2s(smem_u32(sTab) + (uint32_t)s * kBwTabBytes, p.rowcoef + (size_t)b * p.rows_pad + (size_t)t * kBwN, kBwTabBytes,
                                 b_xfull + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp == kBwEpiWarps + 1) {
        // ------------------------------------------------------------ MMA issuer
        const int total = n_items * T;
        const uint32_t stat0 = smem_u32(sStat), str0 = smem_u32(sStr), g0 = smem_u32(sG);
        for (int g = 0; g <= total; ++g) {
            if (g < total) {
                // ---- first GEMM of tile g: S = stationary * streamed^T, K = C
                const int n = g / T, t = g - n * T;
                const int s = g % kBwStages, a = g % kBwSAcc;
                if (t == 0) mbar_wait(b_sfull, n & 1);
                mbar_wait(b_aempty + 8 * a, ((g / kBwSAcc) & 1) ^ 1);
                mbar_wait(b_xfull + 8 * s, (g / kBwStages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t acc = tmem_base + (uint32_t)a * kBwN;
#pragma unroll
                    for (int kb = 0; kb < KBN; ++kb) {
                        const uint64_t da = umma_desc_sw128(stat0 + kb * kBwStatKB);
                        const uint64_t db = umma_desc_sw128(str0 + (uint32_t)s * kStageBytes + kb * kBwStrKB);
#pragma unroll
                        for (int q = 0; q < 4; ++q) umma_f16_rt(acc, da + 2 * q, db + 2 * q, kIdescS, (kb | q) ? 1u : 0u);
                    }
                    umma_commit(b_afull + 8 * a);
                    if (t == T - 1) umma_commit(b_sempty);           // the stationary operand may be replaced
                }
                __syncwarp();
            }
            if (g >= 1) {
                // ---- second GEMM of tile h = g-1: D2 += G * streamed, K = 64 streamed rows
                const int h = g - 1, n = h / T, t = h - n * T;
                const int s = h % kBwStages, gb = h & 1;
                if (t == 0 && n >= 1) mbar_wait(b_dempty, (n - 1) & 1);   // the epilogue has drained the previous item's D2
                mbar_wait(b_gfull + 8 * gb, (h >> 1) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d2 = tmem_base + (uint32_t)kBwD2Col;
                    const uint64_t da = umma_desc_sw128(g0 + (uint32_t)gb * kBwGBytes);
                    const uint64_t db = umma_desc_sw128_mn(str0 + (uint32_t)s * kStageBytes, kBwStrKB);
#pragma unroll
                    for (int q = 0; q < 4; ++q) umma_f16_rt(d2, da + 2 * q, db + (uint64_t)(128 * q), kIdesc2, (t | q) ? 1u : 0u);
                    umma_commit(b_xempty + 8 * s);
                    umma_commit(b_gempty + 8 * gb);
                    if (t == T - 1) umma_commit(b_dfull);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread = TMEM lane, 16 columns per warp and tile
        const int quarter = warp & 3, part = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t g0 = smem_u32(sG);
        const int Wd = p.W;
        int g = 0;
        for (int n = 0; n < n_items; ++n) {
            const int item = blockIdx.x + n * gridDim.x;
            const int b = item / p.per_b, j = item - b * p.per_b;
            float4 rc = make_float4(0.f, 0.f, 0.f, 0.f);
            float fw = 0.f, fh = 0.f, gsum = 0.f;
            if (MODE == MODE_DW) {
                rc = p.rowcoef[(size_t)b * p.rows_pad + j * kBwM + row];
            } else {
                const int pix = j * kBwM + row;
                const int hh = pix / Wd;
                fh = (float)hh;
                fw = (float)(pix - hh * Wd);
            }
            for (int t = 0; t < T; ++t, ++g) {
                const int a = g % kBwSAcc, s = g % kBwStages, gb = g & 1;
                mbar_wait(b_afull + 8 * a, (g / kBwSAcc) & 1);
                tc_fence_after();
                uint32_t r[16];
                tmem_ld16(lane_addr + (uint32_t)(a * kBwN + part * 16), r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_aempty + 8 * a);        // values are in registers: release the accumulator early
                uint32_t o[8];
                if (MODE == MODE_DW) {
                    // columns = 16 consecutive pixels of one image row (W % 16 == 0)
                    const int pix = t * kBwN + part * 16;
                    const int hh = pix / Wd;
                    const float rowterm = fmaf(rc.y, (float)(pix - hh * Wd), fmaf(rc.z, (float)hh, rc.w));
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float v0 = ex2(fmaf(__uint_as_float(r[i]), kLog2e, rc.x)) * fmaf(rc.y, (float)i, rowterm);
                        const float v1 = ex2(fmaf(__uint_as_float(r[i + 1]), kLog2e, rc.x)) * fmaf(rc.y, (float)(i + 1), rowterm);
                        gsum += v0 + v1;
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i >> 1]) : "f"(v1), "f"(v0));
                    }
                } else {
                    // columns = 16 consecutive logit rows of the streamed weight tile: coefficients from the stage's table
                    mbar_wait(b_xfull + 8 * s, (g / kBwStages) & 1);   // (already complete: the MMAs read this stage) acquire the table
                    const float4* tab = reinterpret_cast<const float4*>(sTab + s * kBwTabBytes) + part * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float4 q0 = tab[i], q1 = tab[i + 1];
                        const float v0 = ex2(fmaf(__uint_as_float(r[i]), kLog2e, q0.x)) * fmaf(q0.y, fw, fmaf(q0.z, fh, q0.w));
                        const float v1 = ex2(fmaf(__uint_as_float(r[i + 1]), kLog2e, q1.x)) * fmaf(q1.y, fw, fmaf(q1.z, fh, q1.w));
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i >> 1]) : "f"(v1), "f"(v0));
                    }
                }
                // publish this warp's [32 lanes x 16 columns] of the G tile: row = 128 bytes, 16-byte chunk c at (c ^ (row & 7))
                mbar_wait(b_gempty + 8 * gb, ((g >> 1) & 1) ^ 1);      // the second GEMM that read this buffer has retired
                const uint32_t grow = g0 + (uint32_t)gb * kBwGBytes + (uint32_t)row * 128u;
                const uint32_t sw = (uint32_t)(row & 7);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(grow + (((uint32_t)(2 * part) ^ sw) << 4)), "r"(o[0]), "r"(o[1]),
                             "r"(o[2]), "r"(o[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(grow + (((uint32_t)(2 * part + 1) ^ sw) << 4)), "r"(o[4]), "r"(o[5]),
                             "r"(o[6]), "r"(o[7]) : "memory");
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_gfull + 8 * gb);
            }
            // ---- end of the item: drain D2 (this warp: its lane quarter, C/4 columns)
            mbar_wait(b_dfull, n & 1);
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
                if (p.dbias && grow_i < p.rows_total) atomicAdd(p.dbias + grow_i, gsum);
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
    const size_t smem = 1024 + (size_t)KBN * kBwStatKB + (size_t)kBwStages * KBN * kBwStrKB + 2 * kBwGBytes + kBwStages * kBwTabBytes + 32 * 8;
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
    p.rowcoef = reinterpret_cast<const float4*>(rowcoef_ws);
    p.C = C; p.HW = H * W; p.W = W; p.rows_total = K * D; p.rows_pad = conv_bwd_rows_pad(K, D);
    const long long n = (long long)B * p.rows_pad;
    conv_rowcoef_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(coef, coef_stride, bias, reinterpret_cast<float4*>(rowcoef_ws), B, K, D, p.rows_pad);
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
