// Device-side helpers shared by the xsup_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xsup_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "xsup_b200 kernels are written for sm_100a (B200) only"
#endif

namespace xsup {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegHuge = -1.0e30f;      // finite stand-in for -inf (keeps (m - m') NaN-free)

// ------------------------------------------------------------------ warp reductions
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));   // CREDUX.MAX.F32 (sm_100a)
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));                         // MUFU.EX2
    return r;
}

// ------------------------------------------------------------------ mbarrier + bulk async copy (TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "XSUP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra XSUP_DONE_%=;\n\t"
        "bra XSUP_WAIT_%=;\n\t"
        "XSUP_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst_smem),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ int2 lds_int2(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void stg128_stream(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// ------------------------------------------------------------------ packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2)
// Two independent round-to-nearest fp32 operations per instruction: same results as the scalar forms,
// half the issue slots.  A pair lives in one 64-bit register (even/odd 32-bit register pair).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 ex2_2(f32x2 a) {                 // MUFU has no packed form
    float lo, hi;
    upk2(a, lo, hi);
    return pk2(ex2(lo), ex2(hi));
}

// ------------------------------------------------------------------ 16-byte vector <-> fp32 pairs
template <typename T>
struct Vec;
template <>
struct Vec<float> {
    static constexpr int N = 4;       // elements per 16-byte vector
    static constexpr int P = 2;       // fp32 pairs per vector
    __device__ __forceinline__ static void unpack2(const uint4& u, f32x2 (&x)[2]) {
        x[0] = pk2u(u.x, u.y);
        x[1] = pk2u(u.z, u.w);
    }
    __device__ __forceinline__ static uint4 pack2(const f32x2 (&x)[2]) {
        uint4 u;
        asm("mov.b64 {%0,%1}, %2;" : "=r"(u.x), "=r"(u.y) : "l"(x[0]));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(u.z), "=r"(u.w) : "l"(x[1]));
        return u;
    }
    // lane-local max of U vectors
    template <int U>
    __device__ __forceinline__ static float vmax(const uint4 (&raw)[U]) {
        float m = kNegHuge;
#pragma unroll
        for (int i = 0; i < U; ++i) {
            m = fmaxf(m, fmaxf(__uint_as_float(raw[i].x), __uint_as_float(raw[i].y)));
            m = fmaxf(m, fmaxf(__uint_as_float(raw[i].z), __uint_as_float(raw[i].w)));
        }
        return m;
    }
};
template <>
struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    static constexpr int P = 4;
    // bf16 -> fp32 is a 16-bit left shift; element 2i is the low half of word i
    __device__ __forceinline__ static f32x2 widen(uint32_t w) { return pk2u(w << 16, w & 0xffff0000u); }
    __device__ __forceinline__ static void unpack2(const uint4& u, f32x2 (&x)[4]) {
        x[0] = widen(u.x);
        x[1] = widen(u.y);
        x[2] = widen(u.z);
        x[3] = widen(u.w);
    }
    __device__ __forceinline__ static uint32_t narrow(f32x2 v) {
        float lo, hi;
        upk2(v, lo, hi);
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
    __device__ __forceinline__ static uint4 pack2(const f32x2 (&x)[4]) {
        return make_uint4(narrow(x[0]), narrow(x[1]), narrow(x[2]), narrow(x[3]));
    }
    // max on the packed bf16 pairs (HMNMX2.BF16_V2): no widening needed for the max pass
    template <int U>
    __device__ __forceinline__ static float vmax(const uint4 (&raw)[U]) {
        uint32_t m = 0xff7fff7fu;     // (-bf16 max, -bf16 max)
#pragma unroll
        for (int i = 0; i < U; ++i) {
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m) : "r"(raw[i].x));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m) : "r"(raw[i].y));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m) : "r"(raw[i].z));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m) : "r"(raw[i].w));
        }
        return fmaxf(__uint_as_float(m << 16), __uint_as_float(m & 0xffff0000u));
    }
};

// ------------------------------------------------------------------ tiling of one (b,k) unit
// A unit is D slices of H rows of W elements, contiguous.  It is cut into "tasks" of U 16-byte
// vectors per lane (one warp, 512*U bytes, never straddling a depth slice) and "stages" of
// kTasksPerStage tasks (one bulk copy, one ring slot).
constexpr int kTasksPerStage = 4;
constexpr int kConsumerWarps = 16;
constexpr int kGroups = kConsumerWarps / kTasksPerStage;
constexpr int kMaxU = 8;
constexpr int kMaxD = 256;

struct Tiling {
    int D, H, W;
    int esize;            // bytes per element
    int lpr;              // lanes per row  = W*esize/16   (power of two <= 32)
    int lpr_log2;
    int U;                // vectors per lane per task
    int task_bytes;       // 512*U
    int parts;            // tasks per depth slice
    int parts_log2;       // log2(parts) when parts is a power of two, else -1 (task -> (slice, part) by shift instead of division)
    int rows_per_task;    // 32*U/lpr
    int tasks_per_unit;   // D*parts
    int rounds;           // tasks per warp and stage: small tasks are batched so that a stage (one bulk copy, one handshake) is ~16 KB
    int stages_per_unit;  // ceil(tasks_per_unit / (kTasksPerStage*rounds))
    int stage_bytes;      // kTasksPerStage*rounds*task_bytes
    long long unit_bytes;
};

}  // namespace xsup
