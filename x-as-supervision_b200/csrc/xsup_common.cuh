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

// ------------------------------------------------------------------ 16-byte vector <-> fp32 lanes
template <typename T>
struct Vec;
template <>
struct Vec<float> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void unpack(const uint4& u, float (&f)[4]) {
        f[0] = __uint_as_float(u.x);
        f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z);
        f[3] = __uint_as_float(u.w);
    }
    __device__ __forceinline__ static uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
};
template <>
struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void unpack(const uint4& u, float (&f)[8]) {
        // bf16 -> fp32 is a 16-bit left shift; element 2i is the low half of word i
        f[0] = __uint_as_float(u.x << 16);
        f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16);
        f[3] = __uint_as_float(u.y & 0xffff0000u);
        f[4] = __uint_as_float(u.z << 16);
        f[5] = __uint_as_float(u.z & 0xffff0000u);
        f[6] = __uint_as_float(u.w << 16);
        f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
    __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
    __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
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
    int rows_per_task;    // 32*U/lpr
    int tasks_per_unit;   // D*parts
    int stages_per_unit;  // ceil(tasks_per_unit / kTasksPerStage)
    int stage_bytes;      // kTasksPerStage*task_bytes
    long long unit_bytes;
};

}  // namespace xsup
