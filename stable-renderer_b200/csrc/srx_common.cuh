// srx_common.cuh — shared helpers for libsrx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/srx.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsrx is written for sm_100a (B200) only"
#endif

#define SRX_NO_ID_MAP_INDEX 2048  // default_Gbuffer.frag.glsl:8,150 ; corrmap.py:121,267

// ---------------------------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the C ABI)
// ---------------------------------------------------------------------------------------------------------------
int srx_set_error(int code, const char *fmt, ...);
#define SRX_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return srx_set_error(SRX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                                 __FILE__, __LINE__);                                                     \
    } while (0)
#define SRX_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) return srx_set_error(code, __VA_ARGS__); \
    } while (0)

int srx_sm_count_cached();
// srx_group.cu: in-place exclusive scan of a 0/1 int array (entry -> rank among the set entries, -1 where clear)
long long srx_flags_scan_scratch_ints(long long n);
int srx_flags_to_ranks(int *flags, long long n, int *scratch, int64_t *total, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
// One id pixel = (spriteID, materialID, map_index, vertexID).  128-bit (RGBA_32I) or 64-bit (int16 dumps) loads,
// streamed past L1 (each pixel is read once).
struct IdPx { int s, m, i, v; };

__device__ __forceinline__ IdPx load_id(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return IdPx{r.x, r.y, r.z, r.w};
}
__device__ __forceinline__ IdPx load_id(const short4 *p) {
    int lo, hi;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p));
    return IdPx{(int)(short)(lo & 0xffff), (int)(short)(lo >> 16), (int)(short)(hi & 0xffff), (int)(short)(hi >> 16)};
}

// Two horizontally adjacent id pixels with one load.  RGBA_32I ids use Blackwell's 256-bit global load, streamed past
// L1 and marked evict-first in L2 (each pixel is read exactly once; the accumulator and the latents should stay
// resident instead).  The address must be 32-byte aligned (even pixel index).
__device__ __forceinline__ void load_id_pair(const int4 *p, IdPx &a, IdPx &b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.s), "=r"(a.m), "=r"(a.i), "=r"(a.v), "=r"(b.s), "=r"(b.m), "=r"(b.i), "=r"(b.v) : "l"(p));
}
__device__ __forceinline__ void load_id_pair(const short4 *p, IdPx &a, IdPx &b) {
    int w0, w1, w2, w3;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(p));
    a = IdPx{(int)(short)(w0 & 0xffff), (int)(short)(w0 >> 16), (int)(short)(w1 & 0xffff), (int)(short)(w1 >> 16)};
    b = IdPx{(int)(short)(w2 & 0xffff), (int)(short)(w2 >> 16), (int)(short)(w3 & 0xffff), (int)(short)(w3 >> 16)};
}

// corrmap.py:266-275 — keep rows with map_index != 2048 that are not all-zero
__device__ __forceinline__ bool id_valid(const IdPx &p) {
    return (p.i != SRX_NO_ID_MAP_INDEX) && ((p.s | p.m | p.i | p.v) != 0);
}

// Current-generation key (corresponder.py:331-334): the vertex id after its trip through float32
// (torch.cat promotion, corrmap.py:256-261).  (int64)float(v) is injective on the float values, so two ids share a
// slot exactly when the reference's float keys compare equal.
__device__ __forceinline__ long long vertex_slot(int v) { return (long long)__int2float_rn(v); }

template <typename T> struct XIo;
template <> struct XIo<float> {
    static __device__ __forceinline__ float ld(const float *p) { return *p; }
    static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
};
template <> struct XIo<__half> {
    static __device__ __forceinline__ float ld(const __half *p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half *p, float v) { *p = __float2half_rn(v); }
};
template <> struct XIo<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

// 128-bit float vector reduction into L2 (sm_90+): one REDG.E.ADD.F32x4 per key instead of four scalar atomics.
__device__ __forceinline__ void red_add_f32x4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float *addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add_s64(long long *addr, long long a) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(a) : "memory");
}

// Q31.32 fixed point for the deterministic accumulator: integer addition is associative, so the result does not
// depend on the order in which the atomics land.
#define SRX_FIX_SCALE 4294967296.0 /* 2^32 */
__device__ __forceinline__ long long to_fix(float v) { return __double2ll_rn((double)v * SRX_FIX_SCALE); }
__device__ __forceinline__ double from_fix(long long v) { return (double)v * (1.0 / SRX_FIX_SCALE); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
