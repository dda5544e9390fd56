// vt_mem.cuh -- cache-policy-qualified loads and stores (sm_100a).
//
// The wavefront renderer moves two kinds of data with opposite needs (profiles/r01_v15_wf_shade_*: L2 hit 37-43 %, DRAM
// traffic 1.58x the algorithmic bytes, the kernel waits on loads):
//   * STREAMS -- path records, ray records, queue entries, samples: tens of GB per batch, written once and read once,
//     far larger than the 126 MB L2. They must not displace anything: L1 no-allocate, L2 evict-first, and moved as whole
//     32-byte sectors (256-bit LDG/STG, new on sm_100) so that every fetched sector is fully used.
//   * TABLES -- the 16 MiB noise texture, the environment map, its CDFs and guide tables, the material records: ~25 MB
//     gathered at random all the time. They should stay in L2 for the whole batch: L2 evict-last (createpolicy + cache_hint).
// VT_MEM_HINTS=0 compiles every helper to the plain instruction (A/B builds).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef VT_MEM_HINTS
#define VT_MEM_HINTS 1
#endif

namespace vt {

struct f8 { float4 lo, hi; };      // one 32-byte sector

#if VT_MEM_HINTS
// createpolicy is pure (no operands, no side effects): the compiler keeps one copy per kernel
__device__ __forceinline__ unsigned long long pol_keep()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long pol_stream()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif

// ---- streams ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ f8 ld_stream32(const void* p)                       // p 32-byte aligned
{
    f8 r;
#if VT_MEM_HINTS
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
#else
    r.lo = reinterpret_cast<const float4*>(p)[0]; r.hi = reinterpret_cast<const float4*>(p)[1];
#endif
    return r;
}
__device__ __forceinline__ void st_stream32(void* p, float4 lo, float4 hi)     // p 32-byte aligned
{
#if VT_MEM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
#else
    reinterpret_cast<float4*>(p)[0] = lo; reinterpret_cast<float4*>(p)[1] = hi;
#endif
}
__device__ __forceinline__ float4 ld_stream16(const float4* p)
{
#if VT_MEM_HINTS
    float4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol_stream()));
    return r;
#else
    return *p;
#endif
}
__device__ __forceinline__ int4 ld_stream16(const int4* p)
{
#if VT_MEM_HINTS
    int4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol_stream()));
    return r;
#else
    return *p;
#endif
}
__device__ __forceinline__ float2 ld_stream8(const float2* p)
{
#if VT_MEM_HINTS
    float2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol_stream()));
    return r;
#else
    return *p;
#endif
}
__device__ __forceinline__ int ld_stream4(const int* p)
{
#if VT_MEM_HINTS
    int r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol_stream()));
    return r;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_stream16(float4* p, float4 v)
{
#if VT_MEM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol_stream()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_stream16(int4* p, int4 v)
{
#if VT_MEM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol_stream()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_stream8(float2* p, float2 v)
{
#if VT_MEM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol_stream()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_stream4(int* p, int v)
{
#if VT_MEM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.s32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol_stream()) : "memory");
#else
    *p = v;
#endif
}

// ---- tables (read-only for the lifetime of the kernel: non-coherent path, kept in L2) ---------------------------------
__device__ __forceinline__ float ldg_keep(const float* p)
{
#if VT_MEM_HINTS
    float r;
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol_keep()));
    return r;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float4 ldg_keep(const float4* p)
{
#if VT_MEM_HINTS
    float4 r;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol_keep()));
    return r;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ unsigned short ldg_keep(const unsigned short* p)
{
#if VT_MEM_HINTS
    unsigned short r;
    asm("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol_keep()));
    return r;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void prefetch_keep(const void* p)      // bring the line holding p towards L1, keep it in L2
{
#if VT_MEM_HINTS
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p));
#endif
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

} // namespace vt
