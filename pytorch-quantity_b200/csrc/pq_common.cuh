// pq_common.cuh -- shared helpers for the sm_100a kernels of libpq_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pq_sm100.h"

#define PQ_CUDA_TRY(expr)                        \
    do {                                         \
        cudaError_t _e = (expr);                 \
        if (_e != cudaSuccess) return (int)_e;   \
    } while (0)

namespace pq {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p)
{
    // read-once data: non-coherent path, do not allocate in L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// A multi-tensor launch: up to PQ_MAX_SEGMENTS tensors cut into fixed-size chunks.  The
// table travels by value in the kernel parameters (no device allocation, no extra copy).
struct SegTable {
    const float *ptr[PQ_MAX_SEGMENTS];
    unsigned long long n[PQ_MAX_SEGMENTS];
    unsigned int chunk_end[PQ_MAX_SEGMENTS];  // inclusive prefix sum of chunks per segment
    float param[PQ_MAX_SEGMENTS];             // per-segment scalar (histogram bin width)
    int k;
    unsigned int total_chunks;
};

__device__ __forceinline__ int seg_of_chunk(const SegTable &t, unsigned int chunk)
{
    int lo = 0, hi = t.k - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (chunk < t.chunk_end[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

}  // namespace pq
