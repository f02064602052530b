// pq_gemm_s8.cu -- placeholder until the tcgen05 int8 GEMM / implicit-GEMM conv lands.
#include "pq_common.cuh"

extern "C" int pq_gemm_s8(const int8_t *, const int8_t *, const int32_t *, int, int, int, int, int, int,
                          float *, int8_t *, pq_stream_t)
{
    return PQ_EUNSUPPORTED;
}
extern "C" int pq_conv2d_s8(const int8_t *, const int8_t *, const int32_t *, const pq_conv_desc *, float *,
                            int8_t *, pq_stream_t)
{
    return PQ_EUNSUPPORTED;
}
