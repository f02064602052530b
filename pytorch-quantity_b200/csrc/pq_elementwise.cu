// pq_elementwise.cu -- bandwidth kernels of the simulation path:
//   fakequant  y = clamp(rint(x * 2^bit), lo, hi) [/ 2^bit]   new_quantity_op.py:48-58, 246-257
//   add_clamp  y = clamp(a + b, lo, hi)                        new_quantity_op.py:166-174
//   quantize NCHW fp32 -> NHWC int8 (input quantiser fused with the layout change)
// One fused pass each: 8 B/element (fakequant), 12 B/element (add), 5 B/element (quantise)
// instead of the reference's 3-4 separate ATen kernels.
#include "pq_common.cuh"

namespace pq {

constexpr int kEwThreads = 256;
constexpr int kEwUnroll = 4;

// clamp that propagates NaN like torch.clamp (fminf/fmaxf would drop it)
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

template <bool DEQUANT>
__device__ __forceinline__ float fq1(float x, float scale, float inv_scale, float lo, float hi)
{
    // torch.round == round-half-even == rintf (default rounding mode); the scale is a power of
    // two, so the multiply is exact and the divide equals a multiply by the exact reciprocal.
    float v = clamp_nan(rintf(__fmul_rn(x, scale)), lo, hi);
    return DEQUANT ? __fmul_rn(v, inv_scale) : v;
}

template <bool DEQUANT>
__global__ void __launch_bounds__(kEwThreads)
fakequant_kernel(const float *__restrict__ x, float *__restrict__ y, size_t n, float scale,
                 float inv_scale, float lo, float hi)
{
    // x and y are 16-byte aligned here (host checks); scalar tail handled by the last threads
    const size_t nvec = n >> 2;
    const float4 *xv = reinterpret_cast<const float4 *>(x);
    float4 *yv = reinterpret_cast<float4 *>(y);
    const size_t stride = (size_t)gridDim.x * kEwThreads;
    size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x;
    for (; i + (kEwUnroll - 1) * stride < nvec; i += kEwUnroll * stride) {
        float4 v[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) v[u] = ld_stream_f4(xv + i + u * stride);
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            v[u].x = fq1<DEQUANT>(v[u].x, scale, inv_scale, lo, hi);
            v[u].y = fq1<DEQUANT>(v[u].y, scale, inv_scale, lo, hi);
            v[u].z = fq1<DEQUANT>(v[u].z, scale, inv_scale, lo, hi);
            v[u].w = fq1<DEQUANT>(v[u].w, scale, inv_scale, lo, hi);
            st_stream_f4(yv + i + u * stride, v[u]);
        }
    }
    for (; i < nvec; i += stride) {
        float4 v = ld_stream_f4(xv + i);
        v.x = fq1<DEQUANT>(v.x, scale, inv_scale, lo, hi);
        v.y = fq1<DEQUANT>(v.y, scale, inv_scale, lo, hi);
        v.z = fq1<DEQUANT>(v.z, scale, inv_scale, lo, hi);
        v.w = fq1<DEQUANT>(v.w, scale, inv_scale, lo, hi);
        st_stream_f4(yv + i, v);
    }
    const size_t t = (nvec << 2) + (size_t)blockIdx.x * kEwThreads + threadIdx.x;
    if (t < n) y[t] = fq1<DEQUANT>(x[t], scale, inv_scale, lo, hi);
}

// unaligned fallback (views with odd offsets): same arithmetic, scalar accesses
template <bool DEQUANT>
__global__ void __launch_bounds__(kEwThreads)
fakequant_scalar_kernel(const float *__restrict__ x, float *__restrict__ y, size_t n, float scale,
                        float inv_scale, float lo, float hi)
{
    for (size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kEwThreads)
        y[i] = fq1<DEQUANT>(x[i], scale, inv_scale, lo, hi);
}

__global__ void __launch_bounds__(kEwThreads)
add_clamp_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ y,
                 size_t n, float lo, float hi, int vec_ok)
{
    const size_t stride = (size_t)gridDim.x * kEwThreads;
    size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x;
    if (vec_ok) {
        const size_t nvec = n >> 2;
        const float4 *av = reinterpret_cast<const float4 *>(a);
        const float4 *bv = reinterpret_cast<const float4 *>(b);
        float4 *yv = reinterpret_cast<float4 *>(y);
        for (; i + stride < nvec; i += 2 * stride) {
            float4 p0 = ld_stream_f4(av + i), q0 = ld_stream_f4(bv + i);
            float4 p1 = ld_stream_f4(av + i + stride), q1 = ld_stream_f4(bv + i + stride);
            float4 r0, r1;
            r0.x = clamp_nan(__fadd_rn(p0.x, q0.x), lo, hi); r0.y = clamp_nan(__fadd_rn(p0.y, q0.y), lo, hi);
            r0.z = clamp_nan(__fadd_rn(p0.z, q0.z), lo, hi); r0.w = clamp_nan(__fadd_rn(p0.w, q0.w), lo, hi);
            r1.x = clamp_nan(__fadd_rn(p1.x, q1.x), lo, hi); r1.y = clamp_nan(__fadd_rn(p1.y, q1.y), lo, hi);
            r1.z = clamp_nan(__fadd_rn(p1.z, q1.z), lo, hi); r1.w = clamp_nan(__fadd_rn(p1.w, q1.w), lo, hi);
            st_stream_f4(yv + i, r0);
            st_stream_f4(yv + i + stride, r1);
        }
        for (; i < nvec; i += stride) {
            float4 p = ld_stream_f4(av + i), q = ld_stream_f4(bv + i), r;
            r.x = clamp_nan(__fadd_rn(p.x, q.x), lo, hi); r.y = clamp_nan(__fadd_rn(p.y, q.y), lo, hi);
            r.z = clamp_nan(__fadd_rn(p.z, q.z), lo, hi); r.w = clamp_nan(__fadd_rn(p.w, q.w), lo, hi);
            st_stream_f4(yv + i, r);
        }
        const size_t t = (nvec << 2) + (size_t)blockIdx.x * kEwThreads + threadIdx.x;
        if (t < n) y[t] = clamp_nan(__fadd_rn(a[t], b[t]), lo, hi);
    } else {
        for (; i < n; i += stride) y[i] = clamp_nan(__fadd_rn(a[i], b[i]), lo, hi);
    }
}

// stand-alone RightShift / Sp / DeQuantity (inside NewConv2d these live in the GEMM epilogue)
template <int OP>   // 0: rshift-round-clamp, 1: clamp-scale
__global__ void __launch_bounds__(kEwThreads)
unary_kernel(const float *__restrict__ x, float *__restrict__ y, size_t n, float mul, float lo, float hi)
{
    for (size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kEwThreads) {
        const float v = x[i];
        float r;
        if (OP == 0) {
            const float s = __fmul_rn(v, mul);                                  // x / 2^rs, exact
            const float t = __fadd_rn(s, s > 0.0f ? 0.5f : -0.5f);              // :32-37
            r = (float)max(min((int)t, (int)hi), (int)lo);                      // int32 cast truncates; :41
        } else {
            r = __fmul_rn(clamp_nan(v, lo, hi), mul);
        }
        y[i] = r;
    }
}

// NCHW fp32 -> NHWC int8 with channel padding.  One CTA transposes a [32 channels][32 pixels]
// tile through shared memory: coalesced 128 B fp32 reads along pixels, 32 B int8 writes per
// pixel row along channels (c_pad is a multiple of 16, so rows are 16-byte aligned).
__global__ void __launch_bounds__(256)
quantize_nchw_nhwc_kernel(const float *__restrict__ x, int8_t *__restrict__ q, int C, int HW,
                          int c_pad, float scale)
{
    __shared__ int8_t tile[32][33 + 3];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const float *src = x + (size_t)n * C * HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + j * 8, p = p0 + tx;
        float v = 0.0f;
        if (c < C && p < HW) v = __ldg(src + (size_t)c * HW + p);
        v = clamp_nan(rintf(__fmul_rn(v, scale)), -128.0f, 127.0f);
        tile[ty + j * 8][tx] = (int8_t)(int)v;
    }
    __syncthreads();
    int8_t *dst = q + (size_t)n * HW * c_pad;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = p0 + ty + j * 8, c = c0 + tx;
        if (p < HW && c < c_pad) dst[(size_t)p * c_pad + c] = tile[tx][ty + j * 8];
    }
}

}  // namespace pq

namespace {
unsigned int ew_grid(size_t work_items)
{
    size_t blocks = (work_items + pq::kEwThreads - 1) / pq::kEwThreads;
    size_t cap = (size_t)pq::kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned int)(blocks ? blocks : 1);
}
bool aligned16(const void *p) { return (((unsigned long long)p) & 15ull) == 0; }
}  // namespace

extern "C" int pq_fakequant_f32(const float *x, float *y, size_t n, int bit, float lo, float hi,
                                int dequant, pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!x || !y || x == y) return PQ_EINVAL;
    if (bit < -126 || bit > 126) return PQ_EUNSUPPORTED;
    const float scale = ldexpf(1.0f, bit), inv = ldexpf(1.0f, -bit);
    cudaStream_t s = (cudaStream_t)stream;
    if (aligned16(x) && aligned16(y)) {
        unsigned int g = ew_grid((n >> 2) / pq::kEwUnroll + 1);
        if (dequant) pq::fakequant_kernel<true><<<g, pq::kEwThreads, 0, s>>>(x, y, n, scale, inv, lo, hi);
        else pq::fakequant_kernel<false><<<g, pq::kEwThreads, 0, s>>>(x, y, n, scale, inv, lo, hi);
    } else {
        unsigned int g = ew_grid(n);
        if (dequant) pq::fakequant_scalar_kernel<true><<<g, pq::kEwThreads, 0, s>>>(x, y, n, scale, inv, lo, hi);
        else pq::fakequant_scalar_kernel<false><<<g, pq::kEwThreads, 0, s>>>(x, y, n, scale, inv, lo, hi);
    }
    return (int)cudaGetLastError();
}

extern "C" int pq_add_clamp_f32(const float *a, const float *b, float *y, size_t n, float lo, float hi,
                                pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!a || !b || !y || y == a || y == b) return PQ_EINVAL;
    const int vec_ok = aligned16(a) && aligned16(b) && aligned16(y);
    unsigned int g = ew_grid(vec_ok ? (n >> 3) + 1 : n);
    pq::add_clamp_kernel<<<g, pq::kEwThreads, 0, (cudaStream_t)stream>>>(a, b, y, n, lo, hi, vec_ok);
    return (int)cudaGetLastError();
}

extern "C" int pq_rshift_f32(const float *x, float *y, size_t n, int rs, float lo, float hi, pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!x || !y) return PQ_EINVAL;
    if (rs < -100 || rs > 100) return PQ_EUNSUPPORTED;
    pq::unary_kernel<0><<<ew_grid(n), pq::kEwThreads, 0, (cudaStream_t)stream>>>(x, y, n, ldexpf(1.0f, -rs), lo, hi);
    return (int)cudaGetLastError();
}

extern "C" int pq_clamp_scale_f32(const float *x, float *y, size_t n, float lo, float hi, float scale,
                                  pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!x || !y) return PQ_EINVAL;
    pq::unary_kernel<1><<<ew_grid(n), pq::kEwThreads, 0, (cudaStream_t)stream>>>(x, y, n, scale, lo, hi);
    return (int)cudaGetLastError();
}

extern "C" int pq_quantize_nchw_to_nhwc_s8(const float *x, int8_t *q, int N, int C, int H, int W, int c_pad,
                                           int ib, pq_stream_t stream)
{
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return PQ_EINVAL;
    if (!x || !q) return PQ_EINVAL;
    if (c_pad < C || (c_pad & 15)) return PQ_EUNSUPPORTED;
    if (ib < -126 || ib > 126 || N > 65535) return PQ_EUNSUPPORTED;
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (c_pad + 31) / 32, N);
    if (grid.y > 65535) return PQ_EUNSUPPORTED;
    pq::quantize_nchw_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, q, C, HW, c_pad, ldexpf(1.0f, ib));
    return (int)cudaGetLastError();
}
