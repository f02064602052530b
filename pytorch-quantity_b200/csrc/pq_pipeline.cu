// pq_pipeline.cu -- bandwidth kernels of the int8 inter-layer pipeline: the
// reconstructed model keeps activations as int8 NHWC between tensor-core layers instead of
// round-tripping through fp32 NCHW.  Every kernel is exact integer arithmetic that reproduces the
// reference's fp32 composition (Quantity o ReLU / MaxPool / NewAdd), see include/pq_sm100.h.
#include "pq_common.cuh"

namespace pq {

constexpr int kPipeThreads = 256;

__device__ __forceinline__ uint32_t relu4_s8(uint32_t v)
{
    return __vmaxs4(v, 0u);                      // per-byte signed max with 0
}

__global__ void __launch_bounds__(kPipeThreads)
relu_s8_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t nvec, const int8_t *xt, int8_t *yt, size_t n)
{
    const size_t stride = (size_t)gridDim.x * kPipeThreads;
    for (size_t i = (size_t)blockIdx.x * kPipeThreads + threadIdx.x; i < nvec; i += stride) {
        uint4 v = x[i];
        v.x = relu4_s8(v.x); v.y = relu4_s8(v.y); v.z = relu4_s8(v.z); v.w = relu4_s8(v.w);
        y[i] = v;
    }
    const size_t t = (nvec << 4) + (size_t)blockIdx.x * kPipeThreads + threadIdx.x;
    if (t < n) yt[t] = xt[t] > 0 ? xt[t] : (int8_t)0;
}

// int8 NHWC max-pool: one thread per (output pixel, 16 channels)
__global__ void __launch_bounds__(kPipeThreads)
maxpool_nhwc_s8_kernel(const int8_t *__restrict__ x, int8_t *__restrict__ y, int N, int H, int W, int C, int k,
                       int stride, int pad, int P, int Q, int relu)
{
    const int cv = C >> 4;
    const size_t total = (size_t)N * P * Q * cv;
    for (size_t t = (size_t)blockIdx.x * kPipeThreads + threadIdx.x; t < total; t += (size_t)gridDim.x * kPipeThreads) {
        const int c16 = (int)(t % cv);
        size_t pix = t / cv;
        const int q = (int)(pix % Q); pix /= Q;
        const int pp = (int)(pix % P);
        const size_t n = pix / P;
        const uint32_t init = relu ? 0u : 0x80808080u;           // -128 per byte == -inf padding
        uint4 m = make_uint4(init, init, init, init);
        for (int r = 0; r < k; ++r) {
            const int iy = pp * stride - pad + r;
            if ((unsigned)iy >= (unsigned)H) continue;
            for (int s = 0; s < k; ++s) {
                const int ix = q * stride - pad + s;
                if ((unsigned)ix >= (unsigned)W) continue;
                const uint4 v = *reinterpret_cast<const uint4 *>(x + ((n * H + iy) * W + ix) * C + c16 * 16);
                m.x = __vmaxs4(m.x, v.x); m.y = __vmaxs4(m.y, v.y); m.z = __vmaxs4(m.z, v.z); m.w = __vmaxs4(m.w, v.w);
            }
        }
        *reinterpret_cast<uint4 *>(y + ((n * P + pp) * Q + q) * C + c16 * 16) = m;
    }
}

struct AddParams {
    const void *a, *b;
    int a_is16, b_is16, a_relu, b_relu;
    int out_relu;                  // a following nn.ReLU fused: s = max(s, 0)
    int a_shift, b_shift;          // o_bit - a_bit, o_bit - b_bit  (>= 0)
    int lo, hi;                    // clamp of the real sum to [-128, 127] in units of 2^-o_bit
    int q_shift;                   // q_bit - o_bit
    int16_t *out16;
    int8_t *out8;
};

__device__ __forceinline__ int add_load(const void *p, int is16, size_t i)
{
    return is16 ? (int)reinterpret_cast<const int16_t *>(p)[i] : (int)reinterpret_cast<const int8_t *>(p)[i];
}

// round_half_even(num * 2^sh) saturated to int8
__device__ __forceinline__ int requant_rne(int num, int sh)
{
    int r;
    if (sh >= 0) {
        r = max(-128, min(127, num)) << sh;      // saturate first: |num| <= 2^15, sh <= 15
    } else {
        const int d = -sh;
        r = (num + (1 << (d - 1)) - 1 + ((num >> d) & 1)) >> d;   // arithmetic shift, ties to even
    }
    return max(-128, min(127, r));
}

__device__ __forceinline__ uint4 ld_nc_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16 consecutive operands of vector index v, sign-extended
template <bool IS16>
__device__ __forceinline__ void add_load16(const void *base, size_t v, int (&out)[16])
{
    if (IS16) {
        const uint4 w0 = ld_nc_u4(reinterpret_cast<const uint4 *>(base) + 2 * v);
        const uint4 w1 = ld_nc_u4(reinterpret_cast<const uint4 *>(base) + 2 * v + 1);
        const uint32_t ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            out[2 * j] = (int)(int16_t)(ww[j] & 0xffff);
            out[2 * j + 1] = (int)ww[j] >> 16;
        }
    } else {
        const uint4 w = ld_nc_u4(reinterpret_cast<const uint4 *>(base) + v);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = (int)(int8_t)((ww[j >> 2] >> ((j & 3) * 8)) & 0xff);
    }
}

// 16 elements per thread and iteration: 16 / 32-byte vector loads per operand, two 16-byte stores for the
// int16 sum and one for the int8 requantisation (buffers are 16-byte aligned; scalar tail below)
template <bool A16, bool B16>
__global__ void __launch_bounds__(kPipeThreads)
add_requant_kernel(const AddParams p, size_t n)
{
    const size_t nvec = n >> 4;
    const size_t stride = (size_t)gridDim.x * kPipeThreads;
    for (size_t v = (size_t)blockIdx.x * kPipeThreads + threadIdx.x; v < nvec; v += stride) {
        int av[16], bv[16];
        add_load16<A16>(p.a, v, av);
        add_load16<B16>(p.b, v, bv);
        uint32_t o16[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o8[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int x = av[j], y = bv[j];
            if (p.a_relu) x = max(x, 0);
            if (p.b_relu) y = max(y, 0);
            const int num = max(p.lo, min(p.hi, (x << p.a_shift) + (y << p.b_shift)));   // p.lo = 0 with a fused ReLU
            o16[j >> 1] |= ((uint32_t)num & 0xffffu) << ((j & 1) * 16);
            o8[j >> 2] |= ((uint32_t)requant_rne(num, p.q_shift) & 0xffu) << ((j & 3) * 8);
        }
        if (p.out16) {
            reinterpret_cast<uint4 *>(p.out16)[2 * v] = make_uint4(o16[0], o16[1], o16[2], o16[3]);
            reinterpret_cast<uint4 *>(p.out16)[2 * v + 1] = make_uint4(o16[4], o16[5], o16[6], o16[7]);
        }
        if (p.out8) reinterpret_cast<uint4 *>(p.out8)[v] = make_uint4(o8[0], o8[1], o8[2], o8[3]);
    }
    const size_t t = (nvec << 4) + (size_t)blockIdx.x * kPipeThreads + threadIdx.x;
    if (t < n) {
        int x = add_load(p.a, A16, t), y = add_load(p.b, B16, t);
        if (p.a_relu) x = max(x, 0);
        if (p.b_relu) y = max(y, 0);
        const int num = max(p.lo, min(p.hi, (x << p.a_shift) + (y << p.b_shift)));
        if (p.out16) p.out16[t] = (int16_t)num;
        if (p.out8) p.out8[t] = (int8_t)requant_rne(num, p.q_shift);
    }
}

}  // namespace pq

namespace {
unsigned int pipe_grid(size_t items)
{
    size_t blocks = (items + pq::kPipeThreads - 1) / pq::kPipeThreads;
    const size_t cap = (size_t)pq::kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned int)(blocks ? blocks : 1);
}
bool al16(const void *p) { return (((unsigned long long)p) & 15ull) == 0; }
}  // namespace

extern "C" int pq_relu_s8(const int8_t *x, int8_t *y, size_t n, pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!x || !y) return PQ_EINVAL;
    if (!al16(x) || !al16(y)) return PQ_EALIGN;
    pq::relu_s8_kernel<<<pipe_grid((n >> 4) + 1), pq::kPipeThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(x), reinterpret_cast<uint4 *>(y), n >> 4, x, y, n);
    return (int)cudaGetLastError();
}

extern "C" int pq_maxpool_nhwc_s8(const int8_t *x, int8_t *y, int N, int H, int W, int C, int k, int stride, int pad,
                                  int relu, pq_stream_t stream)
{
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || k <= 0 || stride <= 0 || pad < 0 || !x || !y) return PQ_EINVAL;
    if ((C & 15) || pad >= k) return PQ_EUNSUPPORTED;
    if (!al16(x) || !al16(y)) return PQ_EALIGN;
    const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
    if (P <= 0 || Q <= 0) return PQ_EINVAL;
    pq::maxpool_nhwc_s8_kernel<<<pipe_grid((size_t)N * P * Q * (C >> 4)), pq::kPipeThreads, 0, (cudaStream_t)stream>>>(
        x, y, N, H, W, C, k, stride, pad, P, Q, relu);
    return (int)cudaGetLastError();
}

extern "C" int pq_add_requant(const void *a, int a_is16, int a_bit, int a_relu, const void *b, int b_is16, int b_bit,
                              int b_relu, size_t n, int16_t *out16, int8_t *out8, int q_bit, pq_stream_t stream)
{
    return pq_add_requant_ex(a, a_is16, a_bit, a_relu, b, b_is16, b_bit, b_relu, n, 0, out16, out8, q_bit, stream);
}

extern "C" int pq_add_requant_ex(const void *a, int a_is16, int a_bit, int a_relu, const void *b, int b_is16,
                                 int b_bit, int b_relu, size_t n, int flags, int16_t *out16, int8_t *out8,
                                 int q_bit, pq_stream_t stream)
{
    if (n == 0) return PQ_OK;
    if (!a || !b || (!out16 && !out8)) return PQ_EINVAL;
    if (!al16(a) || !al16(b) || (out16 && !al16(out16)) || (out8 && !al16(out8))) return PQ_EALIGN;
    const int o_bit = a_bit > b_bit ? a_bit : b_bit;
    const int span = o_bit - (a_bit < b_bit ? a_bit : b_bit);
    if (span > 7 || o_bit < 0 || o_bit > 7 || q_bit - o_bit > 15 || o_bit - q_bit > 15) return PQ_EUNSUPPORTED;
    pq::AddParams p;
    p.a = a; p.b = b; p.a_is16 = a_is16; p.b_is16 = b_is16; p.a_relu = a_relu; p.b_relu = b_relu;
    p.a_shift = o_bit - a_bit; p.b_shift = o_bit - b_bit;
    p.out_relu = flags & PQ_FLAG_RELU;
    p.lo = p.out_relu ? 0 : -128 * (1 << o_bit); p.hi = 127 * (1 << o_bit);
    p.q_shift = q_bit - o_bit; p.out16 = out16; p.out8 = out8;
    const unsigned int grid = pipe_grid((n >> 4) + 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (a_is16 && b_is16) pq::add_requant_kernel<true, true><<<grid, pq::kPipeThreads, 0, s>>>(p, n);
    else if (a_is16) pq::add_requant_kernel<true, false><<<grid, pq::kPipeThreads, 0, s>>>(p, n);
    else if (b_is16) pq::add_requant_kernel<false, true><<<grid, pq::kPipeThreads, 0, s>>>(p, n);
    else pq::add_requant_kernel<false, false><<<grid, pq::kPipeThreads, 0, s>>>(p, n);
    return (int)cudaGetLastError();
}
