// requant_microbench.cu -- issue-rate microbenchmark for candidate formulations of the GEMM epilogue's
// RightShift -> saturate -> BiasAdd -> saturate -> pack chain (development tool).
// Each thread keeps 16 accumulators in registers, applies the chain ITER times (the input is perturbed with
// the previous output so nothing can be hoisted) and the kernel reports cycles per 16-element chunk per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o requant_microbench requant_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t pack4(int y0, int y1, int y2, int y3)
{
    uint32_t t, w;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(y3), "r"(y2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(y1), "r"(y0), "r"(t));
    return w;
}
__device__ __forceinline__ uint32_t pack2_s16(int hi, int lo)
{
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
__device__ __forceinline__ int mulhi(int a, int b) { return __mulhi(a, b); }

struct P { int half, sh, mulsh, lo; int bias[16]; uint32_t bias2[8]; int c[16], h[16], l[16]; };

// A: current chain (all ALU pipe)
__device__ __forceinline__ void chain_a(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int r = (acc[j] + p.half + (acc[j] >> 31)) >> p.sh;
        r = max(-128, min(127, r));
        y[j] = __viaddmax_s32(r, p.bias[j], p.lo);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// C: sign and shift through IMAD.HI (fma pipe): t = acc + (hi(acc * 1) + half);  r = hi(t * 2^(32 - sh))
__device__ __forceinline__ void chain_c(const int (&acc)[16], const P &p, uint32_t (&out)[4], int one)
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int s = mulhi(acc[j], one) + p.half;      // IMAD.HI with addend
        int t = acc[j] + s;
        int r = mulhi(t, p.mulsh);
        r = max(-128, min(127, r));
        y[j] = __viaddmax_s32(r, p.bias[j], p.lo);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// D: 32-bit shift as in A, then the saturate / bias / saturate tail on packed s16x2
__device__ __forceinline__ void chain_d(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int r0 = (acc[2 * j] + p.half + (acc[2 * j] >> 31)) >> p.sh;
        int r1 = (acc[2 * j + 1] + p.half + (acc[2 * j + 1] >> 31)) >> p.sh;
        uint32_t v = pack2_s16(r1, r0);
        v = __vmaxs2(__vmins2(v, 0x007f007fu), 0xff80ff80u);
        v = __viaddmax_s16x2(v, p.bias2[j], (uint32_t)p.lo);
        h[j] = __vmins2(v, 0x007f007fu);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __byte_perm(h[2 * j], h[2 * j + 1], 0x6420);
}
// E: C's IMAD.HI front end + D's s16x2 tail
__device__ __forceinline__ void chain_e(const int (&acc)[16], const P &p, uint32_t (&out)[4], int one)
{
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int t0 = acc[2 * j] + (mulhi(acc[2 * j], one) + p.half);
        int t1 = acc[2 * j + 1] + (mulhi(acc[2 * j + 1], one) + p.half);
        uint32_t v = pack2_s16(mulhi(t1, p.mulsh), mulhi(t0, p.mulsh));
        v = __vmaxs2(__vmins2(v, 0x007f007fu), 0xff80ff80u);
        v = __viaddmax_s16x2(v, p.bias2[j], (uint32_t)p.lo);
        h[j] = __vmins2(v, 0x007f007fu);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __byte_perm(h[2 * j], h[2 * j + 1], 0x6420);
}
// F: A with the +half through IMAD (fma pipe), rest ALU
__device__ __forceinline__ void chain_f(const int (&acc)[16], const P &p, uint32_t (&out)[4], int one)
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int t = acc[j] * one + p.half;
        int r = (t + (acc[j] >> 31)) >> p.sh;
        r = max(-128, min(127, r));
        y[j] = __viaddmax_s32(r, p.bias[j], p.lo);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// G: bias folded into the rounding add (per-channel c), both saturations as min(h) / max(l) (three constant vectors)
__device__ __forceinline__ void chain_g(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int t = (acc[j] + p.c[j] + (acc[j] >> 31)) >> p.sh;
        y[j] = max(min(t, p.h[j]), p.l[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// H: G with a fused ReLU: min(h) and max(., 0) in one VIMNMX.RELU
__device__ __forceinline__ void chain_h(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int t = (acc[j] + p.c[j] + (acc[j] >> 31)) >> p.sh;
        y[j] = __vimin_s32_relu(t, p.h[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// I: two constant vectors (c, b): max(b - 128, min(b + 127, t)) as two VIADDMNMX
__device__ __forceinline__ void chain_i(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int t = (acc[j] + p.c[j] + (acc[j] >> 31)) >> p.sh;
        y[j] = __viaddmax_s32(p.bias[j], -128, __viaddmin_s32(p.bias[j], 127, t));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// J: I with a fused ReLU: one VIADDMNMX.RELU
__device__ __forceinline__ void chain_j(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int t = (acc[j] + p.c[j] + (acc[j] >> 31)) >> p.sh;
        y[j] = __viaddmin_s32_relu(p.bias[j], 127, t);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// K: only the shift part (VIADD, LEA.HI.SX32, SHF) + pack: the floor of any formulation that keeps this rounding
__device__ __forceinline__ void chain_k(const int (&acc)[16], const P &p, uint32_t (&out)[4])
{
    int y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = (acc[j] + p.c[j] + (acc[j] >> 31)) >> p.sh;
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = pack4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
}
// (a float tail -- r after the integer shift -> float via the magic-number add -- is NOT exact in general and is
// not a candidate for the product) -- r (after the integer shift) -> float via the magic-number add is NOT exact in general; here
// only to see the fma-pipe rate of a FADD/FMNMX mix (not a candidate for the product)
template <int V>
__global__ void __launch_bounds__(512, 1) bench(const int *in, uint32_t *out, P p, int iters, long long *cycles, int one)
{
    int acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = in[threadIdx.x * 16 + j];
    uint32_t o[4] = {0, 0, 0, 0};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0) chain_a(acc, p, o);
        if (V == 2) chain_c(acc, p, o, one);
        if (V == 3) chain_d(acc, p, o);
        if (V == 4) chain_e(acc, p, o, one);
        if (V == 5) chain_f(acc, p, o, one);
        if (V == 6) chain_g(acc, p, o);
        if (V == 7) chain_h(acc, p, o);
        if (V == 8) chain_i(acc, p, o);
        if (V == 9) chain_j(acc, p, o);
        if (V == 10) chain_k(acc, p, o);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += (int)o[j & 3];                   // 1 extra op per element, the same for every variant
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = o[0] ^ o[1] ^ o[2] ^ o[3];
}

int main()
{
    int *in; uint32_t *out; long long *cyc;
    cudaMalloc(&in, 512 * 16 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    int h[512 * 16];
    for (int i = 0; i < 512 * 16; ++i) h[i] = (i * 2654435761u) >> 8;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    P p; p.sh = 9; p.half = 1 << 8; p.mulsh = 1 << (32 - 9); p.lo = -128;
    for (int j = 0; j < 16; ++j) p.bias[j] = j * 7 - 50;
    for (int j = 0; j < 16; ++j) {
        p.c[j] = p.half + p.bias[j] * (1 << p.sh);
        p.h[j] = p.bias[j] + 127 < 127 ? p.bias[j] + 127 : 127;
        p.l[j] = p.bias[j] - 128 > -128 ? p.bias[j] - 128 : -128;
    }
    for (int j = 0; j < 8; ++j) p.bias2[j] = ((uint32_t)(uint16_t)(int16_t)p.bias[2 * j + 1] << 16) | (uint16_t)(int16_t)p.bias[2 * j];
    const int iters = 20000;
    const char *names[11] = {"A current (ALU only)", "", "C IMAD.HI sign+shift", "D s16x2 tail", "E IMAD.HI + s16x2", "F IMAD +half",
                             "G folded bias, min h max l", "H folded bias, min.relu", "I folded, 2x VIADDMNMX", "J folded, VIADDMNMX.RELU",
                             "K shift + pack only"};
    for (int v = 0; v < 11; ++v) {
        if (v == 1) continue;
        for (int rep = 0; rep < 2; ++rep) {
            switch (v) {
                case 0: bench<0><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 2: bench<2><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 3: bench<3><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 4: bench<4><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 5: bench<5><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 6: bench<6><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 7: bench<7><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 8: bench<8><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 9: bench<9><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
                case 10: bench<10><<<148, 512>>>(in, out, p, iters, cyc, 1); break;
            }
            cudaDeviceSynchronize();
        }
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // 16 warps per SM = 4 per SMSP; cycles per 16-element chunk per warp, and per warp-element per SMSP
        printf("%-24s %8.1f cycles / chunk / warp   %6.2f cycles per warp-element per SMSP  (%s)\n", names[v],
               (double)c / iters, (double)c / iters / 16.0 / 4.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
