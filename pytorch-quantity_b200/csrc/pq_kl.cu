// pq_kl.cu -- KL-divergence threshold search (subsystem 2).
//
// Replaces Quantizer.normalize_distribution / threshold_distribution / compute_kl_divergence
// (common/quantity/quantizer.py:95-174).  One CTA evaluates one candidate threshold T of one
// tensor; grid = (1920 candidates, k tensors).  Everything is fp64 and follows the reference's
// operation order so that the divergences agree to the last few ulps and the arg-min (hence
// the fractional bit) is identical:
//   * products and sums use __dmul_rn/__dadd_rn/__ddiv_rn -- never contracted into FMAs;
//   * every np.sum of the reference is reproduced with numpy's pairwise order (blocks of
//     <=128 with 8 strided accumulators, recursive halving above);
//   * the running tail `threshold_sum - distribution[threshold]` (:108) is a sequential chain
//     computed once per tensor in kl_prepare_kernel.
// Not bandwidth bound (16 KB of input per tensor); reported as microseconds per tensor.
//
// The bin count is the reference's INTERVAL_NUM (tools/configs.yml:23), a run-time value `nbins`; the kernels are
// compiled for two capacities CAP (2048: the default and everything below; 8192 = PQ_HIST_BINS_MAX above) that
// only size the per-thread register arrays and the leaf tables.  Shared memory is dynamic (16 * nbins bytes).
#include "pq_common.cuh"

namespace pq {

constexpr int kTarget = PQ_KL_TARGET_BIN;       // 128
constexpr int kKlThreads = 256;

// workspace layout per tensor, in doubles: P[nbins] | tail[nbins - 128] | KL[nbins - 128]
struct KlGeom {
    int nbins, cand;            // cand = nbins - 128
    int ws_tail, ws_kl, ws_doubles;
};

inline KlGeom kl_geom(int nbins)
{
    KlGeom g;
    g.nbins = nbins;
    g.cand = nbins - kTarget;
    g.ws_tail = nbins;
    g.ws_kl = nbins + g.cand;
    g.ws_doubles = 3 * nbins;
    return g;
}

// numpy pairwise_sum leaf: n <= 128
__device__ __forceinline__ double np_leaf_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 = __dadd_rn(r0, a[i + 0]); r1 = __dadd_rn(r1, a[i + 1]);
        r2 = __dadd_rn(r2, a[i + 2]); r3 = __dadd_rn(r3, a[i + 3]);
        r4 = __dadd_rn(r4, a[i + 4]); r5 = __dadd_rn(r5, a[i + 5]);
        r6 = __dadd_rn(r6, a[i + 6]); r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// Leaves of numpy's recursion for a length-n vector, in left-to-right order.
constexpr int kMaxLeaves = 256;                 // n <= 8192: at most 8192 / 64 + a few leaves
__device__ int np_build_leaves(int n, int *leaf_off, int *leaf_len)
{
    int stack_off[16], stack_len[16], sp = 0, count = 0;
    stack_off[0] = 0; stack_len[0] = n; sp = 1;
    while (sp) {
        --sp;
        const int off = stack_off[sp], len = stack_len[sp];
        if (len <= 128) {
            leaf_off[count] = off; leaf_len[count] = len; ++count;
        } else {
            int n2 = len / 2; n2 -= n2 % 8;
            stack_off[sp] = off + n2; stack_len[sp] = len - n2; ++sp;   // right, popped second
            stack_off[sp] = off; stack_len[sp] = n2; ++sp;             // left, popped first
        }
    }
    return count;
}

// Combine the leaf sums in numpy's recursion order: node = left + right.
__device__ double np_combine(int n, const double *leaf_sum, int &next)
{
    if (n <= 128) return leaf_sum[next++];
    int n2 = n / 2; n2 -= n2 % 8;
    const double l = np_combine(n2, leaf_sum, next);
    const double r = np_combine(n - n2, leaf_sum, next);
    return __dadd_rn(l, r);
}

// Block-cooperative numpy-order sum of a[0..n) held in shared memory.  All threads call it;
// the result is valid in thread 0.  s_off/s_len/s_sum: kMaxLeaves entries each, s_cnt: 1 int.
__device__ double np_block_sum(const double *a, int n, int *s_off, int *s_len, double *s_sum, int *s_cnt)
{
    if (threadIdx.x == 0) *s_cnt = np_build_leaves(n, s_off, s_len);
    __syncthreads();
    const int leaves = *s_cnt;
    if ((int)threadIdx.x < leaves) s_sum[threadIdx.x] = np_leaf_sum(a + s_off[threadIdx.x], s_len[threadIdx.x]);
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0) { int next = 0; total = np_combine(n, s_sum, next); }
    return total;
}

// ---- per tensor: P = float32(counts) / (sum + 1e-12); tail[T] for T = 128..nbins-1 ----------
__global__ void __launch_bounds__(kKlThreads)
kl_prepare_kernel(const double *__restrict__ counts, double *__restrict__ workspace, const KlGeom geo)
{
    extern __shared__ double s_dyn[];                      // [nbins]
    double *s_p = s_dyn;
    __shared__ double s_red[kKlThreads / 32];
    __shared__ int s_off[kMaxLeaves], s_len[kMaxLeaves], s_cnt;
    __shared__ double s_sum[kMaxLeaves];
    __shared__ double s_total;
    const int kBins = geo.nbins, kWsP = 0, kWsTail = geo.ws_tail;
    const double *c = counts + (size_t)blockIdx.x * kBins;
    double *ws = workspace + (size_t)blockIdx.x * geo.ws_doubles;

    // hist.sum(): integer-valued addends, exact in any order below 2^53
    double part = 0.0;
    for (int i = threadIdx.x; i < kBins; i += kKlThreads) part += c[i];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kKlThreads / 32; ++w) s += s_red[w];
        s_total = __dadd_rn(s, 1e-12);                                   // quantizer.py:96
    }
    __syncthreads();
    const double denom = s_total;
    for (int i = threadIdx.x; i < kBins; i += kKlThreads) {
        const double p = __ddiv_rn((double)__double2float_rn(c[i]), denom);   // astype(float32) / float64
        s_p[i] = p;
        ws[kWsP + i] = p;
    }
    __syncthreads();
    // threshold_sum = distribution[128:].sum()  (:100), then the running subtraction (:108)
    const double tail0 = np_block_sum(s_p + kTarget, kBins - kTarget, s_off, s_len, s_sum, &s_cnt);
    if (threadIdx.x == 0) {
        double t = tail0;
        for (int T = kTarget; T < kBins; ++T) {
            ws[kWsTail + T - kTarget] = t;
            t = __dsub_rn(t, s_p[T]);
        }
    }
}

// ---- per (candidate T, tensor): the divergence of quantizer.py:103-161 + :169-174 ----------
template <int CAP>
__global__ void __launch_bounds__(kKlThreads)
kl_candidate_kernel(double *__restrict__ workspace, double *__restrict__ kl_out, const KlGeom geo)
{
    constexpr int kPerThread = CAP / kKlThreads;           // consecutive bins per thread (8 for CAP = 2048)
    extern __shared__ double s_dyn[];                      // [nbins] P, [nbins] terms
    double *s_p = s_dyn;
    double *s_term = s_dyn + geo.nbins;
    __shared__ double s_ev[kTarget];
    __shared__ int s_warp_cnt[kKlThreads / 32];
    __shared__ int s_off[kMaxLeaves], s_len[kMaxLeaves], s_cnt;
    __shared__ double s_sum[kMaxLeaves];
    const int kWsP = 0, kWsTail = geo.ws_tail, kWsKl = geo.ws_kl, kCand = geo.cand;

    const int T = kTarget + blockIdx.x;
    double *ws = workspace + (size_t)blockIdx.y * geo.ws_doubles;
    for (int i = threadIdx.x; i < T; i += kKlThreads) s_p[i] = ws[kWsP + i];
    const double tail = ws[kWsTail + blockIdx.x];
    const double npb = (double)T / (double)kTarget;                      // :112, exact dyadic
    __syncthreads();

    if (threadIdx.x < kTarget) {                                         // :114-126 and :128-148
        const int i = threadIdx.x;
        const double start = __dmul_rn((double)i, npb);
        const double end = __dadd_rn(start, npb);
        const int lu = (int)ceil(start);
        const int rl = (int)floor(end);
        double q = 0.0, count = 1e-12;
        if ((double)lu > start) {
            const double ls = __dsub_rn((double)lu, start);
            q = __dadd_rn(q, __dmul_rn(ls, s_p[lu - 1]));
            if (s_p[lu - 1] != 0.0) count = __dadd_rn(count, ls);
        }
        if ((double)rl < end) {
            const double rs = __dsub_rn(end, (double)rl);
            q = __dadd_rn(q, __dmul_rn(rs, s_p[rl]));
            if (s_p[rl] != 0.0) count = __dadd_rn(count, rs);
        }
        const int len = rl - lu;
        q = __dadd_rn(q, np_leaf_sum(s_p + lu, len > 0 ? len : 0));      // slice length <= 16
        for (int j = lu; j < rl; ++j)
            if (s_p[j] != 0.0) count = __dadd_rn(count, 1.0);
        s_ev[i] = __ddiv_rn(q, count);                                   // expand_value, :149
    }
    __syncthreads();

    // expand + KL terms; each thread owns 8 consecutive bins so the compaction keeps bin order
    double term[kPerThread];
    int nz = 0;
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
        const int j = threadIdx.x * kPerThread + u;
        if (j < T) {
            const double pj = s_p[j];
            const int i = (j * kTarget) / T;                             // the bin whose start <= j
            const double start = __dmul_rn((double)i, npb);
            const double end = __dadd_rn(start, npb);
            double e = 1e-9;                                             // :111
            if (pj != 0.0) {
                if ((double)(j + 1) <= end) {
                    e = __dadd_rn(e, s_ev[i]);                           // interior bin, :159-160
                } else {                                                 // j = floor(end): split bin
                    const double rs = __dsub_rn(end, (double)j);         // bin i, right part :154-156
                    const double ls = __dsub_rn((double)(j + 1), end);   // bin i+1, left part :151-153
                    e = __dadd_rn(e, __dmul_rn(s_ev[i], rs));
                    e = __dadd_rn(e, __dmul_rn(s_ev[i + 1], ls));
                }
            }
            const double t = (j == T - 1) ? __dadd_rn(pj, tail) : pj;    // :104-105
            if (t != 0.0) {                                              // :172-174
                const double ratio = __dadd_rn(__ddiv_rn(t, __dadd_rn(e, 1e-12)), 1e-12);
                term[nz++] = __dmul_rn(t, log(ratio));
            }
        }
    }
    // ordered compaction: exclusive scan of per-thread counts
    int incl = nz;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)(threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) s_warp_cnt[threadIdx.x >> 5] = incl;
    __syncthreads();
    int base = 0, m = 0;
    for (int w = 0; w < kKlThreads / 32; ++w) {
        if (w < (int)(threadIdx.x >> 5)) base += s_warp_cnt[w];
        m += s_warp_cnt[w];
    }
    int pos = base + incl - nz;
#pragma unroll
    for (int u = 0; u < kPerThread; ++u)
        if (u < nz) s_term[pos + u] = term[u];
    __syncthreads();

    const double kl = np_block_sum(s_term, m, s_off, s_len, s_sum, &s_cnt);   // np.sum, :173
    if (threadIdx.x == 0) {
        ws[kWsKl + blockIdx.x] = kl;
        if (kl_out) kl_out[(size_t)blockIdx.y * kCand + blockIdx.x] = kl;
    }
}

// ---- per tensor: first strict minimum below 66666, default nbins - 1  (:99-101, :163-165) -------
__global__ void __launch_bounds__(kKlThreads)
kl_argmin_kernel(const double *__restrict__ workspace, int *__restrict__ threshold, const KlGeom geo)
{
    __shared__ double s_v[kKlThreads];
    __shared__ int s_i[kKlThreads];
    const int kCand = geo.cand, kBins = geo.nbins;
    const double *kl = workspace + (size_t)blockIdx.x * geo.ws_doubles + geo.ws_kl;
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    int bi = kCand;
    for (int i = threadIdx.x; i < kCand; i += kKlThreads) {
        const double v = kl[i];
        if (v < best) { best = v; bi = i; }          // NaN never compares below
    }
    s_v[threadIdx.x] = best; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kKlThreads / 2; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double v = s_v[threadIdx.x + o];
            const int i = s_i[threadIdx.x + o];
            if (v < s_v[threadIdx.x] || (v == s_v[threadIdx.x] && i < s_i[threadIdx.x])) {
                s_v[threadIdx.x] = v; s_i[threadIdx.x] = i;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        threshold[blockIdx.x] = (s_v[0] < 66666.0) ? kTarget + s_i[0] : kBins - 1;
}

}  // namespace pq

extern "C" size_t pq_kl_workspace_doubles(void) { return 3 * PQ_HIST_BINS; }

extern "C" size_t pq_kl_workspace_doubles_n(int nbins) { return nbins > 0 ? (size_t)3 * nbins : 0; }

extern "C" int pq_kl_search_n_f64(const double *counts, int k, int nbins, double *workspace, double *kl,
                                  int *threshold, pq_stream_t stream)
{
    if (k == 0) return PQ_OK;
    if (k < 0 || !counts || !workspace || !threshold) return PQ_EINVAL;
    if (nbins <= PQ_KL_TARGET_BIN) return PQ_EINVAL;       // the reference's loop range(128, nbins) would be empty
    if (k > 65535 || nbins > PQ_HIST_BINS_MAX) return PQ_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const pq::KlGeom geo = pq::kl_geom(nbins);
    const size_t smem_p = (size_t)nbins * 8, smem_c = (size_t)nbins * 16;
    static bool attr_set = false;
    if (!attr_set) {
        PQ_CUDA_TRY(cudaFuncSetAttribute(pq::kl_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PQ_HIST_BINS_MAX * 8));
        PQ_CUDA_TRY(cudaFuncSetAttribute(pq::kl_candidate_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PQ_HIST_BINS_MAX * 16));
        attr_set = true;
    }
    pq::kl_prepare_kernel<<<k, pq::kKlThreads, smem_p, s>>>(counts, workspace, geo);
    if (nbins <= 2048)
        pq::kl_candidate_kernel<2048><<<dim3(geo.cand, k), pq::kKlThreads, smem_c, s>>>(workspace, kl, geo);
    else
        pq::kl_candidate_kernel<8192><<<dim3(geo.cand, k), pq::kKlThreads, smem_c, s>>>(workspace, kl, geo);
    pq::kl_argmin_kernel<<<k, pq::kKlThreads, 0, s>>>(workspace, threshold, geo);
    return (int)cudaGetLastError();
}

extern "C" int pq_kl_search_f64(const double *counts, int k, double *workspace, double *kl,
                                int *threshold, pq_stream_t stream)
{
    return pq_kl_search_n_f64(counts, k, PQ_HIST_BINS, workspace, kl, threshold, stream);
}
