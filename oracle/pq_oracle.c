/*
 * pq_oracle.c -- CPU restatement of the reference's calibration arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pytorch-quantity_b200/)
 * may link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker and as
 * the timed CPU baseline.
 *
 * Parity pinning: the reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the
 * unmodified reference executed in the build container (numpy 2.3.5, torch
 * 2.11): the fixtures under tests/golden/, produced by tests/golden/gen_golden.py.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/quantity/common/quantity/).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: an FMA contraction would
 * change the float64 results the reference computes with separate mul/add).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* distribution_collector.py:77-78 -- m = max(m, max(|max x|, |min x|)) in fp32. */
void pqo_absmax_f32(const float *x, size_t n, float *inout_max)
{
    if (n == 0) return;
    float mx = x[0], mn = x[0];
    for (size_t i = 1; i < n; ++i) {
        if (x[i] > mx) mx = x[i];
        if (x[i] < mn) mn = x[i];
    }
    float a = fabsf(mx), b = fabsf(mn);
    float m = a > b ? a : b;
    if (m > *inout_max) *inout_max = m;
}

/* distribution_collector.py:127-135 -- for x != 0:
 *   idx = min((int32)trunc(fl32(|x| / interval)), nbins-1); hist[idx] += 1.
 * The division is an IEEE fp32 division (numpy float32 / float32).  Quotients
 * >= 2^31 are undefined in the reference (numpy's cast yields INT_MIN and the
 * python list index then fails); here they clamp to nbins-1. */
void pqo_hist_f32(const float *x, size_t n, float interval, int32_t *hist, int nbins)
{
    const float top = (float)(nbins - 1);
    for (size_t i = 0; i < n; ++i) {
        float v = x[i];
        if (v != 0.0f) {
            volatile float q = fabsf(v) / interval; /* volatile: force a rounded fp32 quotient */
            int idx = (q >= top) ? (nbins - 1) : (int)q;
            hist[idx] += 1;
        }
    }
}

/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
 * @TYPE@_pairwise_sum; numpy is a third-party dependency of the reference, pinned
 * by this image at 2.3.5): the order np.sum / ndarray.sum use on a contiguous
 * float64 vector.  Verified bit-for-bit against np.sum in tests/test_oracle.py. */
static double pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

double pqo_pairwise_sum(const double *a, long n) { return pairwise_sum(a, n); }

/* quantizer.py:95-96 -- P = hist.astype(float32) / (hist.sum() + 1e-12), float64
 * result under numpy 2 (NEP 50).  `counts` holds the histogram as doubles
 * (int32 counts, or the float64 group sums of pytorch_quantizer.py:434-445). */
void pqo_normalize(const double *counts, int nbins, double *P)
{
    double s = 0.0;
    /* integer-valued addends below 2^53: any order is exact */
    for (int i = 0; i < nbins; ++i) s += counts[i];
    double denom = s + 1e-12;
    for (int i = 0; i < nbins; ++i) P[i] = (double)(float)counts[i] / denom;
}

/* quantizer.py:98-174 -- KL threshold search.  Returns the chosen threshold bin;
 * if kl_out != NULL it receives the nbins-target_bin divergences.  Scratch is
 * allocated here.  Same statement order, same float64 operations. */
int pqo_kl_search(const double *P, int nbins, int target_bin, double *kl_out)
{
    double min_kl = 66666.0;                                        /* :99  */
    double threshold_sum = pairwise_sum(P + target_bin, nbins - target_bin); /* :100 */
    int target_threshold = nbins - 1;                               /* :101 */
    double *t = (double *)malloc(sizeof(double) * nbins);
    double *e = (double *)malloc(sizeof(double) * nbins);
    double *q = (double *)malloc(sizeof(double) * target_bin);
    double *term = (double *)malloc(sizeof(double) * nbins);

    for (int threshold = target_bin; threshold < nbins; ++threshold) { /* :103 */
        memcpy(t, P, sizeof(double) * threshold);                   /* :104 */
        t[threshold - 1] += threshold_sum;                          /* :105 */
        threshold_sum = threshold_sum - P[threshold];               /* :108 */
        for (int i = 0; i < target_bin; ++i) q[i] = 0.0;            /* :110 */
        for (int i = 0; i < threshold; ++i) e[i] = 1e-9;            /* :111 */
        double num_per_bin = (double)threshold / (double)target_bin; /* :112 */

        for (int i = 0; i < target_bin; ++i) {                      /* :114-126 */
            double start = i * num_per_bin;
            double end = start + num_per_bin;
            long left_upper = (long)ceil(start);
            if ((double)left_upper > start) {
                double left_scale = (double)left_upper - start;
                q[i] += left_scale * P[left_upper - 1];
            }
            long right_lower = (long)floor(end);
            if ((double)right_lower < end) {
                double right_scale = end - (double)right_lower;
                q[i] += right_scale * P[right_lower];
            }
            long len = right_lower - left_upper;
            q[i] += pairwise_sum(P + left_upper, len > 0 ? len : 0);
        }

        for (int i = 0; i < target_bin; ++i) {                      /* :128-160 */
            double start = i * num_per_bin;
            double end = start + num_per_bin;
            double count = 1e-12;
            long left_upper = (long)ceil(start);
            double left_scale = 0.0;
            if ((double)left_upper > start) {
                left_scale = (double)left_upper - start;
                if (P[left_upper - 1] != 0) count += left_scale;
            }
            long right_lower = (long)floor(end);
            double right_scale = 0.0;
            if ((double)right_lower < end) {
                right_scale = end - (double)right_lower;
                if (P[right_lower] != 0) count += right_scale;
            }
            for (long j = left_upper; j < right_lower; ++j)
                if (P[j] != 0) count = count + 1;
            double expand_value = q[i] / count;
            if ((double)left_upper > start)
                if (P[left_upper - 1] != 0) e[left_upper - 1] += expand_value * left_scale;
            if ((double)right_lower < end)
                if (P[right_lower] != 0) e[right_lower] += expand_value * right_scale;
            for (long j = left_upper; j < right_lower; ++j)
                if (P[j] != 0) e[j] += expand_value;
        }

        /* compute_kl_divergence, :169-174 */
        long m = 0;
        for (int j = 0; j < threshold; ++j)
            if (t[j] != 0) term[m++] = t[j] * log(t[j] / (e[j] + 1e-12) + 1e-12);
        double kl = pairwise_sum(term, m);
        if (kl_out) kl_out[threshold - target_bin] = kl;
        if (kl < min_kl) {                                          /* :163-165 */
            min_kl = kl;
            target_threshold = threshold;
        }
    }
    free(t); free(e); free(q); free(term);
    return target_threshold;
}

/* new_quantity_op.py:246-257 -- y = clamp(rint(x * 2^bit), -128, 127) / 2^bit.
 * rintf under the default rounding mode is round-half-even == torch.round. */
void pqo_fakequant_f32(const float *x, float *y, size_t n, int bit)
{
    const float s = ldexpf(1.0f, bit);
    for (size_t i = 0; i < n; ++i) {
        float v = rintf(x[i] * s);
        v = v < -128.0f ? -128.0f : (v > 127.0f ? 127.0f : v);
        y[i] = v / s;
    }
}
