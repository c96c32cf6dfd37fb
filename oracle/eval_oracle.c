/* ORACLE (test infrastructure, NOT product code) — plain-C restatement of
 * /root/reference/reid/evaluator/eva_functions.py:134-184 (`evaluate`) and
 * /root/reference/reid/evaluator/attevaluator.py:44-46 (`cosin_dist`).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Built by oracle/Makefile into
 * oracle/_build/liboracle_eval.so.
 *
 * evaluate: per query a *stable* argsort of the distance row (merge sort on
 * (distance, index)), then the reference's loop: drop same-pid&same-cam, skip
 * queries without a match, clipped cumsum -> CMC, AP = sum(cum/(i+1)*match)/num_rel
 * accumulated in double exactly like numpy's float64 path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void merge_sort_idx(const float *d, int32_t *idx, int32_t *tmp, int n)
{
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* stable: take left on ties (left always has the lower original index) */
                if (d[idx[j]] < d[idx[i]]) tmp[k++] = idx[j++]; else tmp[k++] = idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

/* returns number of valid queries (0 => the reference would assert) */
int grl_oracle_evaluate(const float *dist, int nq, int ng,
                        const int64_t *q_pid, const int64_t *g_pid,
                        const int64_t *q_cam, const int64_t *g_cam,
                        int max_rank, float *cmc_out, double *map_out,
                        double *ap_out /* [nq] or NULL, -1 for invalid */)
{
    if (ng < max_rank) max_rank = ng;
    int32_t *idx = (int32_t *)malloc((size_t)ng * 4), *tmp = (int32_t *)malloc((size_t)ng * 4);
    double *hits = (double *)calloc((size_t)max_rank, sizeof(double));
    double ap_sum = 0.0;
    int nvalid = 0;
    for (int q = 0; q < nq; ++q) {
        const float *d = dist + (size_t)q * ng;
        for (int i = 0; i < ng; ++i) idx[i] = i;
        merge_sort_idx(d, idx, tmp, ng);
        int kept = 0, cum = 0, first = -1;
        double ap = 0.0;
        for (int r = 0; r < ng; ++r) {
            int g = idx[r];
            int same = g_pid[g] == q_pid[q];
            if (same && g_cam[g] == q_cam[q]) continue; /* junk */
            if (same) {
                ++cum;
                if (first < 0) first = kept;
                ap += (double)cum / (double)(kept + 1);
            }
            ++kept;
        }
        if (ap_out) ap_out[q] = -1.0;
        if (cum == 0) continue;
        ++nvalid;
        ap /= (double)cum;
        ap_sum += ap;
        if (ap_out) ap_out[q] = ap;
        for (int r = first; r < max_rank; ++r) hits[r] += 1.0;
    }
    for (int r = 0; r < max_rank; ++r) cmc_out[r] = nvalid ? (float)((float)hits[r] / (double)nvalid) : 0.f;
    *map_out = nvalid ? ap_sum / nvalid : 0.0;
    free(idx); free(tmp); free(hits);
    return nvalid;
}

/* cosin_dist: out[q][g] = -sum_k qf[q][k]*gf[g][k], float accumulate in blocks (CPU baseline only) */
void grl_oracle_neg_dot(const float *qf, const float *gf, int nq, int ng, int dim, float *out)
{
    for (int q = 0; q < nq; ++q)
        for (int g = 0; g < ng; ++g) {
            float acc = 0.f;
            const float *a = qf + (size_t)q * dim, *b = gf + (size_t)g * dim;
            for (int k = 0; k < dim; ++k) acc += a[k] * b[k];
            out[(size_t)q * ng + g] = -acc;
        }
}
