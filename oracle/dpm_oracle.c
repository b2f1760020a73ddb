/*
 * dpm_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, the parity checker).
 *
 * Plain-C restatement of the index-producing ops on DeepPointMap's encoder hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (deeppointmap_b200/) never does.
 *
 * What it follows (paths relative to /root/reference):
 *   - oracle_fps            : Sampler.fps                network/encoder/utils.py:209-270
 *                             (itself a copy of pytorch3d 0.7.4
 *                             sample_farthest_points_naive; start index 0,
 *                             lengths-aware, idx = -1 once K > length)
 *   - oracle_knn            : contract of pytorch3d 0.7.4 knn_points as used at
 *                             network/encoder/utils.py:94,115 (squared L2 by
 *                             direct differences, ascending, only the first
 *                             lengths2 points, zero padding when lengths2 < K)
 *   - oracle_hybrid         : Querier.hybrid_query_t3d   network/encoder/utils.py:112-123
 *   - oracle_ball_query     : contract of pytorch3d 0.7.4 ball_query as used at
 *                             network/encoder/utils.py:100-110 (first K points,
 *                             in index order, with d2 < r2; -1 padded)
 *
 * pytorch3d 0.7.4 itself is not vendored in the reference (requirements.txt:14,
 * README.md:45) and is not installable here, so the kNN contract is restated
 * from its published behaviour and anchored on the reference's own runnable
 * fallback (Querier.hybrid_query, utils.py:75-89): tests/test_oracle_pin.py
 * checks set-equality of neighbour rows against that fallback and index
 * equality of FPS against Sampler.fps.
 *
 * Arithmetic that parity depends on (SURVEY.md section 7 "Hard parts"):
 *   d2 = (dx*dx + dy*dy) + dz*dz in fp32, no FMA contraction (build with
 *   -ffp-contract=off); argmax = first maximum; kNN total order = (d2, index).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float d2_f32(const float *a, const float *b) {
    volatile float dx = a[0] - b[0];
    volatile float dy = a[1] - b[1];
    volatile float dz = a[2] - b[2];
    volatile float xx = dx * dx;
    volatile float yy = dy * dy;
    volatile float zz = dz * dz;
    volatile float s = xx + yy;
    return s + zz;
}

/* non-volatile version for the hot loops; -ffp-contract=off keeps it exact */
static inline float d2_fast(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

/* points: (B, N, D) row-major, D >= 3 (first three columns are xyz)
 * lengths: (B) valid points at the front of each cloud (NULL = N)
 * idx_out: (B, K) int64, -1 padded
 * mind_ws: scratch of N floats per thread (allocated here) */
int oracle_fps(const float *points, int B, int N, int D, const int64_t *lengths, int K,
               int64_t *idx_out) {
    if (B < 0 || N <= 0 || D < 3 || K <= 0) return -1;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *P = points + (size_t)b * N * D;
        int64_t len = lengths ? lengths[b] : N;
        int64_t *out = idx_out + (size_t)b * K;
        for (int k = 0; k < K; ++k) out[k] = -1;
        if (len > N) { rc = -2; continue; }
        if (len <= 0) continue;
        float *mind = (float *)malloc(sizeof(float) * (size_t)len);
        for (int64_t i = 0; i < len; ++i) mind[i] = INFINITY;
        int64_t sel = 0;
        out[0] = 0;
        int64_t kn = len < K ? len : K;
        for (int64_t k = 1; k < kn; ++k) {
            const float sx = P[sel * D], sy = P[sel * D + 1], sz = P[sel * D + 2];
            float best = -1.0f;
            int64_t besti = 0;
            for (int64_t i = 0; i < len; ++i) {
                float d = d2_fast(sx, sy, sz, P[i * D], P[i * D + 1], P[i * D + 2]);
                float m = mind[i];
                m = d < m ? d : m; /* torch.min(dist, closest) */
                mind[i] = m;
                if (m > best) { best = m; besti = i; } /* first maximum */
            }
            sel = besti;
            out[k] = sel;
        }
        free(mind);
    }
    return rc;
}

/* Brute-force K nearest neighbours of p1 (B,S,3-of-D1) in p2[:lengths2] (B,N,3-of-D2).
 * idx (B,S,K) int64 and d2 (B,S,K) fp32, ascending by (d2, idx); slots >= lengths2
 * are zero-filled (pytorch3d convention). */
int oracle_knn(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
               const int64_t *lengths2, int K, int64_t *idx_out, float *d2_out) {
    if (B < 0 || S < 0 || N <= 0 || K <= 0 || D1 < 3 || D2 < 3) return -1;
    if (K > 4096) return -3;
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
        for (int s = 0; s < S; ++s) {
            const float *Q = p1 + ((size_t)b * S + s) * D1;
            const float *P = p2 + (size_t)b * N * D2;
            int64_t len = lengths2 ? lengths2[b] : N;
            if (len > N) len = N;
            const int KK = K;
            float bd[KK]; /* K is bounded by the caller (index_ops.py: K <= 4096) */
            int64_t bi[KK];
            int cnt = 0;
            for (int64_t i = 0; i < len; ++i) {
                float d = d2_fast(Q[0], Q[1], Q[2], P[i * D2], P[i * D2 + 1], P[i * D2 + 2]);
                if (cnt == KK && !(d < bd[KK - 1])) continue; /* ties keep the lower index */
                int pos = cnt < KK ? cnt : KK - 1;
                while (pos > 0 && d < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = i;
                if (cnt < KK) ++cnt;
            }
            int64_t *oi = idx_out + ((size_t)b * S + s) * K;
            float *od = d2_out ? d2_out + ((size_t)b * S + s) * K : NULL;
            for (int k = 0; k < K; ++k) {
                oi[k] = k < cnt ? bi[k] : 0;
                if (od) od[k] = k < cnt ? bd[k] : 0.0f;
            }
        }
    }
    return 0;
}

/* Querier.hybrid_query_t3d: kNN, then every slot with d2 > r2 takes slot 0's index.
 * r2 is the fp32 value the comparison is made against. */
int oracle_hybrid(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                  const int64_t *lengths2, int K, float r2, int64_t *idx_out) {
    float *d2 = (float *)malloc(sizeof(float) * (size_t)B * S * K);
    if (!d2) return -4;
    int rc = oracle_knn(p1, D1, p2, D2, B, S, N, lengths2, K, idx_out, d2);
    if (rc == 0) {
#pragma omp parallel for
        for (int64_t r = 0; r < (int64_t)B * S; ++r) {
            int64_t *oi = idx_out + r * K;
            const float *od = d2 + r * K;
            int64_t first = oi[0];
            for (int k = 0; k < K; ++k)
                if (od[k] > r2) oi[k] = first;
        }
    }
    free(d2);
    return rc;
}

/* pytorch3d ball_query: the first K points of p2[:lengths2] (index order) with
 * d2 < r2; idx -1 / d2 0 padded. */
int oracle_ball_query(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                      const int64_t *lengths2, int K, float r2, int64_t *idx_out, float *d2_out) {
    if (B < 0 || S < 0 || N <= 0 || K <= 0 || D1 < 3 || D2 < 3) return -1;
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
        for (int s = 0; s < S; ++s) {
            const float *Q = p1 + ((size_t)b * S + s) * D1;
            const float *P = p2 + (size_t)b * N * D2;
            int64_t len = lengths2 ? lengths2[b] : N;
            if (len > N) len = N;
            int64_t *oi = idx_out + ((size_t)b * S + s) * K;
            float *od = d2_out ? d2_out + ((size_t)b * S + s) * K : NULL;
            int cnt = 0;
            for (int64_t i = 0; i < len && cnt < K; ++i) {
                float d = d2_fast(Q[0], Q[1], Q[2], P[i * D2], P[i * D2 + 1], P[i * D2 + 2]);
                if (d < r2) {
                    oi[cnt] = i;
                    if (od) od[cnt] = d;
                    ++cnt;
                }
            }
            for (int k = cnt; k < K; ++k) {
                oi[k] = -1;
                if (od) od[k] = 0.0f;
            }
        }
    }
    return 0;
}

/* one exact distance, exported so tests can pin the arithmetic itself */
float oracle_d2(const float *a, const float *b) { return d2_f32(a, b); }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
