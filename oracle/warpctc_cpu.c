/*
 * fp32 CPU restatement of warp-ctc's OpenMP CTC path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  The product (aes_lac_2018_b200) never links it.
 *
 * Restates: the computation behind `criterion(out, targets, out_sizes, target_sizes)` at
 * /root/reference/codes/engine.py:22 and /root/reference/codes/metrics.py:51 when the activations
 * live on the CPU, i.e. third-party SeanNaren/warp-ctc (unpinned HEAD, cloned by
 * /root/reference/docker/Dockerfile:52-66; NOT present under /root/reference, not installable
 * offline).  Written from the published algorithm as specified in SURVEY.md Appendix C / section 8c:
 *   (1) softmax to fp32 *probabilities*, std::log(p) taken at every use;
 *   (2) log-space alpha over the blank-extended labels, only inside the reachable band;
 *   (3) beta includes the emission at t; posterior = exp(alpha+beta - log p - logZ);
 *   (4) grad = p when the per-symbol accumulator is -inf, exactly 0.0, or p == 0;
 *   (5) L + repeats > T  =>  cost 0, gradient rows untouched;
 *   (6) cost = -logZ from the forward pass;
 *   (7) frames t >= T_b keep the caller's zeros;
 *   (8) OpenMP parallel-for over the minibatch, fp32 throughout.
 * PARITY UNPINNED by the reference (no tests there; dependency absent): pinned instead against
 * upstream's known-answer vectors and torch float64 (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK 0
#define ORACLE_BAD_ARG 2
#define ORACLE_NO_MEM 1

static inline float lse2(float a, float b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return log1pf(expf(-fabsf(a - b))) + (a > b ? a : b);
}

/* probabilities of one utterance: probs[t*V + k], rows t < T; input row stride = B*V */
static void utt_softmax(const float *acts, int B, int b, int T, int V, float *probs)
{
    for (int t = 0; t < T; ++t) {
        const float *row = acts + ((size_t)t * B + b) * V;
        float mx = row[0];
        for (int k = 1; k < V; ++k) if (row[k] > mx) mx = row[k];
        float den = 0.f;
        for (int k = 0; k < V; ++k) { float e = expf(row[k] - mx); probs[t * V + k] = e; den += e; }
        for (int k = 0; k < V; ++k) probs[t * V + k] /= den;
    }
}

/* Returns cost of one utterance, writes its gradient rows (t < T) unless grads == NULL. */
static float utt_cost_and_grad(const float *acts, float *grads, int B, int b, int T, int V,
                               const int *lab, int L, int blank, float *scratch)
{
    const int S = 2 * L + 1;
    int repeats = 0;
    for (int i = 1; i < L; ++i) repeats += (lab[i] == lab[i - 1]);
    if (T <= 0 || L + repeats > T) return 0.f;             /* (5) */

    /* scratch layout */
    float *probs = scratch;                                /* T*V */
    float *alpha = probs + (size_t)T * V;                  /* T*S */
    float *beta = alpha + (size_t)T * S;                   /* 2*S rolling */
    float *acc = beta + 2 * S;                             /* V */
    int *ext = (int *)(acc + V);                           /* S */
    int *need_before = ext + S;                            /* S: min frames to stand on s */
    int *need_after = need_before + S;                     /* S: min further frames to finish */

    for (int s = 0; s < S; ++s) ext[s] = (s & 1) ? lab[s >> 1] : blank;
    need_before[0] = 1;
    if (S > 1) need_before[1] = 1;
    for (int s = 2; s < S; ++s) {
        int skip = (ext[s] != blank) && (ext[s] != ext[s - 2]);
        need_before[s] = 1 + (skip ? need_before[s - 2] : need_before[s - 1]);
    }
    need_after[S - 1] = 0;
    if (S > 1) need_after[S - 2] = 0;
    for (int s = S - 3; s >= 0; --s) {
        int skip = (ext[s] != blank) && (ext[s + 2] != ext[s]);
        int via1 = need_after[s + 1], via2 = skip ? need_after[s + 2] : via1;
        need_after[s] = 1 + (via2 < via1 ? via2 : via1);
    }

    utt_softmax(acts, B, b, T, V, probs);                  /* (1) */

    for (size_t i = 0; i < (size_t)T * S; ++i) alpha[i] = -INFINITY;

    /* ---- alpha, banded (2) ---- */
    alpha[0] = logf(probs[blank]);
    if (S > 1) alpha[1] = logf(probs[ext[1]]);
    for (int t = 1; t < T; ++t) {
        const float *pr = probs + (size_t)t * V;
        const float *prev = alpha + (size_t)(t - 1) * S;
        float *cur = alpha + (size_t)t * S;
        for (int s = 0; s < S; ++s) {
            if (need_before[s] > t + 1 || need_after[s] > T - 1 - t) continue;
            float a = prev[s];
            if (s >= 1) a = lse2(a, prev[s - 1]);
            if (s >= 2 && ext[s] != blank && ext[s] != ext[s - 2]) a = lse2(a, prev[s - 2]);
            cur[s] = a + logf(pr[ext[s]]);
        }
    }
    float logz = -INFINITY;
    {
        const float *last = alpha + (size_t)(T - 1) * S;
        for (int s = (S > 1 ? S - 2 : 0); s < S; ++s)
            if (need_before[s] <= T) logz = lse2(logz, last[s]);
    }
    const float cost = -logz;                              /* (6) */
    if (grads == NULL) return cost;

    /* ---- beta (rolling) + gradient (3)(4) ---- */
    float *bcur = beta, *bnext = beta + S;
    for (int t = T - 1; t >= 0; --t) {
        const float *pr = probs + (size_t)t * V;
        const float *al = alpha + (size_t)t * S;
        for (int k = 0; k < V; ++k) acc[k] = -INFINITY;
        for (int s = 0; s < S; ++s) {
            bcur[s] = -INFINITY;
            if (need_before[s] > t + 1 || need_after[s] > T - 1 - t) continue;
            float v;
            if (t == T - 1) {
                v = 0.f;  /* log 1: final states S-1, S-2 (band already restricts to them) */
            } else {
                v = bnext[s];
                if (s + 1 < S) v = lse2(v, bnext[s + 1]);
                if (s + 2 < S && ext[s] != blank && ext[s + 2] != ext[s]) v = lse2(v, bnext[s + 2]);
            }
            v += logf(pr[ext[s]]);
            bcur[s] = v;
            acc[ext[s]] = lse2(acc[ext[s]], al[s] + v);
        }
        float *g = grads + ((size_t)t * B + b) * V;
        for (int k = 0; k < V; ++k) {
            float p = pr[k];
            if (acc[k] == 0.0f || acc[k] == -INFINITY || p == 0.0f)
                g[k] = p;
            else
                g[k] = p - expf(acc[k] - logf(p) - logz);
        }
        float *tmp = bcur; bcur = bnext; bnext = tmp;
    }
    return cost;
}

static size_t utt_scratch_floats(int T, int V, int L)
{
    size_t S = 2 * (size_t)L + 1;
    return (size_t)T * V + (size_t)T * S + 2 * S + V + 3 * S + 16;
}

/*
 * C entry point (the restatement's analogue of warp-ctc's compute_ctc_loss with CTC_CPU):
 * activations [T_max, B, V] dense time-major; gradients same shape, PRE-ZEROED by the caller, or NULL
 * for score-only; flat_labels concatenated; costs[B] written.  num_threads <= 0 => OpenMP default.
 */
int oracle_warpctc_cpu(const float *activations, float *gradients, const int *flat_labels,
                       const int *label_lengths, const int *input_lengths, int alphabet_size,
                       int minibatch, float *costs, int blank_label, int num_threads)
{
    if (!activations || !flat_labels || !label_lengths || !input_lengths || !costs ||
        alphabet_size <= 0 || minibatch <= 0)
        return ORACLE_BAD_ARG;
    const int B = minibatch, V = alphabet_size;
    size_t *lab_off = (size_t *)malloc(sizeof(size_t) * (size_t)B);
    if (!lab_off) return ORACLE_NO_MEM;
    size_t off = 0, max_scratch = 0;
    for (int b = 0; b < B; ++b) {
        lab_off[b] = off;
        off += (size_t)label_lengths[b];
        size_t n = utt_scratch_floats(input_lengths[b], V, label_lengths[b]);
        if (n > max_scratch) max_scratch = n;
    }
    int status = ORACLE_OK;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel
    {
        float *scratch = (float *)malloc(max_scratch * sizeof(float));
        if (!scratch) {
#pragma omp critical
            status = ORACLE_NO_MEM;
        }
#pragma omp barrier
        if (status == ORACLE_OK) {
#pragma omp for schedule(dynamic, 1)
            for (int b = 0; b < B; ++b) {
                costs[b] = utt_cost_and_grad(activations, gradients, B, b, input_lengths[b], V,
                                             flat_labels + lab_off[b], label_lengths[b],
                                             blank_label, scratch);
            }
        }
        free(scratch);
    }
    free(lab_off);
    return status;
}

int oracle_warpctc_cpu_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
