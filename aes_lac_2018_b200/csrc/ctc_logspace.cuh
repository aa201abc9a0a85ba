// ctc_logspace.cuh -- robust fp64 LOG-SPACE forward/backward for the utterances the fast kernel flags.
//
// ctc_fused_kernel keeps alpha/beta in the linear domain with one power-of-two scale per column.  That is
// exact and fast while the states of a column stay within the fp64 exponent range of each other, but a
// confident-and-wrong model (e.g. a blank-collapsed network asked for a long transcript: cost = L x margin,
// thousands of nats) spreads a column over more than e^700 and the states that carry the result underflow.
// The fast kernel detects this (Z^ = 0 -> CTC_B200_UTT_INF_COST, or its forward/backward consistency check
// fails -> CTC_B200_UTT_RANGE).  The detour is taken ON THE DEVICE: this kernel is enqueued right behind the fast
// kernels of every call (same stream, no host round trip, so non-blocking calls get it too); its persistent CTAs
// scan the status words of the batch and redo exactly the flagged utterances (none, almost always: the scan of
// 8192 status words costs a few microseconds).  Same maths as
// warp-ctc's compute_alpha_kernel / compute_betas_and_grad_kernel (log-space, beta includes the emission,
// SURVEY.md Appendix C) but in float64 throughout, so it has no range limit and agrees with the oracle to
// ~1e-12.  One CTA per utterance; alpha rows go to a caller-provided HBM slot (8*T*S bytes per utterance in
// flight), beta is rolled in shared memory; per-symbol log-sum-exp over the precomputed position lists.
// Throughput is secondary here (it is ~20x slower than the fused kernel) -- correctness on hostile inputs is
// the point.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ctc_fused.cuh"

namespace ctcb200 {

struct LogParams {
    const float *acts;
    long long act_stride_t, act_stride_b;
    float *grads;                     // dense [T_max][B][V] or nullptr
    const int *labels, *label_off, *label_len, *act_len;
    int *queue;                       // work counter over the batch (zeroed before the launch)
    int n_utts;                       // utterances 0 .. n_utts-1 are scanned; those flagged RANGE / INF_COST are redone
    float *costs;
    int *status;
    double *alpha_ws;                 // [gridDim.x] slots (one per persistent CTA)
    long long slot_stride;            // doubles per slot (>= T_max * S_max)
    int V, T_max, B, blank, S_max;
    float grad_scale;
};

constexpr int kLogThreads = 256;

__host__ __device__ inline int logspace_smem_bytes(int S_max, int V)
{
    // prev[S], cur[S], ab[S] doubles; lp[V], accs[V] doubles; red[64] doubles; ext[S], pos[S/2+1], off[V+2] ints
    const int S = S_max + 2;
    return (3 * S + 2 * (V + 1) + 64) * 8 + (S + S / 2 + 2 + V + 2) * 4 + 64;
}

__device__ __forceinline__ double lse2(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = fmax(a, b);
    return m + log1p(exp(-fabs(a - b)));
}

// block-wide (kLogThreads) max and sum helpers through shared memory `red`
__device__ __forceinline__ double block_max(double v, double *red, int tid)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double m = red[0];
    for (int w = 1; w < kLogThreads / 32; ++w) m = fmax(m, red[w]);
    return m;
}
__device__ __forceinline__ double block_sum(double v, double *red, int tid)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kLogThreads / 32; ++w) s += red[w];
    return s;
}

// one utterance, whole CTA; returns the CTC_B200_UTT_* bits of the result
__device__ __forceinline__ int logspace_utterance(const LogParams &P, int b, unsigned char *smem)
{
    const int tid = threadIdx.x, NT = kLogThreads;
    const int V = P.V, blank = P.blank;
    const int T = P.act_len[b], L = P.label_len[b], S = 2 * L + 1;
    const int SA = P.S_max + 2;
    double *prev = (double *)smem;                 // [SA]
    double *cur = prev + SA;                       // [SA]
    double *ab = cur + SA;                         // [SA]
    double *lp = ab + SA;                          // [V+1]
    double *accs = lp + (V + 1);                   // [V+1]
    double *red = accs + (V + 1);                  // [64]
    int *ext = (int *)(red + 64);                  // [SA]
    int *pos = ext + SA;                           // [SA/2+2] label state indices grouped by symbol
    int *off = pos + SA / 2 + 2;                   // [V+2]
    const int *lab_g = P.labels + P.label_off[b];
    const float *acts_b = P.acts + (long long)b * P.act_stride_b;
    float *grads_b = P.grads ? P.grads + (long long)b * V : nullptr;
    const long long gst = (long long)P.B * V;
    double *aws = P.alpha_ws + (long long)blockIdx.x * P.slot_stride;

    for (int s = tid; s < S; s += NT) ext[s] = (s & 1) ? lab_g[s >> 1] : blank;
    __syncthreads();
    // per-symbol position lists (ascending state index), symbol `blank` excluded (handled by a block reduction)
    for (int k = tid; k <= V; k += NT) {
        int c = 0;
        if (k < V && k != blank) for (int j = 0; j < L; ++j) c += (ext[2 * j + 1] == k);
        accs[k] = (double)c;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int k = 0; k <= V; ++k) { off[k] = o; o += (int)accs[k]; }
        off[V + 1] = o;
    }
    __syncthreads();
    for (int k = tid; k < V; k += NT) {
        int q = off[k];
        if (k != blank) for (int j = 0; j < L; ++j) if (ext[2 * j + 1] == k) pos[q++] = 2 * j + 1;
    }
    __syncthreads();

    // log-softmax of row t into lp[] (float64), by the whole block
    auto row_logsoftmax = [&](int t) {
        const float *row = acts_b + (long long)t * P.act_stride_t;
        double m = -INFINITY;
        for (int k = tid; k < V; k += NT) m = fmax(m, (double)row[k]);
        m = block_max(m, red, tid);
        if (m == -INFINITY) m = 0.0;
        double s = 0.0;
        for (int k = tid; k < V; k += NT) s += exp((double)row[k] - m);
        s = block_sum(s, red, tid);
        const double ls = log(s);
        for (int k = tid; k < V; k += NT) {
            // warp-ctc form: log of the softmax probability; a probability that underflows fp64 is exactly 0
            const double d = (double)row[k] - m;
            lp[k] = (d < -745.0) ? -INFINITY : d - ls;
        }
        __syncthreads();
    };

    // ---------------- forward ----------------
    for (int s = tid; s < S; s += NT) prev[s] = -INFINITY;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        row_logsoftmax(t);
        for (int s = tid; s < S; s += NT) {
            double a;
            if (t == 0) {
                a = (s <= 1) ? 0.0 : -INFINITY;
            } else {
                a = prev[s];
                if (s >= 1) a = lse2(a, prev[s - 1]);
                if (s >= 2 && (s & 1) && ext[s] != ext[s - 2]) a = lse2(a, prev[s - 2]);
            }
            a += lp[ext[s]];
            cur[s] = a;
            aws[(long long)t * S + s] = a;
        }
        __syncthreads();
        double *tmp = prev; prev = cur; cur = tmp;
    }
    double logz = prev[S - 1];
    if (S > 1) logz = lse2(logz, prev[S - 2]);
    __syncthreads();
    int ustat = 0;
    if (logz == -INFINITY) ustat = UTT_INF_COST;
    else if (!(logz == logz)) ustat = UTT_RANGE;          // NaN activations
    if (tid == 0) P.costs[b] = (float)(-logz);
    if (!grads_b) return ustat;

    // ---------------- backward + gradient ----------------
    for (int t = T - 1; t >= 0; --t) {
        row_logsoftmax(t);
        for (int s = tid; s < S; s += NT) {
            double v;
            if (t == T - 1) {
                v = (s >= S - 2) ? 0.0 : -INFINITY;
            } else {
                v = prev[s];
                if (s + 1 < S) v = lse2(v, prev[s + 1]);
                if (s + 2 < S && (s & 1) && ext[s + 2] != ext[s]) v = lse2(v, prev[s + 2]);
            }
            v += lp[ext[s]];
            cur[s] = v;
            ab[s] = aws[(long long)t * S + s] + v;
        }
        __syncthreads();
        // blank: log-sum-exp over the even states by two block reductions
        double bm = -INFINITY;
        for (int s = 2 * tid; s < S; s += 2 * NT) bm = fmax(bm, ab[s]);
        bm = block_max(bm, red, tid);
        double bs = 0.0;
        if (bm != -INFINITY) for (int s = 2 * tid; s < S; s += 2 * NT) bs += exp(ab[s] - bm);
        bs = block_sum(bs, red, tid);
        // labels: thread k walks the positions of symbol k
        for (int k = tid; k < V; k += NT) {
            double acc;
            if (k == blank) {
                acc = (bm == -INFINITY) ? -INFINITY : bm + log(bs);
            } else {
                acc = -INFINITY;
                for (int q = off[k]; q < off[k + 1]; ++q) acc = lse2(acc, ab[pos[q]]);
            }
            const double lpk = lp[k];
            const double p = exp(lpk);
            double post = 0.0;
            if (acc != -INFINITY && lpk != -INFINITY && logz != -INFINITY) post = exp(acc - lpk - logz);
            grads_b[(long long)t * gst + k] = (float)((p - post) * (double)P.grad_scale);
        }
        __syncthreads();
        double *tmp = prev; prev = cur; cur = tmp;
    }
    for (int t = T + (tid >> 5); t < P.T_max; t += NT / 32)
        for (int k = (tid & 31); k < V; k += 32) grads_b[(long long)t * gst + k] = 0.f;
    return ustat;
}

__global__ void __launch_bounds__(kLogThreads) ctc_logspace_kernel(const LogParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_item;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(P.queue, 1);
        __syncthreads();
        const int b = s_item;
        __syncthreads();
        if (b >= P.n_utts) break;
        const int st = P.status[b];
        if (!(st & (UTT_RANGE | UTT_INF_COST)) || (st & UTT_BAD_LABEL)) continue;
        const int ustat = logspace_utterance(P, b, smem);
        __syncthreads();
        if (threadIdx.x == 0) P.status[b] = ustat | UTT_LOGSPACE | (st & UTT_WIDE);
    }
}

}  // namespace ctcb200
