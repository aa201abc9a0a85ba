// ctc_combine.cuh -- second half of the BIDIRECTIONAL path used for small batches.
//
// With few utterances in flight the bound is the T-serial chain, and ctc_fused_kernel walks it three times
// (alpha, alpha again inside each chunk, beta).  Beta of a CTC problem is alpha of the time- and
// label-reversed problem, so for small batches the host launches ctc_fused_kernel in sweep-only mode with
// 2 CTAs per utterance -- one on (acts, labels), one on (acts reversed in time, labels reversed) -- which run
// concurrently on different SMs and spill every (rescaled) column as fp64 high words.  This kernel then
// forms, fully in parallel over (utterance, frame), the posteriors and gradient rows:
//     beta_t(s) = alpha'_{T-1-t}(S-1-s)            (both include the emission at t)
//     posterior_t(k) = sum_{s: l'_s = k} alpha^_t(s) beta^_t(s) * 2^(Ea[c(t)] + Eb[c(T-1-t)] - Ea_fin) / (Z^ p~_t(k))
// Latency = one sweep + this kernel instead of three sweeps.  Reference maths: SURVEY.md Appendix C; replaces the
// same reference call as ctc_fused.cuh (codes/engine.py:22).  Checks per frame that the posteriors sum to 1 and
// that forward and reversed partition functions agree; failures flag the utterance for the log-space kernel.
#pragma once
#include "ctc_fused.cuh"

namespace ctcb200 {

struct CombineParams {
    const float *acts;
    long long act_stride_t, act_stride_b;
    float *grads;
    const int *labels, *label_off, *label_len, *act_len, *utt_ids;
    int *status;
    const unsigned *col;              // slots [2*n][T_max][NS][NT]; slot u = forward, slot n+u = reversed
    long long col_stride;
    const int *col_exp;
    int col_exp_stride;
    const double *col_z;
    int n;                            // utterances in this launch
    int V, T_max, B, blank;
    float grad_scale;
    int frames_per_cta;
};

constexpr int kCombineThreads = 128;

constexpr int kCombineWarps = kCombineThreads / 32;

__host__ __device__ inline int combine_smem_bytes(int S_pad, int V)
{
    // per warp: prod[S_pad] floats + ptw[V+2] doubles; shared: lab[S_pad/2], pos[S_pad/2], cnt[V+1], off[V+2] ints
    return kCombineWarps * (S_pad * 4 + (V + 2) * 8) + S_pad * 4 + (2 * V + 8) * 4 + 64;
}

// NS, NT = 32*W, K: layout parameters of the sweep variant that produced the columns
template <int NS, int W, int K>
__global__ void __launch_bounds__(kCombineThreads) ctc_combine_kernel(const CombineParams P)
{
    constexpr int NT = 32 * W, SP = NS * NT, CT = kCombineThreads;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = blockIdx.y;
    const int b = P.utt_ids[u];
    if (P.status[b] & (UTT_INFEASIBLE | UTT_BAD_LABEL)) return;      // cost 0 / gradient 0 already written by the sweep
    const int T = P.act_len[b], L = P.label_len[b], S = 2 * L + 1, V = P.V, blank = P.blank;
    double *ptw_all = (double *)smem;                      // [warps][V+2] p~ of the warp's frame
    float *prod_all = (float *)(ptw_all + kCombineWarps * (V + 2));   // [warps][SP]
    int *lab = (int *)(prod_all + kCombineWarps * SP);     // [SP/2]
    int *pos = lab + SP / 2;                               // [SP/2] label state indices grouped by symbol
    int *cnt = pos + SP / 2;                               // [V+1]
    int *off = cnt + V + 1;                                // [V+2]
    double *ptw = ptw_all + warp * (V + 2);
    float *prod = prod_all + warp * SP;
    const int *lab_g = P.labels + P.label_off[b];
    const float *acts_b = P.acts + (long long)b * P.act_stride_b;
    float *grads_b = P.grads + (long long)b * V;
    const long long gst = (long long)P.B * V;
    const unsigned *colA = P.col + (long long)u * P.col_stride;
    const unsigned *colB = P.col + (long long)(P.n + u) * P.col_stride;
    const int *expA = P.col_exp + (long long)u * P.col_exp_stride;
    const int *expB = P.col_exp + (long long)(P.n + u) * P.col_exp_stride;
    const double *zA = P.col_z + (long long)u * 4, *zB = P.col_z + (long long)(P.n + u) * 4;
    const double zhat = zA[0];
    const int Ea_fin = (int)zA[1];
    const bool z_ok = (zhat > 0.0) && (zhat < INFINITY);
    const double inv_z = z_ok ? (1.0 + 2.0 * 0.7213 * 4.76837158203125e-7) / zhat : 0.0;   // two truncated factors per product

    for (int j = tid; j < L; j += CT) lab[j] = lab_g[j];
    __syncthreads();
    for (int k = tid; k <= V; k += CT) {
        int c = 0;
        if (k < V && k != blank) for (int j = 0; j < L; ++j) c += (lab[j] == k);
        cnt[k] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int k = 0; k <= V; ++k) { off[k] = o; o += cnt[k]; }
        off[V + 1] = o;
    }
    __syncthreads();
    for (int k = tid; k < V; k += CT) {
        int q = off[k];
        if (cnt[k]) for (int j = 0; j < L; ++j) if (lab[j] == k) pos[q++] = 2 * j + 1;
    }
    __syncthreads();

    int bad = 0;
    // forward and reversed sweeps must agree on log Z (each was computed from its own 750-step chain)
    if (z_ok && !(fabs(zA[2] - zB[2]) <= 1e-6 * fmax(1.0, fabs(zA[2])))) bad = 1;

    // one WARP per frame (no block barrier inside the loop): softmax row, products, per-symbol sums, gradient row
    const int t_begin = blockIdx.x * P.frames_per_cta, t_end = min(T, t_begin + P.frames_per_cta);
    for (int t = t_begin + warp; t < t_end; t += kCombineWarps) {
        // p~ of the frame, exactly as the sweeps formed it (same exp_wide, same truncation)
        const float *row = acts_b + (long long)t * P.act_stride_t;
        float m = -INFINITY;
        for (int k = lane; k < V; k += 32) m = fmaxf(m, row[k]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m == -INFINITY) m = 0.f;
        double ssum = 0.0;
        for (int k = lane; k < V; k += 32) {
            const double e = __hiloint2double(__double2hiint(exp_wide(row[k] - m)), 0);
            ptw[k] = e;
            ssum += e;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) ssum += shfl_xor_d(ssum, o);
        const float rinv = 1.f / (float)ssum;
        // products alpha^_t(s) * beta^_t(s), scaled
        const double sc = scalbn(inv_z, expA[t / K] + expB[(T - 1 - t) / K] - Ea_fin);
        const unsigned *ra = colA + (long long)t * SP;
        const unsigned *rb = colB + (long long)(T - 1 - t) * SP;
        float btot = 0.f;
        for (int s = lane; s < S; s += 32) {
            const int sr = S - 1 - s;
            const double av = __hiloint2double((int)ra[(s % NS) * NT + s / NS], 0);
            const double bv = __hiloint2double((int)rb[(sr % NS) * NT + sr / NS], 0);
            const float pr = (float)(av * bv * sc);
            prod[s] = pr;
            if (!(s & 1)) btot += pr;                       // (lane parity == state parity: even lanes own the blanks)
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) btot += __shfl_xor_sync(0xffffffffu, btot, o);
        __syncwarp();
        float psum = 0.f;
        for (int k = lane; k < V; k += 32) {
            float acc;
            if (k == blank) acc = btot;
            else {
                acc = 0.f;
                for (int q = off[k]; q < off[k + 1]; ++q) acc += prod[pos[q]];
            }
            const float pt = (float)ptw[k];
            float post = __fdividef(acc, pt);
            post = (pt > 0.f) ? post : 0.f;
            psum += post;
            grads_b[(long long)t * gst + k] = (pt * rinv - post) * P.grad_scale;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        if (z_ok && !(fabsf(psum - 1.f) <= 7e-6f)) bad = 1;
        __syncwarp();
    }
    bad = __syncthreads_or(bad);
    if (bad && tid == 0) atomicOr(&P.status[b], UTT_RANGE);
}

}  // namespace ctcb200
