// ctc_editdist.cuh -- batched Levenshtein distance between decoded transcripts and references (sm_100a).
//
// SURVEY.md section 8(f) row 3, second half: the reference scores every decoded utterance with
// `decoder.wer(transcript, reference)` / `decoder.cer(...)` (/root/reference/codes/decoder.py:49-78, called from
// codes/metrics.py:118 and test.py:83-84): strings on the host, one python-Levenshtein call per utterance, after a
// Python loop turned the argmax into a string.  Here the hypotheses stay where ctc_greedy_decode_kernel left them
// (device token rows + counts) and one warp per utterance produces the integer distance:
//
//   EDIT_TOKENS  plain Levenshtein over the token ids
//   EDIT_CER     both sides lose their space tokens first           (decoder.py:77  s.replace(' ', ''))
//   EDIT_WER     both sides are cut into words at runs of spaces and the distance is over words
//                (decoder.py:58-66: s.split(), words mapped to integers, Lev.distance over those)
//
// and the normaliser the metric divides by (metrics.py:145-160): len(reference) INCLUDING spaces for CER and
// TOKENS, the number of reference words for WER.  Unit costs (insert = delete = substitute = 1), integers
// throughout, so the result is exact; word identity is exact too (token-by-token comparison, no hashing).
//
// Dynamic programme: rows = hypothesis symbols (serial), columns = reference symbols, lane l owning the C
// consecutive columns [l*C, l*C + C) in registers.  Within a row
//     tmp[j] = min(D[i-1][j] + 1, D[i-1][j-1] + (a_i != b_j))                (vertical / diagonal, independent)
//     D[i][j] = min_{k <= j} (tmp[k] + j - k)                                (horizontal chain = prefix minimum)
// so the only serial part across lanes is a prefix-min of (tmp[k] - k): C local steps, one 5-step warp scan,
// C fix-ups.  Integer work on a few KB per utterance -- latency/issue bound, negligible next to the decode pass
// over T*V floats that feeds it; no roofline claim is made for it.
#pragma once
#include <cuda_runtime.h>

namespace ctcb200 {

enum { EDIT_TOKENS = 0, EDIT_CER = 1, EDIT_WER = 2 };

struct EditParams {
    const int *hyp;                   // [B][hyp_stride] decoded tokens
    long long hyp_stride;
    const int *hyp_len;               // [B]
    const int *ref;                   // flat reference tokens
    const int *ref_off;               // [B] start of each reference in `ref`
    const int *ref_len;               // [B]
    int *dist;                        // [B] out
    int *norm;                        // [B] out (may be null)
    int B, space, mode, max_hyp, max_ref;
};

__host__ __device__ inline int editdist_words_cap(int max_hyp, int max_ref)
{
    return (max_hyp + 1) / 2 + (max_ref + 1) / 2 + 2;      // a word needs a separator: at most ceil(len / 2) words
}
__host__ __device__ inline int editdist_smem_bytes(int max_hyp, int max_ref, int mode)
{
    int n = (max_hyp + 1) + (max_ref + 1);
    if (mode == EDIT_WER) n += 2 * editdist_words_cap(max_hyp, max_ref);
    return n * 4;
}

template <int C>
__global__ void __launch_bounds__(32) ctc_edit_distance_kernel(const EditParams P)
{
    extern __shared__ int esm[];
    const int lane = threadIdx.x, b = blockIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int space = P.space;
    int *hs = esm;                                         // [max_hyp + 1] hypothesis symbols
    int *rs = hs + P.max_hyp + 1;                          // [max_ref + 1] reference symbols
    const int *hg = P.hyp + (long long)b * P.hyp_stride;
    const int *rg = P.ref + P.ref_off[b];
    const int m = min(max(P.hyp_len[b], 0), P.max_hyp);
    const int n = min(max(P.ref_len[b], 0), P.max_ref);
    int ms, ns, normv = n;

    if (P.mode == EDIT_TOKENS) {
        for (int i = lane; i < m; i += 32) hs[i] = hg[i];
        for (int i = lane; i < n; i += 32) rs[i] = rg[i];
        ms = m; ns = n;
    } else if (P.mode == EDIT_CER) {
        auto drop_spaces = [&](const int *g, int len, int *out) -> int {
            int base = 0;
            for (int i0 = 0; i0 < len; i0 += 32) {
                const int i = i0 + lane;
                const int tok = (i < len) ? g[i] : space;
                const bool keep = (tok != space);
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                if (keep) out[base + __popc(mask & lt)] = tok;
                base += __popc(mask);
            }
            return base;
        };
        ms = drop_spaces(hg, m, hs);
        ns = drop_spaces(rg, n, rs);
    } else {
        int *wst = rs + P.max_ref + 1;                     // word start (index into its own token row)
        int *wln = wst + editdist_words_cap(P.max_hyp, P.max_ref);
        auto find_words = [&](const int *g, int len, int wbase) -> int {
            int base = wbase;
            for (int i0 = 0; i0 < len; i0 += 32) {
                const int i = i0 + lane;
                const int tok = (i < len) ? g[i] : space;
                const int prv = (i > 0 && i < len) ? g[i - 1] : space;
                const bool st = (tok != space) && (prv == space);
                const unsigned mask = __ballot_sync(0xffffffffu, st);
                if (st) {
                    int e = i + 1;
                    while (e < len && g[e] != space) ++e;
                    const int pos = base + __popc(mask & lt);
                    wst[pos] = i;
                    wln[pos] = e - i;
                }
                base += __popc(mask);
            }
            return base - wbase;
        };
        ms = find_words(hg, m, 0);
        ns = find_words(rg, n, ms);
        normv = ns;
        __syncwarp();
        // word -> integer: the index of the first identical word in (hypothesis words, reference words)
        for (int w = lane; w < ms + ns; w += 32) {
            const int *gw = (w < ms ? hg : rg) + wst[w];
            const int lw = wln[w];
            int id = w;
            for (int u = 0; u < w; ++u) {
                if (wln[u] != lw) continue;
                const int *gu = (u < ms ? hg : rg) + wst[u];
                bool eq = true;
                for (int k = 0; k < lw; ++k)
                    if (gu[k] != gw[k]) { eq = false; break; }
                if (eq) { id = u; break; }
            }
            if (w < ms) hs[w] = id; else rs[w - ms] = id;
        }
    }
    __syncwarp();

    constexpr int INF = 1 << 29;
    int prev[C], rsym[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = lane * C + c;
        prev[c] = j;                                       // D[0][j]
        rsym[c] = (j >= 1 && j <= ns) ? rs[j - 1] : (int)0x80000000;   // never equal to a symbol
    }
    for (int i = 0; i < ms; ++i) {
        const int a = hs[i];
        int up = __shfl_up_sync(0xffffffffu, prev[C - 1], 1);          // D[i][lane*C - 1]
        if (lane == 0) up = INF;
#pragma unroll
        for (int c = C - 1; c >= 1; --c)                   // descending: prev[c-1] is still row i
            prev[c] = min(prev[c] + 1, prev[c - 1] + (a != rsym[c] ? 1 : 0));
        prev[0] = (lane == 0) ? i + 1 : min(prev[0] + 1, up + (a != rsym[0] ? 1 : 0));
#pragma unroll
        for (int c = 1; c < C; ++c) prev[c] = min(prev[c], prev[c - 1] + 1);   // horizontal chain inside the lane
        int v = prev[C - 1] - (lane * C + C - 1);          // ... and across lanes: prefix-min of (value - column)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = min(v, t);
        }
        int carry = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) carry = INF;
        carry += lane * C;
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = min(prev[c], carry + c);
    }
    int res = 0;
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (c == ns % C) res = prev[c];
    res = __shfl_sync(0xffffffffu, res, ns / C);
    if (lane == 0) {
        P.dist[b] = res;
        if (P.norm) P.norm[b] = normv;
    }
}

}  // namespace ctcb200
