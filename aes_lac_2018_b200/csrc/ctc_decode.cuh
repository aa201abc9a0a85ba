// ctc_decode.cuh -- greedy (best-path) CTC decode on the GPU: argmax over the alphabet per frame, collapse
// repeated symbols, drop blanks.  SURVEY.md section 8(f) row 3: the consumer on the other side of the same
// activation tensor.
//
// Replaces the reference's GreedyDecoder.decode (/root/reference/codes/decoder.py:143-160):
//     _, max_probs = torch.max(probs, 2)                       # B x T
//     process_string(): for i < size: c = seq[i]; skip blanks; skip if i != 0 and c == seq[i-1]; keep (c, i)
// which runs a Python loop with one .item() device sync per frame (decoder.py:126-136).
//
// One warp per utterance, 32 frames per iteration.  The tile of 32 x V activations is contiguous in the
// reference's B x T x V layout: it is read with fully coalesced loads (lane + 32*j), issued one tile ahead of
// its use and held in registers, then laid out in shared memory with an odd row stride so that each lane can
// take the argmax of one frame without bank conflicts; keep flags are compacted with ballot/popc.  Each
// activation is read exactly once and only the kept tokens are written: the kernel is HBM-bound (4*V bytes
// per frame).
#pragma once
#include <cuda_runtime.h>

namespace ctcb200 {

struct DecodeParams {
    const float *probs;               // element (b, t, k) at b*stride_b + t*stride_t + k
    long long stride_b, stride_t;
    const int *sizes;                 // DEVICE [B] valid frames per utterance, or nullptr (= T)
    int *tokens;                      // [B][T] kept symbols, first counts[b] entries valid
    int *offsets;                     // [B][T] frame index of each kept symbol, or nullptr
    int *counts;                      // [B]
    int B, T, V, blank;
};

// VCH: the alphabet fits 32*VCH symbols (bounds the register staging: V values per lane per tile)
template <int VCH>
__global__ void __launch_bounds__(32) ctc_greedy_decode_kernel(const DecodeParams P)
{
    extern __shared__ float tile[];                       // [32][RS]
    constexpr int VMAX = 32 * VCH;
    const int lane = threadIdx.x, b = blockIdx.x;
    const int V = P.V, RS = V | 1;                        // odd row stride => conflict-free column walks
    const int size = P.sizes ? min(max(P.sizes[b], 0), P.T) : P.T;
    const float *base = P.probs + (long long)b * P.stride_b;
    int *tok = P.tokens + (long long)b * P.T;
    int *off = P.offsets ? P.offsets + (long long)b * P.T : nullptr;
    const bool dense = (P.stride_t == V);                 // the frames of one utterance are contiguous

    // register staging of the next tile (dense layout): element e = lane + 32*j of the 32*V contiguous floats
    float st[VMAX];
    auto issue = [&](int t0) {
        const float *src = base + (long long)t0 * V + lane;
        if (t0 + 32 <= size) {                            // full tile: no per-element bound
#pragma unroll
            for (int j = 0; j < VMAX; ++j)
                if (j < V) st[j] = __ldg(src + 32 * j);
        } else {
            const int n_el = (size - t0) * V;
#pragma unroll
            for (int j = 0; j < VMAX; ++j)
                st[j] = (j < V && lane + 32 * j < n_el) ? __ldg(src + 32 * j) : 0.f;
        }
    };
    auto stash = [&](int t0) {
        if (dense && (V & 1)) {                           // RS == V: the tile is stored exactly as it lies in HBM
#pragma unroll
            for (int j = 0; j < VMAX; ++j)
                if (j < V) tile[lane + 32 * j] = st[j];
        } else if (dense) {
            int row = lane / V, col = lane % V;
#pragma unroll
            for (int j = 0; j < VMAX; ++j) {
                if (j < V) tile[row * RS + col] = st[j];
                col += 32;
                while (col >= V) { col -= V; ++row; }
            }
        } else {                                          // strided frames (e.g. a T x B x V view): row-by-row copy
            for (int r = 0; r < 32 && t0 + r < size; ++r)
                for (int k = lane; k < V; k += 32) tile[r * RS + k] = base[(long long)(t0 + r) * P.stride_t + k];
        }
    };

    int count = 0, carry = -1;                            // carry = argmax of the previous frame
    if (dense && size > 0) issue(0);
    for (int t0 = 0; t0 < size; t0 += 32) {
        __syncwarp();                                     // previous tile fully consumed
        stash(t0);
        __syncwarp();
        if (dense && t0 + 32 < size) issue(t0 + 32);
        const int t = t0 + lane;
        int best = 0;
        {
            // first maximum wins.  The common case runs a 3-instruction compare/select per symbol; rows that
            // contain a NaN (torch.max treats NaN as maximal, first NaN wins) are redone on a slow path.
            const float *row = tile + lane * RS;
            float bv = row[0];
            bool has_nan = (bv != bv);
#pragma unroll
            for (int k = 1; k < VMAX; ++k) {
                if (k < V) {
                    const float v = row[k];
                    has_nan |= (v != v);
                    const bool gt = v > bv;
                    bv = gt ? v : bv;
                    best = gt ? k : best;
                }
            }
            if (has_nan) {
                best = 0;
                bv = row[0];
                for (int k = 1; k < V; ++k) {
                    const float v = row[k];
                    if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }
                }
            }
        }
        int prev = __shfl_up_sync(0xffffffffu, best, 1);
        if (lane == 0) prev = carry;
        const bool keep = (t < size) && best != P.blank && (t == 0 || best != prev);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            tok[pos] = best;
            if (off) off[pos] = t;
        }
        count += __popc(m);
        carry = __shfl_sync(0xffffffffu, best, 31);
    }
    if (lane == 0) P.counts[b] = count;
}

}  // namespace ctcb200
