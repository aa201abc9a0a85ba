// ctc_warp32.cuh -- the fp32 throughput kernel of the sm_100a CTC engine: one warp per utterance like ctc_warp.cuh,
// but the T-serial alpha / beta recursion runs in SINGLE precision with PER-LANE BLOCK EXPONENTS.
//
// Same result as ctc_warp_kernel / ctc_fused_kernel (cost_b and d cost_b / d acts of the reference's
// `criterion(out, targets, out_sizes, target_sizes)`, reference codes/engine.py:22, codes/metrics.py:51; upstream
// warp-ctc compute_alpha_kernel + compute_betas_and_grad_kernel; maths in SURVEY.md Appendix C).
//
// Why: the fp64 warp kernel is bound by issue slots and by the latency of its dependent fp64 chains at 8-12 resident
// warps per SM (248 instructions per utterance-frame, 60 of them fp64 at half issue rate, 27 more
// only to pair 32-bit high words with a zero low word).  In fp32 the recursion issues at full rate with 4-cycle
// latency, the state needs half the registers (16 warps per SM at NS = 8), every table value is a plain float (no
// pairing, no conversions) and the neighbour exchange is one 32-bit shuffle.
//
// Why it is allowed: SURVEY.md Appendix D scheme F (fp32, ONE scale per column) fails on dynamic range -- inside one
// column alpha^ spans more than the e^87 of a float.  Here every lane (NS consecutive states) carries its own binary
// exponent, chosen once per chunk from the largest magnitude among itself and the lanes mass can reach it from within
// the chunk, and neighbour values cross a lane boundary multiplied by the exact power of two between the two frames.
// tests/proto_f32_bfp.py models exactly this arithmetic on the CPU: <= 6e-7 max |dgrad| against the float64 oracle on
// random and trained-model-like (peaky) activations up to T = 1500, L = 400 once the posteriors of a chunk are
// normalised by the chunk's own mass check (below); confident-and-wrong transcripts with costs above ~1000 nats leave
// the range, fail the check and take the fp64 log-space detour like any other flagged utterance.
//
//  * p~ domain: p~_t(k) = exp(a_t(k) - ref_t) with ref_t = ceil(max_k a_t(k)) (one F2I + REDUX.MAX + I2F per row: any
//    reference in [max, max + 1) keeps p~ in (e^-1, 1]).  Nothing on the chain can overflow: a column only grows by
//    the <= 3-term sums, 3^K per chunk.
//  * lane l: float x[NS], int e: true value = x * 2^e.  Chunk start: e = (max exponent over this lane and the W - 1
//    lanes below [alpha] / above [beta]) - target, then made 64-Lipschitz in the flow direction so that the
//    neighbour factor 2^(e_nb - e) <= 2^64 never overflows.
//  * checkpoint = the fp32 column + the 32 exponents (one more 128-byte row per chunk).
//  * backward: the checkpoint column of lane l is scaled by 2^(ea_l + eb_l - ez) (Z^ = mz * 2^ez), so the products
//    alpha * tb are posteriors in units of mz; the scaled column lives in the frame 2^(ez - eb_l), i.e. its neighbour
//    factor is 2^(eb_l - eb_{l-1}).
//  * mass checks: q(t) = sum_s alpha(t, s) * tb(t, s) = Z^ at every frame.  It is formed at the FIRST and the LAST
//    frame of every chunk.  The first frame's value is (a) DIVIDED OUT of the chunk's posteriors, which removes the
//    rounding drift of the two T-step fp32 product chains (up to ~2e-5 at T = 1500), and (b) compared with the
//    previous chunk's: the beta recursion is linear with positive terms, so mass it loses to underflow anywhere in the
//    chunk never arrives at the first frame.  The last frame's value is compared with the first: mass the recomputed
//    alpha recursion loses (underflow in a lane frame that is fixed for the chunk while a narrow band of live states
//    sweeps through it, overflow of the scaled column) never arrives there.  Either deficit above 4e-6 flags the
//    utterance.  A check at the first frame alone is NOT enough in fp32 (found by the parity fuzz: sigma = 4 logits with
//    T = L + 12 gave a wrong gradient at frames in mid-chunk with a clean first-frame check).
//  * the blank gradient of a frame is minus the sum of the other gradients of the row (sum_k p = sum_k posterior = 1).
//
// Everything else (prologue, product slots grouped by symbol, conflict-free gather, cp.async staging of the backward
// operands, persistent CTAs with an atomic queue) is the organisation of ctc_warp.cuh.
#pragma once
#include <type_traits>

#include "ctc_warp.cuh"

namespace ctcb200 {

constexpr int kW32TargetA = 100;        // forward: group max at 2^100 (growth 3^8 < 2^13 per chunk; range 2^226 below)
constexpr int kW32TargetB = 40;         // backward: alpha_scaled * tb ~ posterior, the two ranges are reciprocal
constexpr int kW32Lip = 64;
constexpr int kW32Dead = -(1 << 28);


struct Warp32Layout {
    int PS;
    int off_prod, off_lab, off_slot, off_cnt, off_off, off_stg, total;
};

__host__ __device__ inline Warp32Layout make_warp32_layout(int NS, int K, int VCH)
{
    Warp32Layout l;
    const int LP = 16 * NS;
    l.PS = LP + 96;                      // labels + segment padding + one dump slot per lane
    int o = 0;
    l.off_prod = o;                      // [K][PS] floats: alpha*tb products of a chunk, grouped by symbol
    l.off_lab = o;                       // [LP] ints   (prologue only: aliases the product rows)
    l.off_slot = o + LP * 4;             // [LP] ints + [32 VCH] masks (prologue only)
    o += K * l.PS * 4;
    l.off_cnt = o;  o += 32 * VCH * 4;
    l.off_off = o;  o += 32 * VCH * 4;
    o = (o + 127) & ~127;
    l.off_stg = o;  o += ((NS + 1) * 32 + K * VCH * 32 + 32) * 4;   // column + exponents | p~ rows | 1/s
    l.total = (o + 15) & ~15;
    return l;
}

// words (4 bytes) of workspace per resident CTA: checkpoint columns with their exponent row, p~ image, 1/s per frame
__host__ __device__ inline long long warp32_slot_words(int NS, int K, int VCH, int T_max)
{
    const long long nC = warp_max_chunks(K, T_max);
    return nC * 32LL * (NS + 1) + (long long)T_max * 32LL * VCH + (((long long)T_max + 31) & ~31LL);
}

__device__ __forceinline__ float exp2i(int n) { return __int_as_float((n + 127) << 23); }   // 2^n, n in [-126, 127]

// Packed single precision (sm_100: FADD2 / FMUL2 / FFMA2 work on an aligned register pair, one issue slot for two
// IEEE fp32 operations).  The kernel is bound by issue slots, the recursion is 40 % of its instructions.
typedef unsigned long long f32x2;
// which variants use the packed recursion: bit NS / 2.  Measured against the scalar form (B = 8192, T = 750, kernel-only,
// M utt/s scalar -> packed): NS = 2: 8.47 -> 8.42, 4: 6.58 -> 6.50, 6: 5.40 -> 5.38, 8: 4.57 -> 4.73, 10: 3.96 -> 3.81,
// 12: 3.33 -> 3.50, 14: 3.02 -> 3.22, 16: 2.57 -> 2.23 (the register pairs push it into spills)
#ifndef CTC_W32_PACKED_MASK
#define CTC_W32_PACKED_MASK 0x00d0      /* NS = 8, 12, 14 */
#endif
// the same for two-slice alphabets (V = 32 .. 63; V = 43, their own register caps): NS = 4: 4.52 -> 4.56, 6: 3.78 -> 3.81,
// 8: 3.08 -> 2.96, 10: 2.62 -> 2.74, 12: 2.30 -> 2.26, 14: 1.98 -> 1.66, 16: 1.64 -> 1.73
#ifndef CTC_W32_PACKED_MASK2
#define CTC_W32_PACKED_MASK2 0x0120     /* NS = 10, 16 */
#endif
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// x[] *= 2^d (d clamped to +-252).  One exact factor when every lane's |d| <= 126 (warp-uniform test: almost always),
// two otherwise.
template <int NS>
__device__ __forceinline__ void scale32(float (&x)[NS], int d)
{
    d = max(-252, min(d, 252));
    if (__all_sync(kFull, (unsigned)(d + 126) <= 252u)) {
        const float f = exp2i(d);
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] *= f;
    } else {
        const int h = d >> 1;
        const float f1 = exp2i(h), f2 = exp2i(d - h);
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] = (x[i] * f1) * f2;
    }
}

// Per-lane block exponent for the next chunk.  UP: mass flows towards higher lanes (alpha), else towards lower (beta).
// W = window of lanes (this one and W - 1 in the direction mass comes from) whose magnitudes the exponent must hold.
// Returns the factor that brings the neighbour's values into this lane's frame (0 on the boundary lane).
template <int NS, int W, bool UP>
__device__ __forceinline__ float rescale32(float (&x)[NS], int &e, int target, int lane)
{
    unsigned key = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) key = max(key, __float_as_uint(x[i]));
    int M = key ? e + (int)(key >> 23) : kW32Dead;           // biased absolute exponent of the lane's max
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {                        // (out-of-range lanes get their own value back)
        const int o = UP ? __shfl_up_sync(kFull, M, d) : __shfl_down_sync(kFull, M, d);
        M = max(M, o);
    }
    int en = M - (target + 127);
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
        const int o = UP ? __shfl_up_sync(kFull, en, d) : __shfl_down_sync(kFull, en, d);
        en = max(en, o - kW32Lip * d);
    }
    if (en < -(1 << 27)) en = e;                             // nothing alive within reach: keep the frame
    scale32<NS>(x, e - en);
    e = en;
    const int enb = UP ? __shfl_up_sync(kFull, e, 1) : __shfl_down_sync(kFull, e, 1);
    const int df = min(enb - e, kW32Lip);
    const bool edge = UP ? (lane == 0) : (lane == 31);
    return (edge || df < -126) ? 0.f : exp2i(df);
}

template <int NS, int K, int VCH, int MAXR>
__global__ void __maxnreg__(MAXR) ctc_warp32_kernel(const FusedParams P)
{
    static_assert(NS % 2 == 0 && NS >= 2 && NS <= 16, "NS must be even, <= 16");
    static_assert(K == 4 || K == 8 || K == 16, "chunk length");
    static_assert(VCH == 1 || VCH == 2, "alphabet slices");
    constexpr int NL = NS / 2, LP = 16 * NS, CW = (NS + 1) * 32;
    // lanes mass can arrive from within one chunk: ceil(2K / NS); the window is the next power of two above it
    constexpr int REACH = (2 * K + NS - 1) / NS;
    constexpr int WIN = REACH >= 16 ? 32 : REACH >= 8 ? 16 : REACH >= 4 ? 8 : REACH >= 2 ? 4 : 2;
    constexpr float kCheckTol = 1e-4f;            // drift of the two T-step product chains against Z^ (divided out)
    constexpr float kJumpTol = 4e-6f;             // mass of a chunk's first frame against the previous chunk's, of its last frame
                                                  // against its first
    constexpr float L2E = 1.4426950408889634f;

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const int V = P.V, blank = P.blank;
    const Warp32Layout lay = make_warp32_layout(NS, K, VCH);
    const int PS = lay.PS;
    float *prod = (float *)(smem + lay.off_prod);
    int *lab_s = (int *)(smem + lay.off_lab);
    int *slot_s = (int *)(smem + lay.off_slot);
    int *cnt_s = (int *)(smem + lay.off_cnt);
    int *off_s = (int *)(smem + lay.off_off);
    const long long gst = (long long)P.B * V;               // gradient row stride (dense)
    const bool want_grad = (P.grads != nullptr);
    const int nCmax = warp_max_chunks(K, P.T_max);
    int slot;
    if (outside_sm_range(P.sm_lo, P.sm_hi, P.queue, P.n_slots, slot)) return;
    float *ckw = (float *)P.ckpt + (long long)slot * (P.ckpt_stride * 2);     // [nCmax][NS + 1][32] column, exponents
    float *imgw = ckw + (long long)nCmax * CW;                                      // [T_max][VCH][32] p~
    float *invw = imgw + (long long)P.T_max * VCH * 32;                             // [T_max] 1 / s_t
    const int bl = blank & 31, bs = blank >> 5;
    bool colok[VCH], wr[VCH];
#pragma unroll
    for (int v = 0; v < VCH; ++v) {
        colok[v] = (lane + 32 * v < V);
        wr[v] = colok[v] && (lane + 32 * v != blank);
    }

    for (int round = 0;; ++round) {
        int item;
        if (P.queue != nullptr) {
            item = 0;
            if (lane == 0) item = atomicAdd(P.queue, 1);
            item = __shfl_sync(kFull, item, 0);
        } else {
            item = (round == 0) ? (int)blockIdx.x : P.n_items;
        }
        if (item >= P.n_items) break;
        const int b = P.utt_ids[item];
        const int T = P.act_len[b];
        const int L = P.label_len[b];
        const int S = 2 * L + 1;
        const int *lab_g = P.labels + P.label_off[b];
        const float *acts_b = P.acts + (long long)b * P.act_stride_b;
        float *grads_b = want_grad ? P.grads + (long long)b * V : nullptr;

        // ---- labels -> shared, repeats, validity ----
        __syncwarp();                                       // previous utterance's readers of the aliased rows are done
        int rep = 0, bad = 0;
        for (int j = lane; j < LP; j += 32) {
            int v = -1;
            if (j < L) {
                v = lab_g[j];
                if (v < 0 || v >= V || v == blank) { bad = 1; v = -1; }
                else if (j > 0 && lab_g[j - 1] == v) rep++;
            }
            lab_s[j] = v;
        }
        rep = __reduce_add_sync(kFull, rep);
        bad = __any_sync(kFull, bad);
        __syncwarp();
        int ustat = 0;
        if (bad) ustat |= UTT_BAD_LABEL;
        if (T <= 0 || L + rep > T) ustat |= UTT_INFEASIBLE;
        if (ustat) {                                        // cost 0, gradient 0 (warp-ctc CPU convention)
            if (lane == 0) { P.costs[b] = 0.f; P.status[b] = ustat; }
            if (want_grad)
                for (int t = 0; t < P.T_max; ++t)
                    for (int k = lane; k < V; k += 32) grads_b[(long long)t * gst + k] = 0.f;
            continue;
        }

        // ---- per-thread label constants ----
        const int j0 = lane * NL;
        int lsrc[NL];                                       // symbol of label j0+jj; padding labels point at the last pad
                                                            // lane, whose p~ is 0
        float msk[NL + 1];                                  // 1 if the skip INTO label j0+jj is allowed
#pragma unroll
        for (int jj = 0; jj <= NL; ++jj) {
            const int j = j0 + jj;
            const int cur = (j < LP) ? lab_s[j] : -1;
            const int prv = (j >= 1 && j - 1 < LP) ? lab_s[j - 1] : -1;
            if (jj < NL) lsrc[jj] = (cur < 0) ? (32 * VCH - 1) : cur;
            msk[jj] = (cur >= 0 && j >= 1 && cur != prv) ? 1.f : 0.f;
        }

        // ---- product slots (see ctc_warp.cuh): products of one frame are stored grouped by symbol, segment starts
        // padded to distinct banks, slots inside a segment assigned greedily so that the stores of one round spread
        // over the banks ----
        int sl[NL];
        int kcnt[VCH], koff[VCH];
        if (want_grad) {
#pragma unroll
            for (int v = 0; v < VCH; ++v) {
                const int k = lane + 32 * v;
                int c = 0;
                if (k < V && k != blank)
                    for (int j = 0; j < L; ++j) c += (lab_s[j] == k);
                cnt_s[k] = c;
            }
            __syncwarp();
            if (lane == 0) {
                unsigned used = 0u;
                int cur = 0;
                for (int k = 0; k < 32 * VCH; ++k) {
                    if ((k & 31) == 0) used = 0u;
                    int o = cur;
                    if (cnt_s[k]) {
                        while ((used >> (o & 31)) & 1u) ++o;
                        used |= 1u << (o & 31);
                        cur = o + cnt_s[k];
                    }
                    off_s[k] = o;
                }
                if (cur > PS - 32) {                        // padded segments do not fit: plain prefix sums
                    int o = 0;
                    for (int k = 0; k < 32 * VCH; ++k) { off_s[k] = o; o += cnt_s[k]; }
                }
            }
            __syncwarp();
            unsigned *free_s = (unsigned *)(slot_s + LP);   // [32 * VCH] free-slot masks (prologue only)
            int big = 0;
#pragma unroll
            for (int v = 0; v < VCH; ++v) {
                const int k = lane + 32 * v;
                kcnt[v] = cnt_s[k];
                koff[v] = off_s[k];
                free_s[k] = (kcnt[v] >= 32) ? 0xffffffffu : ((1u << kcnt[v]) - 1u);
                big |= (kcnt[v] > 32);
            }
            big = __any_sync(kFull, big);
            __syncwarp();
            if (!big) {
                if (lane == 0) {
                    unsigned usedb[NL];
#pragma unroll
                    for (int jj = 0; jj < NL; ++jj) usedb[jj] = 0u;
                    for (int jb = 0; jb < L; jb += NL) {
                        if ((jb & (32 * NL - 1)) == 0) {
#pragma unroll
                            for (int jj = 0; jj < NL; ++jj) usedb[jj] = 0u;
                        }
#pragma unroll
                        for (int jj = 0; jj < NL; ++jj) {
                            const int j = jb + jj;
                            if (j < L) {
                                const int k = lab_s[j];
                                const unsigned fr = free_s[k];
                                const int o = off_s[k];
                                const unsigned banks = __funnelshift_l(fr, fr, o & 31);
                                const unsigned okb = banks & ~usedb[jj];
                                const int bank = __ffs(okb ? okb : banks) - 1;
                                const int r = (bank - o) & 31;
                                free_s[k] = fr & ~(1u << r);
                                usedb[jj] |= 1u << bank;
                                slot_s[j] = o + r;
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < VCH; ++v) {
                    const int k = lane + 32 * v;
                    int q = koff[v];
                    if (kcnt[v])
                        for (int j = 0; j < L; ++j)
                            if (lab_s[j] == k) slot_s[j] = q++;
                }
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int v = 0; v < VCH; ++v) { kcnt[v] = 0; koff[v] = 0; }
        }
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int j = j0 + jj;
            sl[jj] = (want_grad && j < L) ? slot_s[j] : PS - 32 + lane;
        }
        __syncwarp();

        const int nfull = T / K, ntail = T - nfull * K;

        // Raw activations: lane = symbol; the rows of a full chunk are loaded one chunk ahead of their use.
        float xr[K][VCH];
        auto load_rows = [&](auto tag, int t0) {
            constexpr int KK = decltype(tag)::value;
            const float *src = acts_b + (long long)t0 * P.act_stride_t + lane;
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
#pragma unroll
                for (int v = 0; v < VCH; ++v) xr[tt][v] = colok[v] ? __ldg(src + 32 * v) : -INFINITY;
                src += P.act_stride_t;
            }
        };
        auto lookup = [&](const float (&row)[VCH], int src) -> float {
            float v = __shfl_sync(kFull, row[0], src);
            if (VCH == 2) {
                const float v1 = __shfl_sync(kFull, row[VCH - 1], src);
                v = (src & 32) ? v1 : v;
            }
            return v;
        };
        // one alpha step in place (descending i keeps the old neighbours intact).  fu brings the lower lane's top state
        // into this lane's frame.
        // Packed form.  Lane-local labels Lb[j] = a[2j+1] and blanks Bk[j] = a[2j] are paired (j, j + H2), H2 = NL / 2: the
        // "previous label" vector of pair m is then pair m - 1 itself (no repacking), only pair 0 takes the neighbour
        // lane's value.  Same operations in the same order per element as the scalar form: bit-identical results.
        constexpr int H2 = NL / 2;
        constexpr bool PK = (((VCH == 1 ? CTC_W32_PACKED_MASK : CTC_W32_PACKED_MASK2) >> NL) & 1) != 0;   // packed recursion for this variant (measured per NS, one-slice alphabets)
        f32x2 mskp[H2 > 0 ? H2 : 1], msk1p[H2 > 0 ? H2 : 1];   // (msk[m], msk[m+H2]) and the same one label up (beta)
#pragma unroll
        for (int m = 0; m < H2; ++m) { mskp[m] = pk2(msk[m], msk[m + H2]); msk1p[m] = pk2(msk[m + 1], msk[m + H2 + 1]); }
        auto alpha_step = [&](float (&a)[NS], const float (&row)[VCH], float pb, float fu) {
            if constexpr (PK) {
            const float up1 = __shfl_up_sync(kFull, a[NS - 1], 1) * fu;
            float pl[NL];
#pragma unroll
            for (int jj = 0; jj < NL; ++jj) pl[jj] = lookup(row, lsrc[jj]);
            const f32x2 pbb = pk2(pb, pb);
            f32x2 nL[H2 > 0 ? H2 : 1], nB[H2 > 0 ? H2 : 1];
#pragma unroll
            for (int m = 0; m < H2; ++m) {
                const f32x2 PL = pk2(a[2 * m + 1], a[2 * (m + H2) + 1]), PB = pk2(a[2 * m], a[2 * (m + H2)]);
                const f32x2 SL = (m == 0) ? pk2(up1, a[2 * (H2 - 1) + 1]) : pk2(a[2 * (m - 1) + 1], a[2 * (m + H2 - 1) + 1]);
                nL[m] = mul2(fma2(mskp[m], SL, add2(PL, PB)), pk2(pl[m], pl[m + H2]));
                nB[m] = mul2(add2(PB, SL), pbb);
            }
            float oL = 0.f, oB = 0.f;
            if (NL & 1) {                                   // odd label count: the last label / blank stay scalar
                constexpr int j = NL - 1;
                const float p2 = (j >= 1) ? a[2 * j - 1] : up1;
                oL = fmaf(msk[j], p2, a[2 * j + 1] + a[2 * j]) * pl[j];
                oB = (a[2 * j] + p2) * pb;
            }
#pragma unroll
            for (int m = 0; m < H2; ++m) {
                a[2 * m + 1] = lo2(nL[m]); a[2 * (m + H2) + 1] = hi2(nL[m]);
                a[2 * m] = lo2(nB[m]); a[2 * (m + H2)] = hi2(nB[m]);
            }
            if (NL & 1) { a[NS - 1] = oL; a[NS - 2] = oB; }
                    } else {
            const float up1 = __shfl_up_sync(kFull, a[NS - 1], 1) * fu;
#pragma unroll
            for (int i = NS - 1; i >= 0; --i) {
                if (i & 1) {
                    const int jj = i >> 1;
                    const float pl = lookup(row, lsrc[jj]);
                    const float p2 = (i >= 2) ? a[i - 2] : up1;            // (i == 1: msk[0] is 0 on lane 0)
                    a[i] = fmaf(msk[jj], p2, a[i] + a[i - 1]) * pl;
                } else {
                    a[i] = ((i >= 1) ? a[i] + a[i - 1] : a[i] + up1) * pb;
                }
            }
                    }
        };

        // =============================== forward sweep ===============================================
        float a[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) a[i] = 0.f;
        if (lane == 0) a[0] = exp2i(kW32TargetA);           // virtual column t = -1
        int ea = -kW32TargetA;
        float fu = (lane == 0) ? 0.f : 1.f;
        double sprod = 1.0;                                 // product of the row sums of "this lane's" frames (ctc_warp.cuh)
        int sexp = 0, sren = 0;
        unsigned hmax = 0u;                                 // largest p~ bit pattern seen; 1 / s poisoned
        float rcur[K][VCH];

        auto fwd_chunk = [&](auto tag, int t0, int ci, int t0_next) {
            constexpr int KK = decltype(tag)::value;
            {   // states below S - 2(T - t) can no longer reach the end of the transcript: zero them (exact)
                const int lo = S - 2 * (T - t0 + 1);
                if (lo > 0) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) if (lane * NS + i < lo) a[i] = 0.f;
                }
            }
            fu = rescale32<NS, WIN, true>(a, ea, kW32TargetA, lane);
            if (want_grad) {
                float *cp = ckw + (long long)ci * CW + lane;
#pragma unroll
                for (int i = 0; i < NS; ++i) cp[i * 32] = a[i];
                cp[NS * 32] = __int_as_float(ea);
            }
            // p~ of the chunk's rows; lane tt ends up with s_tt
            float sv[KK], pbv[KK];
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                float m = xr[tt][0];
                if (VCH == 2) m = fmaxf(m, xr[tt][VCH - 1]);
                const float ref = (float)__reduce_max_sync(kFull, __float2int_ru(m));
                sv[tt] = 0.f;
#pragma unroll
                for (int v = 0; v < VCH; ++v) {
                    const float p = ex2_approx((xr[tt][v] - ref) * L2E);   // (pad columns: -inf -> 0)
                    rcur[tt][v] = p;
                    hmax = max(hmax, __float_as_uint(p));
                    sv[tt] += p;
                }
                pbv[tt] = __shfl_sync(kFull, (VCH == 2 && bs) ? rcur[tt][VCH - 1] : rcur[tt][0], bl);
            }
            if (t0_next >= 0) load_rows(std::integral_constant<int, K>(), t0_next);
            const float mys = warp_sum_transposed<KK>(sv, lane);
            const float myinv = __frcp_rn(mys);
            hmax = max(hmax, (mys > 0.f) ? 0u : 0x7fc00000u);
            sprod *= (double)((lane < KK) ? mys : 1.f);
            if (++sren == 8) {
                sren = 0;
                const int h = __double2hiint(sprod);
                const int e = ((h >> 20) & 0x7ff) - 1023;   // (NaN from a poisoned row stays NaN)
                sexp += e;
                sprod = __hiloint2double(h - e * (1 << 20), __double2loint(sprod));
            }
            if (want_grad) {
                float *ip = imgw + (long long)t0 * (VCH * 32) + lane;
#pragma unroll
                for (int tt = 0; tt < KK; ++tt)
#pragma unroll
                    for (int v = 0; v < VCH; ++v) ip[(tt * VCH + v) * 32] = rcur[tt][v];
                if (lane < KK) invw[t0 + lane] = myinv;
            }
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) alpha_step(a, rcur[tt], pbv[tt], fu);
        };

        if (nfull > 0) load_rows(std::integral_constant<int, K>(), 0);
        for (int c = 0; c < nfull; ++c)
            fwd_chunk(std::integral_constant<int, K>(), c * K, c, (c + 1 < nfull) ? (c + 1) * K : -1);
        for (int u = 0; u < ntail; ++u) {
            load_rows(std::integral_constant<int, 1>(), nfull * K + u);
            fwd_chunk(std::integral_constant<int, 1>(), nfull * K + u, nfull + u, -1);
        }

        // Z^ = alpha^_{T-1}(S-1) + alpha^_{T-1}(S-2), in the frame of the lane that holds S-1
        const int e_ref = __shfl_sync(kFull, ea, (S - 1) / NS);
        double zloc = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int s = lane * NS + i;
            if (s == S - 1 || s == S - 2) zloc += (double)a[i] * pow2d(max(-1000, min(ea - e_ref, 1000)));
        }
        const double zhat = warp_sum_d(zloc);
        const double lsum = warp_sum_d(log(sprod) + (double)sexp * 0.6931471805599453);
        hmax = __reduce_max_sync(kFull, hmax);
        const bool poisoned = (hmax > 0x3fc00000u);         // p~ above 1.5 (inf / NaN / absurd logits), or an empty row
        const bool z_ok = (zhat > 0.0) && (zhat < INFINITY) && !poisoned;
        if (poisoned || !(zhat == zhat) || zhat == INFINITY) ustat |= UTT_RANGE;
        else if (!z_ok) ustat |= UTT_INF_COST;
        if (lane == 0) {
            const double logz = log(zhat) + (double)e_ref * 0.6931471805599453 - lsum;
            P.costs[b] = z_ok ? (float)(-logz) : ((ustat & UTT_RANGE) ? __int_as_float(0x7fc00000) : INFINITY);
        }
        if (!want_grad) {
            if (lane == 0) P.status[b] = ustat;
            continue;
        }

        // =============================== backward sweep ==============================================
        int ez = 0;
        float inv_mz = 0.f;
        if (z_ok) {
            const int zh = __double2hiint(zhat);
            const int ezl = ((zh >> 20) & 0x7ff) - 1023;
            const double mz = __hiloint2double((zh & 0x000fffff) | 0x3ff00000, __double2loint(zhat));   // [1, 2)
            ez = ezl + e_ref;
            inv_mz = (float)(1.0 / mz);
        }
        float bt[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) bt[i] = (lane * NS + i == S - 1) ? exp2i(kW32TargetB) : 0.f;   // virtual column t = T
        int eb = -kW32TargetB;
        float fd = (lane == 31) ? 0.f : 1.f;
        float chk_dev = 0.f, jmp_dev = 0.f;
        float rq_prev = inv_mz;                             // 1 / (mass at the first frame of the previous chunk); Z^ to start with
        unsigned pmax = 0u;                                 // largest per-symbol product sum seen (must stay below ~2)

        float *stg = (float *)(smem + lay.off_stg);         // [NS + 1][32] column, exponents | [K][VCH][32] rows | [K] 1/s
        auto stage = [&](auto tag, int t0, int ci) {
            constexpr int KK = decltype(tag)::value;
            {
                const char *src = (const char *)(ckw + (long long)ci * CW);
                char *dst = (char *)stg;
#pragma unroll
                for (int o = 0; o < (CW * 4 + 511) / 512; ++o)
                    if (o * 512 + lane * 16 < CW * 4) cp_async16(dst + o * 512 + lane * 16, src + o * 512 + lane * 16);
            }
            {
                const char *src = (const char *)(imgw + (long long)t0 * (VCH * 32));
                char *dst = (char *)(stg + CW);
#pragma unroll
                for (int o = 0; o < (KK * VCH * 128 + 511) / 512; ++o)
                    if (o * 512 + lane * 16 < KK * VCH * 128) cp_async16(dst + o * 512 + lane * 16, src + o * 512 + lane * 16);
            }
            if (lane < KK) cp_async4(stg + CW + K * VCH * 32 + lane, invw + t0 + lane);
            cp_async_commit();
        };
        float myinv = 0.f;
        int ea_c = 0;
        auto unstage = [&](auto tag) {
            constexpr int KK = decltype(tag)::value;
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NS; ++i) a[i] = stg[i * 32 + lane];
            ea_c = __float_as_int(stg[NS * 32 + lane]);
#pragma unroll
            for (int tt = 0; tt < KK; ++tt)
#pragma unroll
                for (int v = 0; v < VCH; ++v) rcur[tt][v] = stg[CW + (tt * VCH + v) * 32 + lane];
            myinv = stg[CW + K * VCH * 32 + (lane & (KK - 1))];
            __syncwarp();                                   // everybody has read: the buffer may be refilled
        };

        auto bwd_chunk = [&](auto tag, int t0, int next, int t0n, int cin) {
            constexpr int KK = decltype(tag)::value;
            unstage(tag);
            if (next == 2) stage(std::integral_constant<int, K>(), t0n, cin);
            else if (next == 1) stage(std::integral_constant<int, 1>(), t0n, cin);
            // posterior scale of this lane: alpha_sc = alpha^ * 2^(eb - ez)
            scale32<NS>(a, ea_c + eb - ez);               // (no clamp: an overflow ends as inf / NaN in the mass check below)
            float fa;
            {
                const int ebl = __shfl_up_sync(kFull, eb, 1);
                const int df = min(eb - ebl, kW32Lip);
                fa = (lane == 0 || df < -126) ? 0.f : exp2i(df);
            }
            float pbv[KK];
#pragma unroll
            for (int tt = 0; tt < KK; ++tt)
                pbv[tt] = __shfl_sync(kFull, (VCH == 2 && bs) ? rcur[tt][VCH - 1] : rcur[tt][0], bl);

            // -- recompute alpha inside the chunk from its checkpoint; keep the label states --
            float av[KK][NL];
            float ab0[NL], ab1[NL];                         // blank states of the chunk's first and last frame (mass checks)
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                alpha_step(a, rcur[tt], pbv[tt], fa);
#pragma unroll
                for (int jj = 0; jj < NL; ++jj) {
                    av[tt][jj] = a[2 * jj + 1];
                    if (tt == 0) ab0[jj] = a[2 * jj];
                    if (tt == KK - 1) ab1[jj] = a[2 * jj];
                }
            }

            // -- beta over the chunk; products alpha * tb go to shared memory grouped by symbol --
            float qf = 0.f, ql = 0.f;                       // this lane's share of sum_s alpha(t, s) * tb(t, s) at t0 and at t1 - 1
            if constexpr (PK) {
#pragma unroll
            for (int tt = KK - 1; tt >= 0; --tt) {
                // what the lane below needs from this one is only  blank 0 + (skip allowed) * label 0  of its last label's
                // successor sum: one shuffle instead of two (msk[0] here is the msk[NL] of the lane below)
                const float dnc = __shfl_down_sync(kFull, fmaf(msk[0], bt[1], bt[0]), 1) * fd;
                float *prow = prod + tt * PS;
                float pl[NL];
#pragma unroll
                for (int jj = 0; jj < NL; ++jj) pl[jj] = lookup(rcur[tt], lsrc[jj]);
                const f32x2 pbb = pk2(pbv[tt], pbv[tt]);
                f32x2 nL[H2 > 0 ? H2 : 1], nB[H2 > 0 ? H2 : 1];
                auto Bk = [&](int j) -> float { return (j < NL) ? bt[2 * (j < NL ? j : 0)] : dnc; };        // blank j (NL: from the lane above, combined)
                auto Lb = [&](int j) -> float { return (j < NL) ? bt[2 * (j < NL ? j : 0) + 1] : 0.f; };    // label j
#pragma unroll
                for (int m = 0; m < H2; ++m) {
                    const f32x2 PL = pk2(Lb(m), Lb(m + H2)), PB = pk2(Bk(m), Bk(m + H2));
                    const f32x2 BN = pk2(Bk(m + 1), Bk(m + H2 + 1)), LN = pk2(Lb(m + 1), Lb(m + H2 + 1));
                    const f32x2 tb = fma2(msk1p[m], LN, add2(PL, BN));
                    const f32x2 pr = mul2(pk2(av[tt][m], av[tt][m + H2]), tb);
                    const f32x2 tbb = add2(PB, PL);
                    const float pr0 = lo2(pr), pr1 = hi2(pr);
                    if (tt == 0) { qf += pr0; qf += pr1; qf = fmaf(ab0[m], lo2(tbb), qf); qf = fmaf(ab0[m + H2], hi2(tbb), qf); }
                    if (tt == KK - 1 && KK > 1) { ql += pr0; ql += pr1; ql = fmaf(ab1[m], lo2(tbb), ql); ql = fmaf(ab1[m + H2], hi2(tbb), ql); }
                    prow[sl[m]] = pr0;
                    prow[sl[m + H2]] = pr1;
                    nL[m] = mul2(tb, pk2(pl[m], pl[m + H2]));
                    nB[m] = mul2(tbb, pbb);
                }
                float oL = 0.f, oB = 0.f;
                if (NL & 1) {
                    constexpr int j = NL - 1;
                    const float tb = bt[2 * j + 1] + dnc;
                    const float pr = av[tt][j] * tb;
                    const float tbb = bt[2 * j] + bt[2 * j + 1];
                    if (tt == 0) { qf += pr; qf = fmaf(ab0[j], tbb, qf); }
                    if (tt == KK - 1 && KK > 1) { ql += pr; ql = fmaf(ab1[j], tbb, ql); }
                    prow[sl[j]] = pr;
                    oL = tb * pl[j];
                    oB = tbb * pbv[tt];
                }
#pragma unroll
                for (int m = 0; m < H2; ++m) {
                    bt[2 * m + 1] = lo2(nL[m]); bt[2 * (m + H2) + 1] = hi2(nL[m]);
                    bt[2 * m] = lo2(nB[m]); bt[2 * (m + H2)] = hi2(nB[m]);
                }
                if (NL & 1) { bt[NS - 1] = oL; bt[NS - 2] = oB; }
            }
            } else {
#pragma unroll
            for (int tt = KK - 1; tt >= 0; --tt) {
                const float dnc = __shfl_down_sync(kFull, fmaf(msk[0], bt[1], bt[0]), 1) * fd;   // (see the packed form)
                float *prow = prod + tt * PS;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    if (i & 1) {
                        const int jj = i >> 1;
                        const float pl = lookup(rcur[tt], lsrc[jj]);
                        const float tb = (i + 2 < NS) ? fmaf(msk[jj + 1], bt[i + 2], bt[i] + bt[i + 1]) : bt[i] + dnc;
                        const float pr = av[tt][jj] * tb;
                        if (tt == 0) qf += pr;
                        if (tt == KK - 1 && KK > 1) ql += pr;
                        prow[sl[jj]] = pr;
                        bt[i] = tb * pl;
                    } else {
                        const float tb = bt[i] + bt[i + 1];
                        if (tt == 0) qf = fmaf(ab0[i >> 1], tb, qf);
                        if (tt == KK - 1 && KK > 1) ql = fmaf(ab1[i >> 1], tb, ql);
                        bt[i] = tb * pbv[tt];
                    }
                }
            }
            }
            __syncwarp();                                   // products visible to the gather

            // -- mass checks.  sum_s alpha(t, s) * tb(t, s) = Z^ at every frame; all terms are positive and both
            //    recursions are linear, so relevant mass lost (flushed, denormal, overflowed) on the way shows as a deficit
            //    where the recursion that lost it ends: the beta recursion at the chunk's FIRST frame (a jump against the
            //    previous chunk's value; mass lost by the forward sweep makes Z^ itself too small and shows the same way),
            //    the recomputed alpha recursion at its LAST frame.  The frames in between need no check of their own:
            //    their products are made of the same alpha and tb values.  The first frame's sum also replaces Z^ for the
            //    chunk's posteriors, which removes the rounding drift of the two T-step fp32 product chains. --
            float scale = 0.f;
            {
                float two[2] = {qf, ql};
                const float mine = warp_sum_transposed<2>(two, lane);       // even lanes: first frame, odd lanes: last
                const float q0 = __shfl_sync(kFull, mine, 0);
                const float q1 = (KK > 1) ? __shfl_sync(kFull, mine, 1) : q0;
                if (z_ok) {
                    const float dev = q0 * inv_mz - 1.f;
                    chk_dev = fmaxf(chk_dev, (dev == dev) ? fabsf(dev) : INFINITY);
                    const float dj = q0 * rq_prev - 1.f;
                    const float rq0 = __frcp_rn(q0);
                    const float dl = q1 * rq0 - 1.f;
                    const float d2 = fmaxf((dj == dj) ? fabsf(dj) : INFINITY, (dl == dl) ? fabsf(dl) : INFINITY);
                    jmp_dev = fmaxf(jmp_dev, d2);
                    rq_prev = rq0;
                    scale = rq0;                                            // posterior = products / (their own sum at t0)
                }
            }

            // -- gather: lane k sums the products of symbol k for the KK frames of the chunk --
            float acc[VCH][KK];
#pragma unroll
            for (int v = 0; v < VCH; ++v) {
                const float *gp = prod + koff[v];
                if constexpr (PK && KK >= 2) {              // frames summed in pairs (FADD2): half the additions of the loop
                    f32x2 ap[KK / 2];
#pragma unroll
                    for (int u = 0; u < KK / 2; ++u) ap[u] = pk2(0.f, 0.f);
                    for (int qq = 0; qq < kcnt[v]; ++qq) {
#pragma unroll
                        for (int u = 0; u < KK / 2; ++u) ap[u] = add2(ap[u], pk2(gp[(2 * u) * PS + qq], gp[(2 * u + 1) * PS + qq]));
                    }
#pragma unroll
                    for (int u = 0; u < KK / 2; ++u) { acc[v][2 * u] = lo2(ap[u]); acc[v][2 * u + 1] = hi2(ap[u]); }
                } else {
#pragma unroll
                    for (int tt = 0; tt < KK; ++tt) acc[v][tt] = 0.f;
                    for (int qq = 0; qq < kcnt[v]; ++qq) {
#pragma unroll
                        for (int tt = 0; tt < KK; ++tt) acc[v][tt] += gp[tt * PS + qq];
                    }
                }
            }
            // gradient rows: lane = symbol (coalesced)
            float gsum[KK];
            float *grow = grads_b + (long long)t0 * gst + lane;
            const float nscale = -scale * P.grad_scale;
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                const float inv_t = __shfl_sync(kFull, myinv, tt) * P.grad_scale;
                gsum[tt] = 0.f;
#pragma unroll
                for (int v = 0; v < VCH; ++v) {
                    pmax = max(pmax, __float_as_uint(acc[v][tt]));
                    const float g = fmaf(rcur[tt][v], inv_t, acc[v][tt] * nscale);
                    if (wr[v]) { grow[32 * v] = g; gsum[tt] += g; }
                }
                grow += gst;
            }
            // the blank entry of frame tt is minus the sum of the row's other entries; written by lane tt
            const float total = warp_sum_transposed<KK>(gsum, lane);
            if (lane < KK) grads_b[(long long)(t0 + lane) * gst + blank] = -total;
            {   // bt holds column t0.  States above 2*t0 + 1 cannot be reached from the start: zero them
                const int hi = 2 * t0 + 1;
                if (hi < S - 1) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) if (lane * NS + i > hi) bt[i] = 0.f;
                }
            }
            fd = rescale32<NS, WIN, false>(bt, eb, kW32TargetB, lane);
            __syncwarp();                                   // gather reads done before the next chunk's products
        };

        if (ntail > 0) stage(std::integral_constant<int, 1>(), T - 1, nfull + ntail - 1);
        else stage(std::integral_constant<int, K>(), (nfull - 1) * K, nfull - 1);
        for (int u = ntail - 1; u >= 0; --u) {
            const int next = (u > 0) ? 1 : (nfull > 0 ? 2 : 0);
            bwd_chunk(std::integral_constant<int, 1>(), nfull * K + u, next,
                      (u > 0) ? nfull * K + u - 1 : (nfull - 1) * K, nfull + u - 1);
        }
        for (int c = nfull - 1; c >= 0; --c)
            bwd_chunk(std::integral_constant<int, K>(), c * K, (c >= 1) ? 2 : 0, (c - 1) * K, c - 1);

        if (!(chk_dev <= kCheckTol) || !(jmp_dev <= kJumpTol) || pmax > 0x40100000u) ustat |= UTT_RANGE;   // (acc = posterior * mz * q < 2; NaN / inf land here)
        if (__any_sync(kFull, ustat & UTT_RANGE)) ustat |= UTT_RANGE;
        if (lane == 0) P.status[b] = ustat;
        if (P.debug && lane == 0) {                         // development aid: what the self-checks saw
            P.debug[b * 16 + 0] = __float_as_uint(chk_dev);
            P.debug[b * 16 + 1] = pmax;
            P.debug[b * 16 + 2] = hmax;
            P.debug[b * 16 + 3] = ustat;
            P.debug[b * 16 + 4] = __float_as_uint(jmp_dev);
        }
        // padded frames get zero gradient
        for (int t = T; t < P.T_max; ++t)
            for (int k = lane; k < V; k += 32) grads_b[(long long)t * gst + k] = 0.f;
    }
    if (P.sm_hi > P.sm_lo && lane == 0) atomicAdd(P.queue + 1, 1);
}

}  // namespace ctcb200
