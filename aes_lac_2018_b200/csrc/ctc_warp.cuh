// ctc_warp.cuh -- the throughput kernel of the sm_100a CTC engine: ONE WARP per utterance, no block barrier,
// (almost) no shared-memory traffic on the T-serial chain.
//
// Same result as ctc_fused_kernel (cost_b and d cost_b / d acts of the reference's
// `criterion(out, targets, out_sizes, target_sizes)`, reference codes/engine.py:22, codes/metrics.py:51; upstream
// warp-ctc compute_alpha_kernel + compute_betas_and_grad_kernel; maths in SURVEY.md Appendix C), reorganised
// around what the round-1 profile showed to be the limiters of that kernel (shared-memory wavefronts, issue
// slots, fp64 instruction count -- profiles/r1_e_ncu_summary.txt):
//
//  * RATIO domain.  Instead of p~ = exp(a - rowmax) the recursion uses r_t(k) = exp(a_t(k) - a_t(blank)), i.e.
//    every column is divided by prod_{tau<=t} p_tau(blank).  r_t(blank) = 1, so a blank state is ONE fp64 add
//    (alpha^_t(s) = alpha^_{t-1}(s) + alpha^_{t-1}(s-1)); a label state stays add + fma + mul.  The blank factors
//    re-enter only the cost: log p_t(blank) = -log s_t with s_t = sum_k r_t(k), which is also the softmax
//    normaliser of the gradient (p_t(k) = r_t(k) / s_t).  No row maximum is needed.
//  * The r table never touches shared memory: lane k holds r_t(k) of the K rows of a chunk in registers (high
//    32 bits of the double: fp32-mantissa accuracy, fp64 exponent range) and a label looks its symbol up with
//    one SHFL.idx (2 per clock per SM, against 1 wavefront per clock for an LDS).
//  * The recomputed alpha columns of the backward sweep stay in REGISTERS (label states only, high words with the
//    posterior scale 2^esc folded into the exponent by one integer add); blank states are not stored at all:
//    the blank posterior is 1 - sum of the others.
//  * No division: posterior_t(k) = sum_{s in pos(k)} alpha^_t(s) * tb_t(s) / Z^, where tb is the beta sum BEFORE
//    it is multiplied by r (beta^_t(s) = tb_t(s) * r_t(l'_s)).
//  * Range self-check once per chunk instead of once per frame: sum_s alpha^_t0(s) * tb_t0(s) must equal Z^ at
//    the first frame t0 of every chunk.  Mass lost to underflow before t0 (alpha) or after it (beta) is
//    missing from every such sum, and loss of alpha mass after the last check makes Z^ itself too small, so
//    the two-sided test |Q_c / Z^ - 1| <= tol at all chunk starts bounds the posterior error of every frame.
//  * Checkpoints are 32-bit (rounded high words), the softmax image is the r row itself (4*32 bytes per frame and
//    32-symbol slice): extra HBM traffic per frame 4*NS*32/K*2 + 2*128 bytes.
//  * An utterance is cut into T / K full chunks of K frames and T % K tail chunks of ONE frame; the chunk bodies
//    are instantiated for both lengths and fully unrolled, so there is no per-frame "is this frame valid" test
//    (those tests and the divergence checks ptxas wraps around them were 15 % of the first version's instructions).
//  * Persistent CTAs: the grid is sized to the resident warps and pulls utterances (longest first) from an
//    atomic queue, so the workspace is per resident CTA, not per utterance.
//
// Thread/state map: lane owns states s = lane*NS + i, i < NS (NS even => even i are blanks).
// tests/proto_ratio.py models exactly this arithmetic on the CPU (<= 1.2e-6 max |dgrad| against the float64 oracle).
#pragma once
#include <type_traits>

#include "ctc_fused.cuh"

namespace ctcb200 {

constexpr int kWarpTargetExp = 192;     // column max after a rescale: 2^192 (headroom 2^831 up, 2^1266 down)
constexpr unsigned kFull = 0xffffffffu;

struct WarpLayout {
    int PS;                              // product row stride (floats); the last 32 slots are per-lane dump slots of padding labels
    int off_prod, off_lab, off_slot, off_cnt, off_off, off_av, off_stg, total;
};

// avs: the recomputed alpha of the label states waits for the beta steps in shared memory instead of registers
__host__ __device__ inline WarpLayout make_warp_layout(int NS, int K, int VCH, int avs)
{
    WarpLayout l;
    const int LP = 16 * NS;
    l.PS = LP + 96;                     // labels + segment padding (< 64) + one dump slot per lane
    int o = 0;
    l.off_prod = o;                      // [K][PS] floats: alpha*tb products of a chunk, grouped by symbol
    l.off_lab = o;                       // [LP] ints   (prologue only: aliases the product rows)
    l.off_slot = o + LP * 4;             // [LP] ints   (prologue only)
    o += K * l.PS * 4;
    l.off_cnt = o;  o += 32 * VCH * 4;
    l.off_off = o;  o += 32 * VCH * 4;
    l.off_av = o;   o += avs ? K * (NS / 2) * 32 * 4 : 0;   // [K][NL][32] high words
    o = (o + 127) & ~127;
    l.off_stg = o;  o += (NS * 32 + K * VCH * 32 + K + 1) * 4;   // staged operands of the next backward chunk
    l.total = (o + 15) & ~15;
    return l;
}

// chunks of an utterance of T frames: T / K full chunks of K frames, then T % K tail chunks of one frame
__host__ __device__ inline int warp_max_chunks(int K, int T_max) { return T_max / K + (K - 1); }

// words (4 bytes) of workspace per resident CTA: checkpoint columns + exponent per chunk, r image + 1/s per frame
__host__ __device__ inline long long warp_slot_words(int NS, int K, int VCH, int T_max)
{
    const long long nC = warp_max_chunks(K, T_max);          // every section is a multiple of 128 bytes
    return nC * 32LL * NS + ((nC + 31) & ~31LL) + (long long)T_max * 32LL * VCH + (((long long)T_max + 31) & ~31LL);
}

__device__ __forceinline__ double hi2d(unsigned hi) { return __hiloint2double((int)hi, 0); }

__device__ __forceinline__ float hi2f(unsigned hi)          // float of a (non-poisoned) ratio high word; below the
{                                                           // float range: (sub)normal garbage <= 2^-126, i.e. 0
    return __int_as_float(__viaddmax_s32((int)hi, -0x38000000, 0) << 3);
}

// r = exp(d) as the high word of a double (20 mantissa bits, rounded) and as the float of that rounded value.
// d < -700: exactly 0.  d > 69 or NaN: poisoned (the utterance is flagged and redone in log space) -- 2^100 per
// frame is what the rescale headroom of an 8-frame chunk can absorb.
__device__ __forceinline__ unsigned ratio_hi(float d, float &rf)
{
    const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.925963033500011e-8f;
    const float yh = d * L2E_HI;
    const float yl = fmaf(d, L2E_LO, fmaf(d, L2E_HI, -yh));
    const float MAGIC = 12582912.f;                         // 1.5 * 2^23
    const float t = yh + MAGIC;
    const float yi = t - MAGIC;                             // nearest integer to yh
    const float fr = (yh - yi) + yl;                        // [-0.5, 0.5]
    const float mf = ex2_approx(fr);                        // [0.707, 1.415]
    const int e = __float_as_int(t) - 0x4B400000;
    // 23 -> 20 mantissa bits, ROUNDED: a truncation would lower every non-blank ratio by 3.4e-7 on average against the
    // blank's exact 1 -- a systematic logit bias that re-weights alignments by their number of label frames and grows
    // with T (found by the parity fuzz: 3e-5 gradient error at L = 1, T = 1200); rounding errors average out
    unsigned hi = (((unsigned)__float_as_int(mf) + 4u) >> 3) + 0x38000000u + ((unsigned)e << 20);
    const bool tiny = (d < -700.f);
    const bool pois = !(d <= 69.f);
    hi = tiny ? 0u : hi;
    hi = pois ? 0x7ff80000u : hi;
    rf = pois ? __int_as_float(0x7fc00000) : hi2f(hi);
    return hi;
}
__device__ __forceinline__ unsigned hi_round(double x)      // high word, rounded to nearest
{
    return (unsigned)__double2hiint(x) + ((unsigned)__double2loint(x) >> 31);
}
__device__ __forceinline__ float warp_sum_f(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// Rescale x[] by an exact power of two so that the warp-wide max has binary exponent kWarpTargetExp.
template <int NS>
__device__ __forceinline__ void warp_rescale(double (&x)[NS], int &E)
{
    unsigned key = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) key = max(key, (unsigned)__double2hiint(x[i]));
    key = __reduce_max_sync(kFull, key);
    const int ex = (int)(key >> 20);
    if (key == 0u || ex >= 0x7ff) return;                   // all zero, or inf/nan (flagged later)
    int sh = kWarpTargetExp - (ex - 1023);
    sh = min(sh, 1023);
    if (sh == 0) return;
    const double f = pow2d(sh);
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] *= f;
    E -= sh;
}

// transposed butterfly: on entry every lane holds K partial values t[0..K), on exit t[0] of lane l is the sum
// over all lanes of value (l & (K-1)).  Fixed order => deterministic.
template <int K>
__device__ __forceinline__ float warp_sum_transposed(float (&t)[K], int lane)
{
#pragma unroll
    for (int m = K / 2; m >= 1; m >>= 1) {
        const bool h = (lane & m) != 0;
#pragma unroll
        for (int u = 0; u < m; ++u) {
            const float send = h ? t[u] : t[u + m];
            const float keep = h ? t[u + m] : t[u];
            t[u] = keep + __shfl_xor_sync(kFull, send, m);
        }
    }
    float v = t[0];
#pragma unroll
    for (int o = K; o < 32; o <<= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

template <int NS, int K, int VCH, int MAXR, int AVS>
__global__ void __maxnreg__(MAXR) ctc_warp_kernel(const FusedParams P)
{
    static_assert(NS % 2 == 0 && NS >= 2 && NS <= 16, "NS must be even, <= 16");
    static_assert(K == 4 || K == 8 || K == 16, "chunk length");
    static_assert(VCH == 1 || VCH == 2, "alphabet slices");
    constexpr int NL = NS / 2, LP = 16 * NS, SP = 32 * NS;
    constexpr float kCheckTol = 5e-6f;
    constexpr double kMeanC = 1.0 + 0.7213 * 4.76837158203125e-7;   // mean of the truncation error of the alpha high words

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const int V = P.V, blank = P.blank;
    const WarpLayout lay = make_warp_layout(NS, K, VCH, AVS);
    int *av_s = (int *)(smem + lay.off_av) + threadIdx.x;    // [K][NL][32]
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
    unsigned *ckw = (unsigned *)P.ckpt + (long long)slot * (P.ckpt_stride * 2);   // [nCmax][NS][32] checkpoint high words
    int *eaw = (int *)(ckw + (long long)nCmax * SP);                                // [nCmax] alpha exponent of the chunk
    unsigned *imgw = (unsigned *)(eaw + ((nCmax + 31) & ~31));                      // [T_max][VCH][32] r high words
    float *invw = (float *)(imgw + (long long)P.T_max * VCH * 32);                  // [T_max] 1 / s_t
    const int bl = blank & 31, bs = blank >> 5;
    const double m_first = (lane == 0) ? 0.0 : 1.0;         // lane 0 has no lower neighbour, lane 31 no upper one:
    const double m_last = (lane == 31) ? 0.0 : 1.0;         //   folded into an fma instead of a select per step
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
        if (P.only_flagged && !(P.status[b] & (UTT_RANGE | UTT_INF_COST))) continue;    // second tier: flagged utterances only
        long long dbg_c0 = 0, dbg_n0 = 0, dbg_c1 = 0;
        if (P.debug && lane == 0) {
            dbg_c0 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_n0));
        }
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
        int ustat = P.only_flagged ? UTT_WIDE : 0;
        if (bad) ustat |= UTT_BAD_LABEL;
        if (T <= 0 || L + rep > T) ustat |= UTT_INFEASIBLE;
        if (ustat & (UTT_BAD_LABEL | UTT_INFEASIBLE)) {     // cost 0, gradient 0 (warp-ctc CPU convention)
            if (lane == 0) { P.costs[b] = 0.f; P.status[b] = ustat; }
            if (want_grad)
                for (int t = 0; t < P.T_max; ++t)
                    for (int k = lane; k < V; k += 32) grads_b[(long long)t * gst + k] = 0.f;
            continue;
        }

        // ---- per-thread label constants ----
        const int j0 = lane * NL;
        int lsrc[NL];                                       // symbol of label j0+jj (= slice*32 + source lane of the lookup);
                                                            // padding labels point at the last pad lane, whose r is 0
        double msk[NL + 1];                                 // 1.0 if the skip INTO label j0+jj is allowed
#pragma unroll
        for (int jj = 0; jj <= NL; ++jj) {
            const int j = j0 + jj;
            const int cur = (j < LP) ? lab_s[j] : -1;
            const int prv = (j >= 1 && j - 1 < LP) ? lab_s[j - 1] : -1;
            if (jj < NL) lsrc[jj] = (cur < 0) ? (32 * VCH - 1) : cur;
            msk[jj] = (cur >= 0 && j >= 1 && cur != prv) ? 1.0 : 0.0;
        }

        // ---- product slots: the alpha*tb products of one frame are stored grouped by symbol, symbol k owning the
        // slots [off[k], off[k] + cnt[k]).  Segment starts are padded so that off[k] mod 32 is distinct for the 32
        // symbols of one pass: at iteration q lane k reads slot off[k] + q and all lanes hit different banks.
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
                if (cur > PS - 32) {                        // padded segments do not fit: plain prefix sums (bank conflicts, same result)
                    int o = 0;
                    for (int k = 0; k < 32 * VCH; ++k) { off_s[k] = o; o += cnt_s[k]; }
                }
            }
            __syncwarp();
            // Which slot of its symbol's segment a label gets decides the bank its product is STORED to: round jj of a
            // frame has lane l store label l*NL + jj, and two labels of a round that land in the same bank cost a
            // replay (random assignment: ~3 wavefronts per store, 11-18 replays per utterance-timestep, a fifth of all
            // shared-memory wavefronts).  One lane assigns the slots greedily in label order -- first free slot of the
            // segment whose bank is still unused in the label's round -- which removes most of them; ~12 instructions
            // per label, once per utterance.  (Segments longer than 32 slots take the plain order.)
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
                        if ((jb & (32 * NL - 1)) == 0) {    // (a round only spans 32 * NL consecutive labels)
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
                                const unsigned banks = __funnelshift_l(fr, fr, o & 31);        // free slots -> their banks
                                const unsigned ok = banks & ~usedb[jj];
                                const int bank = __ffs(ok ? ok : banks) - 1;
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
            sl[jj] = (want_grad && j < L) ? slot_s[j] : PS - 32 + lane;   // (a shared dump slot is a benign write-write race,
                                                                          //  but racecheck rightly reports it)
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
                for (int v = 0; v < VCH; ++v) xr[tt][v] = colok[v] ? __ldg(src + 32 * v) : 0.f;
                src += P.act_stride_t;
            }
        };
        // one label lookup: r of symbol `src` in row `row` (lane src & 31 of slice src >> 5)
        auto lookup = [&](const unsigned (&row)[VCH], int src) -> unsigned {
            unsigned v = __shfl_sync(kFull, row[0], src);
            if (VCH == 2) {
                const unsigned v1 = __shfl_sync(kFull, row[VCH - 1], src);
                v = (src & 32) ? v1 : v;
            }
            return v;
        };
        // one alpha step in place (descending i keeps the old neighbours intact)
        auto alpha_step = [&](double (&a)[NS], const unsigned (&row)[VCH]) {
            const double up1 = shfl_up_d(a[NS - 1]);
#pragma unroll
            for (int i = NS - 1; i >= 0; --i) {
                if (i & 1) {
                    const int jj = i >> 1;
                    const double pl = hi2d(lookup(row, lsrc[jj]));
                    const double p2 = (i >= 2) ? a[i - 2] : up1;           // (i == 1: msk[0] is 0 on lane 0)
                    a[i] = fma(msk[jj], p2, a[i] + a[i - 1]) * pl;
                } else {
                    a[i] = (i >= 1) ? a[i] + a[i - 1] : fma(m_first, up1, a[i]);
                }
            }
        };

        // =============================== forward sweep ===============================================
        double a[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) a[i] = 0.0;
        if (lane == 0) a[0] = pow2d(kWarpTargetExp);        // virtual column t = -1
        int Ea = -kWarpTargetExp;
        // sum_t log s_t: lane tt multiplies the s of "its" frame of every chunk into a running fp64 product (s >= 1: the
        // blank ratio is 1; s < 2^105), renormalised every 8th chunk; ONE log per lane at the end of the sweep
        double sprod = 1.0;
        int sexp = 0, sren = 0;
        unsigned hmax = 0u;                                 // largest ratio high word seen (poison detector)
        unsigned rcur[K][VCH];

        // forward chunk of KK frames starting at t0 (rows already in xr); checkpoint index ci
        auto fwd_chunk = [&](auto tag, int t0, int ci, int t0_next) {
            constexpr int KK = decltype(tag)::value;
            {   // states below S - 2(T - t) can no longer reach the end of the transcript: zero them (exact), which
                // also keeps them out of the column max (see ctc_fused.cuh)
                const int lo = S - 2 * (T - t0 + 1);
                if (lo > 0) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) if (lane * NS + i < lo) a[i] = 0.0;
                }
            }
            warp_rescale<NS>(a, Ea);
            if (want_grad) {
                unsigned *cp = ckw + (long long)ci * SP + lane;
#pragma unroll
                for (int i = 0; i < NS; ++i) cp[i * 32] = hi_round(a[i]);
                if (lane == 0) eaw[ci] = Ea;
            }
            // ratios of the chunk's rows; lane tt ends up with s_tt
            float sv[KK];
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                const float xb = __shfl_sync(kFull, (VCH == 2 && bs) ? xr[tt][VCH - 1] : xr[tt][0], bl);
                sv[tt] = 0.f;
#pragma unroll
                for (int v = 0; v < VCH; ++v) {
                    float rf;
                    const unsigned h = ratio_hi((colok[v] ? xr[tt][v] : -INFINITY) - xb, rf);
                    rcur[tt][v] = h;
                    hmax = max(hmax, h);
                    sv[tt] += rf;
                }
            }
            if (t0_next >= 0) load_rows(std::integral_constant<int, K>(), t0_next);
            const float mys = warp_sum_transposed<KK>(sv, lane);
            const float myinv = __frcp_rn(mys);
            sprod *= (double)((lane < KK) ? mys : 1.f);
            if (++sren == 8) {
                sren = 0;
                const int h = __double2hiint(sprod);
                const int e = ((h >> 20) & 0x7ff) - 1023;   // (NaN from a poisoned row stays NaN)
                sexp += e;
                sprod = __hiloint2double(h - e * (1 << 20), __double2loint(sprod));
            }
            if (want_grad) {
                unsigned *ip = imgw + (long long)t0 * (VCH * 32) + lane;
#pragma unroll
                for (int tt = 0; tt < KK; ++tt)
#pragma unroll
                    for (int v = 0; v < VCH; ++v) ip[(tt * VCH + v) * 32] = rcur[tt][v];
                if (lane < KK) invw[t0 + lane] = myinv;
            }
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) alpha_step(a, rcur[tt]);
        };

        if (nfull > 0) load_rows(std::integral_constant<int, K>(), 0);
        for (int c = 0; c < nfull; ++c)
            fwd_chunk(std::integral_constant<int, K>(), c * K, c, (c + 1 < nfull) ? (c + 1) * K : -1);
        for (int u = 0; u < ntail; ++u) {
            load_rows(std::integral_constant<int, 1>(), nfull * K + u);
            fwd_chunk(std::integral_constant<int, 1>(), nfull * K + u, nfull + u, -1);
        }

        // Z^ = alpha^_{T-1}(S-1) + alpha^_{T-1}(S-2)
        double zloc = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int s = lane * NS + i;
            if (s == S - 1 || s == S - 2) zloc += a[i];
        }
        const double zhat = warp_sum_d(zloc);
        const double lsum = warp_sum_d(log(sprod) + (double)sexp * 0.6931471805599453);
        hmax = __reduce_max_sync(kFull, hmax);
        const int Ea_fin = Ea;
        const bool poisoned = (hmax >= 0x7ff00000u);
        const bool z_ok = (zhat > 0.0) && (zhat < INFINITY) && !poisoned;
        if (poisoned || !(zhat == zhat) || zhat == INFINITY) ustat |= UTT_RANGE;
        else if (!z_ok) ustat |= UTT_INF_COST;
        if (lane == 0) {
            const double logz = log(zhat) + (double)Ea_fin * 0.6931471805599453 - lsum;
            P.costs[b] = z_ok ? (float)(-logz) : ((ustat & UTT_RANGE) ? __int_as_float(0x7fc00000) : INFINITY);
        }
        if (P.debug && lane == 0) dbg_c1 = clock64();
        auto dbg_out = [&]() {
            if (P.debug && lane == 0) {
                long long n1;
                unsigned smid;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                P.debug[b * 16 + 0] = dbg_c1 - dbg_c0;
                P.debug[b * 16 + 1] = clock64() - dbg_c0;
                P.debug[b * 16 + 2] = n1 - dbg_n0;
                P.debug[b * 16 + 3] = smid;
                for (int i = 0; i < 12; ++i) P.debug[b * 16 + 4 + i] = 0;
            }
        };
        if (!want_grad) {
            if (lane == 0) P.status[b] = ustat;
            dbg_out();
            continue;
        }

        // =============================== backward sweep ==============================================
        // Z^ = mz * 2^ez, mz in [1, 2).  Products are formed as alpha^ * tb * 2^esc with esc = Ea_c + Eb - Ea_fin - ez
        // folded into the exponent of the alpha high word; the remaining 1/mz (and the mean of the truncation
        // error of those high words) is one float factor applied per symbol and frame.
        int ez = 0;
        float inv_zm = 0.f;
        double inv_mz_d = 0.0;
        if (z_ok) {
            const int zh = __double2hiint(zhat);
            ez = ((zh >> 20) & 0x7ff) - 1023;
            const double mz = __hiloint2double((zh & 0x000fffff) | 0x3ff00000, __double2loint(zhat));
            inv_mz_d = kMeanC / mz;
            inv_zm = (float)inv_mz_d;
        }
        const float one = z_ok ? 1.f : 0.f;
        double bt[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) bt[i] = (lane * NS + i == S - 1) ? pow2d(kWarpTargetExp) : 0.0;   // virtual column t = T
        int Eb = -kWarpTargetExp;
        float chk_dev = 0.f;

        // Operands of a backward chunk (checkpoint column, r rows, 1/s per frame, chunk exponent): written by this CTA
        // during the forward sweep, fetched back with cp.async (16 bytes per lane: the column and the rows of a chunk
        // are contiguous, so four instructions move all of it) into a shared-memory staging buffer while the
        // PREVIOUS chunk computes, and read out of it (conflict-free, lane-contiguous) when their chunk starts.  No
        // registers are held by the prefetch (an earlier version kept them in registers, ptxas spilled those, and
        // the spill store waited for the load: 10 % of all stall samples) and nothing waits on L2 / DRAM latency at a
        // chunk boundary (loading them "late, into registers that are dead by then" cost 10 % at K = 8, 25 % at K = 4).
        unsigned *stg = (unsigned *)(smem + lay.off_stg);   // [NS][32] column | [K][VCH][32] rows | [K] 1/s | [1] exponent
        auto stage = [&](auto tag, int t0, int ci) {
            constexpr int KK = decltype(tag)::value;
            {   // checkpoint column: NS * 128 bytes
                const char *src = (const char *)(ckw + (long long)ci * SP);
                char *dst = (char *)stg;
#pragma unroll
                for (int o = 0; o < (NS * 128 + 511) / 512; ++o)
                    if (o * 512 + lane * 16 < NS * 128) cp_async16(dst + o * 512 + lane * 16, src + o * 512 + lane * 16);
            }
            {   // r rows: KK * VCH * 128 bytes
                const char *src = (const char *)(imgw + (long long)t0 * (VCH * 32));
                char *dst = (char *)(stg + SP);
#pragma unroll
                for (int o = 0; o < (KK * VCH * 128 + 511) / 512; ++o)
                    if (o * 512 + lane * 16 < KK * VCH * 128) cp_async16(dst + o * 512 + lane * 16, src + o * 512 + lane * 16);
            }
            if (lane < KK) cp_async4(stg + SP + K * VCH * 32 + lane, invw + t0 + lane);
            if (lane == KK) cp_async4(stg + SP + K * VCH * 32 + K, eaw + ci);
            cp_async_commit();
        };
        float myinv = 0.f;
        int ea_c = 0;
        // staging buffer -> registers (a[] gets the raw high words; the posterior scale is applied by the caller)
        auto unstage = [&](auto tag) {
            constexpr int KK = decltype(tag)::value;
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NS; ++i) a[i] = hi2d(stg[i * 32 + lane]);
#pragma unroll
            for (int tt = 0; tt < KK; ++tt)
#pragma unroll
                for (int v = 0; v < VCH; ++v) rcur[tt][v] = stg[SP + (tt * VCH + v) * 32 + lane];
            myinv = __uint_as_float(stg[SP + K * VCH * 32 + (lane & (KK - 1))]);
            ea_c = (int)stg[SP + K * VCH * 32 + K];
            __syncwarp();                                   // everybody has read: the buffer may be refilled
        };

        // backward chunk of KK frames starting at t0 (operands ready); next: 0 none, 1 one-frame chunk, 2 full chunk
        auto bwd_chunk = [&](auto tag, int t0, int next, int t0n, int cin) {
            constexpr int KK = decltype(tag)::value;
            unstage(tag);
            if (next == 2) stage(std::integral_constant<int, K>(), t0n, cin);        // lands while this chunk computes
            else if (next == 1) stage(std::integral_constant<int, 1>(), t0n, cin);
            // The recursion is linear: scaling the checkpoint column by 2^esc (esc = Ea_c + Eb - Ea_fin - ez, the
            // posterior scale of this chunk) scales every recomputed column, so the products alpha * tb come out
            // in units of mz directly.  (Entries pushed below 2^-1022 by the scale would need a ratio product of
            // 2^700 within the chunk to matter again; the range check covers it.)
            {
                int esc = ea_c + Eb - Ea_fin - ez;
                esc = max(-1000, min(esc, 700));
                const double f = pow2d(esc);
#pragma unroll
                for (int i = 0; i < NS; ++i) a[i] *= f;
            }

            // -- recompute alpha inside the chunk from its checkpoint; keep the label states (high words) --
            int av[AVS ? 1 : KK][NL];
            int ab[NL];                                     // blank states of the first frame (range check)
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                alpha_step(a, rcur[tt]);
#pragma unroll
                for (int jj = 0; jj < NL; ++jj) {
                    if (AVS) av_s[(tt * NL + jj) * 32] = __double2hiint(a[2 * jj + 1]);
                    else av[AVS ? 0 : tt][jj] = __double2hiint(a[2 * jj + 1]);
                }
                if (tt == 0) {
#pragma unroll
                    for (int jj = 0; jj < NL; ++jj) ab[jj] = __double2hiint(a[2 * jj]);
                }
            }

            // -- beta over the chunk; products alpha * tb go to shared memory grouped by symbol --
            double q = 0.0;
#pragma unroll
            for (int tt = KK - 1; tt >= 0; --tt) {
                const double dn0 = shfl_down_d(bt[0]), dn1 = shfl_down_d(bt[1]);
                float *prow = prod + tt * PS;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    if (i & 1) {
                        const int jj = i >> 1;
                        const double pl = hi2d(lookup(rcur[tt], lsrc[jj]));
                        const double s1 = (i + 1 < NS) ? bt[i] + bt[i + 1] : fma(m_last, dn0, bt[i]);
                        const double n2 = (i + 2 < NS) ? bt[i + 2] : dn1;  // (msk[NL] is 0 on lane 31)
                        const double tb = fma(msk[jj + 1], n2, s1);
                        const double pr = hi2d((unsigned)(AVS ? av_s[(tt * NL + jj) * 32] : av[AVS ? 0 : tt][jj])) * tb;
                        if (tt == 0) q += pr;
                        prow[sl[jj]] = (float)pr;
                        bt[i] = tb * pl;
                    } else {
                        bt[i] = bt[i] + bt[i + 1];
                        if (tt == 0) q = fma(hi2d((unsigned)ab[i >> 1]), bt[i], q);
                    }
                }
            }
            __syncwarp();                                   // products visible to the gather

            // -- range check: the posterior mass of frame t0 must be 1 --
            if (z_ok) {
                const float dev = (float)(warp_sum_d(q) * inv_mz_d - 1.0);
                chk_dev = fmaxf(chk_dev, (dev == dev) ? fabsf(dev) : INFINITY);
            }

            // -- gather: lane k sums the products of symbol k for the KK frames of the chunk --
            float tot[KK];
            float post[VCH][KK];
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) tot[tt] = 0.f;
#pragma unroll
            for (int v = 0; v < VCH; ++v) {
                float acc[KK];
#pragma unroll
                for (int tt = 0; tt < KK; ++tt) acc[tt] = 0.f;
                const float *gp = prod + koff[v];
                for (int qq = 0; qq < kcnt[v]; ++qq) {
#pragma unroll
                    for (int tt = 0; tt < KK; ++tt) acc[tt] += gp[tt * PS + qq];
                }
#pragma unroll
                for (int tt = 0; tt < KK; ++tt) {
                    post[v][tt] = acc[tt] * inv_zm;
                    tot[tt] += post[v][tt];
                }
            }
            // gradient rows: lane = symbol (coalesced); the blank entry of frame tt is written by lane tt
            float *grow = grads_b + (long long)t0 * gst + lane;
#pragma unroll
            for (int tt = 0; tt < KK; ++tt) {
                const float inv_t = __shfl_sync(kFull, myinv, tt);
#pragma unroll
                for (int v = 0; v < VCH; ++v) {
                    const float g = (hi2f(rcur[tt][v]) * inv_t - post[v][tt]) * P.grad_scale;
                    if (wr[v]) grow[32 * v] = g;
                }
                grow += gst;
            }
            const float total = warp_sum_transposed<KK>(tot, lane);     // lane l: sum over symbols of frame l & (KK-1)
            if (lane < KK) {
                const float g = (myinv - (one - total)) * P.grad_scale; // p(blank) = 1 / s
                grads_b[(long long)(t0 + lane) * gst + blank] = g;
            }
            {   // bt holds column t0.  States above 2*t0 + 1 cannot be reached from the start: zero them
                const int hi = 2 * t0 + 1;
                if (hi < S - 1) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) if (lane * NS + i > hi) bt[i] = 0.0;
                }
            }
            warp_rescale<NS>(bt, Eb);
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

        if (!(chk_dev <= kCheckTol)) ustat |= UTT_RANGE;
        if (__any_sync(kFull, ustat & UTT_RANGE)) ustat |= UTT_RANGE;
        if (lane == 0) P.status[b] = ustat;
        // padded frames get zero gradient
        for (int t = T; t < P.T_max; ++t)
            for (int k = lane; k < V; k += 32) grads_b[(long long)t * gst + k] = 0.f;
        dbg_out();
        if (P.debug && lane == 0) P.debug[b * 16 + 4] = __float_as_uint(chk_dev);   // (what the range self-check saw)
    }
    if (P.sm_hi > P.sm_lo && lane == 0) atomicAdd(P.queue + 1, 1);
}

}  // namespace ctcb200
