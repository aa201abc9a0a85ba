// ctc_head_bwd_tc.cuh -- tensor-core backward passes of the classifier head (tcgen05.mma kind::tf32, 3xTF32 split).
//
// SURVEY.md 8f row 4; reference /root/reference/codes/model.py:177-180, 199-207 (BatchNorm1d + Linear under autograd).
// Both contractions of the backward pass are skinny (V <= 64 on one side) and sit ABOVE the fp32-FMA ridge of a B200
// (29 FMA per 4 bytes of x), so the round-1 FMA kernels were shared-memory / FMA bound at 17-28 % of the DRAM peak
// (profiles/r1_head_ncu_final.txt).  Here they run on the 5th-generation tensor cores with the same error-compensated
// operand split as the forward pass (a = a_hi + a_lo, three products per K-step -- two MMAs in the weight gradient, where hi x [hi | lo] is one -- fp32 accumulation in TMEM):
//
//   head_wgrad_tc_kernel   G^T[h][v] = sum_n (x[n][h] - mu_h) * dl[n][v]        M = 128 features, N = VP classes, K = rows
//        Both operands lie in HBM with the *non*-contracted index contiguous (x[n][h]: h, dl[n][v]: v), i.e. they are
//        MN-major operands.  tcgen05 takes MN-major tf32 operands directly (instruction-descriptor bits 15 / 16), but
//        for 32-bit types only in ONE shared-memory layout: the 128-byte swizzle with 32-byte atoms
//        (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B, Swizzle<2,5,2> on byte addresses): a 128-byte line holds 32
//        consecutive features (classes) of one row, four lines (rows k..k+3) make a 512-byte atom inside which the
//        32-byte chunk c of line r sits at chunk c ^ r; atoms of the next four rows are SBO apart, the next 32 features
//        LBO apart.  (The plain no-swizzle MN-major layout is accepted silently and multiplies zeros -- measured.)
//        A float4 as loaded from HBM (four features of one row) is one 16-byte store into that layout, a quarter-warp
//        writes one whole line: no transpose anywhere, no bank conflicts.
//   head_dgrad_tc_kernel   D^T[h][n] = sum_v W[v][h] * dl[n][v]                  M = 128 features, N = 128 rows, K = VP
//        computed TRANSPOSED on purpose: a TMEM lane is then a feature h and a TMEM column a row n, so in the epilogue
//        the 32 lanes of a warp hold 32 consecutive features of one row -- the loads of x and the stores of dx are
//        128-byte coalesced without any staging, and the per-feature coefficients A_h, B_h, C_h, mu_h of
//        dx = A (dl W) + B + C (x - mu) are four registers per thread for the whole kernel.
//
// Why no TMA for x: every x tile needs an element-wise transform (centre, hi/lo split) before the MMA may read it, so a
// TMA'd raw tile would have to be read back and rewritten by the threads -- with the three operand reads of the MMA that
// is ~1000 shared-memory wavefronts per 16 KB tile against the 745 clocks the tile has at the HBM rate.  Loading through
// registers (LDG.128 -> transform -> STS.128) costs 256 + the MMA's reads and keeps a deeper prefetch (4 tiles per thread).
#pragma once
#include "ctc_head_tc.cuh"

namespace ctcb200 {
namespace tc {

__device__ __forceinline__ float ldg_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// ---------------------------------------------------------------------------------------------------------
// weight gradient
constexpr int kWBM = 128, kWBK = 32;                       // features per CTA, rows per stage
constexpr int kWPF = 4;                                    // register-staged tiles in flight per thread
__host__ __device__ inline int head_wgrad_tc_smem_bytes(int VP) { return 2 * (2 * kWBM * kWBK * 4 + 2 * VP * kWBK * 4) + 1024; }

template <int VP>
__global__ void __launch_bounds__(256, 2) head_wgrad_tc_kernel(const float *__restrict__ x, const float *__restrict__ dl,
                                                               const float *__restrict__ mean, float *__restrict__ part,
                                                               int N, int H, int V, int rows_per_block)
{
    constexpr int A_F4 = kWBM * kWBK / 4;                  // float4 per x tile (one of hi / lo): 1024
    constexpr int B_F = VP * kWBK;                         // floats per dl tile
    constexpr int STAGE_F4 = 2 * A_F4 + 2 * B_F / 4;       // [a_hi | a_lo | b_hi | b_lo]
    constexpr int DQ = B_F / 256;                          // dl values per thread and tile (4 / 8)
    extern __shared__ __align__(1024) unsigned char hsm[];
    __shared__ __align__(8) uint64_t mbar[3];
    __shared__ uint32_t tmem_base_s;
    // (the swizzle is a function of absolute shared-memory address bits: every operand tile starts on a 1024-byte line)
    float4 *const stage0 = (float4 *)(hsm + ((1024u - (smem_u32(hsm) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h0 = blockIdx.x * kWBM;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
    const int nk = (r1 - r0 + kWBK - 1) / kWBK;

    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init(&mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * VP) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the dl tiles: zero once (the pad classes v >= V are never written again)
    for (int s = 0; s < 2; ++s)
        for (int i = tid; i < 2 * B_F / 4; i += 256) stage0[s * STAGE_F4 + 2 * A_F4 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // x: a quarter-warp loads / stores one 128-byte line (32 features of one row); the four quarters take rows k..k+3
    // of one atom; this thread's float4 i of a tile is row ((warp >> 2) + 2 i) * 4 + kq of feature block mb32 = warp & 3
    const int h4l = lane & 7, kq = lane >> 3, mb32 = warp & 3;
    const int hx = h0 + mb32 * 32 + h4l * 4;
    const bool hv = hx < H;
    // float4 slot inside an x tile: [feature block: 4096 B][atom: 512 B][line kq: 128 B][chunk ^ kq: 32 B][half: 16 B]
    const int aoff = mb32 * 256 + (warp >> 2) * 32 + kq * 8 + ((((h4l >> 1) ^ kq)) << 1) + (h4l & 1);
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hv) mu = __ldg((const float4 *)(mean + hx));
    // dl: the tile is 32 * V contiguous floats; element idx = tid + 256 q is row idx / V, class idx % V
    int doff[DQ];                                          // float offset inside a dl tile, -1 = beyond the tile
#pragma unroll
    for (int q = 0; q < DQ; ++q) {
        const int idx = tid + q * 256;
        const int k = idx / V, v = idx - k * V;
        doff[q] = (idx < kWBK * V) ? ((v >> 5) * 1024 + (k >> 2) * 128 + (k & 3) * 32 + ((((v & 31) >> 3) ^ (k & 3)) << 3) + (v & 7)) | (k << 16) : -1;
    }

    auto gload = [&](int kt, float4 (&xr)[4], float (&dq)[DQ]) {
        const int rb = r0 + kt * kWBK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = min(rb + ((warp >> 2) + 2 * i) * 4 + kq, r1 - 1);    // (rows beyond the block meet dl = 0)
            xr[i] = hv ? ldg_stream4(x + (size_t)n * H + hx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float *d = dl + (size_t)rb * V;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            const int idx = tid + q * 256;
            dq[q] = (doff[q] >= 0 && rb + (doff[q] >> 16) < r1) ? __ldg(d + idx) : 0.f;
        }
    };
    // D = f32, A = B = tf32, both MN-major, N >> 3, M >> 4
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(VP >> 3) << 17) |
                               ((uint32_t)(kWBM >> 4) << 24);
    // the same with N = 2 VP: the hi and lo tiles of dl lie behind one another (the lo tile is simply the next class
    // blocks), so a_hi x [b_hi | b_lo] is ONE MMA into 2 VP accumulator columns -- two MMAs per K step instead of three,
    // x_hi read from shared memory once instead of twice; the epilogue adds the two column halves
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((2 * VP) >> 3) << 17) |
                                ((uint32_t)(kWBM >> 4) << 24);

    auto stage_body = [&](int kt, float4 (&xr)[4], float (&dq)[DQ]) {
        const int s = kt & 1;
        float4 *const Ahi = stage0 + s * STAGE_F4, *const Alo = Ahi + A_F4;
        float *const Bhi = (float *)(Alo + A_F4), *const Blo = Bhi + B_F;
        if (kt >= 2) mbar_wait(&mbar[s], (uint32_t)(((kt >> 1) - 1) & 1));     // the MMAs of tile kt-2 have read stage s
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v0 = xr[i].x - mu.x, v1 = xr[i].y - mu.y, v2 = xr[i].z - mu.z, v3 = xr[i].w - mu.w;
            const float a0 = __uint_as_float(tf32_rna(v0)), a1 = __uint_as_float(tf32_rna(v1));
            const float a2 = __uint_as_float(tf32_rna(v2)), a3 = __uint_as_float(tf32_rna(v3));
            Ahi[aoff + i * 64] = make_float4(a0, a1, a2, a3);
            Alo[aoff + i * 64] = make_float4(v0 - a0, v1 - a1, v2 - a2, v3 - a3);
        }
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            if (doff[q] >= 0) {
                const float hi = __uint_as_float(tf32_rna(dq[q]));
                Bhi[doff[q] & 0xffff] = hi;
                Blo[doff[q] & 0xffff] = dq[q] - hi;
            }
        }
        if (kt + kWPF < nk) gload(kt + kWPF, xr, dq);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic-proxy stores -> visible to the MMA
        __syncthreads();
        // (Tried: no block barrier -- each warp counts itself in with an atomic and the eighth to arrive issues the MMAs, so
        //  that no warp waits for the slowest.  Slower on the box: weight gradient 156 -> 193 us, forward 179 -> 211 us.)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), bh = smem_u32(Bhi);
#pragma unroll
            for (int j = 0; j < kWBK / 8; ++j) {           // one MMA = 8 rows (K = 8 tf32) = two 512-byte atoms
                const uint32_t o = j * 1024;
                // LBO = the next 32 features / classes (4096 B), SBO = the next atom of four rows (512 B)
                const uint64_t dah = smem_desc_sw(ah + o, 4096, 512, 1), dal = smem_desc_sw(al + o, 4096, 512, 1);
                const uint64_t dbh = smem_desc_sw(bh + o, 4096, 512, 1);
                mma_tf32(tmem, dah, dbh, idesc2, (kt > 0 || j > 0) ? 1u : 0u);   // columns [0, VP): hi x hi, [VP, 2 VP): hi x lo
                mma_tf32(tmem, dal, dbh, idesc, 1u);                              // columns [0, VP) += lo x hi
            }
            mma_commit(&mbar[s]);
            if (kt == nk - 1) mma_commit(&mbar[2]);
        }
    };

    float4 xa[4], xb[4], xc[4], xd[4];
    float da[DQ], db[DQ], dc[DQ], dd[DQ];
    if (nk > 0) gload(0, xa, da);
    if (nk > 1) gload(1, xb, db);
    if (nk > 2) gload(2, xc, dc);
    if (nk > 3) gload(3, xd, dd);
    for (int kt = 0; kt < nk; kt += kWPF) {
        stage_body(kt, xa, da);
        if (kt + 1 < nk) stage_body(kt + 1, xb, db);
        if (kt + 2 < nk) stage_body(kt + 2, xc, dc);
        if (kt + 3 < nk) stage_body(kt + 3, xd, dd);
    }
    if (nk > 0) mbar_wait(&mbar[2], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warps 0-3, thread = feature (TMEM lane), VP class columns -> part[rb][v][h] (coalesced over h)
    if (warp < 4) {
        const int h = h0 + warp * 32 + lane;
        float *pp = part + (size_t)blockIdx.y * VP * H;
#pragma unroll
        for (int c = 0; c < VP / 16; ++c) {
            uint32_t v[16], w[16];
            if (nk > 0) {
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c * 16, v);
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + VP + c * 16, w);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[j] = 0u; w[j] = 0u; }
            }
            if (h < H) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < V) pp[(size_t)(c * 16 + j) * H + h] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * VP) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// input gradient
constexpr int kDBM = 128, kDBN = 128;                      // features per CTA (TMEM lanes), rows per block (TMEM columns)
__host__ __device__ inline int head_dgrad_tc_smem_bytes(int VP) { return 4 * kDBM * VP * 4; }
__host__ __device__ inline size_t head_dgrad_tc_weight_bytes(int H, int VP) { return (size_t)((H + kDBM - 1) / kDBM) * 2 * kDBM * VP * 4; }

// W[v][h] split into tf32 hi / lo and laid out as the K-major canonical A tiles of the input-gradient MMA:
// tile ht, part (0 hi, 1 lo), 16-byte column c (classes 4c..4c+3), feature m -> one float4
__global__ void head_wt_tc_kernel(const float *__restrict__ W, int H, int V, int VP, float4 *__restrict__ wtc)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = VP / 4, tiles = (H + kDBM - 1) / kDBM;
    if (idx >= tiles * C * kDBM) return;
    const int m = idx % kDBM, c = (idx / kDBM) % C, ht = idx / (kDBM * C);
    const int h = ht * kDBM + m;
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int v = c * 4 + e;
        const float w = (v < V && h < H) ? W[(size_t)v * H + h] : 0.f;
        hi[e] = __uint_as_float(tf32_rna(w));
        lo[e] = __uint_as_float(tf32_rna(w - hi[e]));
    }
    wtc[((size_t)(ht * 2 + 0) * C + c) * kDBM + m] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    wtc[((size_t)(ht * 2 + 1) * C + c) * kDBM + m] = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

template <int VP>
__global__ void __launch_bounds__(256, (VP <= 48) ? 2 : 1)
head_dgrad_tc_kernel(const float *__restrict__ x, const float *__restrict__ dl, const float4 *__restrict__ wtc,
                     const float *__restrict__ coef, float *__restrict__ dx, int N, int H, int V, int blocks_per_cta)
{
    constexpr int T_F4 = kDBM * VP / 4;                    // float4 per operand tile (one of hi / lo)
    constexpr int DQ = kDBN * VP / 256;                    // dl values per thread and block (16 / 32)
    extern __shared__ __align__(1024) unsigned char hsm[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    float4 *const Whi = (float4 *)hsm, *const Wlo = Whi + T_F4;
    float *const dbuf = (float *)(Wlo + T_F4);             // [hi | lo][VP/4 columns][128 rows][4]; ONE buffer: a block's MMAs are
                                                           // waited for before its epilogue anyway, the next tile is staged right after
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h0 = blockIdx.x * kDBM;
    const int nblk = (N + kDBN - 1) / kDBN;
    const int b0 = blockIdx.y * blocks_per_cta, b1 = min(nblk, b0 + blocks_per_cta);
    if (b0 >= b1) return;
    const int nb = b1 - b0;

    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * kDBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // resident weight slice; dl buffers zeroed once (pad classes stay 0)
        const float4 *src = wtc + (size_t)blockIdx.x * 2 * T_F4;
        for (int i = tid; i < 2 * T_F4; i += 256) Whi[i] = __ldg(src + i);
        float4 *d4 = (float4 *)dbuf;
        for (int i = tid; i < 2 * T_F4; i += 256) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // this thread in the epilogue: feature h (TMEM lane 32 * (warp & 3) + lane), rows (warp >> 2) * 64 .. + 63 of a block
    const int q4 = warp & 3, half = warp >> 2;
    const int h = h0 + q4 * 32 + lane;
    const bool hv = h < H;
    float cA = 0.f, cB = 0.f, cC = 0.f, cM = 0.f;
    if (hv) { cA = __ldg(coef + h); cB = __ldg(coef + H + h); cC = __ldg(coef + 2 * H + h); cM = __ldg(coef + 3 * H + h); }

    // dl block = 128 * V contiguous floats; element tid + 256 q is row (tid + 256 q) / V, class (tid + 256 q) % V.  Its
    // float offset in the K-major operand tile is the same for every block: kept as packed 16-bit pairs (the compiler
    // hoists the unpacked form out of the block loop and spills it); 0xffff = beyond the tile
    uint32_t dpk[DQ / 2];
    {
        const int dr = 256 / V, dv = 256 - dr * V;
        int r = tid / V, v = tid - r * V;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            const uint32_t o = (r < kDBN) ? (uint32_t)((v >> 2) * (kDBN * 4) + r * 4 + (v & 3)) : 0xffffu;
            if (q & 1) dpk[q >> 1] |= o << 16; else dpk[q >> 1] = o;
            v += dv; r += dr;
            if (v >= V) { v -= V; ++r; }
        }
    }
    auto load_dl = [&](int blk, float (&dq)[DQ]) {
        const float *d = dl + (size_t)blk * kDBN * V;
        const int lim = min(kDBN, N - blk * kDBN) * V;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            const int idx = tid + q * 256;
            dq[q] = (idx < lim) ? __ldg(d + idx) : 0.f;
        }
    };
    auto stage_dl = [&](const float (&dq)[DQ]) {
        float *const hi = dbuf, *const lo = hi + kDBN * VP;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            uint32_t pk;                                   // (opaque: keeps the unpacked offsets out of the loop-invariant set)
            asm volatile("mov.b32 %0, %1;" : "=r"(pk) : "r"(dpk[q >> 1]));
            const uint32_t o = (q & 1) ? (pk >> 16) : (pk & 0xffffu);
            if (o != 0xffffu) {
                const float a = __uint_as_float(tf32_rna(dq[q]));
                hi[o] = a;
                lo[o] = dq[q] - a;
            }
        }
    };
    // D = f32, A = B = tf32, both K-major, N = 128 rows, M = 128 features
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kDBN >> 3) << 17) | ((uint32_t)(kDBM >> 4) << 24);
    auto issue_mma = [&](int li) {
        int buf;                                           // opaque to the optimiser: otherwise the 48 descriptors of the two
        asm volatile("and.b32 %0, %1, 1;" : "=r"(buf) : "r"(li));   // buffers are hoisted out of the block loop as ~100 live registers
        const uint32_t wh = smem_u32(Whi), wl = smem_u32(Wlo);
        const uint32_t dh = smem_u32(dbuf), dlo = dh + kDBN * VP * 4;
        const uint32_t acc = tmem + buf * kDBN;
#pragma unroll
        for (int j = 0; j < VP / 8; ++j) {                 // K = 8 classes = two 16-byte columns of 128 rows
            const uint32_t o = j * 2 * kDBM * 16;
            const uint64_t awh = smem_desc(wh + o, kDBM * 16, 128), awl = smem_desc(wl + o, kDBM * 16, 128);
            const uint64_t bdh = smem_desc(dh + o, kDBN * 16, 128), bdl = smem_desc(dlo + o, kDBN * 16, 128);
            mma_tf32(acc, awl, bdh, idesc, j > 0 ? 1u : 0u);
            mma_tf32(acc, awh, bdl, idesc, 1u);
            mma_tf32(acc, awh, bdh, idesc, 1u);
        }
        mma_commit(&mbar[buf]);
    };

    float dq[DQ];
    load_dl(b0, dq);
    stage_dl(dq);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_mma(0);
    }

    for (int li = 0; li < nb; ++li) {
        const int blk = b0 + li;
        const int nbase = blk * kDBN + half * 64;
        // the next block's dl tile and the x values of this block's epilogue are requested together; the tile is staged
        // (and its MMAs issued) while the x values are still in flight.  (Holding the tile in registers across the
        // epilogue instead made ptxas spill.)
        if (li + 1 < nb) load_dl(blk + 1, dq);
        float xr[2][32];
        const bool full = (blk + 1) * kDBN <= N;           // CTA-uniform: only the last block of all needs per-row tests
        auto load_x = [&](int c) {
            const float *px = x + (size_t)(nbase + c * 32) * H + h;
            if (full) {
                if (hv) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xr[c][j] = ldg_stream(px + (size_t)j * H);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) xr[c][j] = (hv && nbase + c * 32 + j < N) ? ldg_stream(px + (size_t)j * H) : 0.f;
            }
        };
        load_x(0);
        load_x(1);
        mbar_wait(&mbar[li & 1], (uint32_t)((li >> 1) & 1));       // this block's MMAs are complete: the dl buffer is free
        if (li + 1 < nb) stage_dl(dq);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && li + 1 < nb) issue_mma(li + 1);    // (runs beside this block's epilogue, into the other accumulator)
#pragma unroll
        for (int c = 0; c < 4; ++c) {                      // 16 rows at a time (32 made ptxas spill under the 128-register cap)
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(q4 * 32) << 16) + (li & 1) * kDBN + half * 64 + c * 16, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (hv) {
                float *po = dx + (size_t)(nbase + c * 16) * H + h;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        po[(size_t)j * H] = fmaf(cA, __uint_as_float(v[j]), fmaf(cC, xr[c >> 1][(c & 1) * 16 + j] - cM, cB));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (nbase + c * 16 + j < N)
                            po[(size_t)j * H] = fmaf(cA, __uint_as_float(v[j]), fmaf(cC, xr[c >> 1][(c & 1) * 16 + j] - cM, cB));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * kDBN) : "memory");
}

}  // namespace tc
}  // namespace ctcb200
