// ctc_head_tc.cuh -- tensor-core forward pass of the classifier head: tcgen05.mma kind::tf32 with a 3xTF32 split.
//
// logits[n][v] = sum_h (x[n][h] - mu_h) * Wk[h][v] + bias'[v] is the one contraction on the path.
// It is a skinny product (V <= 64 columns): 2*V flops per 4 bytes of x puts it ABOVE the fp32-FMA ridge of a B200
// (47 TFMA/s needed at the HBM rate, 36 available), so an fp32 FMA kernel cannot be HBM-bound (the first version of this
// pass was one: 42 % of the FMA peak, 1.9-2.2 TB/s), while the
// 5th-generation tensor cores have ~30x the throughput needed.  tf32 alone would round the operands to 10 mantissa
// bits (1e-3 logits; the reference computes in fp32), so each operand is split  a = a_hi + a_lo  (both tf32) and
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (error ~2^-21 relative per product, fp32 accumulation)
// -- three products per K-step (two MMAs: a_hi x [b_hi | b_lo] is one, with N = 2 VP), far below the HBM time of the tile.
//
// One CTA = 128 rows of x (UMMA M = 128, cta_group::1), N = VP columns, accumulator = VP TMEM columns.
// The operands cannot come straight from HBM by TMA: x needs (x - mu) and the hi/lo split first.  So the 256 threads
// load the 128 x 32 tile of a stage into registers (4 float4 each; a quarter-warp takes the 128 contiguous bytes a row
// has in the tile), transform it, and store a_hi and a_lo to shared memory in the K-major SWIZZLE_128B UMMA layout (row
// = 128 bytes = the 32 features of the tile, 16-byte chunk c of row r at chunk c ^ (r & 7), groups of 8 rows 1024 bytes
// apart = SBO; the K = 8 steps of an MMA advance the start address by 32 bytes inside the swizzled row).  The weights
// arrive pre-split and pre-arranged in that layout (head_fold_tc_kernel).
// Two stages: while the tensor core works on stage s (tracked by an mbarrier through tcgen05.commit) the threads
// transform stage s^1 and the global loads of the stage after are in flight in registers.
// Epilogue: each warp reads its 32 TMEM lanes (tcgen05.ld 32x32b: one full logits row per thread), adds the bias,
// optionally normalises the row (softmax, eval mode) and writes it in T x B x V order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctcb200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a protocol error traps (the launch fails loudly) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    for (unsigned it = 0; !mbar_try(b, parity); ++it)
        if (it > (1u << 26)) __trap();
}
// round to tf32 (10 mantissa bits), ties away from zero in magnitude -- what cvt.rna.tf32.f32 does, as two integer
// operations on the ALU pipe instead of a conversion-unit instruction (32 of them per thread and tile)
__device__ __forceinline__ uint32_t tf32_rna(float x)
{
    return ((uint32_t)__float_as_int(x) + 0x1000u) & 0xffffe000u;
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start >> 4 at [0,14),
// leading-dimension byte offset >> 4 at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48), layout 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}

constexpr int kBM = 128, kBK = 32;
constexpr int kPF = 4;                                    // register-staged tiles in flight per thread
constexpr int kThreads = 256;                             // 8 warps: all of them load / transform, warps 0-3 run the epilogue
constexpr int kXPT = kBM * kBK / 4 / kThreads;            // float4 of a tile per thread (4)

__host__ __device__ inline int head_tc_smem_bytes(int VP) { return 2 * (2 * kBM * kBK * 4 + 2 * VP * kBK * 4) + 1024; }
__host__ __device__ inline size_t head_tc_weight_bytes(int H, int VP)
{
    return (size_t)((H + kBK - 1) / kBK) * 2 * kBK * VP * 4;
}
// matrix descriptor of a swizzled operand tile (layout type at bits [61,64): 2 = 128-byte swizzle, 1 = 128-byte swizzle
// with 32-byte atoms)
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type)
{
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ float4 ldg_stream4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// weights, already scaled by gamma * invstd (wk[h][v], head_fold_kernel), split into tf32 hi/lo and laid out as the
// K-major SWIZZLE_128B tiles the MMA reads: tile kt, part (0 hi, 1 lo), class row n = 128 bytes (the 32 features of the
// tile), 16-byte chunk c of row n stored at chunk c ^ (n & 7)
__global__ void head_fold_tc_kernel(const float *__restrict__ wk, int H, int VP, float4 *__restrict__ bc)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nk = (H + kBK - 1) / kBK;
    if (idx >= nk * 8 * VP) return;
    const int n = idx % VP, c = (idx / VP) % 8, kt = idx / (8 * VP);
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = kt * kBK + c * 4 + e;
        const float w = (k < H) ? wk[(size_t)k * VP + n] : 0.f;       // (columns n >= V hold 0 already)
        hi[e] = __uint_as_float(tf32_rna(w));
        lo[e] = __uint_as_float(tf32_rna(w - hi[e]));
    }
    bc[((size_t)(kt * 2 + 0) * VP + n) * 8 + (c ^ (n & 7))] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    bc[((size_t)(kt * 2 + 1) * VP + n) * 8 + (c ^ (n & 7))] = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

template <int VP>
__global__ void __launch_bounds__(kThreads, 2) head_fwd_tc_kernel(const float *__restrict__ x, const float4 *__restrict__ bc,
                                                             const float *__restrict__ bias, const float *__restrict__ mean,
                                                             float *__restrict__ out, int N, int H, int V, int softmax)
{
    constexpr int A_F4 = kBM * kBK / 4;                    // float4 per A tile (one of hi / lo)
    constexpr int B_F4 = VP * kBK / 4;
    constexpr int TCOLS = VP <= 32 ? 64 : 128;             // 2 VP accumulator columns (below), allocated in powers of two
    extern __shared__ __align__(1024) unsigned char hsm_raw[];
    __shared__ __align__(8) uint64_t mbar[3];              // [0], [1]: stage free again; [2]: accumulator complete
    __shared__ uint32_t tmem_base_s;
    constexpr int STAGE_F4 = 2 * A_F4 + 2 * B_F4;          // stage s: [a_hi | a_lo | b_hi | b_lo]
    // (the swizzle is a function of absolute shared-memory address bits: every operand tile starts on a 1024-byte line)
    unsigned char *const hsm = hsm_raw + ((1024u - (smem_u32(hsm_raw) & 1023u)) & 1023u);
    float4 *const stage0 = (float4 *)hsm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * kBM;
    const int nk = (H + kBK - 1) / kBK;

    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init(&mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();                                      // (.sync.aligned below: the warp must be converged after the tid == 0 block)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // A quarter-warp loads the 128 bytes one row has in a tile (32 features) and stores them as the 8 chunks of that row's
    // swizzled line: coalesced 128-byte global segments, conflict-free 16-byte stores.  This thread: chunk h4l of rows
    // i*32 + warp*4 + rq, i = 0..3  (round 1 loaded 32-byte pieces of 16 rows per instruction: 61 % L1 throughput, the
    // x stream at 24 % of the DRAM peak)
    const int h4l = lane & 7, rq = lane >> 3;
    const int rl = warp * 4 + rq;                          // row of the thread inside a group of 32
    const int aoff = rl * 8 + (h4l ^ (rl & 7));            // float4 slot; + i * 256 for the group
    auto gload = [&](int kt, float4 (&xr)[kXPT], float4 &mu) {
        const int k = kt * kBK + h4l * 4;
#pragma unroll
        for (int i = 0; i < kXPT; ++i) {
            const int n = n0 + i * 32 + rl;
            xr[i] = (n < N && k < H) ? ldg_stream4(x + (size_t)n * H + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mu = (k < H) ? __ldg((const float4 *)(mean + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, N >> 3, M >> 4
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(VP >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    // the same with N = 2 VP: the lo weight tile follows the hi tile as the next VP class rows, so x_hi x [w_hi | w_lo] is ONE
    // MMA into 2 VP accumulator columns: two MMAs per K step instead of three, x_hi read from shared memory once, not twice
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * VP) >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    constexpr int BQ = (B_F4 + kThreads - 1) / kThreads;   // weight float4 per thread and part
    // the weight tile (L2-resident) of a stage is requested one stage ahead: its L2 round trip sat exposed between the
    // transform and the MMAs of every stage in round 1
    auto wload = [&](int kt, float4 (&bq)[2 * BQ]) {
#pragma unroll
        for (int q = 0; q < BQ; ++q) {
            if (tid + q * kThreads < B_F4) {
                bq[q] = __ldg(bc + (size_t)(kt * 2 + 0) * B_F4 + tid + q * kThreads);
                bq[BQ + q] = __ldg(bc + (size_t)(kt * 2 + 1) * B_F4 + tid + q * kThreads);
            }
        }
    };

    auto stage_body = [&](int kt, float4 (&xr)[kXPT], float4 &mu, float4 (&bq)[2 * BQ], float4 (&bnext)[2 * BQ]) {
        const int s = kt & 1;
        float4 *const Ahi = stage0 + s * STAGE_F4, *const Alo = Ahi + A_F4, *const Bhi = Alo + A_F4, *const Blo = Bhi + B_F4;
        if (kt + 1 < nk) wload(kt + 1, bnext);
        if (kt >= 2) mbar_wait(&mbar[s], (uint32_t)(((kt >> 1) - 1) & 1));     // the MMAs of tile kt-2 have read stage s
#pragma unroll
        for (int i = 0; i < kXPT; ++i) {
            const float v0 = xr[i].x - mu.x, v1 = xr[i].y - mu.y, v2 = xr[i].z - mu.z, v3 = xr[i].w - mu.w;
            const float h0 = __uint_as_float(tf32_rna(v0)), h1 = __uint_as_float(tf32_rna(v1));
            const float h2 = __uint_as_float(tf32_rna(v2)), h3 = __uint_as_float(tf32_rna(v3));
            Ahi[aoff + i * 256] = make_float4(h0, h1, h2, h3);
            Alo[aoff + i * 256] = make_float4(v0 - h0, v1 - h1, v2 - h2, v3 - h3);      // exact in fp32; the MMA reads its top 19 bits
        }
#pragma unroll
        for (int q = 0; q < BQ; ++q) {
            if (tid + q * kThreads < B_F4) {
                Bhi[tid + q * kThreads] = bq[q];
                Blo[tid + q * kThreads] = bq[BQ + q];
            }
        }
        if (kt + kPF < nk) gload(kt + kPF, xr, mu);        // (this buffer is free again)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic-proxy stores -> visible to the MMA
        __syncthreads();
        // Tried on the box and dropped (T = 750, B = 256, H = 800, V = 29; this version: 176 us):
        //  * no block barrier -- each warp counts itself in with an atomic and the eighth to arrive issues the MMAs, so that
        //    no warp waits for the slowest: 211 us (weight gradient 156 -> 193 us);
        //  * row blocks sized for whole waves (109 instead of 128 rows, 5.95 instead of 5.07 waves): 181 us;
        //  * 256-byte L2 fills on the x loads (DRAM locality of the 128-byte row pieces): 175 us;
        //  * persistent CTAs (two per SM) walking the row blocks with the prefetch running across block boundaries, two
        //    TMEM accumulators and the epilogue of a block deferred into the next block's second stage: 216 us;
        //  * (against the 168 us of the two-MMA version) the weight tiles by bulk async copies -- cp.async.bulk, the TMA
        //    engine -- into a ring of four 8 KB slots, requested two stages ahead and awaited by the MMA-issuing thread,
        //    instead of through registers: 189 us.
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), bh = smem_u32(Bhi);
#pragma unroll
            for (int j = 0; j < kBK / 8; ++j) {            // one MMA covers K = 8 tf32 = 32 bytes inside the 128-byte swizzled rows
                const uint32_t o = j * 32;
                // K-major SWIZZLE_128B: SBO = 1024 B between groups of 8 rows; LBO is not used
                const uint64_t dah = smem_desc_sw(ah + o, 16, 1024, 2), dal = smem_desc_sw(al + o, 16, 1024, 2);
                const uint64_t dbh = smem_desc_sw(bh + o, 16, 1024, 2);
                mma_tf32(tmem, dah, dbh, idesc2, (kt > 0 || j > 0) ? 1u : 0u);  // columns [0, VP): hi x hi, [VP, 2 VP): hi x lo
                mma_tf32(tmem, dal, dbh, idesc, 1u);                             // columns [0, VP) += lo x hi
            }
            mma_commit(&mbar[s]);
            if (kt == nk - 1) mma_commit(&mbar[2]);
        }
    };

    float4 xa[kXPT], xb[kXPT], xc[kXPT], xd[kXPT], ma, mb, mc, md;
    float4 bA[2 * BQ], bB[2 * BQ];
    gload(0, xa, ma);
    wload(0, bA);
    if (nk > 1) gload(1, xb, mb);
    if (nk > 2) gload(2, xc, mc);
    if (nk > 3) gload(3, xd, md);
    for (int kt = 0; kt < nk; kt += kPF) {
        stage_body(kt, xa, ma, bA, bB);
        if (kt + 1 < nk) stage_body(kt + 1, xb, mb, bB, bA);
        if (kt + 2 < nk) stage_body(kt + 2, xc, mc, bA, bB);
        if (kt + 3 < nk) stage_body(kt + 3, xd, md, bB, bA);
    }
    mbar_wait(&mbar[2], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warps 0-3, thread = row (TMEM lane 32*warp + lane), VP columns
    if (warp < 4) {
    float val[VP];
#pragma unroll
    for (int h = 0; h < VP / 16; ++h) {
        uint32_t v[16], w[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + h * 16, v);
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + VP + h * 16, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) val[h * 16 + j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
    }
#pragma unroll
    for (int j = 0; j < VP; ++j) val[j] += __ldg(bias + j);
    if (softmax) {
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < VP; ++j) if (j < V) m = fmaxf(m, val[j]);
        float ssum = 0.f;
#pragma unroll
        for (int j = 0; j < VP; ++j) { val[j] = (j < V) ? expf(val[j] - m) : 0.f; ssum += val[j]; }
        const float inv = 1.f / ssum;
#pragma unroll
        for (int j = 0; j < VP; ++j) val[j] *= inv;
    }
    // The 128 x V block of logits is contiguous in the T x B x V output: stage it in shared memory (the operand stages
    // are idle now; row stride V, odd for both alphabets => conflict-free) and write it with coalesced 16-byte stores.
    float *srow = (float *)hsm + (warp * 32 + lane) * V;
#pragma unroll
    for (int j = 0; j < VP; ++j) if (j < V) srow[j] = val[j];
    }
    float *so = (float *)hsm;
    __syncthreads();
    {
        const int rows = min(kBM, N - n0);
        const int total = rows * V;
        float *o = out + (size_t)n0 * V;                   // (n0 * V * 4 bytes is a multiple of 512)
        if ((((uintptr_t)out) & 15) == 0) {
            for (int i = tid; i < (total >> 2); i += kThreads) ((float4 *)o)[i] = ((const float4 *)so)[i];
            for (int i = (total & ~3) + tid; i < total; i += kThreads) o[i] = so[i];
        } else {
            for (int i = tid; i < total; i += kThreads) o[i] = so[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
}

}  // namespace tc
}  // namespace ctcb200
