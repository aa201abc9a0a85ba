// ctc_head.cu -- the classifier head that produces the CTC activations, fused (sm_100a).
//
// SURVEY.md section 8(f) row 4.  Reference: /root/reference/codes/model.py:177-180 and 205-222
//     fc = SequenceWise(Sequential(BatchNorm1d(H), Linear(H, V, bias=False)));  x = fc(x).transpose(0, 1)
//     (eval mode: F.softmax(x, dim=-1))
// i.e. on the N = T*B rows of the last recurrent layer's output x[N][H] (H = 800, V = 29 / 43):
//     training:  mu, var = batch statistics over the N rows;   y = (x - mu) / sqrt(var + eps) * gamma + beta
//     logits = y W^T                                            (W is V x H)
// PyTorch runs this as 3 kernels over x-sized tensors forward and 5 backward.  Here, with V <= 64 the op is a
// *skinny* product -- 2*V flops per 4 bytes of x, far below the tensor-core ridge and close to the fp32 FMA
// ridge -- so it is organised around reading x as few times as possible.  The forward product runs on the tensor
// cores with an error-compensated tf32 split, and since round 2 so do both backward products (the reference computes in
// fp32; no plain tf32/bf16 rounding is introduced anywhere):
//
//   forward   head_stats_kernel   one pass over x: per-feature sum / sum of squares (shifted by row 0, fp64 merge)
//             head_fold_kernel    BatchNorm's scale folded into the weights: Wk[h][v] = W[v][h] * gamma_h * invstd_h;
//                                 running statistics updated
//             head_bias_kernel    bias'[v] = sum_h W[v][h] * beta_h
//             head_fold_tc_kernel, head_fwd_tc_kernel (ctc_head_tc.cuh)   one pass over x on the TENSOR CORES:
//                                 logits = (x - mu) Wk + bias' as tcgen05.mma kind::tf32 with a 3xTF32 operand split
//                                 (fp32-level accuracy), accumulator in TMEM, bias / softmax in the tcgen05.ld epilogue.
//                                 The mean is subtracted from the staged tile, NOT folded into the bias: with
//                                 |mu| >> sigma the folded form cancels catastrophically in fp32.  Rows are written in
//                                 T x B x V order -- exactly what ctc_fused_kernel reads
//   backward  head_colsum_kernel  s[v] = sum_n dlogits[n][v]
//             head_wgrad_tc_kernel (ctc_head_bwd_tc.cuh)  one pass over x on the tensor cores:
//                                 G[v][h] = sum_n dlogits[n][v] * (x[n][h] - mu_h)  (per-row-block partials, reduced in a
//                                 fixed order => deterministic)
//             head_reduce_kernel, head_finalize_kernel  from G and s alone: dW, dgamma, dbeta and the per-feature coefficients
//                                 of dx = A_h * (dlogits W)[n][h] + B_h + C_h * (x[n][h] - mu_h)   (BatchNorm's backward
//                                 needs only column reductions that are linear in G and s)
//             head_wt_tc_kernel, head_dgrad_tc_kernel (ctc_head_bwd_tc.cuh)  one pass over x, dx as above, the product
//                                 dlogits W on the tensor cores
// => x is read 2x forward and 2x backward and dx written once; nothing else of size N x H exists.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "../../include/ctc.h"
#include "ctc_internal.h"
#include "ctc_head_tc.cuh"
#include "ctc_head_bwd_tc.cuh"

namespace ctcb200 {

// 16-byte read-only load that ptxas may not sink next to its first use: a batch of these is issued back to back, so a
// thread really has the whole batch in flight (with plain __ldg the compiler re-used one register and serialised them)
__device__ __forceinline__ float4 hd_ldg_f4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// ---------------------------------------------------------------------------------------------------------
// batch statistics: sums[0][h] = sum_n (x[n][h] - x[0][h]),  sums[1][h] = sum_n (x[n][h] - x[0][h])^2
// (shifting by the first row removes the cancellation of E[x^2] - E[x]^2 for features with a large mean)
// rows requested per thread before the first is consumed (measured at T = 750, B = 256, H = 800: 8 rows 144 us, 16 rows
// 126 us, 32 rows 143 us)
#ifndef CTC_HEAD_STATS_ROWS
#define CTC_HEAD_STATS_ROWS 16
#endif
constexpr int kStatsRows = CTC_HEAD_STATS_ROWS;
__global__ void __launch_bounds__(256) head_stats_kernel(const float *__restrict__ x, int N, int H, int rows_per_cta,
                                                         double *__restrict__ sums)
{
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    const int H4 = H >> 2;
    for (int h4 = threadIdx.x; h4 < H4; h4 += blockDim.x) {
        const float4 sh = __ldg((const float4 *)x + h4);
        double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        for (int rb = r0; rb < r1; rb += 32) {             // fp32 partials over 32 rows, merged in fp64
            float a[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
            const int re = min(r1, rb + 32);
            for (int r = rb; r < re; r += kStatsRows) {    // kStatsRows independent row loads in flight per thread
                float4 v[kStatsRows];
#pragma unroll
                for (int u = 0; u < kStatsRows; ++u) v[u] = hd_ldg_f4(x + (size_t)min(r + u, re - 1) * H + h4 * 4);
#pragma unroll
                for (int u = 0; u < kStatsRows; ++u) {
                    const bool in = (r + u < re);          // (clamped rows were loaded twice: count them once)
                    const float d0 = in ? v[u].x - sh.x : 0.f, d1 = in ? v[u].y - sh.y : 0.f;
                    const float d2 = in ? v[u].z - sh.z : 0.f, d3 = in ? v[u].w - sh.w : 0.f;
                    a[0] += d0; a[1] += d1; a[2] += d2; a[3] += d3;
                    q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { s1[i] += (double)a[i]; s2[i] += (double)q[i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(&sums[h4 * 4 + i], s1[i]);
            atomicAdd(&sums[H + h4 * 4 + i], s2[i]);
        }
    }
}

// fold BatchNorm into the weights; one thread per feature
__global__ void head_fold_kernel(const float *__restrict__ x, const double *__restrict__ sums, int N, int H, int V, int VP,
                                 const float *__restrict__ gamma, const float *__restrict__ beta,
                                 float *__restrict__ running_mean, float *__restrict__ running_var, float eps,
                                 float momentum, int training, const float *__restrict__ W, float *__restrict__ wk,
                                 float *__restrict__ shift, float *__restrict__ save_mean, float *__restrict__ save_invstd)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const bool first = (blockIdx.y == 0);                  // blockIdx.y = slice of 8 classes; slice 0 also owns the statistics
    double mean, var;
    if (training) {
        const double m1 = sums[h] / N;
        mean = (double)x[h] + m1;
        var = fmax(sums[H + h] / N - m1 * m1, 0.0);        // biased, as BatchNorm normalises with
        if (running_mean && first) {                       // ... and unbiased into the running estimate
            const double unb = (N > 1) ? var * ((double)N / (double)(N - 1)) : var;
            running_mean[h] = (float)((1.0 - momentum) * running_mean[h] + momentum * mean);
            running_var[h] = (float)((1.0 - momentum) * running_var[h] + momentum * unb);
        }
    } else {
        mean = running_mean[h];
        var = running_var[h];
    }
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[h] : 1.0, bt = beta ? (double)beta[h] : 0.0;
    const double scale = g * invstd;
    if (first) {
        shift[h] = (float)bt;
        if (save_mean) save_mean[h] = (float)mean;
        if (save_invstd) save_invstd[h] = (float)invstd;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int v = blockIdx.y * 8 + j;
        o[j] = (v < V) ? (float)((double)W[(size_t)v * H + h] * scale) : 0.f;
    }
    float4 *dst = (float4 *)(wk + (size_t)h * VP + blockIdx.y * 8);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// bias'[v] = sum_h W[v][h] * shift[h]; one CTA per class
__global__ void __launch_bounds__(256) head_bias_kernel(const float *__restrict__ W, const float *__restrict__ shift, int H,
                                                        int V, float *__restrict__ bias)
{
    __shared__ double red[8];
    const int v = blockIdx.x;
    double s = 0.0;
    if (v < V)
        for (int h = threadIdx.x; h < H; h += 256) s += (double)W[(size_t)v * H + h] * (double)shift[h];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        bias[v] = (float)t;                                // classes v >= V get 0
    }
}

// ---------------------------------------------------------------------------------------------------------
// s[v] = sum_n dl[n][v].  The CTA's rows are one flat run of floats; the first NTV = (256 / V) * V threads walk it with
// stride NTV, so that a thread stays on class tid % V and a warp reads consecutive addresses (round 1 gave every warp one
// 116-byte row per load: 30 us for 22 MB); eight loads per thread are in flight.
__global__ void __launch_bounds__(256) head_colsum_kernel(const float *__restrict__ dl, int N, int V, int rows_per_cta,
                                                          double *__restrict__ s)
{
    __shared__ double part[256];
    const int tid = threadIdx.x;
    const int NTV = (256 / V) * V;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    const float *base = dl + (size_t)r0 * V;
    const long long total = (long long)max(r1 - r0, 0) * V;
    double acc = 0.0;
    if (tid < NTV) {
        for (long long f = tid; f < total; f += 8LL * NTV) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (f + (long long)u * NTV < total) ? __ldg(base + f + (long long)u * NTV) : 0.f;
            float a = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) a += v[u];
            acc += (double)a;
        }
    }
    part[tid] = acc;
    __syncthreads();
    if (tid < V) {                                         // threads tid, tid + V, ... hold class tid
        double t = 0.0;
        for (int i = tid; i < NTV; i += V) t += part[i];
        atomicAdd(&s[tid], t);
    }
}

// reduce the row-block partials in a fixed order (deterministic): one thread per (class, feature).
// Leaves gx[v][h] = sum_n dlogits[n][v] * xhat[n][h] in the first slab of `part` and writes dW.
__global__ void __launch_bounds__(128) head_reduce_kernel(float *__restrict__ part, int RB, int VP, const double *__restrict__ s,
                                                          int H, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                          const float *__restrict__ invstd, float *__restrict__ dW)
{
    const int h = blockIdx.x * 128 + threadIdx.x, v = blockIdx.y;
    if (h >= H) return;
    double G = 0.0;
    for (int rb = 0; rb < RB; ++rb) G += (double)part[((size_t)rb * VP + v) * H + h];
    const double gx = G * (double)invstd[h];               // (G is centred: no mu * s[v] term)
    part[(size_t)v * H + h] = (float)gx;
    if (dW) {
        const double g = gamma ? (double)gamma[h] : 1.0, bt = beta ? (double)beta[h] : 0.0;
        dW[(size_t)v * H + h] = (float)(g * gx + bt * s[v]);
    }
}

// one thread per feature: dgamma, dbeta and the coefficients of dx (BatchNorm's backward needs only these column sums)
__global__ void head_finalize_kernel(const float *__restrict__ gxs, const double *__restrict__ s, int N, int H, int V,
                                     const float *__restrict__ W, const float *__restrict__ gamma,
                                     const float *__restrict__ mean, const float *__restrict__ invstd, int training,
                                     float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ coef)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const double mu = mean[h], is = invstd[h];
    const double g = gamma ? (double)gamma[h] : 1.0;
    double db = 0.0, dg = 0.0;
    for (int v = 0; v < V; ++v) {
        const double w = W[(size_t)v * H + h];
        db += s[v] * w;
        dg += w * (double)gxs[(size_t)v * H + h];
    }
    if (dgamma) dgamma[h] = (float)dg;
    if (dbeta) dbeta[h] = (float)db;
    // dx[n][h] = A * (dlogits W)[n][h] + B + C * (x[n][h] - mu)
    const double A = g * is;
    double B = 0.0, C = 0.0;
    if (training) {                                        // batch statistics depend on x
        C = -g * dg * is * is / N;
        B = -A * db / N;
    }
    coef[h] = (float)A; coef[H + h] = (float)B; coef[2 * H + h] = (float)C; coef[3 * H + h] = (float)mu;
}

// ---------------------------------------------------------------------------------------------------------
namespace {

size_t up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int sm_count()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    return sms;
}

struct HeadLayout {
    int VP, VPW, RB, rows_per_block;   // classes padded for the forward / input-gradient MMAs (32, 48, 64) and for the weight gradient (32, 64)
    size_t off_sums, off_wk, off_shift, off_bias, off_s, off_coef, off_mean, off_invstd, off_part, off_bc, off_wtc, total;
};

HeadLayout head_layout(int N, int H, int V)
{
    HeadLayout l;
    l.VP = (V <= 32) ? 32 : (V <= 48) ? 48 : 64;
    l.VPW = (V <= 32) ? 32 : 64;
    const int tiles = (H + 127) / 128;
    int rb = std::max(1, (4 * 148) / tiles);               // at most two waves of two resident CTAs per SM for the weight-gradient pass
    int rows = (N + rb - 1) / rb;
    rows = std::max(32, (rows + 31) / 32 * 32);
    l.rows_per_block = rows;
    l.RB = (N + rows - 1) / rows;
    size_t o = 0;
    l.off_sums = o;   o += up(sizeof(double) * 2 * H);
    l.off_s = o;      o += up(sizeof(double) * 64);
    l.off_wk = o;     o += up(sizeof(float) * (size_t)H * l.VP);
    l.off_shift = o;  o += up(sizeof(float) * H);
    l.off_bias = o;   o += up(sizeof(float) * 64);
    l.off_coef = o;   o += up(sizeof(float) * 4 * H);
    l.off_mean = o;   o += up(sizeof(float) * H);
    l.off_invstd = o; o += up(sizeof(float) * H);
    l.off_part = o;   o += up(sizeof(float) * (size_t)l.RB * l.VPW * H);
    l.off_bc = o;     o += up(tc::head_tc_weight_bytes(H, l.VP));
    l.off_wtc = o;    o += up(tc::head_dgrad_tc_weight_bytes(H, l.VP));
    l.total = o;
    return l;
}

bool ok(cudaError_t e, const char *what, ctcStatus_t &st)
{
    if (e == cudaSuccess) return true;
    ctcb200_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    st = CTC_STATUS_EXECUTION_FAILED;
    return false;
}

ctcStatus_t bad(const char *msg)
{
    ctcb200_set_error(msg);
    return CTC_STATUS_INVALID_VALUE;
}

ctcStatus_t check_shape(int N, int H, int V)
{
    if (N <= 0 || H <= 0 || V <= 0) return bad("non-positive size");
    if (H % 4 != 0) return bad("features must be a multiple of 4");
    if (V > 64) {
        ctcb200_set_error("classes above 64 are not supported by this build");
        return CTC_STATUS_UNKNOWN_ERROR;
    }
    return CTC_STATUS_SUCCESS;
}

}  // namespace
}  // namespace ctcb200

using namespace ctcb200;

extern "C" {

ctcStatus_t ctc_b200_head_workspace_size(int rows, int features, int classes, size_t *size_bytes)
{
    if (!size_bytes) return bad("null pointer argument");
    ctcStatus_t st = check_shape(rows, features, classes);
    if (st != CTC_STATUS_SUCCESS) return st;
    *size_bytes = head_layout(rows, features, classes).total;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_head_forward(const ctcB200HeadForward *c)
{
    if (!c || !c->x || !c->weight || !c->out || !c->workspace) return bad("null pointer argument");
    const int N = c->rows, H = c->features, V = c->classes;
    ctcStatus_t st = check_shape(N, H, V);
    if (st != CTC_STATUS_SUCCESS) return st;
    if (!c->training && (!c->running_mean || !c->running_var)) return bad("eval mode needs the running statistics");
    if (((uintptr_t)c->x | (uintptr_t)c->weight) & 15) return bad("x and weight must be 16-byte aligned");
    const HeadLayout l = head_layout(N, H, V);
    if (c->workspace_bytes < l.total) return bad("workspace too small");
    char *ws = (char *)c->workspace;
    double *sums = (double *)(ws + l.off_sums);
    float *wk = (float *)(ws + l.off_wk), *shift = (float *)(ws + l.off_shift), *bias = (float *)(ws + l.off_bias);
    float *mean = c->save_mean ? c->save_mean : (float *)(ws + l.off_mean);
    cudaStream_t s = (cudaStream_t)c->stream;
    if (c->training) {
        if (!ok(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * H, s), "memset", st)) return st;
        // one balanced wave: as many CTAs as are resident at once (round 1 launched 1.6 waves of equal CTAs), and a
        // block of just enough warps for the H / 4 float4 columns (H = 800: 224 threads, 200 of them busy, instead of 256)
        const int threads = std::min(256, ((H >> 2) + 31) / 32 * 32);
        int per_sm = 4;
        if (!ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, head_stats_kernel, threads, 0), "occupancy", st)) return st;
        const int ctas = std::max(1, std::min((N + 31) / 32, sm_count() * std::max(1, per_sm)));
        const int rows = (N + ctas - 1) / ctas;
        head_stats_kernel<<<(N + rows - 1) / rows, threads, 0, s>>>(c->x, N, H, rows, sums);
        ctcb200_count_launch();
    }
    head_fold_kernel<<<dim3((H + 127) / 128, l.VP / 8), 128, 0, s>>>(c->x, sums, N, H, V, l.VP, c->bn_weight, c->bn_bias, c->running_mean,
                                                    c->running_var, c->eps, c->momentum, c->training, c->weight, wk, shift,
                                                    mean, c->save_invstd);
    head_bias_kernel<<<l.VP, 256, 0, s>>>(c->weight, shift, H, V, bias);
    {
        float4 *bc = (float4 *)(ws + l.off_bc);
        const int nk = (H + tc::kBK - 1) / tc::kBK;
        tc::head_fold_tc_kernel<<<(nk * 8 * l.VP + 127) / 128, 128, 0, s>>>(wk, H, l.VP, bc);
        const int smem = tc::head_tc_smem_bytes(l.VP);
        const int grid = (N + tc::kBM - 1) / tc::kBM;
#define FWD_LAUNCH(VPX) do { \
            if (!ok(cudaFuncSetAttribute(tc::head_fwd_tc_kernel<VPX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attribute", st)) return st; \
            tc::head_fwd_tc_kernel<VPX><<<grid, tc::kThreads, smem, s>>>(c->x, bc, bias, mean, c->out, N, H, V, c->softmax); } while (0)
        if (l.VP == 32) FWD_LAUNCH(32); else if (l.VP == 48) FWD_LAUNCH(48); else FWD_LAUNCH(64);
#undef FWD_LAUNCH
    }
    for (int i = 0; i < 4; ++i) ctcb200_count_launch();    // fold, bias, fold_tc, forward
    if (!ok(cudaGetLastError(), "head forward launch", st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_head_backward(const ctcB200HeadBackward *c)
{
    if (!c || !c->x || !c->dlogits || !c->weight || !c->save_mean || !c->save_invstd || !c->workspace)
        return bad("null pointer argument");
    const int N = c->rows, H = c->features, V = c->classes;
    ctcStatus_t st = check_shape(N, H, V);
    if (st != CTC_STATUS_SUCCESS) return st;
    if (((uintptr_t)c->x | (uintptr_t)c->weight | (uintptr_t)c->dx) & 15) return bad("x, weight and dx must be 16-byte aligned");
    const HeadLayout l = head_layout(N, H, V);
    if (c->workspace_bytes < l.total) return bad("workspace too small");
    char *ws = (char *)c->workspace;
    double *sv = (double *)(ws + l.off_s);
    float *coef = (float *)(ws + l.off_coef), *part = (float *)(ws + l.off_part);
    cudaStream_t s = (cudaStream_t)c->stream;
    if (!ok(cudaMemsetAsync(sv, 0, sizeof(double) * 64, s), "memset", st)) return st;
    {
        const int ctas = std::min((N + 63) / 64, sm_count() * 4);
        const int rows = (N + ctas - 1) / ctas;
        head_colsum_kernel<<<(N + rows - 1) / rows, 256, 0, s>>>(c->dlogits, N, V, rows, sv);
    }
    const dim3 gw((H + 127) / 128, l.RB);
    {
        const int smem = tc::head_wgrad_tc_smem_bytes(l.VPW);
        if (l.VPW == 32) {
            if (!ok(cudaFuncSetAttribute(tc::head_wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attribute", st)) return st;
            tc::head_wgrad_tc_kernel<32><<<gw, 256, smem, s>>>(c->x, c->dlogits, c->save_mean, part, N, H, V, l.rows_per_block);
        } else {
            if (!ok(cudaFuncSetAttribute(tc::head_wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attribute", st)) return st;
            tc::head_wgrad_tc_kernel<64><<<gw, 256, smem, s>>>(c->x, c->dlogits, c->save_mean, part, N, H, V, l.rows_per_block);
        }
    }
    head_reduce_kernel<<<dim3((H + 127) / 128, V), 128, 0, s>>>(part, l.RB, l.VPW, sv, H, c->bn_weight, c->bn_bias,
                                                                c->save_invstd, c->dweight);
    head_finalize_kernel<<<(H + 127) / 128, 128, 0, s>>>(part, sv, N, H, V, c->weight, c->bn_weight, c->save_mean,
                                                        c->save_invstd, c->training, c->dbn_weight, c->dbn_bias, coef);
    ctcb200_count_launch();
    ctcb200_count_launch(); ctcb200_count_launch(); ctcb200_count_launch();
    if (c->dx) {
        float4 *wtc = (float4 *)(ws + l.off_wtc);
        const int tiles = (H + tc::kDBM - 1) / tc::kDBM, nblk = (N + tc::kDBN - 1) / tc::kDBN;
        tc::head_wt_tc_kernel<<<(tiles * (l.VP / 4) * tc::kDBM + 127) / 128, 128, 0, s>>>(c->weight, H, V, l.VP, wtc);
        const int resident = (l.VP <= 48) ? 2 : 1;
        const int groups = std::max(1, std::min(nblk, (sm_count() * resident * 2 + tiles - 1) / tiles));   // two waves of resident CTAs
        const int bpc = (nblk + groups - 1) / groups;
        const dim3 gd(tiles, (nblk + bpc - 1) / bpc);
        const int smem = tc::head_dgrad_tc_smem_bytes(l.VP);
#define DG_LAUNCH(VPX) do { \
            if (!ok(cudaFuncSetAttribute(tc::head_dgrad_tc_kernel<VPX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attribute", st)) return st; \
            tc::head_dgrad_tc_kernel<VPX><<<gd, 256, smem, s>>>(c->x, c->dlogits, wtc, coef, c->dx, N, H, V, bpc); } while (0)
        if (l.VP == 32) DG_LAUNCH(32); else if (l.VP == 48) DG_LAUNCH(48); else DG_LAUNCH(64);
#undef DG_LAUNCH
        ctcb200_count_launch(); ctcb200_count_launch();
    }
    if (!ok(cudaGetLastError(), "head backward launch", st)) return st;
    return CTC_STATUS_SUCCESS;
}

}  // extern "C"
