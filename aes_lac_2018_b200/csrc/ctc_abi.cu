// ctc_abi.cu -- host side of libctc_b200.so: the C ABI declared in include/ctc.h.
//
// Replaces, for the GPU location, what warp-ctc's src/ctc_entrypoint.cu + GpuCTC::cost_and_grad do
// behind `warpctc_pytorch.CTCLoss` (reference train.py:12/179, codes/engine.py:22, codes/metrics.py:51):
// argument validation, per-utterance metadata, workspace carve-up, kernel dispatch, cost read-back.
// Nothing here allocates device memory; everything lives in the caller's workspace.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ctc.h"
#include "ctc_decode.cuh"
#include "ctc_editdist.cuh"
#include "ctc_internal.h"
#include "ctc_logspace.cuh"
#include "ctc_variants.h"

namespace {

using namespace ctcb200;

thread_local std::string g_last_error;
thread_local unsigned long long g_launches = 0;

ctcStatus_t fail(ctcStatus_t st, const std::string &msg)
{
    g_last_error = msg;
    return st;
}

// ---- kernel variants --------------------------------------------------------------------------
// (NS states per thread, W warps per utterance, K timesteps per chunk, VCH 32-symbol alphabet slices).
// SP = 32*NS*W padded states; an utterance with L labels fits when SP >= 2L + 2.
constexpr int kMaxLabelLen = 2047;           // (16, 8): SP = 4096
constexpr int kMaxSmem = 227 * 1024;
constexpr int kBidirMaxB = 160;             // bidirectional (two sweeps + combine) path for batches up to this size ...
constexpr size_t kBidirMaxBytes = 1u << 30;     // ... and up to this many bytes of spilled columns (B = 128, T = 1500,
                                                // L <= 200 -- one eighth of BASELINE configs[3] -- needs 786 MB)
constexpr int kWarpMinB = 640;              // automatic ladder choice: fp32 warp ladder from this batch size (measured
                                            // crossover at T = 750: 0.52 / 0.52 ms at B = 512, 0.59 / 0.87 ms at B = 768;
                                            // profiles/r2_crossover_w32.txt)
constexpr int kWarpMaxLabelLen = 255;       // NS = 16: 512 states hold 2L + 2
constexpr int kMaxLaunches = 32;            // queue counters: launches of one call (fp32 ladder: first and second tier)

const Variant *ladder_table(int ladder, int vch, int *n)
{
    switch (ladder * kMaxVch + (vch - 1)) {
    case 0: return ctc_variants_group0(n);
    case 1: return ctc_variants_group1(n);
    case 2: return ctc_variants_group2(n);
    case 3: return ctc_variants_group3(n);
    case 4: return ctc_variants_group4(n);
    case 5: return ctc_variants_group5(n);
    case 6: return ctc_variants_group6(n);
    case 7: return ctc_variants_group7(n);
    case 8: return ctc_variants_group8(n);
    case 9: return ctc_variants_group9(n);
    }
    *n = 0;
    return nullptr;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Plan {
    int B = 0, T_max = 0, V = 0;
    bool latency = false;
    std::vector<int> meta;                   // [label_off B | label_len B | act_len B | utt_ids B]
    struct Launch { const Variant *v; const Variant *v2 = nullptr;       // v2: fp64 second tier behind an fp32 warp variant
                    int first, count; size_t ckpt_off; long long ckpt_stride; int smem, smem2 = 0;
                    size_t col_off, exp_off, z_off; int exp_stride;      // bidirectional path (column spill) offsets
                    int slots;                                           // warp ladder: workspace slots (= max persistent CTAs)
                    long long frames = 0;                                // sum of the input lengths of the bucket
                    int sm_lo = 0, sm_hi = 0; };                         // warp ladders, several buckets: SM range of this one
    bool bidir = false;
    std::vector<Launch> launches;
    long long total_labels = 0;
    size_t off_meta = 0, off_labels = 0, off_costs = 0, off_status = 0, off_queue = 0, off_ckpt = 0, total = 0, ckpt_bytes = 0;
    int fallback_S = 1;
};

int persistent_grid(const void *kernel, int smem);
double warp_rel_cost(const Variant *v);

// forced_w: 0 = automatic ladder choice; otherwise use the latency ladder entry with that W where possible
ctcStatus_t make_plan(const int *label_lengths, const int *input_lengths, int V, int B, int T_max,
                      bool want_grad, int mode, Plan &plan, bool size_only = false)
{
    if (!label_lengths || !input_lengths) return fail(CTC_STATUS_INVALID_VALUE, "null length array");
    if (V <= 0 || B <= 0 || T_max < 0) return fail(CTC_STATUS_INVALID_VALUE, "non-positive size");
    plan.B = B; plan.T_max = T_max; plan.V = V;
    plan.launches.clear();
    plan.meta.resize((size_t)4 * B);                     // (every entry is written below; Plan objects are reused per thread,
                                                         //  so this does not allocate in the steady state)
    int *label_off = plan.meta.data(), *label_len = label_off + B, *act_len = label_len + B,
        *utt_ids = act_len + B;
    long long off = 0;
    int max_L = 0;
    for (int b = 0; b < B; ++b) {
        const int L = label_lengths[b], T = input_lengths[b];
        if (L < 0 || T < 0) return fail(CTC_STATUS_INVALID_VALUE, "negative length");
        if (T > T_max) return fail(CTC_STATUS_INVALID_VALUE, "input_lengths[b] exceeds max_time");
        if (L > kMaxLabelLen)
            return fail(CTC_STATUS_UNKNOWN_ERROR, "label sequence longer than 2047 is not supported");
        if (off > 0x7fffffffLL - L) return fail(CTC_STATUS_INVALID_VALUE, "too many labels");
        label_off[b] = (int)off; label_len[b] = L; act_len[b] = T;
        off += L;
        max_L = std::max(max_L, L);
    }
    plan.total_labels = off;

    // mode: 0 auto, 1 throughput ladder (16-step chunks), 2 latency ladder, 3 throughput ladder (8-step chunks),
    // 4 warp ladder (ctc_warp.cuh: one warp per utterance, register-resident, persistent CTAs),
    // 5 fp32 warp ladder (ctc_warp32.cuh: the same organisation, single-precision recursion with per-lane exponents).
    // Auto (measured on B200, T=750, L~U{50..200}; profiles/r2_crossover_w32.txt): below kWarpMinB utterances the GPU is
    // far from full with one warp per utterance, so spend more warps per utterance (latency ladder); above, the fp32
    // warp ladder wins.
    int vch = (V + 31) / 32;
    if (vch > kMaxVch)
        return fail(CTC_STATUS_UNKNOWN_ERROR, "alphabet_size above 64 is not supported by this build");
    const int vch_warp = V / 32 + 1;                     // the warp ladder needs one pad lane (r = 0) after the alphabet
    if (mode == 0) mode = (B < kWarpMinB) ? 2 : 5;
    if ((mode == 4 || mode == 5) && (vch_warp > kMaxVch || max_L > kWarpMaxLabelLen)) mode = 3;
    plan.latency = (mode == 2);
    if (mode == 4 || mode == 5) vch = vch_warp;
    int nl = 0;
    const Variant *ladder = ladder_table(mode == 2 ? LADDER_LATENCY : mode == 3 ? LADDER_THROUGHPUT_K8
                                         : mode == 4 ? LADDER_WARP : mode == 5 ? LADDER_WARP32 : LADDER_THROUGHPUT, vch, &nl);

    // bucket utterances by variant (counting sort: big variants first), longest first inside a bucket when the
    // lengths are ragged (tail balance).  O(B) unless T varies.
    thread_local std::vector<int> cls, cls_of_len;       // scratch reused across calls (a fresh 32 KB+ vector per call
    cls.resize(B);                                       //  costs more than the whole pass: mmap / page faults)
    std::vector<int> count(nl, 0);
    std::vector<long long> frames(nl, 0);
    int t_min = 0x7fffffff, t_max = 0;
    // Small batches (bidirectional path): one variant for everybody -- the per-step latency of the latency ladder
    // barely depends on the variant, while every extra bucket costs two more launches and a stream fork/join.
    cls_of_len.resize((size_t)max_L + 1);                // label length -> variant index, once per call
    for (int L = 0, c = 0; L <= max_L; ++L) {
        while (c < nl && ladder[c].max_label() < L) ++c;
        if (c >= nl) return fail(CTC_STATUS_UNKNOWN_ERROR, "no kernel variant for this label length");
        cls_of_len[L] = c;
    }
    // The bidirectional path spills every column of two sweeps: 2 * B * T_max * SP * 4 bytes with SP taken from the
    // LONGEST transcript (one bucket).  Long utterances would turn that into gigabytes of workspace next to the
    // model (B = 96, T = 1500, L = 400: 1.2 GB), so it is only taken while it stays under a fixed budget; above it
    // the checkpointed three-sweep kernel (no spill) runs instead.
    const size_t bidir_spill = sizeof(unsigned) * 2 * (size_t)B * (size_t)T_max * (size_t)ladder[cls_of_len[max_L]].sp();
    const bool one_bucket = (mode == 2) && want_grad && B <= kBidirMaxB && bidir_spill <= kBidirMaxBytes;
    for (int b = 0; b < B; ++b) {
        cls[b] = cls_of_len[one_bucket ? max_L : label_len[b]];
        ++count[cls[b]];
        frames[cls[b]] += act_len[b];
        t_min = std::min(t_min, act_len[b]);
        t_max = std::max(t_max, act_len[b]);
    }
    std::vector<int> start(nl + 1, 0);                   // launch order: descending variant index
    for (int c = nl - 1, o2 = 0; c >= 0; --c) { start[c] = o2; o2 += count[c]; }
    if (!size_only) {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int b = 0; b < B; ++b) utt_ids[fill[cls[b]]++] = b;
        if (t_min != t_max) {
            std::vector<unsigned long long> keys;
            for (int c = 0; c < nl; ++c) {
                if (count[c] < 2) continue;
                keys.resize(count[c]);
                int *seg = utt_ids + start[c];
                for (int i = 0; i < count[c]; ++i)      // descending T, ascending index
                    keys[i] = ((unsigned long long)(unsigned)(0x7fffffff - act_len[seg[i]]) << 32) | (unsigned)seg[i];
                std::sort(keys.begin(), keys.end());
                for (int i = 0; i < count[c]; ++i) seg[i] = (int)(keys[i] & 0xffffffffu);
            }
        }
    }

    size_t o = 0;
    plan.off_meta = o;   o = align_up(o + sizeof(int) * 4 * (size_t)B, 256);
    plan.off_labels = o; o = align_up(o + sizeof(int) * (size_t)std::max<long long>(off, 1), 256);
    plan.off_costs = o;  o = align_up(o + sizeof(float) * (size_t)B, 256);
    plan.off_status = o; o = align_up(o + sizeof(int) * (size_t)B, 256);
    plan.off_queue = o;  o = align_up(o + sizeof(int) * 4 * (kMaxLaunches + 1), 256);   // per launch (+1: the log-space detour): work queue, retired CTAs, claimed slots
    plan.off_ckpt = o;
    size_t ck = 0, bd = 0;
    plan.bidir = one_bucket;
    for (int c = nl - 1; c >= 0; --c) {                  // launch order: descending variant index
        if (count[c] == 0) continue;
        const Variant *v = &ladder[c];
        Plan::Launch l;
        l.v = v; l.first = start[c]; l.count = count[c]; l.frames = frames[c];
        const int nC = (T_max + v->K - 1) / v->K;
        // per CTA: nC checkpoint columns (SP doubles each) followed by nC p~ images
        l.slots = l.count;
        if (v->warp) {                                   // per resident CTA: 32-bit checkpoints, r images, 1/s
            l.slots = l.count;                            // (cut down to the resident warps below)
            long long words = v->slot_words(T_max);
            if (v->warp >= 3) {                          // second tier: the fp64 warp variant of the same NS shares the slots
                int n2 = 0;
                const Variant *t2 = ladder_table(LADDER_WARP, vch, &n2);
                for (int i = 0; i < n2; ++i)
                    if (t2[i].NS == v->NS) l.v2 = &t2[i];
                if (!l.v2) return fail(CTC_STATUS_UNKNOWN_ERROR, "no fp64 second-tier variant for this label length");
                words = std::max(words, l.v2->slot_words(T_max));
                l.smem2 = l.v2->smem_bytes(V, T_max);
            }
            l.ckpt_stride = want_grad ? (words + 1) / 2 : 0;
        } else {
            l.ckpt_stride = want_grad ? (long long)nC * v->sp() + (long long)nC * (pimg_bytes(v->K, V) / 8) : 0;
        }
        // bidirectional path: 2 slots (forward, reversed) of T_max columns of SP high words, exponents, Z
        l.exp_stride = nC + 2;
        l.col_off = bd;  bd = align_up(bd + sizeof(unsigned) * 2 * (size_t)l.count * (size_t)T_max * v->sp(), 256);
        l.exp_off = bd;  bd = align_up(bd + sizeof(int) * 2 * (size_t)l.count * l.exp_stride, 256);
        l.z_off = bd;    bd = align_up(bd + sizeof(double) * 2 * (size_t)l.count * 4, 256);
        l.smem = v->smem_bytes(V, T_max);
        if (l.smem > kMaxSmem)
            return fail(CTC_STATUS_UNKNOWN_ERROR,
                        "alphabet_size / max_time too large for the shared-memory layout of this kernel");
        if (plan.bidir && (!v->combine || combine_smem_bytes(v->sp(), V) > kMaxSmem)) plan.bidir = false;
        plan.launches.push_back(l);
    }
    // Warp ladders: workspace slots are per RESIDENT CTA.  Several buckets in one call: every bucket gets a contiguous
    // range of SMs in proportion to its share of the work, so that all buckets run side by side from the start and
    // finish together (outside_sm_range() in ctc_fused.cuh).  Full-size grids without ranges fill the SMs in launch
    // order: the later buckets only start as earlier CTAs retire and the call ends with a long, half-empty tail
    // (12-16 % below the per-class throughputs); smaller grids that share SMs thrash the instruction cache (2x slower).
    if (!plan.launches.empty() && plan.launches[0].v->warp) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        const int n = (int)plan.launches.size();
        const bool ranges = n > 1 && n <= sms && !std::getenv("CTC_B200_FULL_GRIDS");
        std::vector<int> per_sm(n), nsm(n, 1);
        for (int i = 0; i < n; ++i)
            per_sm[i] = std::max(1, persistent_grid((const void *)plan.launches[i].v->kernel, plan.launches[i].smem) / sms);
        if (ranges) {
            // Estimated finish time of bucket i on m SMs: work / m (a fluid model; counting whole rounds of utterances per
            // resident warp instead was measured to be worse -- warps speed up as their neighbours retire).  Start from
            // one SM each and hand the remaining SMs, one at a time, to the bucket that currently finishes last.
            auto finish = [&](int i, int m) {
                const Plan::Launch &l = plan.launches[i];
                return (double)std::max<long long>(l.frames, 1) * warp_rel_cost(l.v) / m;
            };
            for (int left = sms - n; left > 0; --left) {
                int worst = 0;
                double tw = -1.0;
                for (int i = 0; i < n; ++i) {
                    const double t = finish(i, nsm[i]);
                    if (t > tw) { tw = t; worst = i; }
                }
                ++nsm[worst];
            }
            int used = 0;
            for (int i = 0; i < n; ++i) used += nsm[i];
            for (int i = 0; used < sms; i = (i + 1) % n, ++used) ++nsm[i];   // leftovers: spread
        }
        int cut = 0;
        for (int i = 0; i < n; ++i) {
            Plan::Launch &l = plan.launches[i];
            if (ranges) {
                l.sm_lo = cut; l.sm_hi = cut + nsm[i];
                cut = l.sm_hi;
                l.slots = std::min(l.count, per_sm[i] * nsm[i] + 1);
                if (!size_only && std::getenv("CTC_B200_PLAN_DEBUG"))
                    std::fprintf(stderr, "bucket NS=%d count=%d per_sm=%d SMs [%d,%d) rounds %.2f\n", l.v->NS, l.count, per_sm[i],
                                 l.sm_lo, l.sm_hi, (double)l.count / (per_sm[i] * nsm[i]));
            } else {
                l.slots = std::min(l.count, per_sm[i] * sms);
            }
        }
    }
    for (Plan::Launch &l : plan.launches) {
        l.ckpt_off = ck;
        ck += sizeof(double) * (size_t)l.ckpt_stride * (size_t)l.slots;
    }
    if (plan.bidir) ck = std::max(ck, bd);
    // the checkpoint area doubles as the alpha store of the log-space fallback: keep room for one utterance
    plan.fallback_S = 2 * max_L + 1;
    ck = std::max(ck, sizeof(double) * (size_t)std::max(T_max, 1) * (size_t)plan.fallback_S);
    plan.ckpt_bytes = ck;
    plan.total = align_up(o + ck, 256) + 256;
    return CTC_STATUS_SUCCESS;
}

// Auxiliary streams so that the per-variant launches of one call overlap on the GPU instead of running
// back to back (each variant's grid alone rarely fills 148 SMs).  Forked from / joined to the caller's
// stream with events; created lazily per (thread, device).
constexpr int kAuxStreams = 6;
constexpr int kPipeStreams = 3;               // copy-in, kernels, copy-out
constexpr int kPipeBuffers = 4;               // buffer sets of the host-buffer pipeline
struct AuxStreams {
    bool ready = false;
    cudaStream_t s[kAuxStreams];
    cudaEvent_t fork, join[kAuxStreams];
    cudaStream_t pipe[kPipeStreams];           // host-buffer entry point: H2D / compute / D2H pipeline
    cudaEvent_t pipe_fork, pipe_join[kPipeStreams];
    cudaEvent_t t0, t1;                        // timing events (kernel_ms_host)
    std::vector<cudaEvent_t> ev_in, ev_k, ev_out;      // host-buffer pipeline: per-chunk stage completion events
    bool chunk_events(size_t n)
    {
        while (ev_in.size() < n) {
            cudaEvent_t e[3];
            for (int i = 0; i < 3; ++i)
                if (cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) != cudaSuccess) return false;
            ev_in.push_back(e[0]); ev_k.push_back(e[1]); ev_out.push_back(e[2]);
        }
        return true;
    }
};
thread_local int *g_status_dev_override = nullptr;      // set by the host-buffer entry point around run()
thread_local const int *g_labels_dev_override = nullptr; // ... labels of the slice, already on the device (uploaded once for the batch)
thread_local AuxStreams g_aux[16];

AuxStreams *aux_streams()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    AuxStreams &a = g_aux[dev];
    if (!a.ready) {
        for (int i = 0; i < kAuxStreams; ++i) {
            if (cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < kPipeStreams; ++i) {
            if (cudaStreamCreateWithFlags(&a.pipe[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&a.pipe_join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&a.pipe_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreate(&a.t0) != cudaSuccess || cudaEventCreate(&a.t1) != cudaSuccess) return nullptr;
        a.ready = true;
    }
    return &a;
}

// Per-thread pinned host scratch (grown on demand, kept for the life of the thread; null if the allocation fails, in which
// case the callers fall back to pageable targets).
char *pinned_scratch(size_t bytes)
{
    thread_local char *buf = nullptr;
    thread_local size_t cap = 0;
    if (bytes > cap) {
        if (buf) cudaFreeHost(buf);
        buf = nullptr; cap = 0;
        const size_t want = std::max<size_t>(bytes, 1 << 16);
        void *p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) == cudaSuccess) { buf = (char *)p; cap = want; }
        else (void)cudaGetLastError();
    }
    return buf;
}

bool check(cudaError_t e, const char *what, ctcStatus_t code, ctcStatus_t &out)
{
    if (e == cudaSuccess) return true;
    out = fail(code, std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a launch needs more than any earlier one of the same
// (device, kernel).  The attribute is per-process state, so the record is process-wide and the limit only ever
// grows: a per-thread record let one thread lower the limit under another thread's larger launch.
bool ensure_smem_attr(const void *kernel, int smem, ctcStatus_t &st)
{
    struct Key { const void *k; int dev; int smem; };
    static std::mutex mu;
    static std::vector<Key> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (Key &e : done)
        if (e.k == kernel && e.dev == dev) {
            if (e.smem >= smem) return true;
            if (!check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                       "cudaFuncSetAttribute(smem)", CTC_STATUS_EXECUTION_FAILED, st)) return false;
            e.smem = smem;
            return true;
        }
    if (!check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
               "cudaFuncSetAttribute(smem)", CTC_STATUS_EXECUTION_FAILED, st)) return false;
    // (experiment: one shared-memory carve-out for every kernel -- measured slightly slower than the driver's choice)
    if (std::getenv("CTC_B200_MAX_CARVEOUT") &&
        !check(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
               "cudaFuncSetAttribute(carveout)", CTC_STATUS_EXECUTION_FAILED, st)) return false;
    done.push_back(Key{kernel, dev, smem});
    return true;
}

// Persistent grid of a warp-ladder kernel: resident CTAs per SM (occupancy query, cached per kernel and device) x SMs.
int persistent_grid(const void *kernel, int smem)
{
    struct Key { const void *k; int dev; int smem; int grid; };
    thread_local std::vector<Key> done;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const Key &e : done)
        if (e.k == kernel && e.dev == dev && e.smem == smem) return e.grid;
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32, smem) != cudaSuccess) per_sm = 8;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    const int grid = std::max(1, per_sm) * std::max(1, sms);
    done.push_back(Key{kernel, dev, smem, grid});
    return grid;
}

// Relative cost of one utterance-frame per warp-ladder variant (measured per label class on B200 at B = 8192, T = 750:
// profiles/r2_variant_matrix.txt, profiles/r2_w32_variants.txt).  Only the ratios matter: they size the persistent
// grids of the buckets of one call so that all buckets finish together (below).
double warp_rel_cost(const Variant *v)
{
    static const double c64[9] = {0, 1.20, 1.44, 1.80, 2.21, 2.70, 3.28, 3.80, 4.33};   // NS = 2 .. 16, fp64 recursion
    static const double c32[9] = {0, 1.07, 1.28, 1.54, 1.79, 2.07, 2.46, 2.70, 3.18};   // fp32 recursion
    const int i = std::max(1, std::min(v->NS / 2, 8));
    return v->warp >= 3 ? c32[i] : c64[i];
}

// Enqueue the device-side log-space detour behind the fast kernels of this call (ctc_logspace.cuh): persistent CTAs
// scan the status words and redo the utterances flagged RANGE / INF_COST in fp64 log space.  The fast kernels have
// been ordered before it on `stream`, so their checkpoint area is free: it is cut into alpha slots of
// 8 * T_max * S_max bytes, one per CTA.
ctcStatus_t launch_logspace_detour(const ctcB200Call &c, const Plan &plan, const FusedParams &FP, int *d_queue,
                                   cudaStream_t stream)
{
    ctcStatus_t st = CTC_STATUS_SUCCESS;
    const int B = c.minibatch, V = c.alphabet_size;
    char *ws = (char *)c.workspace;
    const int S_max = plan.fallback_S;
    const size_t slot_doubles = (size_t)c.max_time * (size_t)S_max;
    const size_t n_slots = plan.ckpt_bytes / (slot_doubles * sizeof(double));
    if (n_slots == 0) return fail(CTC_STATUS_EXECUTION_FAILED, "workspace too small for the log-space detour");
    const int smem = logspace_smem_bytes(S_max, V);
    if (smem > kMaxSmem) return fail(CTC_STATUS_UNKNOWN_ERROR, "label sequence too long for the log-space detour");
    if (!ensure_smem_attr((const void *)ctc_logspace_kernel, smem, st)) return st;
    LogParams L;
    L.acts = FP.acts; L.act_stride_t = FP.act_stride_t; L.act_stride_b = FP.act_stride_b;
    L.grads = FP.grads;
    L.labels = FP.labels; L.label_off = FP.label_off; L.label_len = FP.label_len; L.act_len = FP.act_len;
    L.queue = d_queue; L.n_utts = B;
    L.costs = FP.costs; L.status = FP.status;
    L.alpha_ws = (double *)(ws + plan.off_ckpt);
    L.slot_stride = (long long)slot_doubles;
    L.V = V; L.T_max = c.max_time; L.B = B; L.blank = c.blank_label; L.S_max = S_max;
    L.grad_scale = c.grad_scale;
    const int grid = (int)std::min<size_t>(std::min<size_t>(n_slots, 296), (size_t)B);
    ctc_logspace_kernel<<<grid, kLogThreads, smem, stream>>>(L);
    ++g_launches;
    if (!check(cudaGetLastError(), "log-space detour launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t run(const ctcB200Call &c)
{
    if (!c.activations || !c.flat_labels || !c.label_lengths || !c.input_lengths || !c.workspace)
        return fail(CTC_STATUS_INVALID_VALUE, "null pointer argument");
    if (c.alphabet_size <= 0 || c.minibatch <= 0 || c.max_time <= 0)
        return fail(CTC_STATUS_INVALID_VALUE, "non-positive size");
    if (c.blank_label < 0 || c.blank_label >= c.alphabet_size)
        return fail(CTC_STATUS_INVALID_VALUE, "blank_label outside the alphabet");
    if ((c.flags & ~0x70fu) || ((c.flags >> 8) & 0x7) > 5) return fail(CTC_STATUS_INVALID_VALUE, "unknown bits in flags");
    const bool no_sync = (c.flags & CTC_B200_FLAG_NO_SYNC) != 0;
    if (no_sync && (c.costs_host || c.status_host))
        return fail(CTC_STATUS_INVALID_VALUE, "NO_SYNC cannot return host costs/status");
    if (!c.costs_host && !c.costs_device)
        return fail(CTC_STATUS_INVALID_VALUE, "no destination for the costs");

    const int B = c.minibatch, V = c.alphabet_size;
    const bool want_grad = c.gradients != nullptr;
    thread_local Plan plan;                              // reused per thread: no allocation in the steady state
    ctcStatus_t st = make_plan(c.label_lengths, c.input_lengths, V, B, c.max_time, want_grad,
                               (int)((c.flags >> 8) & 0x7), plan);
    if (st != CTC_STATUS_SUCCESS) return st;
    if (plan.total > c.workspace_bytes) return fail(CTC_STATUS_INVALID_VALUE, "workspace too small");

    cudaStream_t stream = (cudaStream_t)c.stream;
    char *ws = (char *)c.workspace;
    int *d_meta = (int *)(ws + plan.off_meta);
    const int *d_labels = g_labels_dev_override ? g_labels_dev_override : (const int *)(ws + plan.off_labels);
    float *d_costs = c.costs_device ? c.costs_device : (float *)(ws + plan.off_costs);
    int *d_status = g_status_dev_override ? g_status_dev_override
                    : (c.status_device ? c.status_device : (int *)(ws + plan.off_status));

    if (!check(cudaMemcpyAsync(d_meta, plan.meta.data(), sizeof(int) * 4 * (size_t)B, cudaMemcpyHostToDevice, stream),
               "H2D metadata", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    if (plan.total_labels > 0 && !g_labels_dev_override &&
        !check(cudaMemcpyAsync((int *)(ws + plan.off_labels), c.flat_labels, sizeof(int) * (size_t)plan.total_labels, cudaMemcpyHostToDevice, stream),
               "H2D labels", CTC_STATUS_MEMOPS_FAILED, st)) return st;

    FusedParams P;
    P.acts = c.activations; P.act_stride_t = c.act_stride_t; P.act_stride_b = c.act_stride_b;
    P.grads = c.gradients;
    P.labels = d_labels; P.label_off = d_meta; P.label_len = d_meta + B; P.act_len = d_meta + 2 * B;
    P.costs = d_costs; P.status = d_status;
    P.V = V; P.T_max = c.max_time; P.B = B; P.blank = c.blank_label;
    P.grad_scale = c.grad_scale;
    P.debug = c.debug_device;
    P.queue = nullptr; P.n_items = 0; P.only_flagged = 0; P.sm_lo = 0; P.sm_hi = 0; P.n_slots = 0;
    int *d_queue = (int *)(ws + plan.off_queue);
    if ((int)plan.launches.size() > kMaxLaunches / 2) return fail(CTC_STATUS_UNKNOWN_ERROR, "too many kernel variants in one call");
    if (!check(cudaMemsetAsync(d_queue, 0, sizeof(int) * 4 * (kMaxLaunches + 1), stream), "queue memset", CTC_STATUS_MEMOPS_FAILED, st))
        return st;

    const bool serial = (c.flags & CTC_B200_FLAG_SERIAL_LAUNCHES) != 0;
    AuxStreams *tim = (c.kernel_ms_host && !no_sync) ? aux_streams() : nullptr;
    if (tim && !check(cudaEventRecord(tim->t0, stream), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    AuxStreams *aux = (plan.launches.size() > 1 && !serial) ? aux_streams() : nullptr;
    if (aux && !check(cudaEventRecord(aux->fork, stream), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    const bool bidir = plan.bidir && want_grad && !(c.flags & CTC_B200_FLAG_NO_BIDIR);
    P.sweep_only = 0; P.n_fwd = 0; P.col = nullptr; P.col_stride = 0; P.col_exp = nullptr; P.col_exp_stride = 0;
    P.col_z = nullptr;
    int n_aux_used = 0, li = 0;
    for (const Plan::Launch &l : plan.launches) {
        P.utt_ids = d_meta + 3 * B + l.first;
        P.ckpt = (double *)(ws + plan.off_ckpt + l.ckpt_off);
        P.ckpt_stride = l.ckpt_stride;
        cudaStream_t ls = stream;
        if (aux && li > 0) {
            const int j = (li - 1) % kAuxStreams;
            ls = aux->s[j];
            if (li - 1 < kAuxStreams) {
                if (!check(cudaStreamWaitEvent(ls, aux->fork, 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
                n_aux_used = li;
            }
        }
        if (!ensure_smem_attr((const void *)l.v->kernel, l.smem, st)) return st;
        if (bidir) {
            // small batch: the T-serial chain is the bound.  Two concurrent forward sweeps per utterance (the problem
            // and its time/label reversal) spill their columns; ctc_combine_kernel forms the gradients in parallel.
            char *area = ws + plan.off_ckpt;
            P.sweep_only = 1; P.n_fwd = l.count;
            P.col = (unsigned *)(area + l.col_off); P.col_stride = (long long)c.max_time * l.v->sp();
            P.col_exp = (int *)(area + l.exp_off); P.col_exp_stride = l.exp_stride;
            P.col_z = (double *)(area + l.z_off);
            l.v->kernel<<<2 * l.count, 32 * l.v->W, l.smem, ls>>>(P);
            ++g_launches;
            if (!check(cudaGetLastError(), "kernel launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
            CombineParams C;
            C.acts = P.acts; C.act_stride_t = P.act_stride_t; C.act_stride_b = P.act_stride_b; C.grads = P.grads;
            C.labels = P.labels; C.label_off = P.label_off; C.label_len = P.label_len; C.act_len = P.act_len;
            C.utt_ids = P.utt_ids; C.status = P.status;
            C.col = P.col; C.col_stride = P.col_stride; C.col_exp = P.col_exp; C.col_exp_stride = P.col_exp_stride;
            C.col_z = P.col_z; C.n = l.count;
            C.V = V; C.T_max = c.max_time; C.B = B; C.blank = c.blank_label; C.grad_scale = c.grad_scale;
            C.frames_per_cta = 16;
            const int csm = combine_smem_bytes(l.v->sp(), V);
            if (!ensure_smem_attr((const void *)l.v->combine, csm, st)) return st;
            dim3 grid((c.max_time + C.frames_per_cta - 1) / C.frames_per_cta, l.count);
            l.v->combine<<<grid, kCombineThreads, csm, ls>>>(C);
        } else if (l.v->warp) {
            P.n_items = l.count;
            int smem_launch = l.smem;
            if (const char *pad = std::getenv("CTC_B200_WARP_PAD_KB")) {          // occupancy experiments only (tools/occupancy_probe.py)
                smem_launch = std::min(kMaxSmem, l.smem + 1024 * std::atoi(pad));
                if (!ensure_smem_attr((const void *)l.v->kernel, smem_launch, st)) return st;
            }
            // with an SM range the grid is full-size (CTAs outside the range retire at once, the others claim a slot)
            const int full = persistent_grid((const void *)l.v->kernel, smem_launch);
            // (CTC_B200_FLAG_SERIAL_LAUNCHES runs the buckets one after the other: no ranges then, each bucket spreads its
            //  planned number of workers over the whole chip)
            const bool ranged = (l.sm_hi > l.sm_lo) && !serial;
            const int grid = ranged ? full : std::min(l.slots, full);
            P.sm_lo = ranged ? l.sm_lo : 0; P.sm_hi = ranged ? l.sm_hi : 0; P.n_slots = l.slots;
            if (ranged && std::getenv("CTC_B200_TEST_EMPTY_RANGES")) {
                P.sm_lo += 100000; P.sm_hi += 100000;      // test hook: no CTA lands in its range, so the last CTA of every grid
            }                                              // has to drain the whole bucket alone (tests/test_gpu_parity.py)
            P.queue = d_queue + 4 * li;
            l.v->kernel<<<grid, 32, smem_launch, ls>>>(P);
            if (l.v2 && !(c.flags & CTC_B200_FLAG_NO_FALLBACK)) {
                // fp32 kernel: utterances that failed its range self-check are redone by the fp64 kernel of the same
                // label class, which scans the status words of the bucket (nothing flagged: a few microseconds)
                if (!check(cudaGetLastError(), "kernel launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
                if (!ensure_smem_attr((const void *)l.v2->kernel, l.smem2, st)) return st;
                P.queue = d_queue + 4 * (kMaxLaunches / 2 + li); P.only_flagged = 1;
                const int full2 = persistent_grid((const void *)l.v2->kernel, l.smem2);
                const int grid2 = ranged ? full2 : std::min(l.slots, full2);
                l.v2->kernel<<<grid2, 32, l.smem2, ls>>>(P);
                P.only_flagged = 0;
                ++g_launches;
            }
            P.queue = nullptr; P.n_items = 0; P.sm_lo = 0; P.sm_hi = 0;
        } else {
            l.v->kernel<<<l.count, 32 * l.v->W, l.smem, ls>>>(P);
        }
        ++g_launches;
        ++li;
        if (!check(cudaGetLastError(), "kernel launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    }
    for (int j = 0; j < n_aux_used && j < kAuxStreams; ++j) {
        if (!check(cudaEventRecord(aux->join[j], aux->s[j]), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        if (!check(cudaStreamWaitEvent(stream, aux->join[j], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    }
    // Out-of-range utterances (Z^ underflowed to 0, the forward/backward consistency check failed, or a genuine +inf
    // cost) are redone in log space by a kernel that finds them itself: no host round trip, so NO_SYNC calls get
    // the detour too.
    if (!(c.flags & CTC_B200_FLAG_NO_FALLBACK)) {
        st = launch_logspace_detour(c, plan, P, d_queue + 4 * kMaxLaunches, stream);
        if (st != CTC_STATUS_SUCCESS) return st;
    }
    if (tim && !check(cudaEventRecord(tim->t1, stream), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    if (no_sync) return CTC_STATUS_SUCCESS;

    // costs and status come back through a per-thread PINNED staging buffer: a D2H copy into pageable memory makes the
    // host wait for that copy, so the two copies + the stream sync were three round trips per blocking call (B = 32: 0.178 ms
    // against 0.140 ms non-blocking); into pinned memory both copies are queued and the host waits once
    thread_local std::vector<int> h_status;
    h_status.resize(B);
    char *pin = pinned_scratch(8 * (size_t)B);
    float *p_costs = pin ? (float *)pin : c.costs_host;
    int *p_status = pin ? (int *)(pin + 4 * (size_t)B) : h_status.data();
    if (c.costs_host &&
        !check(cudaMemcpyAsync(p_costs, d_costs, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, stream),
               "D2H costs", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    if (!check(cudaMemcpyAsync(p_status, d_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, stream),
               "D2H status", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    if (!check(cudaStreamSynchronize(stream), "stream sync", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    if (pin) {
        if (c.costs_host) std::memcpy(c.costs_host, p_costs, sizeof(float) * (size_t)B);
        std::memcpy(h_status.data(), p_status, sizeof(int) * (size_t)B);
    }
    if (tim) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, tim->t0, tim->t1) == cudaSuccess) *c.kernel_ms_host = ms;
    }
    int any = 0;
    for (int b = 0; b < B; ++b) any |= h_status[b];
    if (c.status_host) std::memcpy(c.status_host, h_status.data(), sizeof(int) * (size_t)B);
    // Invalid ARGUMENTS are errors; hostile DATA is not: NaN activations (still CTC_B200_UTT_RANGE after the detour)
    // come back as a NaN cost and NaN gradient rows with the status bit set -- what upstream does, and what lets the
    // reference's trainer skip the step instead of dying (codes/engine.py:27-30).
    if (any & CTC_B200_UTT_BAD_LABEL)
        return fail(CTC_STATUS_INVALID_VALUE, "a label is outside [0, alphabet_size) or equals the blank");
    return CTC_STATUS_SUCCESS;
}


// ---- host-buffer entry point: chunked H2D -> kernels -> D2H pipeline -----------------------------
struct HostPlan {
    int n_chunks = 1, n_buf = 1, Bc = 0;
    size_t inner_ws = 0, acts_bytes = 0, off_costs = 0, off_status = 0, off_labels = 0, off_sets = 0, set_bytes = 0, total = 0;
    long long total_labels = 0;
};

ctcStatus_t make_host_plan(const int *label_lengths, const int *input_lengths, int V, int B, int T, bool want_grad,
                           int n_chunks, HostPlan &hp)
{
    // automatic: slices of 256 utterances (measured optimum on B200, profiles/r2_e2e_sweep.txt: finer slices shorten the
    // fill and drain of the pipeline, coarser ones amortise the per-slice host work of the inner call)
    if (n_chunks <= 0) n_chunks = std::max(1, std::min(64, B / 256));
    n_chunks = std::max(1, std::min(n_chunks, B));
    hp.n_chunks = n_chunks;
    hp.Bc = (B + n_chunks - 1) / n_chunks;
    if (hp.Bc >= 32) hp.Bc = (hp.Bc + 31) / 32 * 32;     // rows of a slice (Bc * V floats) in whole 128-byte lines: the 2-D copies
                                                         // of slices with ragged rows were measured 15 % slower
    hp.n_chunks = (B + hp.Bc - 1) / hp.Bc;
    hp.n_buf = std::min(hp.n_chunks, kPipeBuffers);
    for (int c = 0; c < hp.n_chunks; ++c) {
        const int lo = c * hp.Bc, n = std::min(hp.Bc, B - lo);
        size_t need = 0;
        ctcStatus_t st = ctc_b200_workspace_size(label_lengths + lo, input_lengths + lo, V, n, T, want_grad, &need);
        if (st != CTC_STATUS_SUCCESS) return st;
        hp.inner_ws = std::max(hp.inner_ws, need);
    }
    hp.acts_bytes = align_up(sizeof(float) * (size_t)T * hp.Bc * V, 256);
    size_t o = 0;
    hp.off_costs = o;  o = align_up(o + sizeof(float) * (size_t)B, 256);
    hp.off_status = o; o = align_up(o + sizeof(int) * (size_t)B, 256);
    hp.total_labels = 0;
    for (int b = 0; b < B; ++b) hp.total_labels += std::max(label_lengths[b], 0);
    hp.off_labels = o; o = align_up(o + sizeof(int) * (size_t)std::max<long long>(hp.total_labels, 1), 256);
    hp.off_sets = o;
    hp.set_bytes = hp.acts_bytes * (want_grad ? 2 : 1) + align_up(hp.inner_ws, 256);
    hp.total = o + hp.set_bytes * hp.n_buf + 256;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t run_host(const ctcB200HostCall &c)
{
    if (!c.activations || !c.flat_labels || !c.label_lengths || !c.input_lengths || !c.workspace || !c.costs_host)
        return fail(CTC_STATUS_INVALID_VALUE, "null pointer argument");
    if (c.alphabet_size <= 0 || c.minibatch <= 0 || c.max_time <= 0)
        return fail(CTC_STATUS_INVALID_VALUE, "non-positive size");
    const int B = c.minibatch, V = c.alphabet_size, T = c.max_time;
    const bool want_grad = c.gradients != nullptr;
    HostPlan hp;
    ctcStatus_t st = make_host_plan(c.label_lengths, c.input_lengths, V, B, T, want_grad, c.n_chunks, hp);
    if (st != CTC_STATUS_SUCCESS) return st;
    if (hp.total > c.workspace_bytes) return fail(CTC_STATUS_INVALID_VALUE, "workspace too small");
    AuxStreams *aux = aux_streams();
    if (!aux) return fail(CTC_STATUS_EXECUTION_FAILED, "could not create internal streams");
    cudaStream_t stream = (cudaStream_t)c.stream;
    char *ws = (char *)c.workspace;
    float *d_costs = (float *)(ws + hp.off_costs);
    int *d_status = (int *)(ws + hp.off_status);
    std::vector<long long> lab_off(hp.n_chunks + 1, 0);
    for (int ch = 0; ch < hp.n_chunks; ++ch) {
        long long s = 0;
        const int lo = ch * hp.Bc, n = std::min(hp.Bc, B - lo);
        for (int b = 0; b < n; ++b) s += c.label_lengths[lo + b];
        lab_off[ch + 1] = lab_off[ch] + s;
    }
    // Three stage streams -- copy-in, kernels, copy-out -- chained per chunk by events, over n_buf buffer sets: the
    // H2D engine runs back to back as long as a set is free, whatever the other two stages are doing.  (Round 1 put
    // the three stages of a chunk on ONE stream per set: the next upload into a set waited for the previous
    // download out of it, and the pipeline reached ~60 % of the duplex link rate.)
    cudaStream_t s_in = aux->pipe[0], s_k = aux->pipe[1], s_out = aux->pipe[2];
    if (!aux->chunk_events((size_t)hp.n_chunks)) return fail(CTC_STATUS_EXECUTION_FAILED, "could not create pipeline events");
    if (!check(cudaEventRecord(aux->pipe_fork, stream), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    for (int i = 0; i < kPipeStreams; ++i)
        if (!check(cudaStreamWaitEvent(aux->pipe[i], aux->pipe_fork, 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    // The labels of the whole batch go up ONCE, first thing on the copy-in stream.  A per-slice upload on the kernel
    // stream (what the inner call does on its own) sits in the H2D engine's queue waiting for that stream's events --
    // the download of an earlier slice -- and the activation uploads of the next slices queue behind it: with pinned
    // label arrays that head-of-line blocking cost 6 ms of a 25 ms step (pageable label arrays are copied inline and
    // never showed it).
    int *d_labels_all = (int *)(ws + hp.off_labels);
    if (hp.total_labels > 0 &&
        !check(cudaMemcpyAsync(d_labels_all, c.flat_labels, sizeof(int) * (size_t)hp.total_labels, cudaMemcpyHostToDevice, s_in),
               "H2D labels", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    const size_t row_all = sizeof(float) * (size_t)B * V;
    for (int ch = 0; ch < hp.n_chunks; ++ch) {
        const int lo = ch * hp.Bc, n = std::min(hp.Bc, B - lo);
        const int set = ch % hp.n_buf;
        char *base = ws + hp.off_sets + hp.set_bytes * set;
        float *d_acts = (float *)base;
        float *d_grads = want_grad ? (float *)(base + hp.acts_bytes) : nullptr;
        void *inner = base + hp.acts_bytes * (want_grad ? 2 : 1);
        const size_t row_c = sizeof(float) * (size_t)n * V;
        // copy-in: the set's activation buffer is free once the kernels of its previous tenant are done
        if (ch >= hp.n_buf &&
            !check(cudaStreamWaitEvent(s_in, aux->ev_k[ch - hp.n_buf], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        if (!check(cudaMemcpy2DAsync(d_acts, row_c, c.activations + (size_t)lo * V, row_all, row_c, T,
                                     cudaMemcpyHostToDevice, s_in), "H2D activations", CTC_STATUS_MEMOPS_FAILED, st)) return st;
        if (!check(cudaEventRecord(aux->ev_in[ch], s_in), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        // kernels: need the upload, and the set's gradient buffer drained by the previous tenant's download
        if (!check(cudaStreamWaitEvent(s_k, aux->ev_in[ch], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        if (ch >= hp.n_buf && want_grad &&
            !check(cudaStreamWaitEvent(s_k, aux->ev_out[ch - hp.n_buf], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        ctcB200Call k;
        std::memset(&k, 0, sizeof(k));
        k.activations = d_acts; k.act_stride_t = (long long)n * V; k.act_stride_b = V;
        k.gradients = d_grads;
        k.flat_labels = c.flat_labels + lab_off[ch];
        k.label_lengths = c.label_lengths + lo;
        k.input_lengths = c.input_lengths + lo;
        k.alphabet_size = V; k.minibatch = n; k.max_time = T; k.blank_label = c.blank_label;
        k.grad_scale = c.grad_scale;
        k.costs_device = d_costs + lo;
        k.workspace = inner; k.workspace_bytes = hp.inner_ws;
        k.stream = (CUstream)s_k;
        k.flags = (c.flags & 0x700u) | CTC_B200_FLAG_NO_SYNC;
        g_status_dev_override = d_status + lo;
        g_labels_dev_override = d_labels_all + lab_off[ch];
        st = run(k);
        g_status_dev_override = nullptr;
        g_labels_dev_override = nullptr;
        if (st != CTC_STATUS_SUCCESS) return st;
        if (!check(cudaEventRecord(aux->ev_k[ch], s_k), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        // copy-out
        if (want_grad) {
            if (!check(cudaStreamWaitEvent(s_out, aux->ev_k[ch], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
            if (!check(cudaMemcpy2DAsync(c.gradients + (size_t)lo * V, row_all, d_grads, row_c, row_c, T,
                                         cudaMemcpyDeviceToHost, s_out), "D2H gradients", CTC_STATUS_MEMOPS_FAILED, st)) return st;
            if (!check(cudaEventRecord(aux->ev_out[ch], s_out), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        }
    }
    for (int i = 0; i < kPipeStreams; ++i) {
        if (!check(cudaEventRecord(aux->pipe_join[i], aux->pipe[i]), "event record", CTC_STATUS_EXECUTION_FAILED, st)) return st;
        if (!check(cudaStreamWaitEvent(stream, aux->pipe_join[i], 0), "stream wait", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    }
    std::vector<int> h_status(B);
    if (!check(cudaMemcpyAsync(c.costs_host, d_costs, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, stream),
               "D2H costs", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    if (!check(cudaMemcpyAsync(h_status.data(), d_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, stream),
               "D2H status", CTC_STATUS_MEMOPS_FAILED, st)) return st;
    if (!check(cudaStreamSynchronize(stream), "stream sync", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    // (out-of-range utterances were redone on the device by the detour kernel of their chunk's call)
    int any = 0;
    for (int b = 0; b < B; ++b) any |= h_status[b];
    if (c.status_host) std::memcpy(c.status_host, h_status.data(), sizeof(int) * (size_t)B);
    if (any & CTC_B200_UTT_BAD_LABEL)
        return fail(CTC_STATUS_INVALID_VALUE, "a label is outside [0, alphabet_size) or equals the blank");
    return CTC_STATUS_SUCCESS;
}

// ---- loss glue (include/ctc.h: ctc_b200_reduce_costs / ctc_b200_scale_gradients) ------------------
constexpr int kReduceThreads = 1024;
__global__ void __launch_bounds__(kReduceThreads) reduce_costs_kernel(const float *costs, int n, float scale, int zero_inf,
                                                                       float *loss, int *flag)
{
    __shared__ double part[kReduceThreads / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kReduceThreads) s += (double)costs[i];       // fixed assignment and order
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kReduceThreads / 32; ++w) t += part[w];
        t *= (double)scale;
        const bool inf = (t == INFINITY) || (t == -INFINITY);
        loss[0] = (inf && zero_inf) ? 0.f : (float)t;
        if (flag) flag[0] = (inf && zero_inf) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) scale_gradients_kernel(float4 *g4, float *g, size_t n4, size_t n, float scale_host,
                                                               const float *scale_dev, const int *zero_flag)
{
    float f = scale_host * (scale_dev ? __ldg(scale_dev) : 1.f);
    if (zero_flag && __ldg(zero_flag)) f = 0.f;
    if (f == 1.f) return;                                   // the common case: no pass over the tensor
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (f == 0.f) {                                         // (0 * inf and 0 * NaN must be 0 here: the step is being skipped)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] = 0.f;
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = g4[i];
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        g4[i] = v;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= f;
}

}  // namespace

namespace ctcb200 {
void ctcb200_set_error(const std::string &msg) { g_last_error = msg; }
void ctcb200_count_launch() { ++g_launches; }
}  // namespace ctcb200

extern "C" {

int get_warpctc_version(void) { return 2; }

const char *ctcGetStatusString(ctcStatus_t status)
{
    switch (status) {
    case CTC_STATUS_SUCCESS: return "no error";
    case CTC_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
    case CTC_STATUS_INVALID_VALUE: return "invalid value";
    case CTC_STATUS_EXECUTION_FAILED: return "execution failed";
    case CTC_STATUS_UNKNOWN_ERROR:
    default: return "unknown error";
    }
}

const char *ctc_b200_last_error(void) { return g_last_error.c_str(); }

int ctc_b200_info(unsigned long long *launch_count)
{
    if (launch_count) *launch_count = g_launches;
    return 100;
}

ctcStatus_t ctc_b200_workspace_size(const int *label_lengths, const int *input_lengths, int alphabet_size,
                                    int minibatch, int max_time, int want_gradients, size_t *size_bytes)
{
    if (!size_bytes) return fail(CTC_STATUS_INVALID_VALUE, "null size_bytes");
    // take the max over the ladders so that a forced mode never overruns the workspace
    size_t need = 0;
    for (int mode = 1; mode <= 5; ++mode) {
        thread_local Plan p;
        ctcStatus_t st = make_plan(label_lengths, input_lengths, alphabet_size, minibatch, max_time,
                                   want_gradients != 0, mode, p, /*size_only=*/true);
        if (st != CTC_STATUS_SUCCESS) return st;
        need = std::max(need, p.total);
    }
    *size_bytes = need;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_compute(const ctcB200Call *call)
{
    if (!call) return fail(CTC_STATUS_INVALID_VALUE, "null call");
    return run(*call);
}

ctcStatus_t ctc_b200_reduce_costs(const float *costs_device, int minibatch, float scale, int zero_infinite,
                                  float *loss_device, int *flag_device, CUstream stream)
{
    if (!costs_device || !loss_device || minibatch <= 0) return fail(CTC_STATUS_INVALID_VALUE, "invalid argument");
    reduce_costs_kernel<<<1, kReduceThreads, 0, (cudaStream_t)stream>>>(costs_device, minibatch, scale, zero_infinite,
                                                                        loss_device, flag_device);
    ++g_launches;
    ctcStatus_t st = CTC_STATUS_SUCCESS;
    if (!check(cudaGetLastError(), "reduce_costs launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_scale_gradients(float *gradients, size_t count, float scale_host, const float *scale_device,
                                     const int *zero_flag_device, CUstream stream)
{
    if (!gradients) return fail(CTC_STATUS_INVALID_VALUE, "null gradients");
    if (count == 0) return CTC_STATUS_SUCCESS;
    if (!scale_device && !zero_flag_device && scale_host == 1.f) return CTC_STATUS_SUCCESS;   // nothing to do, no launch
    const bool aligned = (reinterpret_cast<uintptr_t>(gradients) & 15u) == 0;
    const size_t n4 = aligned ? count / 4 : 0;
    const int grid = (int)std::min<size_t>((std::max<size_t>(n4, 1) + 255) / 256, 148 * 8);
    scale_gradients_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(gradients), gradients, n4,
                                                                   count, scale_host, scale_device, zero_flag_device);
    ++g_launches;
    ctcStatus_t st = CTC_STATUS_SUCCESS;
    if (!check(cudaGetLastError(), "scale_gradients launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_workspace_size_host(const int *label_lengths, const int *input_lengths, int alphabet_size,
                                         int minibatch, int max_time, int want_gradients, int n_chunks,
                                         size_t *size_bytes)
{
    if (!size_bytes || !label_lengths || !input_lengths || alphabet_size <= 0 || minibatch <= 0 || max_time <= 0)
        return fail(CTC_STATUS_INVALID_VALUE, "invalid argument");
    HostPlan hp;
    ctcStatus_t st = make_host_plan(label_lengths, input_lengths, alphabet_size, minibatch, max_time,
                                    want_gradients != 0, n_chunks, hp);
    if (st != CTC_STATUS_SUCCESS) return st;
    *size_bytes = hp.total;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_compute_host(const ctcB200HostCall *call)
{
    if (!call) return fail(CTC_STATUS_INVALID_VALUE, "null call");
    return run_host(*call);
}

ctcStatus_t ctc_b200_greedy_decode(const float *probs, long long stride_b, long long stride_t, const int *sizes_device,
                                   int minibatch, int max_time, int alphabet_size, int blank_label, int *tokens_device,
                                   int *offsets_device, int *counts_device, CUstream stream)
{
    if (!probs || !tokens_device || !counts_device) return fail(CTC_STATUS_INVALID_VALUE, "null pointer argument");
    if (minibatch <= 0 || max_time <= 0 || alphabet_size <= 0) return fail(CTC_STATUS_INVALID_VALUE, "non-positive size");
    if (alphabet_size > 128) return fail(CTC_STATUS_UNKNOWN_ERROR, "alphabet_size above 128 is not supported by this build");
    DecodeParams D;
    D.probs = probs; D.stride_b = stride_b; D.stride_t = stride_t; D.sizes = sizes_device;
    D.tokens = tokens_device; D.offsets = offsets_device; D.counts = counts_device;
    D.B = minibatch; D.T = max_time; D.V = alphabet_size; D.blank = blank_label;
    const int smem = 32 * (alphabet_size | 1) * (int)sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (alphabet_size <= 32) ctc_greedy_decode_kernel<1><<<minibatch, 32, smem, s>>>(D);
    else if (alphabet_size <= 64) ctc_greedy_decode_kernel<2><<<minibatch, 32, smem, s>>>(D);
    else ctc_greedy_decode_kernel<4><<<minibatch, 32, smem, s>>>(D);
    ++g_launches;
    ctcStatus_t st = CTC_STATUS_SUCCESS;
    if (!check(cudaGetLastError(), "decode launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t ctc_b200_edit_distance(const int *hyp_tokens_device, long long hyp_stride, const int *hyp_counts_device,
                                   int max_hyp, const int *refs_device, const int *ref_offsets_device,
                                   const int *ref_lengths_device, int max_ref, int minibatch, int space_label, int mode,
                                   int *distances_device, int *normalisers_device, CUstream stream)
{
    if (!hyp_tokens_device || !hyp_counts_device || !refs_device || !ref_offsets_device || !ref_lengths_device ||
        !distances_device)
        return fail(CTC_STATUS_INVALID_VALUE, "null pointer argument");
    if (minibatch <= 0 || max_hyp < 0 || max_ref < 0) return fail(CTC_STATUS_INVALID_VALUE, "non-positive size");
    if (mode != EDIT_TOKENS && mode != EDIT_CER && mode != EDIT_WER)
        return fail(CTC_STATUS_INVALID_VALUE, "mode must be 0 (tokens), 1 (CER) or 2 (WER)");
    if (max_ref > 2047) return fail(CTC_STATUS_UNKNOWN_ERROR, "references above 2047 tokens are not supported by this build");
    EditParams E;
    E.hyp = hyp_tokens_device; E.hyp_stride = hyp_stride; E.hyp_len = hyp_counts_device;
    E.ref = refs_device; E.ref_off = ref_offsets_device; E.ref_len = ref_lengths_device;
    E.dist = distances_device; E.norm = normalisers_device;
    E.B = minibatch; E.space = space_label; E.mode = mode; E.max_hyp = max_hyp; E.max_ref = max_ref;
    const int smem = editdist_smem_bytes(max_hyp, max_ref, mode);
    if (smem > 200 * 1024) return fail(CTC_STATUS_UNKNOWN_ERROR, "hypothesis rows too long for the shared-memory staging");
    cudaStream_t s = (cudaStream_t)stream;
    ctcStatus_t st = CTC_STATUS_SUCCESS;
    const int cols = max_ref + 1;                          // DP columns 0..max_ref, C per lane
#define CTC_EDIT_LAUNCH(C)                                                                                           \
    do {                                                                                                             \
        if (smem > 48 * 1024 &&                                                                                      \
            !check(cudaFuncSetAttribute(ctc_edit_distance_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        smem), "edit distance smem attribute", CTC_STATUS_EXECUTION_FAILED, st))     \
            return st;                                                                                               \
        ctc_edit_distance_kernel<C><<<minibatch, 32, smem, s>>>(E);                                                  \
    } while (0)
    if (cols <= 32 * 4) CTC_EDIT_LAUNCH(4);
    else if (cols <= 32 * 8) CTC_EDIT_LAUNCH(8);
    else if (cols <= 32 * 16) CTC_EDIT_LAUNCH(16);
    else if (cols <= 32 * 32) CTC_EDIT_LAUNCH(32);
    else CTC_EDIT_LAUNCH(64);
#undef CTC_EDIT_LAUNCH
    ++g_launches;
    if (!check(cudaGetLastError(), "edit distance launch", CTC_STATUS_EXECUTION_FAILED, st)) return st;
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t get_workspace_size(const int *const label_lengths, const int *const input_lengths, int alphabet_size,
                               int minibatch, struct ctcOptions info, size_t *size_bytes)
{
    if (!label_lengths || !input_lengths || !size_bytes || alphabet_size <= 0 || minibatch <= 0)
        return fail(CTC_STATUS_INVALID_VALUE, "invalid argument");
    if (info.loc != CTC_GPU)
        return fail(CTC_STATUS_INVALID_VALUE, "libctc_b200 is GPU-only: options.loc must be CTC_GPU");
    int max_t = 0;
    for (int b = 0; b < minibatch; ++b) max_t = std::max(max_t, input_lengths[b]);
    return ctc_b200_workspace_size(label_lengths, input_lengths, alphabet_size, minibatch, std::max(max_t, 1), 1,
                                   size_bytes);
}

ctcStatus_t compute_ctc_loss(const float *const activations, float *gradients, const int *const flat_labels,
                             const int *const label_lengths, const int *const input_lengths, int alphabet_size,
                             int minibatch, float *costs, void *workspace, struct ctcOptions options)
{
    if (!activations || !flat_labels || !label_lengths || !input_lengths || !costs || !workspace ||
        alphabet_size <= 0 || minibatch <= 0)
        return fail(CTC_STATUS_INVALID_VALUE, "invalid argument");
    if (options.loc != CTC_GPU)
        return fail(CTC_STATUS_INVALID_VALUE, "libctc_b200 is GPU-only: options.loc must be CTC_GPU");
    int max_t = 0;
    for (int b = 0; b < minibatch; ++b) max_t = std::max(max_t, input_lengths[b]);
    if (max_t <= 0) {                                   // nothing to align against: all costs 0 (infeasible)
        for (int b = 0; b < minibatch; ++b) costs[b] = 0.f;
        return CTC_STATUS_SUCCESS;
    }
    size_t need = 0;
    ctcStatus_t st = ctc_b200_workspace_size(label_lengths, input_lengths, alphabet_size, minibatch, max_t,
                                             gradients != nullptr, &need);
    if (st != CTC_STATUS_SUCCESS) return st;
    ctcB200Call c;
    std::memset(&c, 0, sizeof(c));
    c.activations = activations;
    c.act_stride_t = (long long)minibatch * alphabet_size;
    c.act_stride_b = alphabet_size;
    c.gradients = gradients;
    c.flat_labels = flat_labels;
    c.label_lengths = label_lengths;
    c.input_lengths = input_lengths;
    c.alphabet_size = alphabet_size;
    c.minibatch = minibatch;
    c.max_time = max_t;
    c.blank_label = options.blank_label;
    c.grad_scale = 1.0f;
    c.costs_host = costs;
    c.workspace = workspace;
    c.workspace_bytes = need;                           // caller promised get_workspace_size() bytes
    c.stream = options.stream;
    return run(c);
}

}  // extern "C"
