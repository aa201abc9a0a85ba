// ctc_variants.h -- table of compiled instantiations of ctc_fused_kernel<NS, W, K, VCH>.
// Each translation unit ctc_variants.cu -DCTC_GROUP=g contributes one group; the groups are compiled in
// parallel by aes_lac_2018_b200/build.py and linked into libctc_b200.so.
#pragma once
#include "ctc_combine.cuh"
#include "ctc_fused.cuh"
#include "ctc_warp.cuh"
#include "ctc_warp32.cuh"

namespace ctcb200 {

struct Variant {
    int NS, W, K, VCH;
    void (*kernel)(const FusedParams);
    void (*combine)(const CombineParams);      // latency ladder only: second half of the bidirectional path
    int warp;                                  // 1, 2: ctc_warp_kernel (persistent, one warp per utterance; ctc_warp.cuh);
                                               // 2 = recomputed alpha staged in shared memory instead of registers;
                                               // 3: ctc_warp32_kernel (fp32 recursion with per-lane exponents; ctc_warp32.cuh)
    int max_label() const { return (32 * NS * W) / 2 - 1; }   // SP = 32*NS*W states must hold 2L+2
    int smem_bytes(int V, int T_max) const
    {
        return warp >= 3 ? make_warp32_layout(NS, K, VCH).total
               : warp ? make_warp_layout(NS, K, VCH, warp == 2).total : make_layout(NS, W, K, V, T_max).total;
    }
    long long slot_words(int T_max) const      // warp ladders: 4-byte words of workspace per resident CTA
    {
        return warp >= 3 ? warp32_slot_words(NS, K, VCH, T_max) : warp_slot_words(NS, K, VCH, T_max);
    }
    int sp() const { return 32 * NS * W; }
};

enum Ladder { LADDER_THROUGHPUT = 0, LADDER_THROUGHPUT_K8 = 1, LADDER_LATENCY = 2, LADDER_WARP = 3, LADDER_WARP32 = 4, NUM_LADDERS = 5 };
constexpr int kMaxVch = 2;                     // alphabets up to 64 symbols (reference: 29 and 43)

// group id = ladder * kMaxVch + (VCH - 1)
const Variant *ctc_variants_group0(int *n);
const Variant *ctc_variants_group1(int *n);
const Variant *ctc_variants_group2(int *n);
const Variant *ctc_variants_group3(int *n);
const Variant *ctc_variants_group4(int *n);
const Variant *ctc_variants_group5(int *n);
const Variant *ctc_variants_group6(int *n);
const Variant *ctc_variants_group7(int *n);
const Variant *ctc_variants_group8(int *n);
const Variant *ctc_variants_group9(int *n);

}  // namespace ctcb200
