// ctc_variants.cu -- instantiates one group of kernel variants (compile with -DCTC_GROUP=0..9).
#include "ctc_variants.h"

#ifndef CTC_GROUP
#error "compile with -DCTC_GROUP=<0..9>"
#endif

namespace ctcb200 {

#define CTC_LADDER (CTC_GROUP / 2)
#define CTC_VCH (CTC_GROUP % 2 + 1)
#if CTC_GROUP / 2 == 2
#define V_(NS, W, K) Variant{NS, W, K, CTC_VCH, ctc_fused_kernel<NS, W, K, CTC_VCH>, ctc_combine_kernel<NS, W, K>, 0}
#elif CTC_GROUP / 2 == 3
#define VW_(NS, K, MAXR) Variant{NS, 1, K, CTC_VCH, ctc_warp_kernel<NS, K, CTC_VCH, MAXR, 0>, nullptr, 1}
#define VWS_(NS, K, MAXR) Variant{NS, 1, K, CTC_VCH, ctc_warp_kernel<NS, K, CTC_VCH, MAXR, 1>, nullptr, 2}
#elif CTC_GROUP / 2 == 4
#define VF_(NS, K, MAXR) Variant{NS, 1, K, CTC_VCH, ctc_warp32_kernel<NS, K, CTC_VCH, MAXR>, nullptr, 3}
#else
#define V_(NS, W, K) Variant{NS, W, K, CTC_VCH, ctc_fused_kernel<NS, W, K, CTC_VCH>, nullptr, 0}
#endif

static const Variant kTable[] = {
#if CTC_LADDER == 0
    // throughput: one warp per utterance, as few states per thread as fit, 16-step chunks
    V_(2, 1, 16), V_(4, 1, 16), V_(6, 1, 16), V_(8, 1, 16), V_(10, 1, 16), V_(12, 1, 16), V_(14, 1, 16), V_(16, 1, 16),
    V_(16, 2, 16), V_(16, 4, 8), V_(16, 8, 4),
#elif CTC_LADDER == 1
    // same with 8-step chunks: half the shared memory per CTA, twice the checkpoint traffic
    V_(2, 1, 8), V_(4, 1, 8), V_(6, 1, 8), V_(8, 1, 8), V_(10, 1, 8), V_(12, 1, 8), V_(14, 1, 8), V_(16, 1, 8),
    V_(16, 2, 8), V_(16, 4, 8), V_(16, 8, 4),
#elif CTC_LADDER == 3
    // one warp per utterance, register-resident (ctc_warp.cuh): (NS, K, register cap per thread)
#ifdef CTC_WARP_TABLE_INC          // kernel experiments: the table comes from a file (tools/build_alt.sh)
#include CTC_WARP_TABLE_INC
#else
    // (measured per label-length class on B200, profiles/r2_variant_matrix.txt)
    VW_(2, 16, 128), VW_(4, 8, 128), VW_(6, 8, 168), VW_(8, 8, 168), VWS_(10, 8, 144), VWS_(12, 8, 168), VW_(14, 8, 232), VW_(16, 8, 255),
#endif
#elif CTC_LADDER == 4
    // one warp per utterance, fp32 recursion with per-lane block exponents (ctc_warp32.cuh): (NS, K, register cap)
#ifdef CTC_WARP32_TABLE_INC        // kernel experiments: the table comes from a file (tools/build_alt.sh)
#include CTC_WARP32_TABLE_INC
#else
    // register caps measured per label class (profiles/r2_w32_variants.txt); the two-slice alphabets (V = 32 .. 63, e.g. the
    // PT-BR alphabet of BASELINE configs[2]) hold twice the row registers and want more room at NS = 4, 10, 16
#if CTC_VCH == 1
    // (NS = 12 with the packed recursion: 152 registers = 13 warps per SM, 3.57 -> 3.61 M utt/s against 168; 144 spills)
    VF_(2, 16, 128), VF_(4, 8, 96), VF_(6, 8, 128), VF_(8, 8, 128), VF_(10, 8, 128), VF_(12, 8, 152), VF_(14, 8, 168), VF_(16, 8, 168),
#else
    VF_(2, 16, 128), VF_(4, 8, 128), VF_(6, 8, 128), VF_(8, 8, 128), VF_(10, 8, 168), VF_(12, 8, 168), VF_(14, 8, 168), VF_(16, 8, 224),
#endif
#endif
#else
    // latency: more warps per utterance, fewer states per thread
    V_(2, 1, 16), V_(2, 2, 16), V_(2, 4, 16), V_(4, 4, 16), V_(4, 8, 16), V_(8, 8, 8), V_(16, 8, 4),
#endif
};

#define CTC_CAT2(a, b) a##b
#define CTC_CAT(a, b) CTC_CAT2(a, b)
const Variant *CTC_CAT(ctc_variants_group, CTC_GROUP)(int *n)
{
    *n = (int)(sizeof(kTable) / sizeof(Variant));
    return kTable;
}

}  // namespace ctcb200
