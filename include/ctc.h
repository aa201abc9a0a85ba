/*
 * ctc.h -- C ABI of libctc_b200.so, the B200-native CTC loss-and-gradient engine.
 *
 * Drop-in boundary.  The reference (igormq/aes-lac-2018) reaches its CTC hot path through the
 * third-party python package `warpctc_pytorch` (imported at /root/reference/train.py:12 and
 * /root/reference/codes/metrics.py:3, called at /root/reference/codes/engine.py:22 and
 * /root/reference/codes/metrics.py:51), whose binding loads `libwarpctc.so` and calls the entry points
 * that warp-ctc's own `include/ctc.h` declares.  warp-ctc is not vendored in /root/reference
 * (cloned at unpinned HEAD by /root/reference/docker/Dockerfile:52-66), so the "reference interface"
 * each declaration below replaces is cited by its upstream name.  The first block keeps those names,
 * argument order, argument meaning, pointer residency and error behaviour, so a binary that was
 * linked against libwarpctc.so resolves the same symbols here.  The second block is the extended
 * entry point the PyTorch front-end (aes_lac_2018_b200/ctc_loss.py) uses.
 *
 * Differences from upstream, all deliberate (DESIGN.md "Boundary"):
 *   - GPU only: options.loc must be CTC_GPU.  CTC_CPU returns CTC_STATUS_INVALID_VALUE -- there is
 *     no CPU fallback in this library.
 *   - The gradient buffer does not have to be pre-zeroed: padded frames (t >= input_lengths[b]) and
 *     infeasible utterances are written as zeros by the kernels.
 *   - Infeasible utterances (L + repeats > T) get cost 0 and zero gradient on the GPU as well (upstream
 *     does this on its CPU path only; its GPU path leaves the cost slot unwritten).
 *   - Labels ARE validated: a label outside [0, alphabet_size) or equal to the blank returns
 *     CTC_STATUS_INVALID_VALUE (upstream reads out of bounds).
 *   - The blank-extended sequence may be up to 4095 states long (L <= 2047); upstream's GPU path stops
 *     at 1280 states (L <= 639) with CTC_STATUS_UNKNOWN_ERROR.  Beyond the limit this library also
 *     returns CTC_STATUS_UNKNOWN_ERROR.
 */
#ifndef CTC_B200_CTC_H
#define CTC_B200_CTC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Forward declaration of the CUDA stream handle so that this header needs no CUDA include.
 * (upstream: `typedef struct CUstream_st* CUstream;` in ctc.h) */
typedef struct CUstream_st *CUstream;

/* upstream ctc.h: ctcStatus_t -- same enumerators, same values. */
typedef enum {
    CTC_STATUS_SUCCESS = 0,
    CTC_STATUS_MEMOPS_FAILED = 1,
    CTC_STATUS_INVALID_VALUE = 2,
    CTC_STATUS_EXECUTION_FAILED = 3,
    CTC_STATUS_UNKNOWN_ERROR = 4
} ctcStatus_t;

/* upstream ctc.h: ctcComputeLocation. */
typedef enum {
    CTC_CPU = 0,
    CTC_GPU = 1
} ctcComputeLocation;

/* upstream ctc.h: struct ctcOptions (passed BY VALUE).  Layout: int, 8-byte union, int. */
struct ctcOptions {
    ctcComputeLocation loc;      /* must be CTC_GPU here */
    union {
        unsigned int num_threads; /* CPU only upstream; ignored here */
        CUstream stream;          /* stream the work is enqueued on (0 = legacy default stream) */
    };
    int blank_label;              /* the reference always uses 0 (train.py:179: CTCLoss() defaults) */
};
#ifndef __cplusplus
typedef struct ctcOptions ctcOptions;
#endif

/* upstream ctc.h: get_warpctc_version().  Returns 2, the API generation this ABI mirrors. */
int get_warpctc_version(void);

/* upstream ctc.h: ctcGetStatusString(). */
const char *ctcGetStatusString(ctcStatus_t status);

/*
 * upstream ctc.h: compute_ctc_loss().  Replaces the call made by warpctc_pytorch's gpu_ctc() binding.
 *
 *   activations   DEVICE, dense time-major [max(input_lengths)][minibatch][alphabet_size], fp32,
 *                 UNNORMALISED (the softmax is internal, reference README.md:168).
 *   gradients     DEVICE, same shape, or NULL to compute the costs only.
 *   flat_labels   HOST, concatenated labels of all utterances (int).
 *   label_lengths HOST [minibatch];  input_lengths HOST [minibatch].
 *   costs         HOST [minibatch]; on return costs[b] = -log p(labels_b | activations_b).
 *   workspace     DEVICE, at least get_workspace_size() bytes; caller-owned; nothing is allocated here.
 *   options       by value; work is enqueued on options.stream and the call BLOCKS until the costs are
 *                 on the host (as upstream does).
 * Errors: null pointers / non-positive sizes / bad labels -> CTC_STATUS_INVALID_VALUE; a CUDA copy that
 * fails -> CTC_STATUS_MEMOPS_FAILED; a launch that fails -> CTC_STATUS_EXECUTION_FAILED; a label
 * sequence beyond the supported length -> CTC_STATUS_UNKNOWN_ERROR.  No exceptions, no errno.
 */
ctcStatus_t compute_ctc_loss(const float *const activations,
                             float *gradients,
                             const int *const flat_labels,
                             const int *const label_lengths,
                             const int *const input_lengths,
                             int alphabet_size,
                             int minibatch,
                             float *costs,
                             void *workspace,
                             struct ctcOptions options);

/* upstream ctc.h: get_workspace_size().  All length arrays are HOST pointers. */
ctcStatus_t get_workspace_size(const int *const label_lengths,
                               const int *const input_lengths,
                               int alphabet_size,
                               int minibatch,
                               struct ctcOptions info,
                               size_t *size_bytes);

/* ------------------------------------------------------------------------------------------------
 * Extended entry point (no upstream counterpart).  It exists to remove the per-step costs SURVEY.md
 * section 8a lists for the reference glue: strided activations (no .contiguous() copy, A1), a folded
 * gradient scale (1/B, task weight; A8), an explicit max_time, optional device-side costs and an
 * optional non-blocking return (A9).
 * ---------------------------------------------------------------------------------------------- */

#define CTC_B200_FLAG_NO_SYNC 0x1u       /* do not synchronise; costs_host/status_host must be NULL.  Out-of-range
                                            utterances are still redone in log space (the detour is a device-side
                                            kernel enqueued behind the fast ones) */
#define CTC_B200_FLAG_SERIAL_LAUNCHES 0x2u /* keep every kernel on `stream` (no internal fork/join streams) */
#define CTC_B200_FLAG_NO_BIDIR 0x8u        /* small batches: keep the three-sweep fused kernel instead of the bidirectional path */
#define CTC_B200_FLAG_NO_FALLBACK 0x4u     /* report out-of-range utterances instead of re-running them in log space */
/* bits 8..10: variant ladder override (0 auto, 1 throughput, 2 latency, 3 throughput with 8-step chunks,
 * 4 warp ladder: one warp per utterance, register-resident, fp64 recursion;
 * 5 fp32 warp ladder: the same organisation with a single-precision recursion and per-lane block exponents) */

typedef struct ctcB200Call {
    const float *activations;   /* DEVICE; element (t,b,k) at t*act_stride_t + b*act_stride_b + k */
    long long act_stride_t;     /* in elements */
    long long act_stride_b;     /* in elements */
    float *gradients;           /* DEVICE dense [max_time][minibatch][alphabet_size], or NULL */
    const int *flat_labels;     /* HOST */
    const int *label_lengths;   /* HOST [minibatch] */
    const int *input_lengths;   /* HOST [minibatch], each <= max_time */
    int alphabet_size;
    int minibatch;
    int max_time;               /* T dimension of activations / gradients */
    int blank_label;
    float grad_scale;           /* gradients are multiplied by this (1.0f = upstream behaviour) */
    float *costs_host;          /* HOST [minibatch] or NULL */
    float *costs_device;        /* DEVICE [minibatch] or NULL (in addition to / instead of costs_host) */
    int *status_host;           /* HOST [minibatch] or NULL: per-utterance CTC_B200_UTT_* bits */
    void *workspace;            /* DEVICE */
    size_t workspace_bytes;
    CUstream stream;
    unsigned int flags;
    long long *debug_device;    /* DEVICE [minibatch][16] or NULL: per-utterance {forward cycles, total cycles,
                                   total ns, SM id, 12 phase cycle counters} -- profiling aid only */
    float *kernel_ms_host;      /* HOST or NULL: device time (CUDA events on `stream`) from the first kernel launch of
                                   this call to the completion of its last one; needs the blocking mode */
    int *status_device;         /* DEVICE [minibatch] or NULL: per-utterance CTC_B200_UTT_* bits left on the device
                                   (what a CTC_B200_FLAG_NO_SYNC caller reads later, or never) */
} ctcB200Call;

/* per-utterance status bits (status_host) */
#define CTC_B200_UTT_INFEASIBLE 0x1     /* L + repeats > T (or T == 0): cost 0, gradient 0 */
#define CTC_B200_UTT_INF_COST 0x2       /* no alignment has non-zero probability: cost = +inf */
#define CTC_B200_UTT_BAD_LABEL 0x4      /* label out of range or equal to blank */
#define CTC_B200_UTT_RANGE 0x8          /* non-finite partition function even in log space (NaN inputs), or out of range
                                           with the fallback disabled */
#define CTC_B200_UTT_LOGSPACE 0x10      /* the utterance exceeded the fp64 linear-domain range of the fused kernel
                                           (or has a +inf cost) and was computed by the fp64 log-space kernel */
#define CTC_B200_UTT_WIDE 0x20          /* informational: the utterance exceeded the range of the fp32 kernel (per-lane block
                                           exponents) and was computed by the fp64 linear-domain kernel */

ctcStatus_t ctc_b200_workspace_size(const int *label_lengths, const int *input_lengths,
                                    int alphabet_size, int minibatch, int max_time,
                                    int want_gradients, size_t *size_bytes);

ctcStatus_t ctc_b200_compute(const ctcB200Call *call);

/*
 * Host-buffer entry point: activations and gradients live in HOST memory (pinned for full overlap).  The
 * batch is cut into `n_chunks` slices along the minibatch axis and streamed through the GPU as a
 * three-stage pipeline (H2D copy of slice i+1 | kernels of slice i | D2H copy of slice i-1) on internal
 * streams forked from `stream`; the call blocks until costs and gradients are on the host.  This is the
 * shape of the reference's CPU-tensor call (warpctc_pytorch cpu_ctc: host activations in, host gradients
 * out) served by the GPU.
 */
typedef struct ctcB200HostCall {
    const float *activations;   /* HOST dense [max_time][minibatch][alphabet_size] */
    float *gradients;           /* HOST same shape, or NULL for costs only */
    const int *flat_labels;     /* HOST */
    const int *label_lengths;   /* HOST [minibatch] */
    const int *input_lengths;   /* HOST [minibatch] */
    int alphabet_size;
    int minibatch;
    int max_time;
    int blank_label;
    float grad_scale;
    float *costs_host;          /* HOST [minibatch] */
    int *status_host;           /* HOST [minibatch] or NULL */
    void *workspace;            /* DEVICE, ctc_b200_workspace_size_host() bytes */
    size_t workspace_bytes;
    CUstream stream;
    int n_chunks;               /* <= 0: automatic */
    unsigned int flags;         /* ladder override bits only */
} ctcB200HostCall;

ctcStatus_t ctc_b200_workspace_size_host(const int *label_lengths, const int *input_lengths,
                                         int alphabet_size, int minibatch, int max_time,
                                         int want_gradients, int n_chunks, size_t *size_bytes);

ctcStatus_t ctc_b200_compute_host(const ctcB200HostCall *call);

/*
 * Greedy (best-path) decode of B x T x V activations or probabilities: per utterance, argmax over the alphabet
 * for each frame t < sizes[b], collapse repeats, drop blanks.  Replaces GreedyDecoder.decode of the reference
 * (/root/reference/codes/decoder.py:143-160: torch.max + a Python loop with one .item() per frame).
 * Everything is DEVICE memory; the call only enqueues one kernel on `stream`.
 *   probs            element (b, t, k) at b*stride_b + t*stride_t + k   (elements)
 *   sizes_device     [minibatch] valid frames, or NULL for max_time
 *   tokens_device    [minibatch][max_time]: the first counts[b] entries of row b are the decoded symbols
 *   offsets_device   [minibatch][max_time] frame index of each decoded symbol, or NULL
 *   counts_device    [minibatch]
 */
ctcStatus_t ctc_b200_greedy_decode(const float *probs, long long stride_b, long long stride_t,
                                   const int *sizes_device, int minibatch, int max_time, int alphabet_size,
                                   int blank_label, int *tokens_device, int *offsets_device, int *counts_device,
                                   CUstream stream);

/*
 * Batched Levenshtein distance between decoded transcripts and their references; everything is DEVICE memory and
 * the call only enqueues one kernel on `stream` (one warp per utterance).  Replaces Decoder.wer / Decoder.cer of
 * the reference (/root/reference/codes/decoder.py:49-78: python-Levenshtein on host strings, called per utterance
 * from codes/metrics.py:118 and test.py:83-84) for hypotheses that ctc_b200_greedy_decode left on the device.
 *   hyp_tokens_device   row b at hyp_tokens_device + b*hyp_stride, hyp_counts_device[b] valid entries (<= max_hyp)
 *   refs_device         concatenated reference token ids; reference b = ref_lengths_device[b] entries starting at
 *                       ref_offsets_device[b]  (<= max_ref <= 2047 each)
 *   mode                0: distance over the token ids as they are
 *                       1: CER -- tokens equal to space_label are removed from both sides first
 *                       2: WER -- both sides are split into words at runs of space_label; distance over words
 *   distances_device    [minibatch] edit distance (unit costs)
 *   normalisers_device  [minibatch] or NULL: what the reference's metric divides by (codes/metrics.py:145-160):
 *                       reference length including spaces (modes 0, 1) / number of reference words (mode 2)
 */
ctcStatus_t ctc_b200_edit_distance(const int *hyp_tokens_device, long long hyp_stride, const int *hyp_counts_device,
                                   int max_hyp, const int *refs_device, const int *ref_offsets_device,
                                   const int *ref_lengths_device, int max_ref, int minibatch, int space_label, int mode,
                                   int *distances_device, int *normalisers_device, CUstream stream);

/*
 * Classifier head that produces the CTC activations: BatchNorm1d(features) followed by Linear(features -> classes,
 * no bias) over the rows = T*B frames of the last recurrent layer.  Replaces `self.fc` of the reference
 * (/root/reference/codes/model.py:177-180, 199-207 and SequenceWiseClassifier, model.py:205-222): 3 + 5 PyTorch
 * kernels over rows x features tensors become two passes over x forward and two backward, the three contractions
 * on the tensor cores (tcgen05 kind::tf32 with an error-compensated operand split: fp32-level results).  All pointers are DEVICE
 * memory, fp32, dense row-major; calls only enqueue work on `stream`.  classes <= 64, features % 4 == 0.
 *   x              [rows][features]           out / dlogits   [rows][classes]  (rows in T x B order: `out` is the
 *                                                               T x B x V tensor the CTC engine reads)
 *   weight         [classes][features]        bn_weight, bn_bias [features] (NULL: 1 / 0)
 *   running_mean, running_var [features]: read in eval mode; updated in training mode (momentum, unbiased variance)
 *                  when non-NULL
 *   training       1: batch statistics (BatchNorm training semantics), 0: running statistics
 *   softmax        1: `out` holds softmax probabilities over the classes (the reference's eval-mode output)
 *   save_mean, save_invstd [features] written by the forward call, consumed by the backward call
 *   dx             [rows][features] or NULL;  dweight [classes][features], dbn_weight, dbn_bias [features] or NULL
 * The backward call takes dlogits = dLoss/dlogits (for the CTC loss: the engine's gradient buffer, as is).
 */
typedef struct {
    const float *x;
    int rows, features, classes;
    const float *weight;
    const float *bn_weight, *bn_bias;
    float *running_mean, *running_var;
    float eps, momentum;
    int training, softmax;
    float *out;
    float *save_mean, *save_invstd;
    void *workspace;
    size_t workspace_bytes;
    CUstream stream;
} ctcB200HeadForward;

typedef struct {
    const float *x, *dlogits;
    int rows, features, classes;
    const float *weight;
    const float *bn_weight, *bn_bias;
    const float *save_mean, *save_invstd;
    int training;
    float *dx, *dweight, *dbn_weight, *dbn_bias;
    void *workspace;
    size_t workspace_bytes;
    CUstream stream;
} ctcB200HeadBackward;

ctcStatus_t ctc_b200_head_workspace_size(int rows, int features, int classes, size_t *size_bytes);
ctcStatus_t ctc_b200_head_forward(const ctcB200HeadForward *call);
ctcStatus_t ctc_b200_head_backward(const ctcB200HeadBackward *call);

/*
 * Loss glue on the device.  The reference wraps the loss call in `_sanitize_loss`
 * (/root/reference/codes/engine.py:19-32: `/ average`, `.sum()`, a host-side `== inf` test that forces a sync,
 * `0 * loss_sum`), multiplies by a task weight (engine.py:77) and lets `_CTC.backward` rescale the whole T x B x V
 * gradient tensor by the incoming scalar (one more pass, engine.py:84).  With the scale folded into
 * ctcB200Call.grad_scale and the cost vector left on the device, two small calls finish the job without a host
 * round trip; both only enqueue work on `stream`.
 *
 * ctc_b200_reduce_costs: loss_device[0] = scale * sum_b costs_device[b] (fp64, fixed order => deterministic).
 *   If zero_infinite is set and the sum is +-inf, loss_device[0] = 0 and flag_device[0] = 1 (else 0): the guard of
 *   engine.py:27-30 (which intends "setting loss value to 0"; its `0 * inf` is NaN).  A NaN sum stays NaN.
 * ctc_b200_scale_gradients: gradients[i] *= f with f = scale_host * (scale_device ? *scale_device : 1), or 0 when
 *   zero_flag_device && *zero_flag_device.  When f is exactly 1 every CTA returns after reading the two scalars,
 *   so the common `loss.backward()` costs no pass over the tensor.
 */
ctcStatus_t ctc_b200_reduce_costs(const float *costs_device, int minibatch, float scale, int zero_infinite,
                                  float *loss_device, int *flag_device, CUstream stream);
ctcStatus_t ctc_b200_scale_gradients(float *gradients, size_t count, float scale_host, const float *scale_device,
                                     const int *zero_flag_device, CUstream stream);

/* Human-readable description of the last failure on the calling thread ("" if none). */
const char *ctc_b200_last_error(void);

/* Build/device facts for logging: returns the compiled SM architecture (100) and, if non-NULL, writes
 * the number of kernel launches issued by this library on the calling thread since load. */
int ctc_b200_info(unsigned long long *launch_count);

#ifdef __cplusplus
}
#endif
#endif /* CTC_B200_CTC_H */
