/* multibox_b200 -- C ABI of the B200-native Multibox hot path.
 *
 * Drop-in boundary for the data-parallel hot path of gvanhorn38/multibox:
 *   - GT -> prior bipartite matching      (reference loss.py:8-53, the body of
 *                                          the tf.py_func at loss.py:81-82)
 *   - multibox loss forward + backward    (reference loss.py:55-117 + TF autodiff)
 *   - detection post-processing           (reference detect.py:74-131, 408-443;
 *                                          eval.py:142-175)
 *
 * Conventions (all functions):
 *   - every pointer is a DEVICE pointer into memory owned by the caller
 *     (PyTorch); the library allocates nothing persistent and keeps no global
 *     state besides the thread-local error string, per-device launch-attribute
 *     caches and a thread-local launch counter per workspace (see MBX_FLAG_PDL);
 *   - tensors are dense, row-major, float32 / int32 unless said otherwise and
 *     16-byte aligned;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises the device;
 *   - the return value is 0 on success, a negative MBX_E_* code for argument
 *     errors, or a positive cudaError_t; mbx_last_error() describes it;
 *   - data-dependent failures (NaN / -inf cost entries, infeasible matrices:
 *     the cases where scipy.optimize.linear_sum_assignment raises ValueError at
 *     reference loss.py:40) are reported asynchronously in a device status
 *     word (MBX_STATUS_* bits) that the host mirror reads back and turns into
 *     the same ValueError;
 *   - functions are re-entrant across streams as long as each in-flight call
 *     has its own workspace.
 */
#ifndef MULTIBOX_B200_H
#define MULTIBOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBX_VERSION 100

/* argument errors (negative return values) */
#define MBX_E_ARG        (-1)   /* null / misaligned pointer, bad size          */
#define MBX_E_TOO_LARGE  (-2)   /* P or M beyond what one CTA's shared memory holds */
#define MBX_E_WORKSPACE  (-3)   /* workspace too small                          */

/* device status word bits */
#define MBX_STATUS_INVALID_COST  1u  /* NaN or -inf cost entry (scipy: "matrix contains invalid numeric entries") */
#define MBX_STATUS_INFEASIBLE    2u  /* no finite assignment (scipy: "cost matrix is infeasible") */
#define MBX_STATUS_BAD_NUM_GT    4u  /* num_gt[b] outside [0, M] (clamped)      */
#define MBX_STATUS_AR_TIMEOUT    8u  /* fused loss all-reduce: a peer rank never arrived */

#define MBX_MAX_PEERS 8              /* GPUs of one NVLink/NVSwitch box */
#define MBX_MAX_HEADS 8              /* detection heads (the reference has 6: 8x8, 6x6, 4x4, 3x3, 2x2, 1x1) */

/* flags */
#define MBX_FLAG_LOGITS        1u   /* `confidences` holds logits; the kernel applies the
                                       sigmoid of reference model.py:322 and the confidence
                                       gradient is returned w.r.t. the logits */
#define MBX_FLAG_BOUNDARY      2u   /* strict py_func boundary (reference loss.py:81): locations
                                       already have the prior added, confidences already have
                                       +1e-10 added; `priors` is ignored */
#define MBX_FLAG_AR_DEFERRED   8u   /* mbx_match_loss_allreduce: publish this step's sums, complete an
                                       EARLIER step's reduction (the previous one; 12 back with MBX_FLAG_PDL):
                                       no waiting for slower peers, no NVLink access inside the kernel */
#define MBX_FLAG_STATIC        16u  /* keep the static image -> CTA assignment even when the batch exceeds
                                       the resident CTAs (default then: heavy-first dynamic scheduling) */
#define MBX_FLAG_HOST_RESULTS  32u  /* `results` is mapped pinned HOST memory that the caller polls: the
                                       launch sequence word results[15] is published after a system-scope
                                       fence (costs ~1 us; without the flag word 15 is still written last) */
#define MBX_FLAG_PDL           64u  /* programmatic dependent launch: the caller promises that the INPUT tensors
                                       do not come from the kernel that precedes this call on `stream` (e.g. the
                                       previous step, or inputs staged in pinned host memory).  The kernel may then
                                       start while that preceding kernel is still running and waits for it
                                       (griddepcontrol.wait) only before its first write to outputs / workspace:
                                       consecutive steps overlap launch latency, tail and head.  Results are
                                       identical.  Works with static and dynamic image scheduling (the library keeps
                                       a per-thread launch counter per workspace to alternate between two scheduler
                                       slots) and with the fused all-reduce (deferred mode then reports the global
                                       sums two steps back).  Ignored with stacked_gt / n_stacked (a scan kernel
                                       runs in front) and MBX_FLAG_GENERIC.  mbx_detect honours it as well. */
#define MBX_FLAG_GENERIC       4u   /* force the generic shared-memory matching kernel (any P) instead
                                       of the register-resident family (tuning / testing) */
#define MBX_FLAG_WARPS_SHIFT   8    /* bits 8..15: force CTA size in warps (0 = heuristic) */
#define MBX_FLAG_COLS_SHIFT    16   /* bits 16..23: force columns per thread of the register-resident
                                       kernel (0 = heuristic) */

int         mbx_version(void);
const char *mbx_last_error(void);

/* Number of SMs / max dynamic shared memory the library sees on the current device. */
int mbx_device_info(int *sm_count, int *max_smem_per_block);

/* ------------------------------------------------------------------------- *
 * Matching (+ fused loss forward/backward)
 * ------------------------------------------------------------------------- */

/* Bytes of scratch mbx_match_loss needs for a batch of B images.  The first
 * use of a workspace must see it zero-filled; the library leaves it reusable. */
size_t mbx_match_workspace_bytes(int B, int P, int M);

/* One pass of the training hot path over a batch.
 *
 * Replaces: reference loss.py:8-53 (compute_assignments: log terms, cost
 * matrix, scipy linear_sum_assignment, mask + stacked GT) and, when the loss
 * outputs are requested, loss.py:67-74,88-101 plus the gradients TF autodiff
 * derives from them (and model.py:322 with MBX_FLAG_LOGITS).
 *
 * Inputs
 *   locations    [B,P,4]  predicted offsets (or absolute boxes with MBX_FLAG_BOUNDARY)
 *   confidences  [B,P]    post-sigmoid confidences (logits with MBX_FLAG_LOGITS)
 *   gt_bboxes    [B,M,4]  zero-padded ground truth (reference inputs.py:346-348)
 *   num_gt       [B]      int32, real GT count per image, 0 <= n <= M <= P
 *   priors       [P,4]    prior boxes (NULL with MBX_FLAG_BOUNDARY)
 *   alpha                 LOCATION_LOSS_ALPHA
 * Outputs (each may be NULL = not wanted)
 *   mask            [B,P]   int32 0/1            (reference: assignment_partitions)
 *   matched_gt_idx  [B,P]   int32 GT index or -1
 *   stacked_gt      [cap,4] matched GT rows in (image, ascending prior) order
 *                           (reference: stacked_gt_bboxes); cap >= sum(num_gt)
 *   d_locations     [B,P,4] dL/d locations
 *   d_confidences   [B,P]   dL/d confidences (d logits with MBX_FLAG_LOGITS)
 *   confidences_out [B,P]   sigmoid(logits) (MBX_FLAG_LOGITS only)
 *   results         [16]    float32 words: [0] location_loss, [1] confidence_loss,
 *                           [2] status word (exact small integer), [3] number of
 *                           matched priors (exact while < 2^24), [4..7] the two
 *                           losses as float64 (2 x 8 bytes), [8..11] the two losses
 *                           summed over all ranks as float64 (== [4..7] unless
 *                           mbx_match_loss_allreduce is used with world > 1),
 *                           [12],[13] the same as float32, [14] the step index the
 *                           global sums belong to
 *                           [15] launch sequence number (uint32 bits, never 0), stored last;
 *                           with MBX_FLAG_HOST_RESULTS after a system-scope fence, so that a
 *                           caller that passed mapped pinned HOST memory as `results` may poll
 *                           it instead of synchronising the stream
 *   n_stacked       [1]     int32 number of rows written to stacked_gt
 * The status word is also OR-ed into results[2]; it is 0 when every image was
 * solved.  grads/loss outputs are produced iff `results` is non-NULL.
 */
int mbx_match_loss(const float *locations, const float *confidences,
                   const float *gt_bboxes, const int32_t *num_gt,
                   const float *priors, int B, int P, int M, float alpha,
                   unsigned flags,
                   int32_t *mask, int32_t *matched_gt_idx,
                   float *stacked_gt, int32_t *n_stacked,
                   float *d_locations, float *d_confidences,
                   float *confidences_out, float *results,
                   void *workspace, size_t workspace_bytes, void *stream);

/* Ragged ground truth (CSR) instead of the zero-padded [B,M,4] block the reference's input
 * pipeline builds (inputs.py:340-348, eval_inputs.py:78-93): image b owns rows
 * gt_row_offsets[b] .. gt_row_offsets[b+1]-1 of gt_flat [N,4]; gt_row_offsets is int32 [B+1],
 * non-decreasing.  M is only the per-image capacity (an image with more than M rows trips
 * MBX_STATUS_BAD_NUM_GT and is clamped).  Saves the 16*(M - n) padding bytes per image; every
 * other argument and every result is as in mbx_match_loss. */
int mbx_match_loss_ragged(const float *locations, const float *confidences,
                          const float *gt_flat, const int32_t *gt_row_offsets,
                          const float *priors, int B, int P, int M, float alpha,
                          unsigned flags,
                          int32_t *mask, int32_t *matched_gt_idx,
                          float *stacked_gt, int32_t *n_stacked,
                          float *d_locations, float *d_confidences,
                          float *confidences_out, float *results,
                          void *workspace, size_t workspace_bytes, void *stream);

/* Head-layout inputs: the training hot path fed straight from the detection heads' conv
 * outputs, replacing the reshape + tf.concat (+ tf.sigmoid with MBX_FLAG_LOGITS) of reference
 * model.py:295-322.  Head h (grid g_h x g_h, K_h boxes per cell; the reference has the six grids
 * 8,6,4,3,2,1 in this order) contributes head_priors[h] = g_h*g_h*K_h consecutive priors; its
 * NHWC conv outputs [B,g,g,K*4] / [B,g,g,K] are exactly [B,head_priors[h],4] / [B,head_priors[h]].
 * Gradients come back in the same per-head layouts (all heads or none), so neither the
 * concatenated [B,P,4] / [B,P] tensors nor their gradients ever exist in HBM.  mask /
 * matched_gt_idx / confidences_out stay in the concatenated prior order [B,P]. */
typedef struct mbx_heads {
    int32_t num_heads;                          /* 1 .. MBX_MAX_HEADS */
    int32_t head_priors[MBX_MAX_HEADS];         /* sum == P */
    const float *locations[MBX_MAX_HEADS];      /* [B, head_priors[h], 4] */
    const float *confidences[MBX_MAX_HEADS];    /* [B, head_priors[h]] (logits with MBX_FLAG_LOGITS) */
    float *d_locations[MBX_MAX_HEADS];          /* NULL = gradient not wanted */
    float *d_confidences[MBX_MAX_HEADS];
} mbx_heads;

/* gt: padded (gt_row_offsets == NULL, num_gt given) or ragged (gt_row_offsets given, num_gt ignored).
 * MBX_FLAG_BOUNDARY is not available (the heads produce offsets, not absolute boxes). */
int mbx_match_loss_heads(const mbx_heads *heads,
                         const float *gt_bboxes, const int32_t *num_gt, const int32_t *gt_row_offsets,
                         const float *priors, int B, int P, int M, float alpha,
                         unsigned flags,
                         int32_t *mask, int32_t *matched_gt_idx,
                         float *stacked_gt, int32_t *n_stacked,
                         float *confidences_out, float *results,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Same as mbx_match_loss, with the SUM all-reduce of the two loss scalars over the
 * `world` GPUs of one NVLink box FUSED into the kernel (reference semantics: the losses
 * are batch sums, loss.py:100-101; the batch is sharded by image, one process per GPU).
 * Messages are four 8-byte words {32 payload bits, 32-bit tag = step + 1} per rank and step, valid exactly
 * when the tag matches (no system-scope fence, no remote atomic, no acknowledgement).  Blocking mode
 * (default): the last CTA stores its words into every rank's table through peer memory (NVLink P2P
 * stores), waits for the other ranks' words of the same step and adds them in rank order; results[8..13]
 * then hold the global sums, bit-identical on every rank.  Every rank must call this once per step, in
 * step order.
 *   peer_buffers [world]  HOST array of device pointers: rank r's symmetric buffer of
 *                         mbx_allreduce_buffer_bytes() bytes (zero-filled once), mapped
 *                         into this process (CUDA IPC / VMM; torch symmetric memory)
 * With MBX_FLAG_AR_DEFERRED the matching kernel does no NVLink access at all: the step leaves its sums in its
 * OWN outbox (local stores) and completes an EARLIER step's reduction instead -- the previous one, or the one
 * 12 steps back under MBX_FLAG_PDL (mbx_allreduce_config) -- from its own table, into which a one-warp relay
 * kernel on a side stream (enqueued by this call every 8 steps) forwards every rank's outbox words; words the
 * relay has not delivered (CUDA graph replays, a relay that exited idle) are pulled from the peers' outboxes
 * with loads over NVLink instead.  results[14] = index of the step the global sums belong to, -1 = none yet.
 * No rank waits for a slower peer inside the step; mbx_allreduce_flush completes the newest step on demand.
 * A rank that never arrives trips MBX_STATUS_AR_TIMEOUT (~2 s) instead of hanging; the condition is
 * sticky (later steps report it at once) until the buffers are zero-filled again with no step in flight. */
size_t mbx_allreduce_buffer_bytes(void);
/* Process-wide tuning of the deferred mode (call it identically on every rank, with no step in flight, before
 * the first deferred step on a buffer or after zero-filling the buffers): `pdl_lag` = how many steps back the
 * reduction completed by a MBX_FLAG_PDL step lies (default 12; without PDL it is always 1), `relay_batch` = how
 * many steps' sums the relay kernel forwards over NVLink in one burst under PDL (default 8 = a relay's life). */
int mbx_allreduce_config(int pdl_lag, int relay_batch);
int mbx_allreduce_flush(float *results, void *workspace, size_t workspace_bytes,
                        const unsigned long long *peer_buffers, int world, int rank, void *stream);
int mbx_match_loss_allreduce(const float *locations, const float *confidences,
                             const float *gt_bboxes, const int32_t *num_gt,
                             const float *priors, int B, int P, int M, float alpha,
                             unsigned flags,
                             int32_t *mask, int32_t *matched_gt_idx,
                             float *stacked_gt, int32_t *n_stacked,
                             float *d_locations, float *d_confidences,
                             float *confidences_out, float *results,
                             void *workspace, size_t workspace_bytes,
                             const unsigned long long *peer_buffers, int world, int rank,
                             void *stream);

/* Prepared launches.  A plan holds the argument list of mbx_match_loss_allreduce (world = 1, rank = 0,
 * peer_buffers = NULL: of mbx_match_loss), checked when it is launched exactly as the direct call checks
 * it; mbx_match_plan_launch(plan, stream) then enqueues one step with a two-argument foreign call.  For
 * training loops whose step (a few microseconds at batch 32) is shorter than the host's marshalling of
 * 24 arguments.  The plan is owned by the caller (destroy it; it holds no device resources) and may be
 * launched any number of times, on any stream, while the pointers it was created with stay valid. */
typedef struct mbx_match_plan mbx_match_plan;
int  mbx_match_plan_create(mbx_match_plan **plan,
                           const float *locations, const float *confidences,
                           const float *gt_bboxes, const int32_t *num_gt,
                           const float *priors, int B, int P, int M, float alpha,
                           unsigned flags,
                           int32_t *mask, int32_t *matched_gt_idx,
                           float *stacked_gt, int32_t *n_stacked,
                           float *d_locations, float *d_confidences,
                           float *confidences_out, float *results,
                           void *workspace, size_t workspace_bytes,
                           const unsigned long long *peer_buffers, int world, int rank);
int  mbx_match_plan_launch(const mbx_match_plan *plan, void *stream);
/* Host-buffer step in ONE foreign call: cudaMemcpyAsync(dev_dst, host_src, nbytes, HostToDevice, stream)
 * followed by mbx_match_plan_launch(plan, stream).  `host_src` is the caller's (pinned) packed input
 * buffer, `dev_dst` the device buffer the plan's input pointers point into.  Steps enqueued on different
 * streams overlap one step's copy with another step's kernel (multibox_b200/loss.py MultiboxLossStep
 * (own_stream=True)). */
int  mbx_match_plan_launch_staged(const mbx_match_plan *plan, const void *host_src, void *dev_dst,
                                  size_t nbytes, void *stream);
void mbx_match_plan_destroy(mbx_match_plan *plan);

/* ------------------------------------------------------------------------- *
 * Detection post-processing
 * ------------------------------------------------------------------------- */

size_t mbx_detect_workspace_bytes(int B, int P, int k_max);

/* One pass of detection post-processing over a batch of patches.
 *
 * Replaces the per-image loop body of reference detect.py:408-436: decode
 * (:412), clip (:413), filter_proposals (:74-104), descending confidence sort
 * and top max_to_keep (:423-427; ties broken as numpy's stable argsort followed
 * by reversal does: equal confidences in DESCENDING prior index), optional
 * greedy NMS on the kept boxes (extension: no reference counterpart), and
 * convert_proposals (:106-131, float64).  With restrictions [0,0,1,1] and
 * nms_iou < 0 it is also the decode/sort/top-k of reference eval.py:146-167.
 * A confidence of -inf marks an empty slot: such a prior is never a proposal (used by pooled
 * multi-patch candidate lists; no sigmoid output is -inf, so reference inputs are unaffected).
 *
 * Inputs
 *   locations    [B,P,4], confidences [B,P] (logits with MBX_FLAG_LOGITS), priors [P,4]
 *                (priors NULL = the locations are absolute boxes already; nothing is staged)
 *   restrictions [B,4]   float32 x1,y1,x2,y2 limits (NULL = [0,0,1,1] everywhere)
 *   max_to_keep  [B]     int32 (NULL = k_max everywhere); clamped to k_max
 *   offsets      [B,2]   int32 (y,x) patch offset    } NULL = identity conversion
 *   patch_dims   [B,2]   int32 (h,w)                 }
 *   image_dims   [B,2]   int32 (h,w)                 }
 *   is_flipped   [B]     int32                       }
 *   nms_iou              IoU threshold; < 0 disables NMS
 * Outputs (padded to k_max per image; rows >= count are zero / -1)
 *   out_boxes    [B,k_max,4] float64 image coordinates (NULL = not wanted)
 *   out_patch_boxes [B,k_max,4] float32 decoded+clipped patch coordinates (NULL ok)
 *   out_scores   [B,k_max]   float32
 *   out_prior_idx[B,k_max]   int32 original prior index
 *   out_count    [B]         int32 detections kept
 */
int mbx_detect(const float *locations, const float *confidences, const float *priors,
               const float *restrictions, const int32_t *max_to_keep,
               const int32_t *offsets, const int32_t *patch_dims,
               const int32_t *image_dims, const int32_t *is_flipped,
               int B, int P, int k_max, float nms_iou, unsigned flags,
               double *out_boxes, float *out_patch_boxes, float *out_scores,
               int32_t *out_prior_idx, int32_t *out_count,
               void *workspace, size_t workspace_bytes, void *stream);

/* mbx_detect fed straight from the detection heads' conv outputs (see mbx_heads above): decode,
 * sigmoid (MBX_FLAG_LOGITS), filter, top-k, NMS and conversion without the reshape + concat of
 * reference model.py:295-322.  Only heads->locations / confidences / head_priors are read. */
int mbx_detect_heads(const mbx_heads *heads, const float *priors,
                     const float *restrictions, const int32_t *max_to_keep,
                     const int32_t *offsets, const int32_t *patch_dims,
                     const int32_t *image_dims, const int32_t *is_flipped,
                     int B, int P, int k_max, float nms_iou, unsigned flags,
                     double *out_boxes, float *out_patch_boxes, float *out_scores,
                     int32_t *out_prior_idx, int32_t *out_count,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Batched filter_proposals (reference detect.py:74-104): order-preserving
 * compaction of the boxes that lie inside the restriction rectangle.
 *   bboxes [B,P,4], confidences [B,P], restrictions [B,4] (NULL = the reference's
 *   default [0.1,0.1,0.9,0.9]) -> out_bboxes [B,P,4], out_confidences [B,P],
 *   out_idx [B,P] int32 original index (NULL ok), out_count [B] int32.
 * Rows >= out_count[b] are left untouched. */
int mbx_filter_proposals(const float *bboxes, const float *confidences, const float *restrictions,
                         int B, int P, float *out_bboxes, float *out_confidences,
                         int32_t *out_idx, int32_t *out_count, void *stream);

/* Batched convert_proposals (reference detect.py:106-131), float64 output:
 *   bboxes [B,K,4] f32, offsets [B,2] (y,x), patch_dims [B,2] (h,w), image_dims [B,2]
 *   (h,w) int32, is_flipped [B] int32 (NULL = 0), counts [B] int32 (NULL = K; rows >=
 *   count are written as 0) -> out_boxes [B,K,4] float64. */
int mbx_convert_proposals(const float *bboxes, const int32_t *offsets, const int32_t *patch_dims,
                          const int32_t *image_dims, const int32_t *is_flipped, const int32_t *counts,
                          int B, int K, double *out_boxes, void *stream);

/* ------------------------------------------------------------------------- *
 * Diagnostics (used by the parity tests; not on the hot path)
 * ------------------------------------------------------------------------- */

/* out[i] = the kernels' float32 log of in[i] (bit-compatible with numpy's np.log
 * on float32, which reference loss.py:21,25 calls). */
int mbx_debug_nplog(const float *in, float *out, long long n, void *stream);

/* Adds to *mismatches (device uint64, caller-zeroed) the number of float32 bit patterns in
 * [first_bits, first_bits+count) whose square root in the cost kernel differs from sqrt.rn
 * (what numpy's np.linalg.norm / np.sqrt compute on the host). */
int mbx_debug_sqrt_mismatches(unsigned first_bits, unsigned count, unsigned long long *mismatches,
                              void *stream);

/* Adds to *violations (device uint64, caller-zeroed) the number of float32 bit patterns x in
 * [first_bits, first_bits+count) for which the fast hardware log the kernels feed into the CHEAP cost
 * bound differs from the exact (numpy-compatible) log by more than the bound's error analysis allows:
 * |__logf(x) - log_numpy(x)| > 2^-21 + 2^-19 |__logf(x)| (multibox_b200/csrc/mbx_bound.h). */
int mbx_debug_fastlog_violations(unsigned first_bits, unsigned count, unsigned long long *violations,
                                 void *stream);

/* The cost matrix of reference loss.py:33-35 for ONE image, as the matching
 * kernel evaluates it on the fly: loc [P,4] absolute boxes, conf [P] (epsilon
 * added), gt [n,4] -> C [P,n] float64 row-major. */
int mbx_debug_cost_matrix(const float *loc, const float *conf, const float *gt, int P, int n,
                          float alpha, double *C, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIBOX_B200_H */
