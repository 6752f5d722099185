/*
 * dagb200.h -- C ABI of libdagb200.so: B200 (sm_100a) kernels for the DAG-loss hot path of
 * ictnlp/DASpeech (DASpeech/custom_ops).
 *
 * This is the drop-in boundary.  The reference exposes the path through a pybind11 module
 * `dag_loss_fn` with four functions taking torch::Tensor (DASpeech/custom_ops/dag_loss.cpp:19-29);
 * each entry point below replaces one of them with plain device pointers, sizes and a CUDA stream
 * (no torch types).  The Python host layer (daspeech_b200/custom_ops/dag_loss.py) re-creates the
 * reference's operator surface (dag_loss.py:66-299) on top of these calls via ctypes; INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; tensors are row-major contiguous
 *     unless a stride argument is given (strides are in ELEMENTS);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued
 *     asynchronously on it, nothing synchronises the host;
 *   - return value: 0 on success, <0 argument error (DAGB200_E*), >0 a cudaError_t; a human readable
 *     message for the calling thread's last failure is returned by dagb200_last_error();
 *   - lengths are int64 (as in the reference, dag_loss.cu:332);
 *   - `status` (nullable, int32[B]) receives per-sample device-side precondition violations that the
 *     reference turns into CUDA_KERNEL_ASSERT (dag_loss.cu:68-69, dag_best_alignment.cu:67-70,118):
 *     0 ok, DAGB200_ST_* otherwise.  Offending samples get -inf alpha/beta (loss -inf) and an all -1
 *     path instead of a sticky device assert.
 */
#ifndef DAGB200_H
#define DAGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAGB200_VERSION 100

/* element types */
#define DAGB200_F32 0
#define DAGB200_F16 1
#define DAGB200_BF16 2 /* extension: the reference rejects bf16 (logsoftmax_gather.cu:46-58) */
#define DAGB200_F64 3

/* argument errors */
#define DAGB200_EINVAL (-1)   /* bad shape / null pointer / bad config */
#define DAGB200_EDTYPE (-2)   /* unsupported element type */
#define DAGB200_ELIMIT (-3)   /* size beyond what the kernels index */
#define DAGB200_EWORKSPACE (-4)

/* per-sample device status */
#define DAGB200_ST_OK 0
#define DAGB200_ST_LEN_LT2 1      /* target/output length < 2            (dag_loss.cu:68)   */
#define DAGB200_ST_GRAPH_SMALL 2  /* output_length < target_length       (dag_loss.cu:69)   */
#define DAGB200_ST_TOO_SHORT 3    /* (Tn-1)*T+1 < O                      (dag_best_alignment.cu:69) */
#define DAGB200_ST_NO_PATH 4      /* Viterbi end cell unreachable        (dag_best_alignment.cu:118) */

int dagb200_version(void);
const char *dagb200_last_error(void);

/* Process-wide switch: 1 = always run the exact log-domain kernels (one exp per lattice edge, the reference's
 * arithmetic); 0 (default) = blocked tensor-core kernels for fp32 lattices (flush-to-zero contract). */
void dagb200_set_exact(int on);
int dagb200_get_exact(void);

/* Measurement aid (off by default): with dagb200_set_profile(1) every entry point records CUDA events on its
 * stream around each kernel it launches; dagb200_get_profile() synchronises on them and returns the durations of
 * the most recent launches in ms: [0] transition-tile precompute, [1] alpha/beta recurrences, [2] grad_match,
 * [3] grad_links, [4] Viterbi; -1 where nothing was recorded.                                              */
void dagb200_set_profile(int on);
int dagb200_get_profile(float *ms, int n);

/* Replaces `logsoftmax_gather` (dag_loss.cpp:22, logsoftmax_gather.cu:313-377).
 *   logits  [B][L][V] contiguous, dtype F32/F16/BF16/F64; OVERWRITTEN with softmax probabilities
 *           iff require_gradient != 0 (the reference's in-place contract, dag_loss.py:272-274)
 *   idx     int64, addressed idx[b*isb + l*isl + s*iss]; the stride-0 expanded view the criterion
 *           passes (nat_dag_loss.py:127: targets.unsqueeze(1).expand(-1, L, -1)) is isl = 0
 *   out     float32 (float64 for F64 logits), addressed out[b*osb + l*osl + s*oss]; the host layer
 *           passes a [B][S][L]-contiguous buffer (osb=S*L, osl=1, oss=L) and returns it as a
 *           [B,L,S]-shaped view so the criterion's transpose(1,2).contiguous() costs nothing.   */
int dagb200_logsoftmax_gather(void *logits, int dtype,
                              const int64_t *idx, int64_t isb, int64_t isl, int64_t iss,
                              void *out, int64_t osb, int64_t osl, int64_t oss,
                              int B, int L, int V, int S, int require_gradient, void *stream);

/* Replaces the torch ops of DagLogsoftmaxGatherFunc.backward (dag_loss.py:293-295):
 *   grad_in = probs * (-sum_s gout[s]);  grad_in[idx[s]] += gout[s]
 * `probs_inout` [B][L][V] holds the saved probabilities and receives grad_in (same buffer, as the
 * reference).  gout is float32 (float64 for F64) addressed gout[b*gsb + l*gsl + s*gss].         */
int dagb200_logsoftmax_gather_backward(void *probs_inout, int dtype,
                                       const int64_t *idx, int64_t isb, int64_t isl, int64_t iss,
                                       const void *gout, int64_t gsb, int64_t gsl, int64_t gss,
                                       int B, int L, int V, int S, void *stream);

/* Replaces `dag_loss` (dag_loss.cpp:19, dag_loss.cu:313-375).
 *   match [B][M][L], links [B][L][T] (links[b][i][k] = log P(i -> i+k+1)), dtype F32 or F64
 *   alpha, beta [B][M][L] outputs, fully written (-inf outside the computed wedge); beta is
 *   computed only when require_gradient != 0 (else zero-filled, exactly what the reference returns)
 *   config: 1..4 accepted for signature compatibility (reference tile selector), ignored.
 *   workspace: device scratch of dagb200_dag_loss_workspace_bytes(B,M,L,T) bytes.  With it (fp32 only) the
 *   blocked tensor-core recurrences run; with workspace == NULL (or fp64, or a lattice beyond their
 *   shared-memory limits) the exact log-domain kernels run.                                        */
size_t dagb200_dag_loss_workspace_bytes(int B, int M, int L, int T);
int dagb200_dag_loss(const void *match, const void *links,
                     const int64_t *output_length, const int64_t *target_length,
                     void *alpha, void *beta, int dtype,
                     int B, int M, int L, int T, int require_gradient, int config,
                     void *workspace, size_t workspace_bytes,
                     int32_t *status, void *stream);

/* Replaces `dag_loss_backward` (dag_loss.cpp:20, dag_loss.cu:518-571).
 *   grad_output [B]; grad_match [B][M][L] and grad_links [B][L][T] are fully written.            */
int dagb200_dag_loss_backward(const void *grad_output, const void *alpha, const void *beta,
                              const void *match, const void *links,
                              const int64_t *output_length, const int64_t *target_length,
                              void *grad_match, void *grad_links, int dtype,
                              int B, int M, int L, int T, int config1, int config2, void *stream);

/* Same contract as dagb200_dag_loss_backward, plus a device scratch of
 * dagb200_dag_loss_backward_workspace_bytes(B,M,L,T) bytes.  With it (fp32, not in exact mode) the backward runs as
 * three kernels (dag_grad4.cu): one streaming pass that writes grad_match and bf16 hi/lo operand planes of exp(alpha),
 * exp(beta) with one integer frame per (row, 32-vertex block), the per-block-pair frame maxima, and a tcgen05
 * contraction over the target index for grad_links.  With workspace == NULL it is exactly dagb200_dag_loss_backward. */
size_t dagb200_dag_loss_backward_workspace_bytes(int B, int M, int L, int T);
int dagb200_dag_loss_backward_ws(const void *grad_output, const void *alpha, const void *beta,
                                 const void *match, const void *links,
                                 const int64_t *output_length, const int64_t *target_length,
                                 void *grad_match, void *grad_links, int dtype,
                                 int B, int M, int L, int T, int config1, int config2,
                                 void *workspace, size_t workspace_bytes, void *stream);

/* Replaces `dag_best_alignment` (dag_loss.cpp:21, dag_best_alignment.cu:209-253).
 *   alpha [B][M][L] (max-plus scores) may be NULL when the caller discards it (the reference's
 *   Python wrapper does, dag_loss.py:227-230); path int32 [B][L], -1 = vertex not on the path.
 *   config 1..4 reproduces the reference's tie-break for TRANS_BLOCK_SIZE 4/8/16/32.
 *   workspace: dagb200_best_alignment_workspace_bytes(B,M,L,T) bytes of device scratch.          */
size_t dagb200_best_alignment_workspace_bytes(int B, int M, int L, int T);
int dagb200_dag_best_alignment(const void *match, const void *links,
                               const int64_t *output_length, const int64_t *target_length,
                               void *alpha, int32_t *path, int dtype,
                               int B, int M, int L, int T, int config,
                               void *workspace, size_t workspace_bytes,
                               int32_t *status, void *stream);

/* GLAT pass: dagb200_logsoftmax_gather plus the arg-max over the vocabulary of every row (int64 [B][L], first index of
 * the maximum), which the criterion computes with a separate pass over the logits right before the gather
 * (criterions/nat_dag_loss.py:209 `pred_tokens = word_ins_out.argmax(-1)`, then :213).  Same arguments otherwise;
 * fp32 / fp16 / bf16 logits.                                                                                        */
int dagb200_logsoftmax_gather_argmax(void *logits, int dtype, const int64_t *select_idx, int64_t isb, int64_t isl,
                                     int64_t iss, void *out, int64_t osb, int64_t osl, int64_t oss,
                                     int64_t *argmax, int B, int L, int V, int S, int require_gradient, void *stream);

/* Next row of the path: the alignment posterior the S2S criterion derives from the two lattices,
 * replacing five torch ops (criterions/s2s_dag_fastspeech2_loss.py:259-261 with custom_ops/dag_loss.py:303-311):
 *   score[b][t][j] = exp(alpha + beta - logsumexp_j(alpha + beta)), 0 for rows without a finite cell (the reference's
 *   NaN -> 0), written in out_dtype (DAGB200_F32 / F16 / BF16 = the feature dtype, :261).  alpha / beta fp32 [B][M][L]. */
int dagb200_dag_posterior(const float *alpha, const float *beta, void *score, int out_dtype,
                          int B, int M, int L, void *stream);

/* Next row: GLAT force-emit masking of the emission plane between logsoftmax_gather and dag_loss,
 * replacing five torch ops (criterions/nat_dag_loss.py:130-132).  match / out fp32 [B][M][L]; matchmask bool [B][M][L]
 * (1 = the cell the glancing alignment chose), keep_word_mask bool [B][L] (1 = glanced vertex).
 *   backward == 0: out = glanced ? (matchmask ? match : -inf) : match
 *   backward != 0: out = glanced ? 0 : match   (match = incoming gradient; matchmask may be NULL)                      */
int dagb200_glat_force_emit(const float *match, const unsigned char *matchmask, const unsigned char *keep_word_mask,
                            float *out, int B, int M, int L, int backward, void *stream);

/* Next row (remainder): what the glancing pass derives from the Viterbi path, replacing a
 * [B][M+1][L] zero fill + scatter + slice, a gather and a masked reduction (criterions/nat_dag_loss.py:223-227):
 *   matchmask[b][t][j] = (path[b][j] == t)                      bool [B][M][L], written exactly once
 *   oracle[b][j]       = tgt_tokens[b][max(path[b][j], 0)]      int64 [B][L]        (nullable)
 *   path64[b][j]       = path[b][j]                             int64 [B][L]        (nullable; the wrapper's .to(long))
 *   align_mask[b][j]   = path[b][j] >= 0                        bool [B][L]         (nullable)
 *   same_num[b]        = #{j : path >= 0 and pred_tokens[b][j] == oracle[b][j]}   int64 [B]   (nullable, needs pred_tokens)
 * path is the int32 [B][L] output of dagb200_dag_best_alignment; tgt_tokens int64 addressed tgt[b*tsb + t*tss];
 * pred_tokens int64 [B][L] contiguous (the arg-max of the vocabulary logits, nat_dag_loss.py:209).                     */
int dagb200_glat_alignment(const int32_t *path, const int64_t *tgt_tokens, int64_t tsb, int64_t tss,
                           const int64_t *pred_tokens, unsigned char *matchmask, int64_t *oracle, int64_t *path64,
                           unsigned char *align_mask, int64_t *same_num, int B, int M, int L, void *stream);

/* Next row: inference decoding over the DAG, replacing the Python walks over `.tolist()`-ed
 * back-pointers of models/s2s_conformer_dag_fastspeech2.py:211-304.
 *   vertex_logit [B][L] f32  = max_y log P(y | v_j)   (:216 unreduced_logits)
 *   vertex_token [B][L] i64  = argmax_y               (:216 unreduced_tokens)
 *   links        [B][L][T]   banded transitions (the model's extract_links output)
 * Outputs: out_tokens int64 [B][L] (pad-filled), out_vertices int32 [B][L] (vertex of every emitted feature, -1 fill),
 * out_lengths int32 [B][2] = (number of tokens, number of features).
 * decode_lookahead: strategies "greedy" (beta = 0) and "lookahead" (:218-244); tokens start with the token of vertex 0
 * (<bos>), which carries no feature.
 * decode_viterbi_finish: strategies "viterbi" / "jointviterbi" (:245-304) AFTER the max-plus recurrence, which is the
 * lattice of dagb200_dag_best_alignment run on a length-independent emission plane (daspeech_b200/decode.py):
 * lattice [B][S+1][L] with row s+1 = the reference's scores[s], S = max_length; end transition, length penalty
 * (s+1)^viterbibeta, first-maximum length, backtrace (smallest source vertex on ties, as torch.max) and token
 * de-duplication.  path_scratch: int32 [B][S].                                                                        */
int dagb200_decode_lookahead(const float *links, const float *vertex_logit, const int64_t *vertex_token,
                             const int64_t *output_length, float beta, int64_t pad, int B, int L, int T,
                             int64_t *out_tokens, int32_t *out_vertices, int32_t *out_lengths, void *stream);
int dagb200_decode_viterbi_finish(const float *lattice, const float *links, const int64_t *vertex_token,
                                  const int64_t *output_length, float viterbibeta, int64_t pad, int B, int S, int L, int T,
                                  int64_t *out_tokens, int32_t *out_vertices, int32_t *out_lengths, int32_t *path_scratch,
                                  void *stream);

/* ---- Transition log-probabilities from the link heads -----------------------------------------
 * Forward of extract_links + extract_valid_links (DASpeech/models/s2t_conformer_dag.py:171-212, 140-155) for the banded
 * form (max_transition_length != -1):  links[b][i][k] = logsumexp_c( log_softmax_k( q[b,i,c,:].key[b,i+k+1,c,:] / sqrt(F) )
 * + log_gates[b,i,c] ), successors j = i+k+1 < output_length[b], -inf elsewhere.
 *   q, key: fp32 [B][L][H][F] (reshaped outputs of query_linear / key_linear); log_gates: fp32 [B][L][H];
 *   output_length: int64 [B] (non-pad positions); links: fp32 [B][L][T], every element written; workspace:
 *   dagb200_extract_links_workspace_bytes(B,L,H,F) bytes of device scratch (operands converted once into the MMA
 *   layout, per-head row normalisers).
 *   F a multiple of 16 in [16, 128], H <= 64.  tcgen05 (bf16 hi/lo split, fp32 accumulate); no [B,L,L,H] temporary. */
size_t dagb200_extract_links_workspace_bytes(int B, int L, int H, int F);
int dagb200_extract_links(const float *q, const float *key, const float *log_gates, const int64_t *output_length,
                          float *links, int B, int L, int H, int F, int T, void *workspace, size_t workspace_bytes,
                          void *stream);

/* ---- Data-parallel gradient exchange over NVLink peer memory (daspeech_b200/csrc/xchg.cu) ---------------------------
 * Replaces fairseq legacy_distributed_data_parallel.py:76-165 (one flat gradient buffer, divided by the world size,
 * all-reduced) as called from trainer.py:928.  One process per GPU of one node.  The bytes move on the copy engines
 * (peer-to-peer cudaMemcpyAsync), the sum of each 1/world slice runs as one short kernel on the rank that owns the
 * slice; all ranks end with bit-identical buffers.
 *   peer_alloc/free   : a cudaMalloc allocation (zero-filled) that can be exported; the library owns nothing else.
 *   peer_export/open  : 64-byte cudaIpcMemHandle of an allocation / map a peer's allocation into this process.
 *   grad_exchange_create(bufs, flags, stagings, numel, rank, world, &h), entry [rank] local, the others opened:
 *       bufs[p]  float[numel] of rank p, numel a multiple of 4;  flags[p] int32[32] of rank p, zero-initialised;
 *       stagings[p] float[(world-1) * grad_exchange_slice(numel, world)] of rank p.
 *   grad_exchange(h, stream): enqueue one exchange; stream-ordered, never blocks the host.
 *   grad_exchange_status(h, &epoch): 0, or the barrier epoch at which a peer failed to arrive within ~4 s
 *       (synchronises the device; diagnostic).                                                                        */
int dagb200_peer_alloc(size_t bytes, void **ptr);
int dagb200_peer_free(void *ptr);
int dagb200_peer_export(void *ptr, void *handle64);
int dagb200_peer_open(const void *handle64, void **ptr);
int dagb200_peer_close(void *ptr);
size_t dagb200_grad_exchange_slice(size_t numel, int world);
int dagb200_grad_exchange_create(void *const *bufs, void *const *flags, void *const *stagings, size_t numel, int rank,
                                 int world, void **handle);
int dagb200_grad_exchange(void *handle, void *stream);
int dagb200_grad_exchange_status(void *handle, int *timed_out_epoch);
int dagb200_grad_exchange_phases(void *handle, float *ms5);   /* measurement aid: push, barrier, reduce, push, barrier */
int dagb200_grad_exchange_destroy(void *handle);
/* In-switch form: mc = MULTICAST address of a buffer every rank has bound to the same multicast object (NVLink SHARP;
 * the caller maps it and synchronises the ranks before and after).  One kernel of `ctas` thread blocks: this rank's 1/world
 * slice of mc[0 .. numel) becomes (sum over ranks) / world in every rank's copy (multimem.ld_reduce / multimem.st). */
int dagb200_grad_exchange_nvls(void *mc, size_t numel, int rank, int world, int ctas, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DAGB200_H */
