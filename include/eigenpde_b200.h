/*
 * eigenpde_b200.h -- C ABI of the B200 (sm_100a) loss+gradient step for EigenPDE-NN.
 *
 * The reference (zwpku/EigenPDE-NN) has no FFI/plugin interface: its hot path is the
 * Python method eigen_solver.update_step (src/pytorch_eigen_solver.py:191-236) with the
 * helpers fun_and_grad_on_data (:150-173) and penalty_term (:177-189).  This header is the
 * boundary a replacement library exports for that path; every entry point names the
 * reference lines it replaces.  Plain C: pointers and sizes only, no torch / C++ types.
 *
 * Conventions
 *   - All pointers named d_* are DEVICE pointers (CUDA global memory); h_* are host pointers.
 *   - Calls are asynchronous on `stream` (a cudaStream_t passed as void*); nothing in the
 *     library synchronises the host.  One in-flight call per workspace; calls on different workspaces may run
 *     concurrently from different threads (the library keeps no mutable global state).
 *   - Return value: 0 = ok; negative = EPDE_ERR_* (argument / unsupported configuration);
 *     positive = a cudaError_t reported by the runtime.  The library never aborts or throws.
 *   - The library allocates nothing persistent: the caller owns every buffer and queries the
 *     workspace size first.  No global mutable state; entry points are re-entrant.
 *   - Parameters are one flat fp32 vector in `model.parameters()` order of the reference's
 *     network_arch.MyNet (src/network_arch.py:25,90): for each net i = 0..k-1, for each
 *     layer l = 1..L: W_l row-major [n_l][n_{l-1}], then b_l [n_l].  Gradients use the same
 *     layout.
 *   - Samples: X is the device mirror of data_set.X_vec (src/data_set.py:21) as fp32 with a
 *     row stride of d_pad floats (d_pad % 4 == 0, X 16-byte aligned, padding = 0); `w` is
 *     data_set.weights (src/data_set.py:22) as fp32; `idx` are the minibatch row indices of
 *     data_set.generate_minibatch (src/data_set.py:67,77), or NULL for rows 0..B-1.
 */
#ifndef EIGENPDE_B200_H
#define EIGENPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPDE_ABI_VERSION 4      /* 2: epde_desc.metric, epde_md_align; 3: EPDE_FLAG_FFMA2 replaces epde_set_fused_impl,
                                   device-side minibatch generator; 4: epde_step_sharded, EPDE_FLAG_GENERIC,
                                   epde_gather_ahead / epde_step_ahead, epde_minibatch_indices_device_split */
#define EPDE_MAX_LAYERS 8   /* Linear layers per net */
#define EPDE_MAX_K 8        /* eigenfunctions (nets) */

/* activation ids: names accepted by the reference (src/network_arch.py:27-32, params.cfg:22) */
enum {
  EPDE_ACT_TANH = 0,
  EPDE_ACT_SIGMOID = 1,
  EPDE_ACT_SOFTPLUS = 2,
  EPDE_ACT_LOGSIGMOID = 3,
  EPDE_ACT_ELU = 4,       /* also CELU (alpha = 1), the reference's fallback */
  EPDE_ACT_RELU = 5,
  EPDE_ACT_TANHSHRINK = 6
};

enum {
  EPDE_OK = 0,
  EPDE_ERR_NULL = -1,         /* required pointer is NULL */
  EPDE_ERR_DESC = -2,         /* malformed descriptor (sizes out of range) */
  EPDE_ERR_UNSUPPORTED = -3,  /* valid descriptor the library has no kernel for */
  EPDE_ERR_WORKSPACE = -4,    /* workspace too small */
  EPDE_ERR_ALIGN = -5,        /* pointer / stride alignment */
  EPDE_ERR_ARG = -6           /* other invalid argument (B <= 0, ...) */
};

#define EPDE_FLAG_SORT 1u     /* sort_eigvals_in_training (src/pytorch_eigen_solver.py:217-221) */
#define EPDE_FLAG_DETERMINISTIC 2u  /* kept for ABI compatibility: both fused kernel families are bit-identical run to
                                     run unconditionally (fixed accumulation order); the generic path adds its fp64
                                     partials with atomics (order-dependent only at the 1e-16 level, invisible after
                                     the fp32 conversion) */
#define EPDE_FLAG_FFMA2 4u        /* kernel family of the fused path (reference architecture d -> 20 -> 20 -> 20 -> 1,
                                     Tanh): set = FP32 FFMA2 kernels (sample-pair threads, weights broadcast from shared
                                     memory); clear (default) = tcgen05 kernels (thread = TMEM lane = sample).  Both
                                     compute the same quantities to the same tolerance.  Must be the same in the
                                     epde_step_partial / epde_step_finish pair of a step.  With a per-row metric
                                     (epde_desc.metric, d <= 40) the fused path always runs the FFMA2 kernels. */
#define EPDE_FLAG_GENERIC 8u      /* never take the fused path: layer-wise generic family even for the reference
                                     architecture (testing / A-B measurements) */

typedef struct epde_desc {
  int32_t d;                            /* input dimension n_0 (data_set.tot_dim) */
  int32_t d_pad;                        /* row stride of X in floats */
  int32_t k;                            /* number of nets / eigenfunctions */
  int32_t n_layers;                     /* L = number of Linear layers (arch_list length - 1) */
  int32_t widths[EPDE_MAX_LAYERS + 1];  /* n_0 .. n_L, n_L must be 1 (src/pytorch_eigen_solver.py:352) */
  int32_t act_id;                       /* EPDE_ACT_* */
  uint32_t flags;                       /* EPDE_FLAG_* */
  const float* metric;                  /* NULL, or device pointer to a per-ROW symmetric metric [K][d][d] that replaces
                                           diag(diag_coeff) in the energy: e = grad f^T M grad f.  Used for the MD
                                           alignment layer (src/data_set.py:126-160): X then holds the ALIGNED rows and
                                           metric the J diag(c) J^T written by epde_md_align */
} epde_desc;

/* out_scalars layout (fp64, device): what update_step returns at :236 plus the moments */
#define EPDE_OUT_LOSS 0
#define EPDE_OUT_NONPEN 1
#define EPDE_OUT_PENALTY 2
#define EPDE_OUT_TOTW 3
#define EPDE_OUT_EIG(k) 4                 /* k sorted eigenvalue estimates */
#define EPDE_OUT_CVEC(k) (4 + (k))        /* k permutation entries (stored as doubles) */
#define EPDE_OUT_MEAN(k) (4 + 2 * (k))    /* k weighted means   (:172) */
#define EPDE_OUT_VAR(k) (4 + 3 * (k))     /* k weighted variances (:173) */
#define EPDE_OUT_COUNT(k) (4 + 4 * (k))

int epde_abi_version(void);
const char* epde_error_string(int code);

/* number of parameters of all k nets (sum of p.numel(), src/pytorch_eigen_solver.py:382) */
int epde_param_count(const epde_desc* desc, int64_t* out);

/* number of batch-global partial sums exchanged between the two phases:
 * [W, S1(k), E(k), S2(k(k+1)/2, row-major lower triangle)] -- the fp64 quantities behind
 * src/pytorch_eigen_solver.py:169-173,215 and :186 */
int epde_num_sums(const epde_desc* desc, int32_t* out);

/* 1 if fused kernels cover this descriptor, 0 if the generic layer-wise kernels are used; negative on a malformed
 * descriptor.  Fused: d -> 20 -> 20 -> 20 -> 1 with d <= 60 and k <= 6; Tanh on the tcgen05 kernels (or the FFMA2 kernels with
 * EPDE_FLAG_FFMA2 / a per-row metric with d <= 40), Softplus / LogSigmoid / ELU (= CELU) / ReLU on the FFMA2 kernels
 * (src/network_arch.py:27-34 takes any torch.nn activation); everything else - other widths or depths, Sigmoid, Tanhshrink,
 * EPDE_FLAG_GENERIC - runs layer-wise. */
int epde_has_fused_path(const epde_desc* desc);

int epde_workspace_bytes(const epde_desc* desc, int64_t B, size_t* out);

/*
 * Phase 1 of update_step: forward of the k nets, input gradients and the batch sums.
 * Replaces fun_and_grad_on_data (src/pytorch_eigen_solver.py:150-173: model forward :157,
 * autograd.grad :164, weights :166-169) and the reductions of :215 and :186 for the
 * samples of THIS rank.  Writes epde_num_sums() doubles to d_sums.
 */
int epde_step_partial(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx,
                      int64_t B, const float* d_params, const float* d_diag_coeff,
                      void* d_workspace, size_t workspace_bytes, double* d_sums, void* stream);

/*
 * Phase 2: given the batch-global sums (d_sums, after an allreduce when the batch is
 * sharded), form eig_vals / cvec / non_penalty_loss / penalty / loss
 * (src/pytorch_eigen_solver.py:215-229, 177-189) and back-propagate into the parameters
 * (loss.backward(), :233) for the samples of THIS rank.  d_grad receives this rank's
 * additive share of d loss / d theta (sum over ranks = full gradient).
 * h_eig_w: k host doubles (Param.eig_w, only the first k are used, :224).
 * Must be called with the same X/w/idx/B/params/workspace as the preceding epde_step_partial.
 */
int epde_step_finish(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx,
                     int64_t B, const float* d_params, const float* d_diag_coeff,
                     double beta, double alpha, const double* h_eig_w, const double* d_sums,
                     void* d_workspace, size_t workspace_bytes,
                     double* d_out_scalars, float* d_grad, void* stream);

/* Single-GPU convenience: epde_step_partial + epde_step_finish with no exchange in between.
 * Replaces the whole of update_step up to and including loss.backward() (:212-233). */
int epde_step(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx,
              int64_t B, const float* d_params, const float* d_diag_coeff,
              double beta, double alpha, const double* h_eig_w,
              void* d_workspace, size_t workspace_bytes,
              double* d_out_scalars, float* d_grad, void* stream);

/*
 * Forward only: y[B][k] = MyNet.forward on the selected rows (src/network_arch.py:108-117),
 * used by the stage-end re-normalisation (src/pytorch_eigen_solver.py:120-141) and by
 * utils/eval_nn_on_*.py.  d_y is fp32 [B][k] row-major.  The samples are processed in slabs of at
 * most 2^18, so the workspace has to be epde_workspace_bytes(desc, min(B, 262144)) large.
 */
int epde_forward(const epde_desc* desc, const float* d_X, const int64_t* d_idx, int64_t B,
                 const float* d_params, const float* d_diag_coeff, void* d_workspace, size_t workspace_bytes,
                 float* d_y, void* stream);

/*
 * Host-side, bit-exact minibatch index stream (src/data_set.py:67 with the cum_weights of
 * :31): B draws of min(floor(u*K), K-1) from CPython's MT19937, continuing from the
 * generator state (`h_mt` 624 words + position `*h_pos`, i.e. random.getstate()[1]) and
 * leaving the advanced state behind, so Python's `random` can resume from it.
 */
int epde_minibatch_indices(uint32_t* h_mt, int32_t* h_pos, int64_t K, int64_t B, int64_t* h_idx_out);

/*
 * The same stream generated ON THE DEVICE (src/data_set.py:31,67,77): d_state holds the generator
 * as 625 uint32 (624 state words + the position, i.e. random.getstate()[1]); the call consumes 2 * B_total words,
 * writes the indices of draws [out_lo, out_lo + out_n) to d_idx[0 .. out_n) -- a rank of a sharded step passes its slice
 * of the global batch, every rank advances the whole stream -- and leaves the advanced state in d_state exactly as CPython
 * would show it after the same draws.  d_raw: scratch of 2 * B_total uint32 (8-byte aligned).  Asynchronous on `stream`;
 * the generator itself is one CTA (the recurrence is sequential: 623 words per synchronisation, about 0.2 ms per 2^20
 * draws), so callers overlap it with the previous step on a second stream.
 */
int epde_minibatch_indices_device(uint32_t* d_state, int64_t K, int64_t B_total, int64_t out_lo, int64_t out_n,
                                  uint32_t* d_raw, int64_t* d_idx, void* stream);
/*
 * The sequential recurrence on several SMs at once: d_states holds `parts` generators of 625 uint32 each, generator q
 * positioned (epde_mt_jump_device) at draw q * n_sub of a run of n_total draws; CTA q generates the draws
 * [q n_sub, min((q + 1) n_sub, n_total)) and all n_total indices go to d_idx.  (parts - 1) * n_sub < n_total <= parts * n_sub.
 * d_raw: scratch of 2 * n_total uint32.  The states are left advanced past their own parts (of no further use: the caller
 * takes the next step's generator from another jump).  Same stream (src/data_set.py:31,67,77), bit-exact.
 */
int epde_minibatch_indices_device_split(uint32_t* d_states, int32_t parts, int64_t K, int64_t n_sub, int64_t n_total,
                                        uint32_t* d_raw, int64_t* d_idx, void* stream);

/*
 * Jump-ahead for the sharded stream (csrc/epde_mt_jump.cu).  MT19937 is linear over GF(2): the generator J words ahead is a
 * fixed XOR-combination, given by g_J(x) = x^J mod (characteristic polynomial), of the next 20 560 words.
 *   epde_mt_jump_poly: host, g_J as 624 uint32 (about 0.2 s; depends only on J -- compute once per slice offset / batch size).
 *   epde_mt_jump_device: from the generator d_state_in (625 uint32) produce, for each of the n_polys <= 32 polynomials in
 *     d_polys (n_polys x 624 uint32, device), the generator advanced by that jump into h_state_out[k] (host array of device
 *     pointers, 625 uint32 each: the 624-word window at the target position, position word 0 -- the same stream position in
 *     a different representation than CPython's (block, position) pair).  d_window: scratch of 20 560 uint32.
 * A rank of a sharded step jumps to its slice (2 * lo words) and to the next step (2 * B words) and generates only its own
 * draws with epde_minibatch_indices_device.
 */
int epde_mt_jump_poly(int64_t J, uint32_t* h_poly624);
int epde_mt_jump_device(const uint32_t* d_state_in, const uint32_t* d_polys, int32_t n_polys, uint32_t* const* h_state_out,
                        uint32_t* d_window, void* stream);

/* CPython random.seed(int) -> MT19937 state (init_by_array over the 32-bit words of |seed|);
 * `h_key` holds the little-endian 32-bit words of the absolute seed. */
int epde_mt_seed(const uint32_t* h_key, int32_t key_len, uint32_t* h_mt, int32_t* h_pos);

/*
 * Fused Adam over the flat parameter vector with torch.optim.Adam defaults
 * (betas 0.9/0.999, eps 1e-8, no weight decay; src/pytorch_eigen_solver.py:234,379),
 * all in fp32 on the device.  `step` is the 1-based step count after this update.
 */
int epde_adam_step(float* d_params, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                   int64_t n, double lr, int64_t step, void* stream);

/*
 * Device-side pieces of the per-step bookkeeping of `train` (src/pytorch_eigen_solver.py:280-284), so a
 * training loop can run without a host round trip per step:
 *   epde_average_accumulate: d_avg[net i] += d_params[net cvec[i]]   (record_model_parameters, :114-118;
 *                            cvec is read from d_out_scalars on the device), d_avg is fp64 [param_count]
 *   epde_eig_stats_accumulate: d_stats[0..k) += eig, d_stats[k..2k) += eig^2   (:280-282)
 */
int epde_average_accumulate(const epde_desc* desc, double* d_avg, const float* d_params,
                            const double* d_out_scalars, void* stream);
int epde_eig_stats_accumulate(const epde_desc* desc, double* d_stats, const double* d_out_scalars, void* stream);
/*
 * epde_adam_step + epde_average_accumulate + epde_eig_stats_accumulate in ONE launch with launch-invariant arguments
 * (CUDA-graph replay of a whole training step, :277-284): d_state[0] = learning rate, d_state[1] = number of optimiser
 * steps taken so far (fp64, incremented by the call).
 */
int epde_train_epilogue(const epde_desc* desc, float* d_params, const float* d_grad, float* d_m, float* d_v,
                        double* d_state, double* d_avg, double* d_stats, const double* d_out_scalars, void* stream);
/*
 * The same with fp64 master parameters and fp64 Adam moments -- what the reference holds (`model.double()` :374-376,
 * torch.optim.Adam :379): d_params64 is updated in fp64 with torch's update order, d_params32 receives the fp32 copy
 * the next step's kernels read, d_avg accumulates the fp64 parameters in cvec order (:114-118).
 */
int epde_train_epilogue64(const epde_desc* desc, double* d_params64, float* d_params32, const float* d_grad,
                          double* d_m, double* d_v, double* d_state, double* d_avg, double* d_stats,
                          const double* d_out_scalars, void* stream);

/*
 * One-shot all-reduce (sum) of a small vector over NVLink peer memory, for the two exchanges of the sharded step
 * (the 1 + 2k + k(k+1)/2 fp64 partial sums between the phases, the fp32 gradient after phase 2).
 * d_peer_bufs: DEVICE array of `world` pointers, entry p = rank p's symmetric buffer mapped into this process
 * (e.g. torch.distributed._symmetric_memory: rendezvous(...).buffer_ptrs_dev); every buffer is
 * 256 + 2 * world * slot_bytes bytes and zero-filled before the first call.  epoch = 1, 2, 3, ... must advance by one
 * per call on every rank; world <= 32 (the 256-byte flag block holds 2 parities x 32 flags).  In place (src == dst) is allowed.  A peer that never arrives traps after about a minute.
 */
typedef struct epde_peer {
  void* const* d_peer_bufs;   /* DEVICE array of `world` pointers to the ranks' symmetric buffers (see below) */
  int32_t rank, world;        /* world <= 32 */
  uint32_t* d_epoch;          /* device counter, zero before the first call; shared with epde_peer_allreduce_graph calls */
  int64_t slot_bytes;         /* per rank and parity; must hold the gradient (4 * epde_param_count bytes) */
} epde_peer;
/*
 * The whole sharded step in one call ("issued from the kernel epilogue by the last-finishing CTA"): as
 * epde_step on this rank's slice of the global minibatch, with the two exchanges done INSIDE the step's own kernels over
 * NVLink peer memory - the fp64 sums by the single CTA that reduces the phase-1 partials and computes the coefficients,
 * the gradient by the last CTA of the final reduction kernel to finish.  No exchange launch, no host involvement; on
 * return (stream order) d_out_scalars / d_grad hold the results for the union batch, bit-identical on all ranks.
 * Tensor-core family only: EPDE_ERR_UNSUPPORTED otherwise (use epde_step_partial / allreduce / epde_step_finish).
 * Replaces src/pytorch_eigen_solver.py:191-233 on each rank of a data-parallel run.
 */
int epde_step_sharded(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx, int64_t B,
                      const float* d_params, const float* d_diag, double beta, double alpha, const double* h_eig_w,
                      void* d_ws, size_t ws_bytes, double* d_out_scalars, float* d_grad, const epde_peer* peer,
                      void* stream);

/*
 * Gather look-ahead (tensor-core family; EPDE_ERR_UNSUPPORTED otherwise).  The minibatch of step i + 1 does not depend on
 * the result of step i (src/data_set.py:67), and the gather of its rows X[idx[n]][:] is the one HBM-bound kernel of the
 * step, while the two long kernels of step i are not.  epde_gather_ahead, called on a SECOND (lower-priority) stream while
 * step i runs, fills set `slot` (0 / 1) of the workspace's gathered-row buffers; epde_step_ahead is epde_step (peer == NULL)
 * or epde_step_sharded on set `slot`, skipping its own gather if `pregathered`, and keeping `reserve_sms` SMs free of its
 * persistent phase-2 kernel so that the concurrent gather for step i + 1 has somewhere to run.  Phase 1 is bound by HBM
 * itself (it writes the stash), so the gather should start behind it: event_phase1 (a cudaEvent_t, or NULL) is recorded on
 * `stream` between the phases for the gather's stream to wait on.  The caller orders the two streams with events (set
 * `slot` is read by step i until its last kernel, and written by the gather for step i + 2) and alternates the slots.
 */
int epde_gather_ahead(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx, int64_t B,
                      void* d_ws, size_t ws_bytes, int32_t slot, void* stream);
int epde_step_ahead(const epde_desc* desc, const float* d_X, const float* d_w, const int64_t* d_idx, int64_t B,
                    const float* d_params, const float* d_diag, double beta, double alpha, const double* h_eig_w,
                    void* d_ws, size_t ws_bytes, double* d_out_scalars, float* d_grad, int32_t slot, int32_t pregathered,
                    int32_t reserve_sms, const epde_peer* peer, void* event_phase1, void* stream);

int epde_peer_allreduce(void* const* d_peer_bufs, int rank, int world, uint32_t epoch, const void* src, void* dst,
                        int64_t n, int is_f64, int64_t slot_bytes, void* stream);
/* The same with the epoch kept in device memory (*d_epoch, zero before the first call, advanced by one per call): the
 * launch arguments never change, so the exchange can be captured into a CUDA graph with the rest of the step. */
int epde_peer_allreduce_graph(void* const* d_peer_bufs, int rank, int world, uint32_t* d_epoch, const void* src, void* dst,
                              int64_t n, int is_f64, int64_t slot_bytes, void* stream);

/*
 * MD pre-processing layer `MD_data_set.align` (src/data_set.py:126-153; called from every net's forward,
 * src/network_arch.py:62): Kabsch alignment of each configuration (n_atoms x 3 coordinates per row of d_X) to the
 * reference configuration ref_x[m][3] over the atoms ref_index[m] (file format: load_ref_state :112-124).
 * The layer has no parameters, so it is evaluated once per data set: d_Xa receives the aligned rows (same layout
 * as d_X) and d_metric, if not NULL, the per-row metric M = J diag(d_diag) J^T [K][d][d] of the layer's Jacobian,
 * with which the step reproduces the reference's input gradients through the layer (epde_desc.metric).
 * ref_x / ref_index are HOST arrays; m <= 21, n_atoms <= 21.
 */
int epde_md_align(int64_t K, int32_t n_atoms, int32_t d_pad, const float* d_X, const double* ref_x,
                  const int32_t* ref_index, int32_t m, const float* d_diag, float* d_Xa, float* d_metric, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EIGENPDE_B200_H */
