/* mrl_b200.h - C ABI of the B200-native modular_rl policy-update path.
 *
 * The reference (ddlau/modular_rl) has no FFI: its hot path is three compiled Theano
 * functions per updater plus scipy scans, called from Python.  Each entry point below
 * replaces one of those call sites; citations are file:line under /root/reference.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mrl_last_error() gives the
 *     message for the calling thread.  No exceptions cross the ABI.  Soft failures of the
 *     algorithm (zero gradient, line-search failure) are reported as output flags, not errors.
 *   - `loc` says where a caller buffer lives: MRL_HOST (pageable or pinned) or MRL_DEVICE.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls enqueue on it;
 *     only functions that return scalars to the host synchronise it.
 *   - caller memory is borrowed for the duration of the call and never freed by the library.
 *   - one mrl_net / mrl_batch belongs to one (process, GPU).
 */
#ifndef MRL_B200_H
#define MRL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MRL_HOST 0
#define MRL_DEVICE 1
/* dtypes of caller arrays */
#define MRL_F32 0
#define MRL_F64 1
#define MRL_I32 2
#define MRL_I64 3
/* distribution heads: DiagGauss (core.py:402-438), Categorical (core.py:339-365), scalar value (agentzoo.py:53-59) */
#define MRL_GAUSS 0
#define MRL_CATEGORICAL 1
#define MRL_VALUE 2
/* hidden activations (Keras names, agentzoo.py:20-23) */
#define MRL_TANH 0
#define MRL_RELU 1
#define MRL_SIGMOID 2

typedef struct mrl_net mrl_net;
typedef struct mrl_batch mrl_batch;
typedef struct mrl_comm mrl_comm;

const char* mrl_last_error(void);
int mrl_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long mrl_launch_count(void);

/* measurement hooks (bench.py): per-kernel-class CUDA-event timing on the launching stream, and the FP32
 * FMA peak of the device.  mrl_profile_read synchronises the device; arrays have mrl_profile_kinds() entries. */
int mrl_profile_enable(int on);
int mrl_profile_kinds(void);
const char* mrl_profile_kind_name(int kind);
int mrl_profile_read(double* ms_out, long long* count_out);
int mrl_measure_fp32_tflops(int device, double* tflops_out);
/* measured mma.sync m16n8k8 TF32 throughput (dense TFLOP/s): the pipe of the register-chain kernels */
int mrl_measure_mma_tf32_tflops(int device, double* tflops_out);
/* measured tcgen05.mma kind::tf32 throughput (dense TFLOP/s, TS mode, M = 128, N = 128): the pipe of the layer-1 GEMMs
 * and of the Fisher-vector chain (mlp_l1_tc.cu, mlp_fvp_tc.cu); every algorithmic product costs three of these */
int mrl_measure_tcgen05_tf32_tflops(int device, double* tflops_out);

/* ---------------------------------------------------------------- batch (paths, flattened)
 * Replaces the `concat([path[k] for path in paths])` host copies of trpo.py:74-77,
 * ppo.py:61-64, core.py:653-654 and the per-call numpy->Theano downcast (keras_theano_setup.py:9). */
int mrl_batch_create(mrl_batch** out, int device, int ob_dim, int with_time_feature);
int mrl_batch_destroy(mrl_batch* b);
/* observations [N x ob_dim], row-major with leading dimension ld (elements) */
int mrl_batch_set_obs(mrl_batch* b, const void* ob, int dtype, long long ld, long long N, int loc, void* stream);
/* CSR trajectory structure: offsets int64[n_paths+1] (offsets[n_paths] == N), terminated uint8[n_paths].
 * Also materialises NnVf.preproc's feature t/timestep_limit (core.py:659-660) when the batch was
 * created with_time_feature. */
int mrl_batch_set_paths(mrl_batch* b, const long long* offsets, const unsigned char* terminated, int n_paths,
                        double timestep_limit, int loc, void* stream);
/* action [N] (int, Categorical) or [N x d] (float, DiagGauss); advantage [N]; oldprob [N x K] or [N x 2d]
 * = the args of trpo.py:66.  adv may be NULL to keep the advantages mrl_batch_gae left on the device. */
int mrl_batch_set_policy_inputs(mrl_batch* b, int head, int dout, const void* act, int act_dtype,
                                const void* adv, int adv_dtype, const void* oldprob, int oldprob_dtype,
                                int loc, void* stream);
/* regression target [N] for the value net, already mixed by the caller (core.py:624) */
int mrl_batch_set_vf_target(mrl_batch* b, const void* y, int dtype, int loc, void* stream);
/* target = mixfrac * return + (1 - mixfrac) * ypred_old (core.py:622-624) from what mrl_batch_gae and
 * mrl_net_predict_into_baseline left on the device (ypred_old == the GAE baseline: same theta) */
int mrl_batch_mix_vf_target(mrl_batch* b, double mixfrac, void* stream);
/* re-bind only the advantage column from the float32 advantages mrl_batch_gae left on the device */
int mrl_batch_refresh_advantages(mrl_batch* b, void* stream);
/* global timestep count when the batch is one shard of a data-parallel job (default: local N) */
int mrl_batch_set_global_n(mrl_batch* b, long long n_global);
long long mrl_batch_size(const mrl_batch* b);
/* within-path time index (int32[N]) computed on the device - the bit-exact integer contract */
int mrl_batch_get_time_index(mrl_batch* b, int* out, int loc, void* stream);

/* compute_advantage (core.py:63-105): reward [N]; baseline [N] (NULL: use the values left by
 * mrl_net_predict_into_baseline); writes return/advantage (float64, host or device, may be NULL) and keeps
 * float32 advantages on the device for the policy update.  standardize != 0 applies (adv-mean)/std with
 * population std and no epsilon (core.py:100-105); on a sharded batch the moments are merged over `comm`. */
int mrl_batch_gae(mrl_batch* b, const void* reward, int reward_dtype, const void* baseline, int baseline_dtype,
                  double gamma, double lam, int standardize, mrl_comm* comm, double* ret_out, double* adv_out,
                  int loc, void* stream);

/* ---------------------------------------------------------------- stand-alone scans */
/* misc_utils.discount over a CSR batch + GAE deltas, no batch object (core.py:69-75) */
int mrl_gae(const void* reward, int reward_dtype, const void* baseline, int baseline_dtype,
            const long long* offsets, const unsigned char* terminated, int n_paths, long long N, double gamma,
            double lam, double* ret_out, double* adv_out, int loc, void* stream);
/* in place (x-mean)/std, population std; stats_out[3] = {n, mean, M2} (may be NULL) */
int mrl_standardize(double* x, long long N, double* stats_out, int loc, void* stream);
/* ZFilter over N consecutive samples of dimension d (filters.py:30-38, running_stat.py:9-30): sample t is
 * normalised with running statistics that include samples 0..t and the incoming state {n, M[d], S[d]},
 * which is updated in place (host memory). */
int mrl_zfilter_scan(const void* x, int x_dtype, long long N, int d, double* state_n, double* state_M,
                     double* state_S, int demean, int destd, double clip, void* y, int y_dtype, int loc,
                     void* stream);

/* ---------------------------------------------------------------- network (policy or value MLP)
 * dims[0..n_layers] = input dim, hidden sizes..., output dim (agentzoo.py:25-61).  Parameters are the
 * reference's flat vector: per Dense layer [kernel(in,out) C-order, bias], then logstd (core.py:518-557). */
int mrl_net_create(mrl_net** out, int device, int n_layers, const int* dims, int head, int activation);
int mrl_net_destroy(mrl_net* n);
long long mrl_net_num_params(const mrl_net* n);
int mrl_net_set_params(mrl_net* n, const void* theta, int dtype, int loc, void* stream); /* SetFromFlat, core.py:526-540 */
int mrl_net_get_params(mrl_net* n, float* theta, int loc, void* stream);                  /* GetFlat, core.py:518-523 */
int mrl_net_set_comm(mrl_net* n, mrl_comm* comm);

/* net output for every timestep of the batch, row-major [N x dout]: means (Gauss; std = exp(logstd) is
 * state independent), softmax probabilities (Categorical) or values.  core.py:270, core.py:608 */
int mrl_net_forward(mrl_net* n, mrl_batch* b, float* out, int loc, void* stream);
/* NnVf.predict over the whole batch, result kept on the device as the GAE baseline (core.py:70) */
int mrl_net_predict_into_baseline(mrl_net* n, mrl_batch* b, void* stream);
/* compute_losses (trpo.py:69, ppo.py:57): out = {surr, kl, ent} */
int mrl_net_losses(mrl_net* n, mrl_batch* b, double out[3], void* stream);
/* compute_policy_gradient (trpo.py:68): g[P]; losses may be NULL */
int mrl_net_policy_gradient(mrl_net* n, mrl_batch* b, float* g, int loc, double losses[3], void* stream);
/* compute_fisher_vector_product (trpo.py:70), WITHOUT the cg_damping term (trpo.py:86-92 adds it) */
int mrl_net_fvp(mrl_net* n, mrl_batch* b, const float* v, float* out, int loc, void* stream);
/* compute_lossgrad (ppo.py:56): pensurr and its flat gradient; losses = {surr, kl, ent} */
int mrl_net_ppo_lossgrad(mrl_net* n, mrl_batch* b, double kl_coeff, double kl_cutoff, int reverse_kl,
                         double* pensurr, double* g, double losses[3], void* stream);
/* NnRegression f_losses / f_lossgrad (core.py:613-617,670-671): losses = {loss, mse, l2}; g may be NULL */
int mrl_net_vf_lossgrad(mrl_net* n, mrl_batch* b, double l2coeff, double losses[3], double* g, void* stream);

/* TrpoUpdater.__call__ after the path concat (trpo.py:80-140): gradient, CG (trpo.py:165-200), step scaling
 * (trpo.py:119-124), backtracking line search (trpo.py:143-159), parameter update or rollback.
 * stats = {surr_before, surr_after, kl_before, kl_after, ent_before, ent_after};
 * info  = {skipped (zero gradient, trpo.py:102), linesearch_success, accepted_backtrack_index,
 *          cg_iterations_run, fvp_calls, loss_passes} */
typedef struct {
  double cg_damping, max_kl, residual_tol, accept_ratio;
  int cg_iters, max_backtracks;
} mrl_trpo_cfg;
int mrl_net_trpo_step(mrl_net* n, mrl_batch* b, const mrl_trpo_cfg* cfg, double stats[6], int info[6],
                      void* stream);
/* debugging / tests: last step direction and full step (float64[P], host) */
int mrl_net_get_trpo_vectors(mrl_net* n, double* stepdir, double* fullstep, double* scalars /* shs,lm,rate,rdotr */);

/* ---------------------------------------------------------------- data-parallel communicator (NCCL)
 * The reference is single-process (core.py:123-124 raises on `parallel`).  Sharding the timestep batch
 * adds one sum over ranks after each batch reduction; see DESIGN.md "Multi-GPU". */
int mrl_comm_unique_id(char id_out[128]);
int mrl_comm_create(mrl_comm** out, const char id[128], int rank, int world, int device);
int mrl_comm_destroy(mrl_comm* c);
int mrl_comm_allreduce_f64(mrl_comm* c, double* buf, long long n, void* stream); /* in place, device memory */
/* Peer-memory transport over NVLink / NVSwitch (CUDA IPC) for vectors of <= max_doubles: every rank exports its
 * receive buffer, the caller all-gathers the 64-byte handles (any side channel; torch.distributed in this repo)
 * and every rank connects.  Afterwards the slab reduce PUSHES each rank's gradient / Fvp partial straight into
 * its peers' buffers and a second kernel sums the slots in rank order (bit-identical on all ranks); NCCL stays
 * the fallback for longer vectors.  All operations on one communicator must be enqueued on one stream. */
int mrl_comm_p2p_export(mrl_comm* c, long long max_doubles, char handle_out[64]);
int mrl_comm_p2p_connect(mrl_comm* c, const char* handles /* [world][64] */);
int mrl_comm_p2p_enable(mrl_comm* c, int on);   /* after every rank connected successfully */

/* ---------------------------------------------------------------- PpoSgdUpdater (ppo.py:115-258), minibatches on the device
 * mrl_batch_gather: dst <- rows idx[0..n) of src (observations + policy side inputs), a device gather from the
 *   resident batch (the reference slices numpy arrays per minibatch, ppo.py:194-199); idx lives where `loc` says.
 * mrl_net_ppo_sgd_step: one `train` call (ppo.py:162-167): penalised-surrogate loss + gradient on the minibatch and
 *   the Adam update of adam_updates (ppo.py:231-258), no host synchronisation; the losses evaluated before the step
 *   are added to a running sum that mrl_net_ppo_sgd_read returns as a mean (and resets). */
int mrl_batch_gather(mrl_batch* dst, const mrl_batch* src, const int* idx, int n, int loc, void* stream);
int mrl_net_ppo_sgd_step(mrl_net* net, mrl_batch* minibatch, double kl_coeff, double kl_cutoff, int reverse_kl,
                         double stepsize, double beta1, double beta2, double epsilon, void* stream);
int mrl_net_ppo_sgd_read(mrl_net* net, double losses[3], long long* count, void* stream);
int mrl_net_adam_reset(mrl_net* net, void* stream);

/* ---------------------------------------------------------------- population forward (cross-entropy method)
 * cem.py:43-44 scores every candidate theta of an iteration by one rollout, one candidate at a time; this entry point
 * evaluates the net of ALL candidates at once: member m has its own flat theta (Dense kernels and biases in the
 * reference's order, no logstd; row stride ld_theta) and its own observation; out[m] is the raw output of the linear
 * last layer (agentzoo.py:63-81).  thetas / obs / out live where `loc` says. */
int mrl_population_forward(int device, int n_layers, const int* dims, int activation, const float* thetas,
                           long long ld_theta, const float* obs, int n_members, float* out, int loc, void* stream);

#ifdef __cplusplus
}
#endif
#endif
