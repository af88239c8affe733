/*
 * xtb200.h -- C-ABI of libxtb200.so, the B200 (sm_100a) engine behind XingTian's
 * Model / Algorithm plugin API.
 *
 * The reference (huawei-noah/xingtian, /root/reference) has NO native boundary:
 * its hot path is Python composing TensorFlow-1.15 ops.  Each entry point below
 * names the reference Python interface it replaces (file:line, relative to the
 * reference root).  All pointers are plain device pointers unless the name ends
 * in `_host`; sizes are element counts unless they say bytes; `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Storage is
 * caller-owned (the Python host allocates it as torch.Tensor storage); the
 * library never frees caller memory.  Every function returns 0 on success and a
 * negative xtb_status otherwise; xtb_last_error() gives the thread-local message
 * (the Python shim raises RuntimeError with it -- reference behaviour: exceptions
 * propagate, xt/train.py:159-178).
 */
#ifndef XTB200_H_
#define XTB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XTB_VERSION 100

enum xtb_status {
  XTB_OK = 0,
  XTB_ERR_ARG = -1,     /* bad argument / unsupported configuration */
  XTB_ERR_CUDA = -2,    /* a CUDA runtime call or kernel launch failed */
  XTB_ERR_STATE = -3,   /* object not bound / wrong call order */
  XTB_ERR_NOMEM = -4
};

enum xtb_layer_kind { XTB_CONV = 0, XTB_DENSE = 1 };
enum xtb_act { XTB_ACT_NONE = 0, XTB_ACT_RELU = 1, XTB_ACT_TANH = 2 };

#define XTB_MAX_LAYERS 16

/* One Keras layer of the reference networks.
 * conv : Conv2D(cout, (k,k), strides=(stride,stride), padding = pad_same ? 'same' : 'valid')
 *        NHWC activations, HWIO kernel  (xt/model/model_utils.py:91-97,
 *        xt/model/impala/impala_cnn_opt.py:120-137, xt/model/dqn/dqn_cnn.py:49-51)
 * dense: Dense(n) on the HWC-flattened source (model_utils.py:83-88, :63-65) */
typedef struct xtb_layer_desc {
  int32_t kind;      /* xtb_layer_kind */
  int32_t src;       /* tensor id of the input: 0 = observation, i+1 = output of layer i */
  int32_t act;       /* xtb_act */
  int32_t k;         /* conv kernel size */
  int32_t stride;    /* conv stride (1, 2 or 4) */
  int32_t cout;      /* conv filters / dense units */
  int32_t pad_same;  /* conv: 1 = TF 'SAME' padding, 0 = 'VALID' */
} xtb_layer_desc;

typedef struct xtb_net_desc {
  int32_t input_u8;        /* 1: uint8 observation decoded as x*scale (model_utils.py:187-189) */
  float scale;             /* 1/255 for uint8 Atari frames, 1 for float observations */
  int32_t in_h, in_w, in_c; /* observation HWC; vectors use 1,1,dim */
  int32_t n_layers;
  xtb_layer_desc layers[XTB_MAX_LAYERS];
} xtb_net_desc;

typedef struct xtb_net xtb_net;

/* ---- library ---------------------------------------------------------------- */
int xtb_version(void);
const char* xtb_last_error(void);
/* Number of kernels this library has launched in the calling process (all threads). */
long long xtb_launch_count(void);
/* Number of CUDA-graph replays (fused training loop / rollout inference) the library has issued. */
long long xtb_graph_replay_count(void);

/* ---- network: replaces XTModel's TF graph (xt/model/model.py:30-127) ---------- */
/* Flat fp32 parameter layout: per layer, kernel [K,N] (HWIO flattened) then bias [N];
 * identical to iterating TFVariables' ordered dict (xt/model/tf_utils.py:84-102). */
int xtb_net_create(const xtb_net_desc* desc, int max_batch, xtb_net** out);
void xtb_net_destroy(xtb_net* net);
long long xtb_net_param_count(const xtb_net* net);
/* offsets (in floats) of layer `layer`'s kernel and bias inside the flat buffer, and K,N */
int xtb_net_layer_params(const xtb_net* net, int layer, long long* kernel_off, long long* bias_off,
                         int* k_rows, int* n_cols);
/* floats per sample of tensor `t` (0 = observation) */
int xtb_net_tensor_size(const xtb_net* net, int t);
size_t xtb_net_workspace_bytes(const xtb_net* net);
/* params/grads: [param_count] floats; workspace: xtb_net_workspace_bytes() bytes */
int xtb_net_bind(xtb_net* net, float* params, float* grads, void* workspace, size_t workspace_bytes);
/* same, with the initialisation work (workspace clear, weight blobs) ordered on `stream`; returns after it completed */
int xtb_net_bind_stream(xtb_net* net, float* params, float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* Every tensor the tensor-core kernels read is kept as two bf16 planes (hi = bf16(x), lo = bf16(x-hi)) next
 * to its fp32 copy.  The planes of the bound parameters are refreshed by xtb_net_bind, by the fused training
 * loops after each optimiser step, and by this call -- which a host that writes the parameter buffer itself
 * (set_weights, target-network sync, an external optimiser step) must make before the next forward. */
int xtb_net_sync_weights(xtb_net* net, void* stream);
/* activation / activation-gradient buffer of tensor t (t>=1), [batch, tensor_size] floats */
float* xtb_net_tensor(xtb_net* net, int t);
float* xtb_net_tensor_grad(xtb_net* net, int t);
/* Forward over `batch` samples. obs: [rows, H,W,C] uint8 or float; if gather_idx != NULL sample b
 * reads observation row gather_idx[b] (the minibatch gather of xt/model/ppo/ppo.py:123-128 folded
 * into the first layer's loads).  `params` NULL = the bound parameters (else e.g. a target net,
 * xt/algorithm/dqn/dqn.py:57). */
int xtb_net_forward(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                    int batch, void* stream);
/* Backward: consumes d(loss)/d(pre-activation) already written into xtb_net_tensor_grad() of the
 * head tensors, zeroes and fills the bound grads buffer.  Replaces
 * optimizer.compute_gradients (xt/model/ppo/ppo.py:99). */
int xtb_net_backward(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                     const int32_t* head_tensors, int n_heads, void* stream);

/* ---- policy head: replaces CategoricalDist (xt/model/tf_dist.py:89-130) and
 *      PPO.predict's fetches (xt/model/ppo/ppo.py:104-109) ---------------------- */
/* action = argmax(logits - log(-log u)); u from `uniforms` [batch,adim] if non-NULL, else
 * Philox4x32-10(seed, offset).  logp = log-softmax(logits)[action]. */
int xtb_categorical_sample(const float* logits, int batch, int adim, const float* uniforms,
                           uint64_t seed, uint64_t offset, int32_t* action, float* logp, void* stream);
/* argmax over the last axis (DQN greedy action, xt/algorithm/algorithm.py:124-135) */
int xtb_argmax(const float* q, int batch, int adim, int32_t* action, void* stream);

/* ---- GAE: replaces PPO.data_proc (xt/agent/ppo/ppo.py:77-106) ------------------- */
/* value [E,T+1], reward [E,T], done [E,T] (uint8) -> adv, old_value, target_value [E,T].
 * sign_clip != 0 applies np.sign to rewards (xt/agent/ppo/atari_ppo.py:46). */
int xtb_gae(const float* value, const float* reward, const uint8_t* done, int n_env, int n_step,
            float gamma, float lam, int sign_clip, float* adv, float* old_value, float* target_value,
            void* stream);

/* ---- PPO loss: replaces actor_loss_with_entropy + critic_loss
 *      (xt/model/ppo/__init__.py:4-25, combined xt/model/ppo/ppo.py:87-92) --------- */
typedef struct xtb_ppo_hyper {
  float clip_ratio;   /* LOSS_CLIPPING */
  float ent_coef;     /* ENTROPY_LOSS */
  float vf_clip;      /* VF_CLIP */
  float critic_coef;  /* CRITIC_LOSS_COEF */
} xtb_ppo_hyper;
/* logits [B,A], v [B]; rollout arrays are indexed through gather_idx (NULL = identity).
 * inv_count = 1/B_global (mean over the *global* minibatch when sharded).
 * Writes dlogits [B,A], dv [B]; atomically adds this minibatch's loss to *loss_out. */
int xtb_ppo_loss_grad(const float* logits, const float* v, const int32_t* gather_idx,
                      const int32_t* action, const float* old_logp, const float* adv,
                      const float* old_v, const float* target_v, int batch, int adim,
                      const xtb_ppo_hyper* hp, float inv_count, float* dlogits, float* dv,
                      float* loss_out, void* stream);

/* ---- V-trace: replaces vtrace.from_logic_outputs + vtrace_loss
 *      (xt/model/impala/vtrace.py:39-115, impala_cnn_opt.py:188-196, :299-351) ------ */
/* Flat env-major inputs [n_traj*step_len,...]; the last step of every trajectory is dropped
 * from the loss and its baseline is the bootstrap (split_batches drop_last).  Writes
 * dlogits [N,A], dbaseline [N] (zeros on dropped rows), optional vs/pg_adv [N], adds loss. */
int xtb_vtrace_loss_grad(const float* tp_logits, const float* baseline, const float* bp_logits,
                         const int32_t* action, const uint8_t* done, const float* reward,
                         int n_traj, int step_len, int adim, float gamma, float* dlogits,
                         float* dbaseline, float* vs_out, float* pg_adv_out, float* loss_out,
                         void* stream);

/* ---- DQN TD target + MSE: replaces DQN.train's target loop and Keras 'mse'
 *      (xt/algorithm/dqn/dqn.py:79-97, xt/model/dqn/dqn_cnn.py:60-61) ---------------- */
/* q [B,A] online Q(s); q_next_target [B,A]; q_next_online NULL or [B,A] (double DQN).
 * Writes dq [B,A] = 2/(B*A)*(q[b,a]-y) on the taken action, td target y [B] (optional), adds loss. */
int xtb_dqn_loss_grad(const float* q, const float* q_next_target, const float* q_next_online,
                      const int32_t* action, const float* reward, const uint8_t* done, int batch,
                      int adim, float gamma, float inv_count, float* dq, float* y_out,
                      float* loss_out, void* stream);

/* Same target/loss with the options BASELINE.json's north_star names (defaults = the reference): rows may be indexed
 * through idx (minibatch rows of a replay ring); disc != NULL is a per-row bootstrap discount (gamma^n of an n-step
 * return, 0 = the window hit a terminal step); huber_delta > 0 selects the Huber loss instead of the squared error. */
int xtb_dqn_td_loss_grad(const float* q, const float* q_next_target, const float* q_next_online, const int32_t* idx,
                         const int32_t* action, const float* reward, const uint8_t* done, const float* disc, int batch,
                         int adim, float gamma, float huber_delta, float inv_count, float* dq, float* y_out,
                         float* loss_out, void* stream);
/* n-step returns over env-major trajectories [n_env][n_step] (north_star "n-step TD-target kernel"; the reference's
 * DQN is 1-step, xt/algorithm/dqn/dqn.py:86-97): ret = sum_{k<m} gamma^k r_{t+k}, m = steps to the first terminal
 * (inclusive), n, or the end of the segment; disc = gamma^m or 0 after a terminal; last = row whose next-state
 * bootstraps; done_n = window contains a terminal. */
int xtb_nstep_returns(const float* reward, const uint8_t* done, int n_env, int n_step, int n, float gamma, float* ret,
                      float* disc, int32_t* last, uint8_t* done_n, void* stream);

/* Keras model.train_on_batch(states, y) with loss='mse' (xt/model/model.py:77-82,
 * xt/model/dqn/dqn_cnn.py:60-61): loss = mean over B*A of (q-y)^2, dq = 2*(q-y)*inv_count. */
int xtb_mse_loss_grad(const float* q, const float* y, int batch, int adim, float inv_count, float* dq,
                      float* loss_out, void* stream);

/* ---- optimiser: replaces tf.train.AdamOptimizer + clip_by_global_norm
 *      (xt/model/ppo/ppo.py:97-102, impala_cnn_opt.py:198-217) and Keras
 *      Adam(clipnorm) (xt/model/dqn/dqn_cnn.py:60) ----------------------------------- */
typedef struct xtb_adam xtb_adam;
enum xtb_clip_mode { XTB_CLIP_NONE = 0, XTB_CLIP_GLOBAL_NORM = 1, XTB_CLIP_PER_TENSOR = 2 };
/* m, v: [count] floats (caller-owned, zero-initialised by this call); seg_offsets: n_seg+1
 * boundaries of the tensors inside the flat buffer (used by XTB_CLIP_PER_TENSOR). */
int xtb_adam_create(long long count, float lr, float beta1, float beta2, float eps, int clip_mode,
                    float clip, const long long* seg_offsets, int n_seg, float* m, float* v,
                    xtb_adam** out);
void xtb_adam_destroy(xtb_adam* opt);
/* grad_scale multiplies the gradient before clipping (1 normally). After the call
 * *xtb_adam_grad_norm() holds the pre-clip global norm (device float). */
int xtb_adam_step(xtb_adam* opt, float* params, const float* grads, float grad_scale, void* stream);
/* Same step on a bound network's parameters/gradients; the kernel also refreshes the weights' bf16 planes. */
int xtb_adam_step_net(xtb_adam* opt, xtb_net* net, float grad_scale, void* stream);
const float* xtb_adam_grad_norm(const xtb_adam* opt);
int xtb_adam_set_lr(xtb_adam* opt, float lr);
/* Switch the optimiser handle to tf.train.RMSPropOptimizer(lr, decay, epsilon, centered=True) (momentum 0), the
 * `opt_type: rmsprop` branch of xt/model/impala/impala_cnn_opt.py:205-206: the `m` buffer of xtb_adam_create becomes the
 * mean-square slot and `mean_grad` (count floats) the mean-gradient slot; this call sets them to ones / zeros as TF
 * initialises them; `v` is unused.  Clipping, chunking and the weight-blob refresh are those of the Adam step. */
int xtb_opt_use_rmsprop(xtb_adam* opt, float* mean_grad, float decay, float epsilon);

/* ---- fused learner loops -------------------------------------------------------- */
/* PPO.train (xt/model/ppo/ppo.py:111-132): for every minibatch slice of `perm`
 * (device int32 [n_epoch*n_sample], the host-generated np.random.shuffle order) run
 * forward, loss, backward, clip, Adam.  loss_per_step: device float [n_epoch*ceil(N/B)].
 * All launches go to `stream` (capturable into a CUDA graph; the library caches one per
 * (n_sample, batch) when use_graph != 0). */
typedef struct xtb_ppo_rollout {
  const void* obs;            /* [N, H,W,C] uint8 / float */
  const int32_t* action;      /* [N] */
  const float* old_logp;      /* [N] */
  const float* adv;           /* [N] */
  const float* old_v;         /* [N] */
  const float* target_v;      /* [N] */
} xtb_ppo_rollout;
int xtb_ppo_train(xtb_net* net, xtb_adam* opt, const xtb_ppo_rollout* ro, int n_sample,
                  int batch_size, int n_epoch, const int32_t* perm, const xtb_ppo_hyper* hp,
                  int pi_tensor, int v_tensor, float* loss_per_step, int use_graph, void* stream);

/* ImpalaCnnOpt.train (xt/model/impala/impala_cnn_opt.py:251-265) as one captured step: forward over n_sample =
 * k*step_len env-major rows (rows gather_idx[b] of obs when non-NULL), in-graph V-trace + summed losses, backward,
 * clip_by_global_norm + Adam.  *loss_out += loss (device float). */
int xtb_impala_train(xtb_net* net, xtb_adam* opt, const void* obs, const int32_t* gather_idx, const float* bp_logits,
                     const int32_t* action, const uint8_t* done, const float* reward, int n_sample, int step_len,
                     float gamma, int logit_tensor, int base_tensor, float* loss_out, int use_graph, void* stream);
/* DQN.train (xt/algorithm/dqn/dqn.py:61-103) as one captured step over rows idx[0..n_sample) of a device replay ring:
 * target forward on next_obs, optional double-DQN online forward (qn_o != NULL), online forward on obs, TD target +
 * loss gradient (see xtb_dqn_td_loss_grad for disc / huber_delta), backward, per-tensor clip + Adam.
 * qn_t, qn_o: scratch [n_sample, adim].  *loss_out += loss. */
int xtb_dqn_train(xtb_net* net, xtb_net* target, xtb_adam* opt, const void* obs, const void* next_obs,
                  const int32_t* idx, const int32_t* action, const float* reward, const uint8_t* done,
                  const float* disc, int n_sample, float gamma, float huber_delta, int q_tensor, float* qn_t,
                  float* qn_o, float* loss_out, int use_graph, void* stream);

/* Rollout inference: for t in [0,n_step): forward over n_env observations (row e of step t is
 * obs[step_idx[t*n_env+e]], NULL = rows t*n_env..), sample actions with Philox(seed, *offset_dev + t), and
 * write action/logp/value time-major [n_step, n_env]; finally *offset_dev += n_step.  This is the batched
 * replacement of the per-explorer batch-1 PPO.predict calls (xt/agent/ppo/ppo.py:35-45,
 * xt/algorithm/ppo/ppo.py:87-95); use_graph caches one CUDA graph per (buffers, n_env, n_step). */
int xtb_ppo_rollout_infer(xtb_net* net, const void* obs, const int32_t* step_idx, int n_env, int n_step,
                          int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                          int32_t* action, float* logp, float* value, int use_graph, void* stream);

/* PPO.predict (xt/model/ppo/ppo.py:104-109) on host buffers in one call: staged H2D of `obs_host` (pageable,
 * obs_bytes) into `obs_dev`, xtb_ppo_rollout_infer with n_step = 1 writing the packed block
 * out_dev = [action int32 | logp f32 | value f32] x n_env, one D2H into `out_host` (pinned) and a stream
 * synchronise: when it returns, out_host holds the step's results. */
int xtb_ppo_predict_host(xtb_net* net, const void* obs_host, size_t obs_bytes, void* obs_dev, int n_env,
                         int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                         float* out_dev, float* out_host, int use_graph, void* stream);

/* The same call for actors whose predict() also returns the logits (ImpalaCnnOpt.predict,
 * xt/model/impala/impala_cnn_opt.py:267-277: [logits, baseline, action]): additionally copies the [n_env, A] logits of
 * tensor `pi_tensor` into `logits_host` (pinned; NULL = skip) before the synchronise. */
int xtb_actor_predict_host(xtb_net* net, const void* obs_host, size_t obs_bytes, void* obs_dev, int n_env,
                           int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                           float* out_dev, float* out_host, float* logits_host, int use_graph, void* stream);

/* Data-parallel hook: called between backward and the optimiser with the flat
 * gradient bucket; must SUM it over ranks on `stream` (e.g. ncclAllReduce).  Called once with
 * grads == NULL before the loop: must return the world size.  While a hook is installed the
 * loss/gradient scale becomes 1/(world*B_local) and CUDA-graph replay is disabled.
 * Reference precedent: xt/framework/trainer.py:82-92 (shared-memory gradient averaging). */
typedef int (*xtb_grad_hook)(void* user, float* grads, long long count, void* stream);
int xtb_set_grad_hook(xtb_grad_hook hook, void* user);

/* Data-parallel communicator owned by the library (precedent zeus/trainer/trainer_tf.py:187-203): NCCL is
 * resolved with dlopen (`nccl_path` NULL = "libnccl.so.2").  Rank 0 calls xtb_comm_unique_id and ships the 128 bytes to
 * the other ranks (engine.py uses torch.distributed); every rank then calls xtb_comm_create.  With a communicator
 * installed (xtb_set_grad_comm) the fused training loops all-reduce the flat gradient bucket themselves -- the large
 * dense weight gradient as soon as it is final, on a side stream under the rest of the backward pass -- and stay inside
 * the CUDA graph. */
typedef struct xtb_comm xtb_comm;
int xtb_comm_unique_id(const char* nccl_path, void* id128);
int xtb_comm_create(const char* nccl_path, const void* id128, int rank, int world, xtb_comm** out);
void xtb_comm_destroy(xtb_comm* comm);
int xtb_comm_world(const xtb_comm* comm);
int xtb_set_grad_comm(xtb_comm* comm);
int xtb_comm_allreduce(xtb_comm* comm, float* buf, long long count, void* stream);

/* Launch one kernel of one layer alone (which: 0 forward, 1 weight gradient) on the tensors
 * currently in the workspace -- measurement hook for bench.py's roofline object. */
int xtb_net_bench_layer(xtb_net* net, int layer, int which, const void* obs, const int32_t* gather_idx,
                        int batch, void* stream);

/* Kernel-path selection: 1 (default) = tcgen05 tensor-core kernels wherever the shape is covered,
 * 0 = fp32 CUDA-core kernels only (also XTB_TC=0 in the environment).  For A/B parity tests. */
int xtb_set_tc_mode(int mode);
/* 1 (default): xtb_ppo_train evaluates both heads, the loss and their backward in one fused kernel;
 * 0: layer-by-layer (also XTB_FUSE_HEADS=0).  Affects graphs captured after the call. */
int xtb_set_fuse_heads(int on);
int xtb_get_tc_mode(void);
/* Self-test of the tcgen05 GEMM core on plain fp32 matrices (sizes multiples of 8):
 * mode 0: C = A[M,K] B[K,N]; mode 1: C = A[M,K] Bt[N,K]^T; mode 2: C = At[K,M]^T B[K,N]. */
int xtb_tc_gemm_test(int mode, const float* a, const float* b, float* c, int M, int N, int K, int ksplit,
                     void* stream);

/* ---- host <-> device staging (pinned ring replacing feed_dict copies) --- */
void* xtb_pinned_alloc(size_t bytes);
void xtb_pinned_free(void* p);
int xtb_copy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
/* Pageable host memory -> device: worker threads memcpy 256 KiB chunks into a pinned ring while the caller
 * enqueues one async copy per staged chunk on `stream`.  On return `src_host` has been consumed (it may be
 * reused); the device side is ordered on `stream`.  Replaces the single-threaded feed_dict staging of
 * sess.run (xt/model/ppo/ppo.py:104-132).  XTB_STAGE_THREADS (default 4) sizes the pool. */
int xtb_copy_h2d_staged(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int xtb_copy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int xtb_stream_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XTB200_H_ */
