// rl_kernels.cuh -- trajectory post-processing and loss kernels (HBM / latency bound).
//   gae_kernel          : xt/agent/ppo/ppo.py:77-106      one warp per env, affine reverse scan
//   sample_kernel       : xt/model/tf_dist.py:89-130      Gumbel-max + log-prob
//   ppo_loss_kernel     : xt/model/ppo/__init__.py:4-25   loss + dlogits/dv, block-reduced loss
//   vtrace_kernel       : xt/model/impala/vtrace.py:39-115 + impala_cnn_opt.py:299-351
//   dqn_loss_kernel     : xt/algorithm/dqn/dqn.py:79-97 + Keras mse
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_f32.cuh"

namespace xtb {

constexpr int MAX_ADIM = 32;

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level sum -> one atomicAdd per block
__device__ inline void block_atomic_add(float v, float* out) {
  __shared__ float red[32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    float s = lane < nw ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(out, s);
  }
}

// ------------------------------------------------------------------------------------------
// GAE.  adv_t = delta_t + c_t * adv_{t+1},  c_t = (1-done_t)*gamma*lam, adv_T = 0.
// Each lane owns `chunk` consecutive steps; per-lane affine map (a,b): adv_in -> a*adv_in+b is
// composed right-to-left, exclusive-scanned across lanes with shuffles, then replayed.
// ------------------------------------------------------------------------------------------
__global__ void gae_kernel(const float* __restrict__ value, const float* __restrict__ reward,
                           const uint8_t* __restrict__ done, int n_env, int T, float gamma, float lam,
                           int sign_clip, float* __restrict__ adv, float* __restrict__ old_v,
                           float* __restrict__ target_v) {
  pdl_wait(); pdl_trigger();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n_env) return;
  const float* V = value + (long long)warp * (T + 1);
  const float* R = reward + (long long)warp * T;
  const uint8_t* D = done + (long long)warp * T;
  float* A = adv + (long long)warp * T;
  float* OV = old_v + (long long)warp * T;
  float* TV = target_v + (long long)warp * T;
  int chunk = (T + 31) / 32;
  int t0 = lane * chunk, t1 = min(T, t0 + chunk);
  // pass 1: compose the chunk's affine map, walking backwards
  float a = 1.f, b = 0.f;
  for (int t = t1 - 1; t >= t0; --t) {
    float r = R[t];
    if (sign_clip) r = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
    float disc = D[t] ? 0.f : gamma;
    float delta = r + disc * V[t + 1] - V[t];
    float c = disc * lam;
    // adv_t = delta + c*adv_{t+1};  adv_{t+1} = a*x + b  =>  adv_t = (c*a)*x + (delta + c*b)
    b = delta + c * b;
    a = c * a;
  }
  // suffix scan over lanes: incoming value for lane l = result of lanes l+1..31 applied to 0
  // inclusive suffix composition F_l = f_l o F_{l+1}
  float fa = a, fb = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float oa = __shfl_down_sync(0xffffffffu, fa, o);
    float ob = __shfl_down_sync(0xffffffffu, fb, o);
    if (lane + o < 32) {  // F = F o other : x -> fa*(oa*x+ob)+fb
      fb = fa * ob + fb;
      fa = fa * oa;
    }
  }
  float incoming = __shfl_down_sync(0xffffffffu, fb, 1);  // F_{l+1}(0) = fb_{l+1}
  if (lane == 31) incoming = 0.f;
  // pass 2: replay
  float nxt = incoming;
  for (int t = t1 - 1; t >= t0; --t) {
    float r = R[t];
    if (sign_clip) r = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
    float disc = D[t] ? 0.f : gamma;
    float v = V[t];
    float delta = r + disc * V[t + 1] - v;
    float ad = delta + disc * lam * nxt;
    A[t] = ad;
    OV[t] = v;
    TV[t] = ad + v;
    nxt = ad;
  }
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10
// ------------------------------------------------------------------------------------------
__device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void sample_kernel(const float* __restrict__ logits, int B, int A,
                              const float* __restrict__ uniforms, uint64_t seed, uint64_t offset,
                              int32_t* __restrict__ action, float* __restrict__ logp) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = logits + (long long)b * A;
  float mx = -INFINITY;
  for (int i = 0; i < A; i++) mx = fmaxf(mx, l[i]);
  float z = 0.f;
  for (int i = 0; i < A; i++) z += expf(l[i] - mx);
  float lz = logf(z);
  float best = -INFINITY; int bi = 0;
  uint32_t c[4] = {0, 0, 0, 0};
  for (int i = 0; i < A; i++) {
    float u;
    if (uniforms) {
      u = uniforms[(long long)b * A + i];
    } else {
      if ((i & 3) == 0) {
        c[0] = (uint32_t)b; c[1] = (uint32_t)(i >> 2);
        c[2] = (uint32_t)(offset & 0xffffffffu); c[3] = (uint32_t)(offset >> 32);
        philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
      }
      u = (float)(c[i & 3] >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f;
    }
    float g = -logf(-logf(u));
    float s = l[i] + g;
    if (s > best) { best = s; bi = i; }   // first maximum wins, like np.argmax
  }
  action[b] = bi;
  logp[b] = l[bi] - mx - lz;
}

// rollout variant: Philox offset = *offset_dev + t_add (device-resident counter so that CUDA-graph replays
// draw fresh noise), and the value head output is copied out alongside.
__global__ void sample_rollout_kernel(const float* __restrict__ logits, const float* __restrict__ v_in, int B, int A,
                                      uint64_t seed, const unsigned long long* __restrict__ offset_dev, int t_add,
                                      int32_t* __restrict__ action, float* __restrict__ logp, float* __restrict__ v_out) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint64_t offset = (uint64_t)(*offset_dev) + (uint64_t)t_add;
  const float* l = logits + (long long)b * A;
  float mx = -INFINITY;
  for (int i = 0; i < A; i++) mx = fmaxf(mx, l[i]);
  float z = 0.f;
  for (int i = 0; i < A; i++) z += expf(l[i] - mx);
  float lz = logf(z);
  float best = -INFINITY; int bi = 0;
  uint32_t c[4] = {0, 0, 0, 0};
  for (int i = 0; i < A; i++) {
    if ((i & 3) == 0) {
      c[0] = (uint32_t)b; c[1] = (uint32_t)(i >> 2);
      c[2] = (uint32_t)(offset & 0xffffffffu); c[3] = (uint32_t)(offset >> 32);
      philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
    }
    float u = (float)(c[i & 3] >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f;
    float s = l[i] - logf(-logf(u));
    if (s > best) { best = s; bi = i; }
  }
  action[b] = bi;
  logp[b] = l[bi] - mx - lz;
  v_out[b] = v_in[b];
}
__global__ void bump_counter_kernel(unsigned long long* ctr, int add) {
  pdl_wait(); pdl_trigger(); *ctr += (unsigned long long)add; }

__global__ void argmax_kernel(const float* __restrict__ q, int B, int A, int32_t* __restrict__ action) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = q + (long long)b * A;
  float best = l[0]; int bi = 0;
  for (int i = 1; i < A; i++) if (l[i] > best) { best = l[i]; bi = i; }
  action[b] = bi;
}

// ------------------------------------------------------------------------------------------
// PPO loss + gradient wrt (logits, v).  One thread per sample.
// ------------------------------------------------------------------------------------------
struct PpoHyperDev { float clip_ratio, ent_coef, vf_clip, critic_coef; };

__global__ void ppo_loss_kernel(const float* __restrict__ logits, const float* __restrict__ v,
                                const int32_t* __restrict__ idx, const int32_t* __restrict__ action,
                                const float* __restrict__ old_logp, const float* __restrict__ adv,
                                const float* __restrict__ old_v, const float* __restrict__ target_v,
                                int B, int A, PpoHyperDev hp, float inv_count,
                                float* __restrict__ dlogits, float* __restrict__ dv,
                                float* __restrict__ loss_out) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float lsum = 0.f;
  if (b < B) {
    int r = idx ? idx[b] : b;
    const float* l = logits + (long long)b * A;
    float lg[MAX_ADIM];
    float mx = -INFINITY;
    for (int i = 0; i < A; i++) { lg[i] = l[i]; mx = fmaxf(mx, lg[i]); }
    float z = 0.f;
    for (int i = 0; i < A; i++) z += expf(lg[i] - mx);
    float lz = logf(z);
    float H = 0.f;
    for (int i = 0; i < A; i++) { float rl = lg[i] - mx; H += (expf(rl) / z) * (lz - rl); }
    int a = action[r];
    float logp_a = lg[a] - mx - lz;
    float ratio = expf(logp_a - old_logp[r]);
    float ad = adv[r];
    float s1 = ratio * ad;
    float rc = fminf(fmaxf(ratio, 1.f - hp.clip_ratio), 1.f + hp.clip_ratio);
    float s2 = rc * ad;
    float surr = fminf(s1, s2);
    // d surr / d logp_a : through s1 when s1<=s2, else through the clip (zero outside the range)
    float dsurr;
    if (s1 <= s2) dsurr = ratio * ad;
    else dsurr = (ratio >= 1.f - hp.clip_ratio && ratio <= 1.f + hp.clip_ratio) ? ratio * ad : 0.f;
    float vv = v[b], R = target_v[r], ov = old_v[r];
    float l1 = (vv - R) * (vv - R);
    float dcl = fminf(fmaxf(vv - ov, -hp.vf_clip), hp.vf_clip);
    float vc = ov + dcl;
    float l2 = (vc - R) * (vc - R);
    float dvl;
    if (l1 >= l2) dvl = 2.f * (vv - R);
    else dvl = (vv - ov >= -hp.vf_clip && vv - ov <= hp.vf_clip) ? 2.f * (vc - R) : 0.f;
    lsum = (-surr - hp.ent_coef * H + hp.critic_coef * 0.5f * fmaxf(l1, l2)) * inv_count;
    dv[b] = hp.critic_coef * 0.5f * dvl * inv_count;
    for (int i = 0; i < A; i++) {
      float rl = lg[i] - mx;
      float p = expf(rl) / z;
      float logp_i = rl - lz;
      float dlogp = ((i == a) ? 1.f : 0.f) - p;             // d logp_a / d l_i
      float dH = -p * (logp_i + H);                           // d H / d l_i
      dlogits[(long long)b * A + i] = (-dsurr * dlogp - hp.ent_coef * dH) * inv_count;
    }
  }
  block_atomic_add(lsum, loss_out);
}

// ------------------------------------------------------------------------------------------
// V-trace + IMPALA loss.  One warp per trajectory, lanes own consecutive time chunks.
// loss = sum xent*pg_adv + 0.25*sum (vs-V)^2 - 0.01*sum H    (sums, not means)
// ------------------------------------------------------------------------------------------
__device__ inline void softmax_stats(const float* l, int A, float& mx, float& lz) {
  mx = -INFINITY;
  for (int i = 0; i < A; i++) mx = fmaxf(mx, l[i]);
  float z = 0.f;
  for (int i = 0; i < A; i++) z += expf(l[i] - mx);
  lz = logf(z);
}

__global__ void vtrace_kernel(const float* __restrict__ tp_logits, const float* __restrict__ baseline,
                              const float* __restrict__ bp_logits, const int32_t* __restrict__ action,
                              const uint8_t* __restrict__ done, const float* __restrict__ reward,
                              int n_traj, int S, int A, float gamma, float* __restrict__ dlogits,
                              float* __restrict__ dbaseline, float* __restrict__ vs_out,
                              float* __restrict__ pg_out, float* __restrict__ loss_out) {
  pdl_wait(); pdl_trigger();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  float lsum = 0.f;
  if (warp < n_traj) {
    const long long base = (long long)warp * S;
    const int T = S - 1;                       // drop_last
    const float boot = baseline[base + S - 1];
    int chunk = (T + 31) / 32;
    int t0 = lane * chunk, t1 = min(T, t0 + chunk);
    // pass 1: chunk affine map of acc_t = delta_t + (disc_t*c_t)*acc_{t+1}
    float a = 1.f, b = 0.f;
    for (int t = t1 - 1; t >= t0; --t) {
      long long n = base + t;
      float mx, lz, bmx, blz;
      softmax_stats(tp_logits + n * A, A, mx, lz);
      softmax_stats(bp_logits + n * A, A, bmx, blz);
      int ac = action[n];
      float tlp = tp_logits[n * A + ac] - mx - lz;
      float blp = bp_logits[n * A + ac] - bmx - blz;
      float rho = expf(tlp - blp);
      float crho = fminf(1.f, rho), cs = fminf(1.f, rho);
      float disc = done[n] ? 0.f : gamma;
      float r = fminf(fmaxf(reward[n], -1.f), 1.f);
      float V = baseline[n];
      float Vn = (t == T - 1) ? boot : baseline[n + 1];
      float delta = crho * (r + disc * Vn - V);
      float c = disc * cs;
      b = delta + c * b;
      a = c * a;
    }
    float fa = a, fb = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float oa = __shfl_down_sync(0xffffffffu, fa, o);
      float ob = __shfl_down_sync(0xffffffffu, fb, o);
      if (lane + o < 32) { fb = fa * ob + fb; fa = fa * oa; }
    }
    float incoming = __shfl_down_sync(0xffffffffu, fb, 1);
    if (lane == 31) incoming = 0.f;
    // vs of the first step of the next lane's chunk (needed for pg_adv at our last step)
    // vs_{t1} = V_{t1} + acc_{t1};  acc_{t1} == incoming.
    float acc_next = incoming;
    for (int t = t1 - 1; t >= t0; --t) {
      long long n = base + t;
      float mx, lz, bmx, blz;
      const float* tl = tp_logits + n * A;
      softmax_stats(tl, A, mx, lz);
      softmax_stats(bp_logits + n * A, A, bmx, blz);
      int ac = action[n];
      float tlp = tl[ac] - mx - lz;
      float blp = bp_logits[n * A + ac] - bmx - blz;
      float rho = expf(tlp - blp);
      float crho = fminf(1.f, rho), cs = fminf(1.f, rho), cpg = fminf(1.f, rho);
      float disc = done[n] ? 0.f : gamma;
      float r = fminf(fmaxf(reward[n], -1.f), 1.f);
      float V = baseline[n];
      float Vn = (t == T - 1) ? boot : baseline[n + 1];
      float vs_next = (t == T - 1) ? boot : (Vn + acc_next);
      float delta = crho * (r + disc * Vn - V);
      float acc = delta + disc * cs * acc_next;
      float vs = V + acc;
      float pg = cpg * (r + disc * vs_next - V);
      if (vs_out) vs_out[n] = vs;
      if (pg_out) pg_out[n] = pg;
      // entropy and grads
      float H = 0.f;
      for (int i = 0; i < A; i++) { float lp = tl[i] - mx - lz; H -= expf(lp) * lp; }
      for (int i = 0; i < A; i++) {
        float lp = tl[i] - mx - lz;
        float p = expf(lp);
        float d = pg * (p - ((i == ac) ? 1.f : 0.f)) + 0.01f * p * (lp + H);
        dlogits[n * A + i] = d;
      }
      dbaseline[n] = 0.5f * (V - vs);
      lsum += (-tlp) * pg + 0.25f * (vs - V) * (vs - V) - 0.01f * H;
      acc_next = acc;
    }
    if (lane == 0) {
      long long n = base + S - 1;
      for (int i = 0; i < A; i++) dlogits[n * A + i] = 0.f;
      dbaseline[n] = 0.f;
      if (vs_out) vs_out[n] = 0.f;
      if (pg_out) pg_out[n] = 0.f;
    }
  }
  block_atomic_add(lsum, loss_out);
}

// ------------------------------------------------------------------------------------------
// DQN TD target + loss gradient.  Reference semantics (xt/algorithm/dqn/dqn.py:79-97 + Keras 'mse', mean over B*A):
// y = r (+ gamma * max_a' Q_target(s') unless done), loss = mean (Q - y)^2 over B*A with only the taken action non-zero.
// Extensions named by BASELINE.json's north_star, off by default:
//   disc != NULL : per-sample bootstrap discount (gamma^n of an n-step return built by nstep_kernel; 0 = terminated)
//   huber > 0    : Huber loss with that delta instead of the squared error (gradient clip(diff, -delta, delta))
//   idx != NULL  : action / reward / done / disc are indexed through idx (minibatch rows of a replay ring)
// ------------------------------------------------------------------------------------------
__global__ void dqn_loss_kernel(const float* __restrict__ q, const float* __restrict__ qn_t,
                                const float* __restrict__ qn_o, const int32_t* __restrict__ idx,
                                const int32_t* __restrict__ action, const float* __restrict__ reward,
                                const uint8_t* __restrict__ done, const float* __restrict__ disc,
                                int B, int A, float gamma, float huber, float inv_count, float* __restrict__ dq,
                                float* __restrict__ y_out, float* __restrict__ loss_out) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float lsum = 0.f;
  if (b < B) {
    const int r = idx ? idx[b] : b;
    const float* t = qn_t + (long long)b * A;
    float mq;
    if (qn_o) {
      const float* o = qn_o + (long long)b * A;
      float best = o[0]; int bi = 0;
      for (int i = 1; i < A; i++) if (o[i] > best) { best = o[i]; bi = i; }
      mq = t[bi];
    } else {
      mq = t[0];
      for (int i = 1; i < A; i++) mq = fmaxf(mq, t[i]);
    }
    const float g = disc ? disc[r] : gamma;
    float y = done[r] ? reward[r] : reward[r] + g * mq;
    int a = action[r];
    float diff = q[(long long)b * A + a] - y;
    float grad, l;
    if (huber > 0.f) {
      const float ad = fabsf(diff);
      l = ad <= huber ? 0.5f * diff * diff : huber * (ad - 0.5f * huber);
      grad = fminf(fmaxf(diff, -huber), huber);
    } else { l = diff * diff; grad = 2.f * diff; }
    for (int i = 0; i < A; i++) dq[(long long)b * A + i] = (i == a) ? grad * inv_count : 0.f;
    if (y_out) y_out[b] = y;
    lsum = l * inv_count;
  }
  block_atomic_add(lsum, loss_out);
}

// n-step returns over env-major trajectories [E][T] (north_star: "n-step TD-target kernel over a device-resident rollout
// buffer"): R_t = sum_{k<m} gamma^k r_{t+k}, m = steps until the first done (inclusive) or n or the end of the segment;
// disc_t = gamma^m, or 0 when the window ended in a terminal step; last_t = t + m - 1 (row whose next-state bootstraps);
// done_n_t = 1 when the window contains a terminal step.  One thread per (env, t); rows of a warp are consecutive t.
__global__ void nstep_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ done, int E, int T, int n, float gamma,
                             float* __restrict__ ret, float* __restrict__ disc, int32_t* __restrict__ last, uint8_t* __restrict__ done_n) {
  pdl_wait(); pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * T) return;
  const int e = i / T, t = i - e * T;
  float acc = 0.f, g = 1.f;
  int k = 0; bool term = false;
  for (; k < n && t + k < T; k++) {
    acc = fmaf(g, reward[e * T + t + k], acc);
    g *= gamma;
    if (done[e * T + t + k]) { term = true; k++; break; }
  }
  ret[i] = acc;
  disc[i] = term ? 0.f : g;
  last[i] = i + k - 1;
  done_n[i] = term ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// Fused PPO heads: pi/v dense heads + categorical loss + their backward in ONE kernel (K4/K7 of the
// kernel work-list: the logits never leave the chip between the head GEMM and the loss).
//   logits = h_pi @ Wpi + bpi,  v = h_v @ Wv + bv            (xt/model/model_utils.py:63-65)
//   loss, dlogits, dv                                        (xt/model/ppo/__init__.py:4-25)
//   dWpi,dbpi,dWv,dbv += ...   (atomics into the flat gradient bucket)
//   g(h) = (dlogits @ Wpi^T [+ dv * Wv^T]) * act'(h)         -> gradient wrt the hidden pre-activation
// One warp per sample, lane owns hidden units lane, lane+32, ...  (K <= 32*HEAD_KPL, A <= HEAD_AMAX)
// ------------------------------------------------------------------------------------------
// template: HEAD_KPL hidden units per lane (K <= 32*HEAD_KPL), HEAD_AMAX >= A

struct PpoHeadsArgs {
  const float* h_pi; const float* h_v;     // [B,K] hidden activations (may alias)
  float* g_pi; float* g_v;                 // [B,K] gradient wrt hidden pre-activation (may alias)
  __nv_bfloat16* gp_hi; long long gp_lo; __nv_bfloat16* gv_hi; long long gv_lo;   // their batch-planar bf16 planes (bp_gemm.cuh), may be NULL
  int pitch;                               // rows per feature chunk of those planes
  const float* w_pi; const float* b_pi; const float* w_v; const float* b_v;
  // Parameter-gradient partial sums, one slab of `slab` floats per block, reduced afterwards in block order by
  // bp::grad_reduce_kernel (no atomics: the step is bitwise reproducible).  Slab layout (K hidden units, A actions):
  // [dW_pi K*A | dW_v K | dbh_pi K | dbh_v K | db_pi A | db_v 1 | loss 1]; dbh_* = bias gradients of the layers that
  // produced h_pi / h_v (column sums of g).
  float* part; int slab;
  const int32_t* idx; const int32_t* action; const float* old_logp; const float* adv; const float* old_v; const float* target_v;
  float* logits_out; float* v_out;
  int B, K, A, act_pi, act_v, shared;
  PpoHyperDev hp; float inv_count;
};

template <int HEAD_KPL, int HEAD_AMAX>
__global__ void __launch_bounds__(256) ppo_heads_kernel(PpoHeadsArgs a) {
  pdl_wait(); pdl_trigger();
  extern __shared__ float sh_dw[];          // [nwarp][nacc] per-warp partial sums in slab layout
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int K = a.K, A = a.A, kpl = K / 32;
  const int nacc = K * A + 3 * K + A + 2;
  float wpi[HEAD_KPL][HEAD_AMAX], wv[HEAD_KPL];
  float dwp[HEAD_KPL][HEAD_AMAX], dwv[HEAD_KPL];
#pragma unroll
  for (int j = 0; j < HEAD_KPL; j++) {
    wv[j] = 0.f; dwv[j] = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) { wpi[j][i] = 0.f; dwp[j][i] = 0.f; }
    if (j < kpl) {
      int k = lane + 32 * j;
      wv[j] = a.w_v[k];
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) if (i < A) wpi[j][i] = a.w_pi[k * A + i];
    }
  }
  float dbp[HEAD_AMAX]; float dbv = 0.f, lsum = 0.f;
  float bh_pi[HEAD_KPL], bh_v[HEAD_KPL];
#pragma unroll
  for (int j = 0; j < HEAD_KPL; j++) { bh_pi[j] = 0.f; bh_v[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < HEAD_AMAX; i++) dbp[i] = 0.f;
  for (int b = blockIdx.x * nwarp + warp; b < a.B; b += gridDim.x * nwarp) {
    float hp_[HEAD_KPL], hv_[HEAD_KPL];
    float acc[HEAD_AMAX + 1];
#pragma unroll
    for (int i = 0; i <= HEAD_AMAX; i++) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_KPL; j++) {
      hp_[j] = 0.f; hv_[j] = 0.f;
      if (j < kpl) {
        int k = lane + 32 * j;
        hp_[j] = a.h_pi[(long long)b * K + k];
        hv_[j] = a.shared ? hp_[j] : a.h_v[(long long)b * K + k];
#pragma unroll
        for (int i = 0; i < HEAD_AMAX; i++) acc[i] = fmaf(hp_[j], wpi[j][i], acc[i]);
        acc[HEAD_AMAX] = fmaf(hv_[j], wv[j], acc[HEAD_AMAX]);
      }
    }
#pragma unroll
    for (int i = 0; i <= HEAD_AMAX; i++) acc[i] = warp_sum(acc[i]);
    // ---- loss + d(logits), d(v): every lane computes the same values
    float lg[HEAD_AMAX]; float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) { lg[i] = (i < A) ? acc[i] + a.b_pi[i] : -INFINITY; mx = fmaxf(mx, lg[i]); }
    float vv = acc[HEAD_AMAX] + a.b_v[0];
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) if (i < A) z += expf(lg[i] - mx);
    float lz = logf(z), H = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) if (i < A) { float rl = lg[i] - mx; H += (expf(rl) / z) * (lz - rl); }
    int r = a.idx ? a.idx[b] : b;
    int ac = a.action[r];
    float logp_a = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) if (i == ac) logp_a = lg[i] - mx - lz;
    float ratio = expf(logp_a - a.old_logp[r]);
    float ad = a.adv[r];
    float s1 = ratio * ad;
    float s2 = fminf(fmaxf(ratio, 1.f - a.hp.clip_ratio), 1.f + a.hp.clip_ratio) * ad;
    float surr = fminf(s1, s2);
    float dsurr = (s1 <= s2) ? ratio * ad
                             : ((ratio >= 1.f - a.hp.clip_ratio && ratio <= 1.f + a.hp.clip_ratio) ? ratio * ad : 0.f);
    float R = a.target_v[r], ov = a.old_v[r];
    float l1 = (vv - R) * (vv - R);
    float vc = ov + fminf(fmaxf(vv - ov, -a.hp.vf_clip), a.hp.vf_clip);
    float l2 = (vc - R) * (vc - R);
    float dvl = (l1 >= l2) ? 2.f * (vv - R)
                           : ((vv - ov >= -a.hp.vf_clip && vv - ov <= a.hp.vf_clip) ? 2.f * (vc - R) : 0.f);
    float dv = a.hp.critic_coef * 0.5f * dvl * a.inv_count;
    float dl[HEAD_AMAX];
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) {
      dl[i] = 0.f;
      if (i < A) {
        float rl = lg[i] - mx, p = expf(rl) / z;
        dl[i] = (-dsurr * (((i == ac) ? 1.f : 0.f) - p) + a.hp.ent_coef * p * (rl - lz + H)) * a.inv_count;
      }
    }
    if (lane == 0) {
      lsum += (-surr - a.hp.ent_coef * H + a.hp.critic_coef * 0.5f * fmaxf(l1, l2)) * a.inv_count;
      if (a.logits_out) for (int i = 0; i < A; i++) a.logits_out[(long long)b * A + i] = lg[i];
      if (a.v_out) a.v_out[b] = vv;
      dbv += dv;
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) dbp[i] += dl[i];
    }
    // ---- head weight gradients (registers) and gradient wrt the hidden units
#pragma unroll
    for (int j = 0; j < HEAD_KPL; j++) {
      if (j < kpl) {
        int k = lane + 32 * j;
        float gp = 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_AMAX; i++) { dwp[j][i] = fmaf(hp_[j], dl[i], dwp[j][i]); gp = fmaf(dl[i], wpi[j][i], gp); }
        dwv[j] = fmaf(hv_[j], dv, dwv[j]);
        float gv = dv * wv[j];
        long long eo = (long long)b * K + k;
        if (a.shared) {
          float r = (gp + gv) * act_grad_from_out(a.act_pi, hp_[j]);
          a.g_pi[eo] = r; f32_store_plane(a.gp_hi, a.gp_lo, ((long long)(k >> 3) * a.pitch + b) * 8 + (k & 7), r);
          bh_pi[j] += r;
        } else {
          float r1 = gp * act_grad_from_out(a.act_pi, hp_[j]), r2 = gv * act_grad_from_out(a.act_v, hv_[j]);
          a.g_pi[eo] = r1; f32_store_plane(a.gp_hi, a.gp_lo, ((long long)(k >> 3) * a.pitch + b) * 8 + (k & 7), r1);
          a.g_v[eo] = r2; f32_store_plane(a.gv_hi, a.gv_lo, ((long long)(k >> 3) * a.pitch + b) * 8 + (k & 7), r2);
          bh_pi[j] += r1; bh_v[j] += r2;
        }
      }
    }
  }
  // plane rows [B, round16(B)) are read by the weight-gradient K loop (16 samples per MMA step): keep them zero
  for (int b = a.B + blockIdx.x * nwarp + warp; b < ((a.B + 15) & ~15); b += gridDim.x * nwarp) {
#pragma unroll
    for (int j = 0; j < HEAD_KPL; j++) {
      if (j < kpl) {
        int k = lane + 32 * j;
        long long e = ((long long)(k >> 3) * a.pitch + b) * 8 + (k & 7);
        f32_store_plane(a.gp_hi, a.gp_lo, e, 0.f);
        if (!a.shared) f32_store_plane(a.gv_hi, a.gv_lo, e, 0.f);
      }
    }
  }
  // ---- every warp parks its sums in its own row; the rows are added in warp order and written to this block's slab
  float* row = sh_dw + (size_t)warp * nacc;
#pragma unroll
  for (int j = 0; j < HEAD_KPL; j++) {
    if (j < kpl) {
      int k = lane + 32 * j;
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) if (i < A) row[k * A + i] = dwp[j][i];
      row[K * A + k] = dwv[j];
      row[K * A + K + k] = bh_pi[j];
      row[K * A + 2 * K + k] = a.shared ? 0.f : bh_v[j];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < HEAD_AMAX; i++) if (i < A) row[K * A + 3 * K + i] = dbp[i];
    row[K * A + 3 * K + A] = dbv;
    row[K * A + 3 * K + A + 1] = lsum;
  }
  __syncthreads();
  float* out = a.part + (size_t)blockIdx.x * a.slab;
  for (int e = threadIdx.x; e < nacc; e += blockDim.x) {
    float t = sh_dw[e];
    for (int w = 1; w < nwarp; w++) t += sh_dw[(size_t)w * nacc + e];
    out[e] = t;
  }
}

// Inference heads: logits / value of both dense heads + Gumbel-max sampling in one kernel (one warp per sample).
// Philox offset = *offset_dev + t_add (device-resident counter, see sample_rollout_kernel).
template <int HEAD_KPL, int HEAD_AMAX>
__global__ void __launch_bounds__(256)
ppo_infer_heads_kernel(const float* __restrict__ h_pi, const float* __restrict__ h_v, const float* __restrict__ w_pi,
                       const float* __restrict__ b_pi, const float* __restrict__ w_v, const float* __restrict__ b_v,
                       int B, int K, int A, uint64_t seed, const unsigned long long* __restrict__ offset_dev, int t_add,
                       int32_t* __restrict__ action, float* __restrict__ logp, float* __restrict__ v_out,
                       float* __restrict__ logits_out) {
  pdl_wait(); pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int kpl = K / 32;
  const uint64_t offset = (uint64_t)(*offset_dev) + (uint64_t)t_add;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    float acc[HEAD_AMAX + 1];
#pragma unroll
    for (int i = 0; i <= HEAD_AMAX; i++) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_KPL; j++) {
      if (j < kpl) {
        int k = lane + 32 * j;
        float hp = h_pi[(long long)b * K + k], hv = h_v[(long long)b * K + k];
#pragma unroll
        for (int i = 0; i < HEAD_AMAX; i++) if (i < A) acc[i] = fmaf(hp, w_pi[k * A + i], acc[i]);
        acc[HEAD_AMAX] = fmaf(hv, w_v[k], acc[HEAD_AMAX]);
      }
    }
#pragma unroll
    for (int i = 0; i <= HEAD_AMAX; i++) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
      float lg[HEAD_AMAX]; float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) { lg[i] = (i < A) ? acc[i] + b_pi[i] : -INFINITY; mx = fmaxf(mx, lg[i]); }
      float z = 0.f;
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) if (i < A) z += expf(lg[i] - mx);
      float lz = logf(z), best = -INFINITY; int bi = 0;
      uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) {
        if (i < A) {
          if ((i & 3) == 0) {
            c[0] = (uint32_t)b; c[1] = (uint32_t)(i >> 2);
            c[2] = (uint32_t)(offset & 0xffffffffu); c[3] = (uint32_t)(offset >> 32);
            philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
          }
          float u = (float)(c[i & 3] >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f;
          float sc = lg[i] - logf(-logf(u));
          if (sc > best) { best = sc; bi = i; }
          if (logits_out) logits_out[(long long)b * A + i] = lg[i];
        }
      }
      float la = 0.f;
#pragma unroll
      for (int i = 0; i < HEAD_AMAX; i++) if (i == bi) la = lg[i];
      action[b] = bi;
      logp[b] = la - mx - lz;
      v_out[b] = acc[HEAD_AMAX] + b_v[0];
    }
  }
}

// Keras train_on_batch(states, y) with loss='mse': mean over B*A of (q-y)^2
__global__ void mse_loss_kernel(const float* __restrict__ q, const float* __restrict__ y, long long n,
                                float inv_count, float* __restrict__ dq, float* __restrict__ loss_out) {
  pdl_wait(); pdl_trigger();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float lsum = 0.f;
  if (i < n) {
    float d = q[i] - y[i];
    dq[i] = 2.f * d * inv_count;
    lsum = d * d * inv_count;
  }
  block_atomic_add(lsum, loss_out);
}

}  // namespace xtb
