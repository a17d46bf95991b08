// brb_policy.cu — fused forward pass of the PPO actor-critic for the rollout (sm_100a).
//
// The reference trains stable_baselines3.PPO("MlpPolicy", env) (src/sb_rl.py:63-71): SB3's ActorCriticPolicy with
// pi = [64, 64], vf = [64, 64], tanh, a state-independent log_std; the exported policy (RobotMovePolicy.tflite) shows the
// same graph: 6 matmuls, 4 tanh, 3 outputs (actions, values, log-prob).  collect_rollouts calls it once per env step:
//   actions, values, log_probs = policy.forward(obs)            [SB3 on_policy_algorithm.collect_rollouts, third party]
// Here that call is ONE launch for all N robots: both 6-64-64 towers, the Gaussian sample, its log-probability, the value
// and the copy of the action clipped to the action space that the env receives.  One thread = one robot: the first hidden
// layer lives in registers, the 64x64 second-layer weights are read from shared memory as broadcast float4 (every lane of
// a warp reads the same weight), and the second hidden layer is consumed by the output layer as it is produced, so no
// activation ever goes to memory.  FP32 throughout (the reference trains in fp32 on the CPU).
//
// Parameter block (one flat fp32 array, SB3 state-dict order, see ppo.py pack_params):
//   pi.W1[64][6] pi.b1[64] pi.W2[64][64] pi.b2[64] | vf.W1[64][6] vf.b1[64] vf.W2[64][64] vf.b2[64] |
//   action_net.W[2][64] action_net.b[2] | value_net.W[1][64] value_net.b[1] | log_std[2]
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brb.h"

#define PH 64                       // hidden width
#define PIN 6                       // observation size
#define TOWER (PH * PIN + PH + PH * PH + PH)
#define OFF_PI 0
#define OFF_VF TOWER
#define OFF_AW (2 * TOWER)
#define OFF_AB (OFF_AW + 2 * PH)
#define OFF_VW (OFF_AB + 2)
#define OFF_VB (OFF_VW + PH)
#define OFF_LS (OFF_VB + 1)
#define NPARAM (OFF_LS + 2)         // 9,413 = BRB_POLICY_NPARAM

static_assert(NPARAM == BRB_POLICY_NPARAM, "parameter block layout");

// tanh to ~2e-7 absolute: 1 - 2 / (exp(2x) + 1) with the fast exponential and reciprocal (the towers' pre-activations are
// O(1); torch.tanh differs by less than the fp32 rounding of the following layer)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

// one 6-64-64 tower; two rows of W3 are applied to the second hidden layer as it is produced (the critic passes its single
// row twice and ignores the second result, so that actor and critic share one instantiation of this code and run one after
// the other: inlined side by side the compiler interleaves them and 2 x 64 first-layer registers spill).
// sp = shared-memory copy of the parameter block.
__device__ __forceinline__ void tower(const float *__restrict__ sp, int off, int off_w3a, int off_w3b, float b3a, float b3b,
                                      const float (&x)[PIN], float &outa, float &outb) {
  float h1[PH];
  const float *W1 = sp + off, *b1 = W1 + PH * PIN, *W2 = b1 + PH, *b2 = W2 + PH * PH;
#pragma unroll
  for (int k = 0; k < PH; k++) {
    float a = b1[k];
#pragma unroll
    for (int i = 0; i < PIN; i++) a = fmaf(W1[k * PIN + i], x[i], a);
    h1[k] = tanh_fast(a);
  }
  float oa = b3a, ob = b3b;
#pragma unroll 1
  for (int j = 0; j < PH; j += 4) {                 // four output units at a time: four independent FMA chains
    float a0 = b2[j], a1 = b2[j + 1], a2 = b2[j + 2], a3 = b2[j + 3];
    const float4 *r0 = reinterpret_cast<const float4 *>(W2 + (j + 0) * PH), *r1 = reinterpret_cast<const float4 *>(W2 + (j + 1) * PH);
    const float4 *r2 = reinterpret_cast<const float4 *>(W2 + (j + 2) * PH), *r3 = reinterpret_cast<const float4 *>(W2 + (j + 3) * PH);
#pragma unroll
    for (int k4 = 0; k4 < PH / 4; k4++) {
      const float4 w0 = r0[k4], w1 = r1[k4], w2 = r2[k4], w3 = r3[k4];
      const float u0 = h1[4 * k4], u1 = h1[4 * k4 + 1], u2 = h1[4 * k4 + 2], u3 = h1[4 * k4 + 3];
      a0 = fmaf(w0.x, u0, a0); a0 = fmaf(w0.y, u1, a0); a0 = fmaf(w0.z, u2, a0); a0 = fmaf(w0.w, u3, a0);
      a1 = fmaf(w1.x, u0, a1); a1 = fmaf(w1.y, u1, a1); a1 = fmaf(w1.z, u2, a1); a1 = fmaf(w1.w, u3, a1);
      a2 = fmaf(w2.x, u0, a2); a2 = fmaf(w2.y, u1, a2); a2 = fmaf(w2.z, u2, a2); a2 = fmaf(w2.w, u3, a2);
      a3 = fmaf(w3.x, u0, a3); a3 = fmaf(w3.y, u1, a3); a3 = fmaf(w3.z, u2, a3); a3 = fmaf(w3.w, u3, a3);
    }
    const float g0 = tanh_fast(a0), g1 = tanh_fast(a1), g2 = tanh_fast(a2), g3 = tanh_fast(a3);
    const float *wa = sp + off_w3a + j, *wb = sp + off_w3b + j;
    oa = fmaf(wa[0], g0, oa); oa = fmaf(wa[1], g1, oa); oa = fmaf(wa[2], g2, oa); oa = fmaf(wa[3], g3, oa);
    ob = fmaf(wb[0], g0, ob); ob = fmaf(wb[1], g1, ob); ob = fmaf(wb[2], g2, ob); ob = fmaf(wb[3], g3, ob);
  }
  outa = oa; outb = ob;
}

// actions = mean + noise * exp(log_std) (noise == NULL: deterministic, actions = mean); log_prob of the UNCLIPPED action
// (SB3 stores the unclipped sample in the rollout buffer and clips only what the env receives).
__global__ void __launch_bounds__(128) brb_policy_act_kernel(const float *__restrict__ params, const float *__restrict__ obs,
                                                             const float *__restrict__ noise, long long n, float *__restrict__ actions,
                                                             float *__restrict__ actions_clipped, float *__restrict__ values,
                                                             float *__restrict__ logp) {
  __shared__ __align__(16) float sp[(NPARAM + 3) & ~3];
  for (int k = threadIdx.x; k < NPARAM; k += blockDim.x) sp[k] = params[k];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x[PIN];
#pragma unroll
    for (int k = 0; k < PIN; k++) x[k] = obs[i * PIN + k];
    float res[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 1
    for (int t = actions ? 0 : 1; t < 2; t++)        // actions == NULL: critic only (values of terminal observations); t = 0 actor (rows 0, 1 of action_net), t = 1 critic (value_net's row twice)
      tower(sp, t ? OFF_VF : OFF_PI, t ? OFF_VW : OFF_AW, t ? OFF_VW : OFF_AW + PH, sp[t ? OFF_VB : OFF_AB], sp[t ? OFF_VB : OFF_AB + 1], x,
            res[t][0], res[t][1]);
    const float mean[2] = {res[0][0], res[0][1]}, v[1] = {res[1][0]};
    values[i] = v[0];
    if (!actions) continue;
    float lp = 0.f;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float ls = sp[OFF_LS + k];
      const float a = noise ? fmaf(noise[i * 2 + k], expf(ls), mean[k]) : mean[k];
      const float d = a - mean[k];
      lp += -(d * d) / (2.f * expf(2.f * ls)) - ls - 0.91893853320467274f;     // 0.5 log(2 pi)
      actions[i * 2 + k] = a;
      if (actions_clipped) actions_clipped[i * 2 + k] = fminf(1.f, fmaxf(-1.f, a));
    }
    logp[i] = lp;
  }
}

extern "C" int brb_policy_act(const float *params, const float *obs, const float *noise, int64_t n, float *actions, float *actions_clipped,
                              float *values, float *logp, void *stream) {
  if (!params || !obs || !values || (actions && !logp) || n < 0) return BRB_EINVAL;
  if (n == 0) return BRB_OK;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  const long long tiles = (n + 127) / 128, cap = (long long)sms * 8;      // grid-stride: a few CTAs per SM re-use their weight copy
  brb_policy_act_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 128, 0, (cudaStream_t)stream>>>(params, obs, noise, n, actions, actions_clipped,
                                                                                               values, logp);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}

// ================================================================================================================
// PPO minibatch gradient: forward + clipped-surrogate / value loss + backward of both towers, fused.
//
// Replaces, for one minibatch of SB3's PPO.train() (third party; reference src/sb_rl.py:63-71 uses its defaults):
//   values, log_prob, entropy = policy.evaluate_actions(obs[idx], actions[idx])
//   ratio = exp(log_prob - old_log_prob[idx]);  adv = normalised advantages[idx]
//   policy_loss = -min(adv ratio, adv clamp(ratio, 1 - c, 1 + c)).mean();  value_loss = mse(returns[idx], values)
//   loss = policy_loss + vf_coef value_loss - ent_coef entropy.mean();  loss.backward()
// Output: d loss / d params in the layout of the parameter block (accumulated with atomics into a zeroed buffer) and
// the sums behind SB3's logged policy_loss / value_loss / approx_kl / clip_fraction.
//
// One kernel instance per tower (actor / critic).  A CTA (128 threads) walks over tiles of 128 samples.  Phases that are
// per sample (forward, d tanh, W2' dz2) run one sample per thread with the 64-wide vectors in registers and the weights
// broadcast from shared memory; phases that reduce over samples (dW = dz' h) are small register-tiled GEMMs over the
// tile through shared memory, each thread owning a fixed block of the weight gradient in registers for the whole launch.
// Tile rows have a stride of 68 floats: 16-byte aligned, and a float4 access at the same column of 8 consecutive rows
// touches 32 distinct banks.
#define TS 68
#define PPO_SMEM_FLOATS (PH * PIN + PH + PH * PH + PH + 2 * PH + PH * PH + 2 * 128 * TS + 128 * 8 + 128 * 2)

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

template <int ACTOR>
__global__ void __launch_bounds__(128, 2) brb_ppo_grad_kernel(const float *__restrict__ params, const float *__restrict__ obs,
                                                              const float *__restrict__ act, const float *__restrict__ oldlogp,
                                                              const float *__restrict__ adv, const float *__restrict__ ret,
                                                              const long long *__restrict__ idx, long long mb,
                                                              const float *__restrict__ adv_stats, float clip, float vf_coef, float ent_coef,
                                                              float *__restrict__ grad, float *__restrict__ stats) {
  constexpr int NOUT = ACTOR ? 2 : 1;
  constexpr int OFF = ACTOR ? OFF_PI : OFF_VF, OFFW3 = ACTOR ? OFF_AW : OFF_VW, OFFB3 = ACTOR ? OFF_AB : OFF_VB;
  extern __shared__ __align__(16) float sm[];
  float *W1 = sm, *b1 = W1 + PH * PIN, *W2 = b1 + PH, *b2 = W2 + PH * PH, *W3 = b2 + PH, *W2T = W3 + 2 * PH;
  float *A = W2T + PH * PH, *Bt = A + 128 * TS, *X = Bt + 128 * TS, *D = X + 128 * 8;
  const int tid = threadIdx.x;
  for (int k = tid; k < PH * PIN + PH + PH * PH + PH; k += 128) sm[k] = params[OFF + k];        // W1 b1 W2 b2 are contiguous
  for (int k = tid; k < 2 * PH; k += 128) W3[k] = k < NOUT * PH ? params[OFFW3 + k] : 0.f;
  __syncthreads();
  for (int k = tid; k < PH * PH; k += 128) W2T[(k & 63) * PH + (k >> 6)] = W2[k];
  __syncthreads();
  const float inv_mb = 1.f / (float)mb;
  const float amean = adv_stats[0], arstd = adv_stats[1];
  float ls[2] = {0.f, 0.f}, ivar[2] = {1.f, 1.f};
  if (ACTOR) {
#pragma unroll
    for (int k = 0; k < 2; k++) { ls[k] = params[OFF_LS + k]; ivar[k] = expf(-2.f * ls[k]); }
  }
  float b3[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; o++) b3[o] = params[OFFB3 + o];

  // weight-gradient blocks owned by this thread for the whole launch
  float gW2[4][8], gb2[4] = {0.f, 0.f, 0.f, 0.f}, gW1[3] = {0.f, 0.f, 0.f}, gb1 = 0.f, gW3 = 0.f, gb3 = 0.f, gls[2] = {0.f, 0.f};
  float st0 = 0.f, st1 = 0.f, st2 = 0.f;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int q = 0; q < 8; q++) gW2[r][q] = 0.f;
  const int j0 = 4 * (tid >> 3), kA = 4 * (tid & 7), kB = kA + 32;      // dW2 block: rows j0..j0+3, columns kA..kA+3 and kB..kB+3
  const int w3o = tid >> 6, w3j = tid & 63;                              // dW3 element
  const int w1k = tid & 63, w1i = 3 * (tid >> 6);                        // dW1 elements (k, i0..i0+2)

  const long long ntiles = (mb + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s = tile * 128 + tid;
    const bool valid = s < mb;
    const long long i = idx[valid ? s : 0];
    // ---- P1: forward, one sample per thread
    float x[PIN];
#pragma unroll
    for (int k = 0; k < PIN; k++) { x[k] = obs[i * PIN + k]; X[tid * 8 + k] = x[k]; }
    float dout[NOUT];
    {
      float h1[PH];
#pragma unroll
      for (int k = 0; k < PH; k++) {
        float a = b1[k];
#pragma unroll
        for (int q = 0; q < PIN; q++) a = fmaf(W1[k * PIN + q], x[q], a);
        h1[k] = tanh_fast(a);
      }
#pragma unroll
      for (int k4 = 0; k4 < PH / 4; k4++)
        *reinterpret_cast<float4 *>(Bt + tid * TS + 4 * k4) = make_float4(h1[4 * k4], h1[4 * k4 + 1], h1[4 * k4 + 2], h1[4 * k4 + 3]);
      float out[NOUT];
#pragma unroll
      for (int o = 0; o < NOUT; o++) out[o] = b3[o];
#pragma unroll 1
      for (int j = 0; j < PH; j += 4) {
        float a0 = b2[j], a1 = b2[j + 1], a2 = b2[j + 2], a3 = b2[j + 3];
        const float4 *r0 = reinterpret_cast<const float4 *>(W2 + (j + 0) * PH), *r1 = reinterpret_cast<const float4 *>(W2 + (j + 1) * PH);
        const float4 *r2 = reinterpret_cast<const float4 *>(W2 + (j + 2) * PH), *r3 = reinterpret_cast<const float4 *>(W2 + (j + 3) * PH);
#pragma unroll
        for (int k4 = 0; k4 < PH / 4; k4++) {
          const float4 w0 = r0[k4], w1 = r1[k4], w2 = r2[k4], w3 = r3[k4];
          const float u0 = h1[4 * k4], u1 = h1[4 * k4 + 1], u2 = h1[4 * k4 + 2], u3 = h1[4 * k4 + 3];
          a0 = fmaf(w0.x, u0, a0); a0 = fmaf(w0.y, u1, a0); a0 = fmaf(w0.z, u2, a0); a0 = fmaf(w0.w, u3, a0);
          a1 = fmaf(w1.x, u0, a1); a1 = fmaf(w1.y, u1, a1); a1 = fmaf(w1.z, u2, a1); a1 = fmaf(w1.w, u3, a1);
          a2 = fmaf(w2.x, u0, a2); a2 = fmaf(w2.y, u1, a2); a2 = fmaf(w2.z, u2, a2); a2 = fmaf(w2.w, u3, a2);
          a3 = fmaf(w3.x, u0, a3); a3 = fmaf(w3.y, u1, a3); a3 = fmaf(w3.z, u2, a3); a3 = fmaf(w3.w, u3, a3);
        }
        const float g0 = tanh_fast(a0), g1 = tanh_fast(a1), g2 = tanh_fast(a2), g3 = tanh_fast(a3);
        *reinterpret_cast<float4 *>(A + tid * TS + j) = make_float4(g0, g1, g2, g3);
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
          const float *w = W3 + o * PH + j;
          out[o] = fmaf(w[0], g0, out[o]); out[o] = fmaf(w[1], g1, out[o]); out[o] = fmaf(w[2], g2, out[o]); out[o] = fmaf(w[3], g3, out[o]);
        }
      }
      // ---- loss and its derivative with respect to the tower outputs
      const float m = valid ? inv_mb : 0.f;
      if (ACTOR) {
        float d[2], lp = 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          d[k] = act[i * 2 + k] - out[k];
          lp += -(d[k] * d[k]) * (0.5f * ivar[k]) - ls[k] - 0.91893853320467274f;
        }
        const float lr = lp - oldlogp[i], ratio = expf(lr);
        const float an = (adv[i] - amean) * arstd;
        const float s1 = an * ratio, s2 = an * fminf(1.f + clip, fmaxf(1.f - clip, ratio));
        const bool inside = ratio >= 1.f - clip && ratio <= 1.f + clip;
        const float g = (inside || s1 < s2) ? -an * ratio * m : 0.f;        // d(-min(s1, s2)) / d log_prob / mb
#pragma unroll
        for (int k = 0; k < 2; k++) {
          dout[k] = g * d[k] * ivar[k];                                      // d log_prob / d mean_k = (a - mean) / var
          gls[k] += g * (d[k] * d[k] * ivar[k] - 1.f);                       // d log_prob / d log_std_k
        }
        st0 += -fminf(s1, s2) * m;
        st1 += ((ratio - 1.f) - lr) * m;
        st2 += (fabsf(ratio - 1.f) > clip ? 1.f : 0.f) * m;
      } else {
        const float diff = out[0] - ret[i];
        dout[0] = vf_coef * 2.f * diff * m;
        st0 += diff * diff * m;
      }
#pragma unroll
      for (int o = 0; o < 2; o++) D[tid * 2 + o] = o < NOUT ? dout[o < NOUT ? o : 0] : 0.f;
    }
    __syncthreads();
    // ---- P2: dW3 += dout' h2, db3 += sum dout
    if (w3o < NOUT) {
      float a = 0.f;
#pragma unroll 8
      for (int q = 0; q < 128; q++) a = fmaf(D[q * 2 + w3o], A[q * TS + w3j], a);
      gW3 += a;
    }
    if (tid < NOUT) {
      float a = 0.f;
      for (int q = 0; q < 128; q++) a += D[q * 2 + tid];
      gb3 += a;
    }
    __syncthreads();
    // ---- P3: dz2 = (W3' dout) (1 - h2^2), kept in registers and written over h2
    float dz2[PH];
#pragma unroll
    for (int k4 = 0; k4 < PH / 4; k4++) {
      const float4 h = *reinterpret_cast<const float4 *>(A + tid * TS + 4 * k4);
      const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float dh = 0.f;
#pragma unroll
        for (int o = 0; o < NOUT; o++) dh = fmaf(W3[o * PH + 4 * k4 + r], dout[o], dh);
        dz2[4 * k4 + r] = dh * (1.f - hh[r] * hh[r]);
      }
      *reinterpret_cast<float4 *>(A + tid * TS + 4 * k4) = make_float4(dz2[4 * k4], dz2[4 * k4 + 1], dz2[4 * k4 + 2], dz2[4 * k4 + 3]);
    }
    __syncthreads();
    // ---- P4: dW2 += dz2' h1 (4 x 8 block per thread), db2 += sum dz2
#pragma unroll 4
    for (int q = 0; q < 128; q++) {
      const float4 a = *reinterpret_cast<const float4 *>(A + q * TS + j0);
      const float4 u = *reinterpret_cast<const float4 *>(Bt + q * TS + kA), v = *reinterpret_cast<const float4 *>(Bt + q * TS + kB);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 8; c++) gW2[r][c] = fmaf(aa[r], bb[c], gW2[r][c]);
        gb2[r] += aa[r];
      }
    }
    __syncthreads();
    // ---- P5: dz1 = (W2' dz2) (1 - h1^2), written over dz2's tile
#pragma unroll 1
    for (int k = 0; k < PH; k += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4 *r0 = reinterpret_cast<const float4 *>(W2T + (k + 0) * PH), *r1 = reinterpret_cast<const float4 *>(W2T + (k + 1) * PH);
      const float4 *r2 = reinterpret_cast<const float4 *>(W2T + (k + 2) * PH), *r3 = reinterpret_cast<const float4 *>(W2T + (k + 3) * PH);
#pragma unroll
      for (int q4 = 0; q4 < PH / 4; q4++) {
        const float4 w0 = r0[q4], w1 = r1[q4], w2 = r2[q4], w3 = r3[q4];
        const float u0 = dz2[4 * q4], u1 = dz2[4 * q4 + 1], u2 = dz2[4 * q4 + 2], u3 = dz2[4 * q4 + 3];
        a0 = fmaf(w0.x, u0, a0); a0 = fmaf(w0.y, u1, a0); a0 = fmaf(w0.z, u2, a0); a0 = fmaf(w0.w, u3, a0);
        a1 = fmaf(w1.x, u0, a1); a1 = fmaf(w1.y, u1, a1); a1 = fmaf(w1.z, u2, a1); a1 = fmaf(w1.w, u3, a1);
        a2 = fmaf(w2.x, u0, a2); a2 = fmaf(w2.y, u1, a2); a2 = fmaf(w2.z, u2, a2); a2 = fmaf(w2.w, u3, a2);
        a3 = fmaf(w3.x, u0, a3); a3 = fmaf(w3.y, u1, a3); a3 = fmaf(w3.z, u2, a3); a3 = fmaf(w3.w, u3, a3);
      }
      const float4 h = *reinterpret_cast<const float4 *>(Bt + tid * TS + k);
      *reinterpret_cast<float4 *>(A + tid * TS + k) =
          make_float4(a0 * (1.f - h.x * h.x), a1 * (1.f - h.y * h.y), a2 * (1.f - h.z * h.z), a3 * (1.f - h.w * h.w));
    }
    __syncthreads();
    // ---- P6: dW1 += dz1' x, db1 += sum dz1
    {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, sb = 0.f;
#pragma unroll 8
      for (int q = 0; q < 128; q++) {
        const float dz = A[q * TS + w1k];
        a0 = fmaf(dz, X[q * 8 + w1i], a0); a1 = fmaf(dz, X[q * 8 + w1i + 1], a1); a2 = fmaf(dz, X[q * 8 + w1i + 2], a2);
        sb += dz;
      }
      gW1[0] += a0; gW1[1] += a1; gW1[2] += a2; gb1 += sb;
    }
    __syncthreads();
  }

  // ---- flush this CTA's partial gradient
#pragma unroll
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      atomicAdd(grad + OFF + PH * PIN + PH + (j0 + r) * PH + kA + c, gW2[r][c]);
      atomicAdd(grad + OFF + PH * PIN + PH + (j0 + r) * PH + kB + c, gW2[r][4 + c]);
    }
    if ((tid & 7) == 0) atomicAdd(grad + OFF + PH * PIN + PH + PH * PH + j0 + r, gb2[r]);
  }
#pragma unroll
  for (int r = 0; r < 3; r++) atomicAdd(grad + OFF + w1k * PIN + w1i + r, gW1[r]);
  if (tid < PH) atomicAdd(grad + OFF + PH * PIN + tid, gb1);
  if (w3o < NOUT) atomicAdd(grad + OFFW3 + w3o * PH + w3j, gW3);
  if (tid < NOUT) atomicAdd(grad + OFFB3 + tid, gb3);
  const float r0 = warp_sum_f(st0), r1 = warp_sum_f(st1), r2 = warp_sum_f(st2), l0 = warp_sum_f(gls[0]), l1 = warp_sum_f(gls[1]);
  if ((tid & 31) == 0) {
    if (ACTOR) {
      atomicAdd(stats + 0, r0); atomicAdd(stats + 2, r1); atomicAdd(stats + 3, r2);
      atomicAdd(grad + OFF_LS, l0); atomicAdd(grad + OFF_LS + 1, l1);
    } else {
      atomicAdd(stats + 1, r0);
    }
  }
  if (ACTOR && blockIdx.x == 0 && tid < 2 && ent_coef != 0.f) atomicAdd(grad + OFF_LS + tid, -ent_coef);   // entropy = sum(log_std) + const
}

extern "C" int brb_ppo_grad(const float *params, const float *obs, const float *actions, const float *old_logp, const float *adv,
                            const float *returns, const int64_t *idx, int64_t mb, const float *adv_stats, float clip_range, float vf_coef,
                            float ent_coef, float *grad, float *stats, void *stream) {
  if (!params || !obs || !actions || !old_logp || !adv || !returns || !idx || !adv_stats || !grad || !stats || mb <= 0) return BRB_EINVAL;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  const size_t smem = PPO_SMEM_FLOATS * sizeof(float);
  if (cudaFuncSetAttribute(brb_ppo_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(brb_ppo_grad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  const long long tiles = (mb + 127) / 128, cap = 2LL * sms;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  cudaStream_t s = (cudaStream_t)stream;
  brb_ppo_grad_kernel<1><<<grid, 128, smem, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats, clip_range,
                                                 vf_coef, ent_coef, grad, stats);
  brb_ppo_grad_kernel<0><<<grid, 128, smem, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats, clip_range,
                                                 vf_coef, ent_coef, grad, stats);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}
