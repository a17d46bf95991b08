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
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/brb.h"

#include "brb_policy_layout.h"

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
                                                             float *__restrict__ logp, const uint8_t *__restrict__ mask) {
  __shared__ __align__(16) float sp[(NPARAM + 3) & ~3];
  for (int k = threadIdx.x; k < NPARAM; k += blockDim.x) sp[k] = params[k];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) { values[i] = 0.f; continue; }      // masked critic call: only the flagged rows are evaluated
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
                                                                                               values, logp, nullptr);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}

// Critic only, and only for the rows whose mask byte is non-zero (values of the others = 0): V(terminal_observation) for the
// TimeLimit bootstrap, where at most N / max_episode_steps rows per step are truncated.  No host-side test of "any truncated?"
// (that would put a sync into every step of the rollout): unflagged warps leave after one byte load per lane.
extern "C" int brb_policy_value_masked(const float *params, const float *obs, const uint8_t *mask, int64_t n, float *values, void *stream) {
  if (!params || !obs || !mask || !values || n < 0) return BRB_EINVAL;
  if (n == 0) return BRB_OK;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  const long long tiles = (n + 127) / 128, cap = (long long)sms * 8;
  brb_policy_act_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 128, 0, (cudaStream_t)stream>>>(params, obs, nullptr, n, nullptr, nullptr, values,
                                                                                               nullptr, mask);
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
// One kernel instance per tower (actor / critic).  A CTA (128 threads) walks over tiles of 128 samples and every 64-wide
// phase is a register-tiled GEMM over the tile: activations are stored TRANSPOSED in shared memory (row = hidden unit,
// column = sample, row stride 132 floats = 33 x 16 B, so float4 reads of 8 consecutive rows at one column hit 32 distinct
// banks), each thread owns an 8 (units) x 8 (samples) block of the 64 x 128 result, and one float4 of weights plus one of
// activations feed 16 FMAs.  Weight gradients are owned by fixed threads in registers for the whole launch and flushed
// with one atomicAdd per element per CTA.  (A first version ran the forward / backward matrix-vector products one sample
// per thread with the weights broadcast from shared memory: 4 FMAs per weight float4, LSU pipe 41 %, FMA pipe 45 %,
// 1.00 ms per tower per 1M samples; this one: LSU 21 %, FMA 52 %, 0.87 ms.)
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

#define TS2 132
#define PPO2_SMEM_FLOATS (PH * PIN + PH + PH * PH + PH + 2 * PH + PH * PH + 2 * PH * TS2 + 8 * TS2 + 2 * TS2)

// acc[r][c] += sum_q Wq[q][r0 + r] * T[q][s_c],  s_c = sA..sA+3, sB..sB+3   (Wq: [64][64], row = reduction index)
__device__ __forceinline__ void tile_gemm(const float *__restrict__ Wq, const float *__restrict__ T, int r0, int sA, int sB, float (&acc)[8][8]) {
#pragma unroll 2
  for (int q = 0; q < PH; q++) {
    const float4 a0 = *reinterpret_cast<const float4 *>(Wq + q * PH + r0), a1 = *reinterpret_cast<const float4 *>(Wq + q * PH + r0 + 4);
    const float4 u = *reinterpret_cast<const float4 *>(T + q * TS2 + sA), v = *reinterpret_cast<const float4 *>(T + q * TS2 + sB);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c < 8; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
  }
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
  float *H1T = W2T + PH * PH, *H2T = H1T + PH * TS2, *XT = H2T + PH * TS2, *D = XT + 8 * TS2;
  const int tid = threadIdx.x;
  for (int k = tid; k < PH * PIN + PH + PH * PH + PH; k += 128) sm[k] = params[OFF + k];
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

  float gW2[4][8], gb2[4] = {0.f, 0.f, 0.f, 0.f}, gW1[3] = {0.f, 0.f, 0.f}, gb1 = 0.f, gW3 = 0.f, gb3 = 0.f, gls[2] = {0.f, 0.f};
  float st0 = 0.f, st1 = 0.f, st2 = 0.f;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int q = 0; q < 8; q++) gW2[r][q] = 0.f;
  const int r0 = 8 * (tid >> 4), sA = 4 * (tid & 15), sB = sA + 64;      // 8 x 8 block of a 64 x 128 tile result
  const int j0 = 4 * (tid >> 3), kk = tid & 7;                           // dW2 block: rows j0..j0+3, columns kk + 8c (consecutive tile rows
                                                                         // across the lanes of a quarter-warp: conflict-free float4 reads)
  const int w3o = tid >> 6, w3j = tid & 63;
  const int w1k = tid & 63, w1i = 3 * (tid >> 6);

  const long long ntiles = (mb + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s = tile * 128 + tid;
    const bool valid = s < mb;
    const long long i = idx[valid ? s : 0];
    // ---- A: first layer, one sample per thread, written transposed
    {
      float x[PIN];
#pragma unroll
      for (int k = 0; k < PIN; k++) { x[k] = obs[i * PIN + k]; XT[k * TS2 + tid] = x[k]; }
#pragma unroll 8
      for (int k = 0; k < PH; k++) {
        float a = b1[k];
#pragma unroll
        for (int q = 0; q < PIN; q++) a = fmaf(W1[k * PIN + q], x[q], a);
        H1T[k * TS2 + tid] = tanh_fast(a);
      }
    }
    __syncthreads();
    // ---- B: h2 = tanh(W2 h1 + b2) for the tile
    {
      float acc[8][8];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
      tile_gemm(W2T, H1T, r0, sA, sB, acc);
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float bb = b2[r0 + r];
        *reinterpret_cast<float4 *>(H2T + (r0 + r) * TS2 + sA) =
            make_float4(tanh_fast(acc[r][0] + bb), tanh_fast(acc[r][1] + bb), tanh_fast(acc[r][2] + bb), tanh_fast(acc[r][3] + bb));
        *reinterpret_cast<float4 *>(H2T + (r0 + r) * TS2 + sB) =
            make_float4(tanh_fast(acc[r][4] + bb), tanh_fast(acc[r][5] + bb), tanh_fast(acc[r][6] + bb), tanh_fast(acc[r][7] + bb));
      }
    }
    __syncthreads();
    // ---- C: output layer and loss, one sample per thread
    {
      float out[NOUT], dout[NOUT];
#pragma unroll
      for (int o = 0; o < NOUT; o++) out[o] = b3[o];
#pragma unroll 8
      for (int j = 0; j < PH; j++) {
        const float h = H2T[j * TS2 + tid];
#pragma unroll
        for (int o = 0; o < NOUT; o++) out[o] = fmaf(W3[o * PH + j], h, out[o]);
      }
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
        const float g = (inside || s1 < s2) ? -an * ratio * m : 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          dout[k] = g * d[k] * ivar[k];
          gls[k] += g * (d[k] * d[k] * ivar[k] - 1.f);
        }
        st0 += -fminf(s1, s2) * m;
        st1 += ((ratio - 1.f) - lr) * m;
        st2 += (fabsf(ratio - 1.f) > clip ? 1.f : 0.f) * m;
      } else {
        const float diff = out[0] - ret[i];
        dout[0] = vf_coef * 2.f * diff * m;
        st0 += diff * diff * m;
      }
      D[tid] = dout[0];
      D[TS2 + tid] = NOUT > 1 ? dout[NOUT - 1] : 0.f;
    }
    __syncthreads();
    // ---- D: dW3 += dout h2', db3 += sum dout
    if (w3o < NOUT) {
      float a = 0.f;
#pragma unroll 4
      for (int q = 0; q < 128; q += 4) {
        const float4 dd = *reinterpret_cast<const float4 *>(D + w3o * TS2 + q), hh = *reinterpret_cast<const float4 *>(H2T + w3j * TS2 + q);
        a = fmaf(dd.x, hh.x, a); a = fmaf(dd.y, hh.y, a); a = fmaf(dd.z, hh.z, a); a = fmaf(dd.w, hh.w, a);
      }
      gW3 += a;
    }
    if (tid < NOUT) {
      float a = 0.f;
      for (int q = 0; q < 128; q++) a += D[tid * TS2 + q];
      gb3 += a;
    }
    __syncthreads();
    // ---- D2: dz2 = (W3' dout)(1 - h2^2), in place over h2 (each thread its own 8 x 8 block)
    {
      const float4 d0A = *reinterpret_cast<const float4 *>(D + sA), d0B = *reinterpret_cast<const float4 *>(D + sB);
      const float4 d1A = *reinterpret_cast<const float4 *>(D + TS2 + sA), d1B = *reinterpret_cast<const float4 *>(D + TS2 + sB);
      const float d0[8] = {d0A.x, d0A.y, d0A.z, d0A.w, d0B.x, d0B.y, d0B.z, d0B.w}, d1[8] = {d1A.x, d1A.y, d1A.z, d1A.w, d1B.x, d1B.y, d1B.z, d1B.w};
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float w0 = W3[r0 + r], w1 = W3[PH + r0 + r];       // second row is zero for the critic
        float4 hA = *reinterpret_cast<const float4 *>(H2T + (r0 + r) * TS2 + sA), hB = *reinterpret_cast<const float4 *>(H2T + (r0 + r) * TS2 + sB);
        float h[8] = {hA.x, hA.y, hA.z, hA.w, hB.x, hB.y, hB.z, hB.w};
#pragma unroll
        for (int c = 0; c < 8; c++) h[c] = fmaf(w1, d1[c], w0 * d0[c]) * (1.f - h[c] * h[c]);
        *reinterpret_cast<float4 *>(H2T + (r0 + r) * TS2 + sA) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(H2T + (r0 + r) * TS2 + sB) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
    __syncthreads();
    // ---- E: dW2 += dz2 h1' (4 x 8 block per thread, reduction over the samples), db2 += sum dz2
#pragma unroll 2
    for (int q = 0; q < 128; q += 4) {
      float4 a[4], b[8];
#pragma unroll
      for (int r = 0; r < 4; r++) a[r] = *reinterpret_cast<const float4 *>(H2T + (j0 + r) * TS2 + q);
#pragma unroll
      for (int c = 0; c < 8; c++) b[c] = *reinterpret_cast<const float4 *>(H1T + (kk + 8 * c) * TS2 + q);
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
          gW2[r][c] = fmaf(a[r].x, b[c].x, gW2[r][c]); gW2[r][c] = fmaf(a[r].y, b[c].y, gW2[r][c]);
          gW2[r][c] = fmaf(a[r].z, b[c].z, gW2[r][c]); gW2[r][c] = fmaf(a[r].w, b[c].w, gW2[r][c]);
        }
        gb2[r] += (a[r].x + a[r].y) + (a[r].z + a[r].w);
      }
    }
    __syncthreads();
    // ---- F: dz1 = (W2' dz2)(1 - h1^2), in place over h1
    {
      float acc[8][8];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
      tile_gemm(W2, H2T, r0, sA, sB, acc);
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float4 hA = *reinterpret_cast<const float4 *>(H1T + (r0 + r) * TS2 + sA), hB = *reinterpret_cast<const float4 *>(H1T + (r0 + r) * TS2 + sB);
        *reinterpret_cast<float4 *>(H1T + (r0 + r) * TS2 + sA) =
            make_float4(acc[r][0] * (1.f - hA.x * hA.x), acc[r][1] * (1.f - hA.y * hA.y), acc[r][2] * (1.f - hA.z * hA.z), acc[r][3] * (1.f - hA.w * hA.w));
        *reinterpret_cast<float4 *>(H1T + (r0 + r) * TS2 + sB) =
            make_float4(acc[r][4] * (1.f - hB.x * hB.x), acc[r][5] * (1.f - hB.y * hB.y), acc[r][6] * (1.f - hB.z * hB.z), acc[r][7] * (1.f - hB.w * hB.w));
      }
    }
    __syncthreads();
    // ---- G: dW1 += dz1 x', db1 += sum dz1
    {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, sb = 0.f;
#pragma unroll 4
      for (int q = 0; q < 128; q += 4) {
        const float4 dz = *reinterpret_cast<const float4 *>(H1T + w1k * TS2 + q);
        const float4 x0 = *reinterpret_cast<const float4 *>(XT + w1i * TS2 + q), x1 = *reinterpret_cast<const float4 *>(XT + (w1i + 1) * TS2 + q);
        const float4 x2 = *reinterpret_cast<const float4 *>(XT + (w1i + 2) * TS2 + q);
        a0 = fmaf(dz.x, x0.x, a0); a0 = fmaf(dz.y, x0.y, a0); a0 = fmaf(dz.z, x0.z, a0); a0 = fmaf(dz.w, x0.w, a0);
        a1 = fmaf(dz.x, x1.x, a1); a1 = fmaf(dz.y, x1.y, a1); a1 = fmaf(dz.z, x1.z, a1); a1 = fmaf(dz.w, x1.w, a1);
        a2 = fmaf(dz.x, x2.x, a2); a2 = fmaf(dz.y, x2.y, a2); a2 = fmaf(dz.z, x2.z, a2); a2 = fmaf(dz.w, x2.w, a2);
        sb += (dz.x + dz.y) + (dz.z + dz.w);
      }
      gW1[0] += a0; gW1[1] += a1; gW1[2] += a2; gb1 += sb;
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int c = 0; c < 8; c++) atomicAdd(grad + OFF + PH * PIN + PH + (j0 + r) * PH + kk + 8 * c, gW2[r][c]);
    if (kk == 0) atomicAdd(grad + OFF + PH * PIN + PH + PH * PH + j0 + r, gb2[r]);
  }
#pragma unroll
  for (int r = 0; r < 3; r++) atomicAdd(grad + OFF + w1k * PIN + w1i + r, gW1[r]);
  if (tid < PH) atomicAdd(grad + OFF + PH * PIN + tid, gb1);
  if (w3o < NOUT) atomicAdd(grad + OFFW3 + w3o * PH + w3j, gW3);
  if (tid < NOUT) atomicAdd(grad + OFFB3 + tid, gb3);
  const float q0 = warp_sum_f(st0), q1 = warp_sum_f(st1), q2 = warp_sum_f(st2), l0 = warp_sum_f(gls[0]), l1 = warp_sum_f(gls[1]);
  if ((tid & 31) == 0) {
    if (ACTOR) {
      atomicAdd(stats + 0, q0); atomicAdd(stats + 2, q1); atomicAdd(stats + 3, q2);
      atomicAdd(grad + OFF_LS, l0); atomicAdd(grad + OFF_LS + 1, l1);
    } else {
      atomicAdd(stats + 1, q0);
    }
  }
  if (ACTOR && blockIdx.x == 0 && tid < 2 && ent_coef != 0.f) atomicAdd(grad + OFF_LS + tid, -ent_coef);
}

// tensor-core version (brb_policy_tc.cu)
extern "C" int brb_ppo_grad_tc_launch(const float *params, const float *obs, const float *actions, const float *old_logp, const float *adv,
                                      const float *returns, const int64_t *idx, int64_t mb, const float *adv_stats, float clip_range, float vf_coef,
                                      float ent_coef, float *grad, float *stats, int *fault, int sms, cudaStream_t s);
static int *g_tc_fault[16];      // per device: set to 1 by the tcgen05 kernel if a (bounded) pipeline wait timed out

extern "C" int brb_ppo_tc_fault(int device) {
  if (device < 0 || device >= 16 || !g_tc_fault[device]) return 0;
  int v = 0;
  if (cudaMemcpy(&v, g_tc_fault[device], sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
  return v;
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
  const bool no_tc = getenv("BRB_PPO_NO_TC") != nullptr;            // A/B measurements only: the FFMA version of the same gradient
  if (!no_tc && dev < 16) {
    if (!g_tc_fault[dev]) {
      if (cudaMalloc(&g_tc_fault[dev], sizeof(int)) != cudaSuccess || cudaMemset(g_tc_fault[dev], 0, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return BRB_ENOMEM; }
    }
    return brb_ppo_grad_tc_launch(params, obs, actions, old_logp, adv, returns, idx, mb, adv_stats, clip_range, vf_coef, ent_coef, grad, stats,
                                  g_tc_fault[dev], sms, (cudaStream_t)stream);
  }
  const long long tiles = (mb + 127) / 128, cap = 2LL * sms;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = PPO2_SMEM_FLOATS * sizeof(float);
  if (cudaFuncSetAttribute(brb_ppo_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(brb_ppo_grad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  brb_ppo_grad_kernel<1><<<grid, 128, smem, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats,
                                                     clip_range, vf_coef, ent_coef, grad, stats);
  brb_ppo_grad_kernel<0><<<grid, 128, smem, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats,
                                                     clip_range, vf_coef, ent_coef, grad, stats);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}

// ================================================================================================================
// One optimiser step on the flat parameter block, fused: gradient averaging over the ranks, global-norm clipping and Adam.
//
// Replaces, for one minibatch of SB3's PPO.train() (third party; reference src/sb_rl.py:63-71 runs it with SB3's defaults):
//   th.nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)      total = ||g||_2 ; g *= min(1, max_norm / (total + 1e-6))
//   policy.optimizer.step()                                              torch.optim.Adam (no weight decay, no amsgrad)
// which in PyTorch is ~25 small launches (per-parameter norms, stack, norm, clamp, foreach mul, foreach Adam pieces) and
// a host round trip per tensor list.  The parameter block is 9,413 floats, so one CTA does it in one launch:
//   g = grad * grad_scale (1 / world after the all-reduce sum); block-reduce sum g^2; clip; m, v, p updated in place;
//   grad is zeroed for the next minibatch's accumulation (saves the memset launch).
// `step` is the 1-based Adam step count.  All pointers are device pointers.
__global__ void __launch_bounds__(1024) brb_adam_clip_kernel(float *__restrict__ params, float *__restrict__ grad, float *__restrict__ m,
                                                             float *__restrict__ v, int n, float lr, float beta1, float beta2, float eps,
                                                             float bias1, float bias2_sqrt, float max_norm, float grad_scale,
                                                             float *__restrict__ norm_out) {
  __shared__ float red[32];
  __shared__ float coef_s;
  float ss = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) { const float g = grad[k] * grad_scale; ss = fmaf(g, g, ss); }
  ss = warp_sum_f(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum_f(t);
    if (threadIdx.x == 0) {
      const float total = sqrtf(t);
      coef_s = max_norm > 0.f ? fminf(1.f, max_norm / (total + 1e-6f)) : 1.f;
      if (norm_out) *norm_out = total;
    }
  }
  __syncthreads();
  const float coef = coef_s * grad_scale, step_size = lr / bias1;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float g = grad[k] * coef;
    const float mk = beta1 * m[k] + (1.f - beta1) * g;
    const float vk = beta2 * v[k] + (1.f - beta2) * g * g;
    m[k] = mk; v[k] = vk;
    params[k] -= step_size * mk / (sqrtf(vk) / bias2_sqrt + eps);
    grad[k] = 0.f;
  }
}

extern "C" int brb_adam_clip_step(float *params, float *grad, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
                                  int64_t step, float max_grad_norm, float grad_scale, float *norm_out, void *stream) {
  if (!params || !grad || !m || !v || n <= 0 || n > (1 << 24) || step < 1) return BRB_EINVAL;
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  brb_adam_clip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, m, v, (int)n, lr, beta1, beta2, eps, (float)b1, (float)sqrt(b2),
                                                             max_grad_norm, grad_scale, norm_out);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}

// ================================================================================================================
// Minibatch order of one epoch (SB3 RolloutBuffer.get: np.random.permutation; third party, reference src/sb_rl.py:63-71).
// torch.randperm sorts random keys: 1.6 ms for the 16.8M samples of a 1M-env x 16-step rollout, ten times per update = 15 % of the update.
// Here out[k] = P(k) for a keyed bijection P of [0, n): a balanced Feistel network on ceil(log2 n) bits (rounded up to even) with a
// multiply-xorshift round function, cycle-walked into [0, n) (a point that lands outside is pushed through P again; the domain is
// < 4n, so that takes < 4 rounds on average).  One pass, 8 bytes written per sample, no sort.
__device__ __forceinline__ uint32_t perm_round(uint32_t x, uint32_t k) {
  x = (x ^ k) * 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  return x ^ (x >> 16);
}
__global__ void brb_permutation_kernel(long long *__restrict__ out, long long n, int half_bits, uint32_t k0, uint32_t k1) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  const uint32_t mask = (1u << half_bits) - 1u;
  unsigned long long x = (unsigned long long)tid;
  do {
    uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int round = 0; round < 6; round++) {
      const uint32_t t = l ^ (perm_round(r, k0 + 0x632BE5ABu * (uint32_t)round + (k1 << (round & 7))) & mask);
      l = r; r = t;
    }
    x = ((unsigned long long)l << half_bits) | r;
  } while (x >= (unsigned long long)n);
  out[tid] = (long long)x;
}

extern "C" int brb_random_permutation(int64_t *out, int64_t n, uint64_t seed, void *stream) {
  if (!out || n <= 0 || n > (1ll << 40)) return BRB_EINVAL;
  int bits = 2;
  while ((1ll << bits) < n) bits++;
  bits += bits & 1;                      // balanced halves
  const uint64_t z = (seed + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;     // splitmix-style key schedule
  brb_permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((long long *)out, n, bits / 2, (uint32_t)(z >> 32), (uint32_t)z ^ (uint32_t)(z >> 29));
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}

// ================================================================================================================
// Gradient all-reduce + clipping + Adam as ONE kernel over NVLink peer memory (one process per GPU).
//
// The only exchange step of data-parallel PPO is the sum of the 9,413-float gradient over the ranks, once per minibatch
//  — 37.7 KB, pure latency.  Calling NCCL for it costs a separate launch on its own stream plus the event
// hand-offs around it (measured ~1 ms per minibatch in round 1 against a 6.5 ms gradient kernel; the tensor-core gradient
// kernel made that 30 % of the update).  Here every rank owns a small symmetric block in its own HBM:
//     grad[2][n]   the gradient the rank's brb_ppo_grad accumulates (double-buffered by the parity of the step)
//     flag[world]  flag[r] = last step whose gradient rank r has finished
// The blocks are exported with cudaIpcGetMemHandle and opened by every peer (NVSwitch: every GPU loads from every other at
// full bandwidth).  One CTA per rank then does, in a single launch behind the gradient kernels on the same stream:
//     publish   st.release.sys  peers' flag[me] = step          (the gradient kernels before this launch have completed)
//     wait      ld.acquire.sys  own flag[r] >= step for all r   (bounded spin: a dead peer raises `fault`, never a hang)
//     reduce    g[k] = sum over ranks r = 0..world-1 of peer_r.grad[step & 1][k]   — P2P loads, SAME order on every rank, so the
//               replicas stay bit-identical without a broadcast
//     update    global-norm clipping + Adam on the flat parameter block (as brb_adam_clip_step)
//     re-arm    zero own grad[(step + 1) & 1]: every peer has passed step - 1's barrier, so nobody reads that buffer any more
struct BrbComm {
  int rank, world, device;
  int64_t n;
  float *block;              // own block (cudaMalloc): grad[2][n] then flags
  float *peer[8];            // every rank's block as mapped here (peer[rank] == block)
  unsigned *flags;           // own flags
  int *fault;                // device int
  float **d_peer;            // device copy of peer[]
};

static size_t comm_block_bytes(int64_t n) { return (size_t)(2 * n) * sizeof(float) + 64 * sizeof(unsigned); }

extern "C" int brb_comm_create(int rank, int world, int device, int64_t n, BrbComm **out) {
  if (!out || world < 1 || world > 8 || rank < 0 || rank >= world || n <= 0) return BRB_EINVAL;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return BRB_ECUDA; }
  BrbComm *c = (BrbComm *)calloc(1, sizeof(BrbComm));
  c->rank = rank; c->world = world; c->device = device; c->n = n;
  int rc = BRB_OK;
  if (cudaMalloc(&c->block, comm_block_bytes(n)) != cudaSuccess || cudaMemset(c->block, 0, comm_block_bytes(n)) != cudaSuccess ||
      cudaMalloc(&c->fault, sizeof(int)) != cudaSuccess || cudaMemset(c->fault, 0, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&c->d_peer, 8 * sizeof(float *)) != cudaSuccess) {
    cudaGetLastError();
    rc = BRB_ENOMEM;
  }
  c->flags = (unsigned *)(c->block + 2 * n);
  c->peer[rank] = c->block;
  cudaDeviceSynchronize();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (rc != BRB_OK) { free(c); return rc; }
  *out = c;
  return BRB_OK;
}

// 64-byte IPC handle of the own block (to be all-gathered by the host side)
extern "C" int brb_comm_export(BrbComm *c, void *handle64) {
  if (!c || !handle64) return BRB_EINVAL;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, c->block) != cudaSuccess) { cudaGetLastError(); return BRB_ECUDA; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
  memcpy(handle64, &h, 64);
  return BRB_OK;
}

// handles: world x 64 bytes in rank order (the own entry is ignored)
extern "C" int brb_comm_open(BrbComm *c, const void *handles) {
  if (!c || !handles) return BRB_EINVAL;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return BRB_ECUDA; }
  int rc = BRB_OK;
  for (int r = 0; r < c->world && rc == BRB_OK; r++) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * r, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); rc = BRB_ECUDA; }
    c->peer[r] = (float *)p;
  }
  if (rc == BRB_OK && cudaMemcpy(c->d_peer, c->peer, 8 * sizeof(float *), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); rc = BRB_ECUDA; }
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  return rc;
}

extern "C" void brb_comm_destroy(BrbComm *c) {
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; r++)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->d_peer); cudaFree(c->fault); cudaFree(c->block);
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  free(c);
}

// device pointer of the gradient buffer brb_ppo_grad must accumulate into for optimiser step `step` (1-based)
extern "C" float *brb_comm_grad(BrbComm *c, int64_t step) { return c ? c->block + (size_t)(step & 1) * c->n : nullptr; }

extern "C" int brb_comm_fault(BrbComm *c) {
  if (!c) return -1;
  int v = 0, prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  const cudaError_t e = cudaMemcpy(&v, c->fault, sizeof(int), cudaMemcpyDeviceToHost);
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return v;
}

__global__ void __launch_bounds__(1024) brb_allreduce_adam_kernel(float *const *__restrict__ peer, int rank, int world, int n, unsigned step,
                                                                  float *__restrict__ params, float *__restrict__ m, float *__restrict__ v,
                                                                  float lr, float beta1, float beta2, float eps, float bias1, float bias2_sqrt,
                                                                  float max_norm, float *__restrict__ norm_out, int *__restrict__ fault) {
  __shared__ float red[32];
  __shared__ float coef_s;
  __shared__ int ok_s;
  const int tid = threadIdx.x;
  if (tid == 0) ok_s = 1;
  __syncthreads();
  if (tid < world) {
    // publish: my gradient for `step` is complete (the gradient kernels ran before this launch on the same stream)
    unsigned *theirs = reinterpret_cast<unsigned *>(peer[tid] + 2 * (size_t)n) + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(step) : "memory");
    // wait for rank `tid`'s gradient
    const unsigned *mine = reinterpret_cast<const unsigned *>(peer[rank] + 2 * (size_t)n) + tid;
    bool got = false;
    for (long long spin = 0; spin < (1ll << 27); spin++) {
      unsigned f;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(mine) : "memory");
      if ((int)(f - step) >= 0) { got = true; break; }
      __nanosleep(64);
    }
    if (!got) { ok_s = 0; atomicExch(fault, 1); }
  }
  __syncthreads();
  const size_t off = (size_t)(step & 1u) * (size_t)n;
  const float scale = 1.f / (float)world;
  // reduce in rank order (identical on every rank) — the sum stays in registers: n <= 10 * 1024
  float g[10];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 10; j++) {
    const int k = tid + j * 1024;
    float a = 0.f;
    if (k < n && ok_s) {
      for (int r = 0; r < world; r++) a += __ldcv(peer[r] + off + k);      // volatile-class load: never served from a stale L1 line
    }
    g[j] = a * scale;
    ss = fmaf(g[j], g[j], ss);
  }
  ss = warp_sum_f(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid < 32) {
    float t = red[tid];
    t = warp_sum_f(t);
    if (tid == 0) {
      const float total = sqrtf(t);
      coef_s = max_norm > 0.f ? fminf(1.f, max_norm / (total + 1e-6f)) : 1.f;
      if (norm_out) *norm_out = total;
    }
  }
  __syncthreads();
  const float coef = coef_s, step_size = lr / bias1;
  float *next = peer[rank] + (size_t)((step + 1u) & 1u) * (size_t)n;
#pragma unroll
  for (int j = 0; j < 10; j++) {
    const int k = tid + j * 1024;
    if (k < n) {
      if (ok_s) {
        const float gk = g[j] * coef;
        const float mk = beta1 * m[k] + (1.f - beta1) * gk;
        const float vk = beta2 * v[k] + (1.f - beta2) * gk * gk;
        m[k] = mk; v[k] = vk;
        params[k] -= step_size * mk / (sqrtf(vk) / bias2_sqrt + eps);
      }
      next[k] = 0.f;      // every peer has passed the barrier of step - 1, i.e. finished reading this buffer
    }
  }
}

// One optimiser step across `world` ranks: peer-memory all-reduce (mean) of the gradient in brb_comm_grad(c, step), clipping, Adam.
extern "C" int brb_comm_allreduce_adam(BrbComm *c, float *params, float *m, float *v, float lr, float beta1, float beta2, float eps,
                                       int64_t step, float max_grad_norm, float *norm_out, void *stream) {
  if (!c || !params || !m || !v || step < 1 || c->n > 10 * 1024) return BRB_EINVAL;
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  brb_allreduce_adam_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(c->d_peer, c->rank, c->world, (int)c->n, (unsigned)step, params, m, v, lr, beta1,
                                                                  beta2, eps, (float)b1, (float)sqrt(b2), max_grad_norm, norm_out, c->fault);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}
