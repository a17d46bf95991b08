// brb_policy_tc.cu — the PPO minibatch gradient on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
// Same contract as brb_ppo_grad_kernel in brb_policy.cu (forward + clipped-surrogate / value loss + backward of one tower for the
// samples idx[0..mb) of the rollout buffer; SB3 PPO.train(), third party, reference src/sb_rl.py:63-71), but every 64-wide
// contraction runs as tcgen05 MMAs.  ncu on the FFMA version showed the condition the north star sets for tensor cores: the
// update is contraction-bound on the FP32 pipe (FMA pipe 52 %, 32 TFLOP/s, 58 % of a training iteration).
//
// One CTA = 128 threads = one tile of 128 samples at a time (thread s <-> sample s <-> TMEM lane s), persistent over tiles.
//   P1  h1 = tanh(W1 x + b1)                      CUDA cores (K = 6), written to shared memory as bf16 hi/lo pairs
//   M1  Z2[128x64]   = H1 W2'                      tcgen05  M=128 N=64 K=64   (A K-major, B K-major)
//   P2  h2 = tanh(z2 + b2); out = W3 h2 + b3; loss; dout                       (tcgen05.ld: one accumulator row per thread)
//   M2  dW3[64x8]   += H2' dOut                    tcgen05  M=64  N=8  K=128  (A MN-major, B MN-major), accumulates over tiles
//   P3  dz2 = (W3' dout)(1 - h2^2)                 in registers, then over H2 in shared memory
//   M3  dH1[128x64]  = dZ2 W2                      tcgen05  M=128 N=64 K=64   (A K-major, B MN-major: the same W2 bytes)
//       dW2|db2[64x72] += dZ2' [H1 | 1]            tcgen05  M=64  N=72 K=128  (both MN-major: the same H1 / dZ2 bytes)
//   P4  dz1 = dh1 (1 - h1^2)                       in place over H1
//   M4  dW1|db1[64x8] += dZ1' [X | 1]              tcgen05  M=64  N=8  K=128
// and one pass of atomics per CTA at the end.  Biases ride along as a constant-one input feature, so no reduction over the
// samples is done on the CUDA cores at all (the FFMA version spent a third of its time there).
//
// Precision: fp32 values are split into bf16 hi + bf16 lo (16 mantissa bits together) and every product is three MMAs
// (hi hi + lo hi + hi lo) accumulated in fp32 in TMEM: relative error ~2^-16 per product, which keeps the gradient within the
// 2e-4 (of its max norm) bound of tests/test_gpu_ppo.py against autograd; a single bf16 or tf32 pass does not.  bench.py states
// "bf16x3" for this kernel.  The tensor pipe has room for the three passes: the kernel is bound by the CUDA-core phases.
//
// Shared-memory operand layout: no-swizzle UMMA "core matrices" of 8 rows x 16 bytes.  A [sample][feature] array is stored as
// core matrices [sample block][feature block], rows = samples.  The SAME bytes serve as a K-major operand (samples = M,
// features = K) and as an MN-major operand (features = M or N, samples = K): only the descriptor's two strides swap.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brb.h"
#include "brb_policy_layout.h"

namespace {

constexpr int A_FB = 9;                                  // feature blocks of buffer A: 8 x 8 hidden units + the constant-one block
constexpr int A_BYTES = 16 * A_FB * 128, B_BYTES = 16 * 8 * 128, W_BYTES = 8 * 8 * 128, V_BYTES = 16 * 128;
constexpr int OFF_AH = 0, OFF_AL = OFF_AH + A_BYTES, OFF_BH = OFF_AL + A_BYTES, OFF_BL = OFF_BH + B_BYTES;
constexpr int OFF_WH = OFF_BL + B_BYTES, OFF_WL = OFF_WH + W_BYTES, OFF_XH = OFF_WL + W_BYTES, OFF_XL = OFF_XH + V_BYTES;
constexpr int OFF_DH = OFF_XL + V_BYTES, OFF_DL = OFF_DH + V_BYTES, OFF_W1B = OFF_DL + V_BYTES;     // fp32 [64][8] = W1 row, b1, 0
constexpr int OFF_B2 = OFF_W1B + 64 * 8 * 4, OFF_W3 = OFF_B2 + 64 * 4, OFF_BAR = OFF_W3 + 2 * 64 * 4, OFF_TMEM = OFF_BAR + 8;
constexpr int TC_SMEM = OFF_TMEM + 8;
constexpr uint32_t C_Z2 = 0, C_DH1 = 64, C_DW2 = 128, C_DW3 = 200, C_DW1 = 208, TMEM_COLS = 256;
static_assert(TC_SMEM <= 110 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// UMMA shared-memory descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), version 1 (Blackwell)
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, majors, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t mk_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// bounded wait (a broken pipeline must not hang the GPU): returns false after ~1 s
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 24); spin++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// 16 consecutive accumulator columns of this thread's TMEM lane -> v[0..15]; asynchronous: tmem_ld_wait() before v is used
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float *v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                 "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  tmem_ld16_async(taddr, v);
  tmem_ld_wait();
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = __uint_as_float(r[k]);
}
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}
// (a, b) -> packed bf16 hi pair and bf16 lo pair (value = hi + lo to 16 mantissa bits); a in the low half (lower address)
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
// 8 consecutive features of sample s -> one 16-byte row of the core matrix (sample block s >> 3, feature block fb)
__device__ __forceinline__ void store8(uint8_t *hi_buf, uint8_t *lo_buf, int nfb, int s, int fb, const float *v) {
  uint4 h, l;
  split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y); split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
  const int off = ((s >> 3) * nfb + fb) * 128 + (s & 7) * 16;
  *reinterpret_cast<uint4 *>(hi_buf + off) = h;
  *reinterpret_cast<uint4 *>(lo_buf + off) = l;
}
__device__ __forceinline__ void load8(const uint8_t *hi_buf, const uint8_t *lo_buf, int nfb, int s, int fb, float *v) {
  const int off = ((s >> 3) * nfb + fb) * 128 + (s & 7) * 16;
  const uint4 h = *reinterpret_cast<const uint4 *>(hi_buf + off), l = *reinterpret_cast<const uint4 *>(lo_buf + off);
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    v[2 * k] = __uint_as_float(hh[k] << 16) + __uint_as_float(ll[k] << 16);
    v[2 * k + 1] = __uint_as_float(hh[k] & 0xFFFF0000u) + __uint_as_float(ll[k] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
// all threads: generic-proxy shared-memory writes and tcgen05.ld reads are ordered before the MMAs the elected thread issues next
__device__ __forceinline__ void publish_to_mma() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// the three bf16 passes of one fp32-grade product: hi hi, lo hi, hi lo.  `first` = overwrite the accumulator with the first pass.
template <int KSTEPS>
__device__ __forceinline__ void umma3(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_hi, uint32_t b_lo,
                                      uint32_t b_step, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, bool first) {
#pragma unroll
  for (int pass = 0; pass < 3; pass++) {
    const uint32_t a = pass == 1 ? a_lo : a_hi, b = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int k = 0; k < KSTEPS; k++)
      umma(d, mk_desc(a + k * a_step, a_lbo, a_sbo), mk_desc(b + k * b_step, b_lbo, b_sbo), idesc, (first && pass == 0 && k == 0) ? 0u : 1u);
  }
}

template <int ACTOR>
__global__ void __launch_bounds__(128, 2) brb_ppo_grad_tc_kernel(const float *__restrict__ params, const float *__restrict__ obs,
                                                                   const float *__restrict__ act, const float *__restrict__ oldlogp,
                                                                   const float *__restrict__ adv, const float *__restrict__ ret,
                                                                   const long long *__restrict__ idx, long long mb,
                                                                   const float *__restrict__ adv_stats, float clip, float vf_coef, float ent_coef,
                                                                   float *__restrict__ grad, float *__restrict__ stats, int *__restrict__ fault) {
  constexpr int NOUT = ACTOR ? 2 : 1;
  constexpr int OFF = ACTOR ? OFF_PI : OFF_VF, OFFW3 = ACTOR ? OFF_AW : OFF_VW, OFFB3 = ACTOR ? OFF_AB : OFF_VB;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t *AH = sm + OFF_AH, *AL = sm + OFF_AL, *BH = sm + OFF_BH, *BL = sm + OFF_BL, *WH = sm + OFF_WH, *WL = sm + OFF_WL;
  uint8_t *XH = sm + OFF_XH, *XL = sm + OFF_XL, *DH = sm + OFF_DH, *DL = sm + OFF_DL;
  float *W1B = reinterpret_cast<float *>(sm + OFF_W1B), *B2 = reinterpret_cast<float *>(sm + OFF_B2), *W3 = reinterpret_cast<float *>(sm + OFF_W3);
  uint64_t *bar_p = reinterpret_cast<uint64_t *>(sm + OFF_BAR);
  uint32_t *tmem_p = reinterpret_cast<uint32_t *>(sm + OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- one-time setup: weights into shared memory (W2 as bf16 hi/lo core matrices [out block][in block], rows = outputs)
  for (int e = tid; e < PH * PH; e += 128) {
    const int o = e >> 6, i = e & 63;
    const float w = params[OFF + PH * PIN + PH + e];
    const __nv_bfloat16 h = __float2bfloat16_rn(w), l = __float2bfloat16_rn(w - __bfloat162float(h));
    const int off = ((o >> 3) * 8 + (i >> 3)) * 128 + (o & 7) * 16 + (i & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(WH + off) = h;
    *reinterpret_cast<__nv_bfloat16 *>(WL + off) = l;
  }
  for (int e = tid; e < PH * 8; e += 128) {
    const int k = e >> 3, i = e & 7;
    W1B[e] = i < PIN ? params[OFF + k * PIN + i] : (i == PIN ? params[OFF + PH * PIN + k] : 0.f);
  }
  if (tid < PH) B2[tid] = params[OFF + PH * PIN + PH + PH * PH + tid];
  W3[tid] = tid < NOUT * PH ? params[OFFW3 + tid] : 0.f;
  {   // constant-one feature block of buffer A (feature 64 = 1, 65..71 = 0): its column of dZ2' [H1 | 1] is db2
    const int off = ((tid >> 3) * A_FB + 8) * 128 + (tid & 7) * 16;
    *reinterpret_cast<uint4 *>(AH + off) = make_uint4(0x00003F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(AL + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_p)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_p)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  publish_to_mma();
  const uint32_t tmem = *tmem_p, bar = smem_u32(bar_p);
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);          // this warp's 32 TMEM lanes
  const uint32_t aAH = smem_u32(AH), aAL = smem_u32(AL), aBH = smem_u32(BH), aBL = smem_u32(BL), aWH = smem_u32(WH), aWL = smem_u32(WL);
  const uint32_t aXH = smem_u32(XH), aXL = smem_u32(XL), aDH = smem_u32(DH), aDL = smem_u32(DL);

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
  float gb3[2] = {0.f, 0.f}, gls[2] = {0.f, 0.f}, st0 = 0.f, st1 = 0.f, st2 = 0.f;
  uint32_t parity = 0;
  bool first_tile = true, pending = false, ok = true;

  const long long ntiles = (mb + 127) / 128;
  // sample data is gathered through idx (random rows of the rollout buffer): every gather is issued one phase ahead of its use, so
  // that its latency hides behind an MMA wait (with two warps per scheduler nothing else would cover it: ncu long_scoreboard 2.1)
  long long i_next = 0;
  float x_next[PIN];
  {
    const long long s0 = (long long)blockIdx.x * 128 + tid;
    i_next = idx[s0 < mb ? s0 : 0];
#pragma unroll
    for (int k = 0; k < PIN; k++) x_next[k] = obs[i_next * PIN + k];
  }
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s = tile * 128 + tid;
    const bool valid = s < mb;
    const long long i = i_next;
    // ---- P1: first layer on the CUDA cores, H1 -> buffer A, [x | 1 | 0] -> X
    float x[8];
#pragma unroll
    for (int k = 0; k < PIN; k++) x[k] = x_next[k];
    x[6] = 1.f; x[7] = 0.f;
    {   // next tile's row index (its observation is fetched during P3)
      const long long sn = (tile + gridDim.x) * 128 + tid;
      i_next = idx[sn < mb ? sn : 0];
    }
    // this tile's loss inputs, used in P2
    float a_act[2] = {0.f, 0.f}, a_oldlp = 0.f, a_adv = 0.f, a_ret = 0.f;
    if (ACTOR) { a_act[0] = act[i * 2]; a_act[1] = act[i * 2 + 1]; a_oldlp = oldlogp[i]; a_adv = adv[i]; }
    else a_ret = ret[i];
    if (pending) { ok &= mbar_wait(bar, parity); parity ^= 1u; asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); pending = false; }   // M4 of the previous tile read A and X
#pragma unroll
    for (int fb = 0; fb < 8; fb++) {
      float h[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float4 w0 = *reinterpret_cast<const float4 *>(W1B + (fb * 8 + q) * 8), w1 = *reinterpret_cast<const float4 *>(W1B + (fb * 8 + q) * 8 + 4);
        float a = w1.z;
        a = fmaf(w0.x, x[0], a); a = fmaf(w0.y, x[1], a); a = fmaf(w0.z, x[2], a); a = fmaf(w0.w, x[3], a); a = fmaf(w1.x, x[4], a); a = fmaf(w1.y, x[5], a);
        h[q] = tanh_fast(a);
      }
      store8(AH, AL, A_FB, tid, fb, h);
    }
    store8(XH, XL, 1, tid, 0, x);
    publish_to_mma();
    if (tid == 0)   // M1: Z2 = H1 W2'   A = buffer A K-major (next 8 features +128, next 8 samples +1152), B = W2 K-major (next 8 ins +128, next 8 outs +1024)
    {
      umma3<4>(tmem + C_Z2, aAH, aAL, 256, 128, A_FB * 128, aWH, aWL, 256, 128, 1024, mk_idesc(128, 64, 0, 0), true);
      umma_commit(bar);
    }
    // ---- P2: second layer activation, output layer, loss
    ok &= mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float h2[PH], out[NOUT], dout[2] = {0.f, 0.f};
#pragma unroll
    for (int o = 0; o < NOUT; o++) out[o] = b3[o];
#pragma unroll
    for (int c = 0; c < PH; c += 16) tmem_ld16_async(tlane + C_Z2 + c, h2 + c);      // the whole accumulator row in flight, one wait
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < PH; j++) {
      const float h = tanh_fast(h2[j] + B2[j]);
      h2[j] = h;
#pragma unroll
      for (int o = 0; o < NOUT; o++) out[o] = fmaf(W3[o * PH + j], h, out[o]);
    }
    {
      const float m = valid ? inv_mb : 0.f;
      if (ACTOR) {
        float d[2], lp = 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          d[k] = a_act[k] - out[k];
          lp += -(d[k] * d[k]) * (0.5f * ivar[k]) - ls[k] - 0.91893853320467274f;
        }
        const float lr = lp - a_oldlp, ratio = expf(lr);
        const float an = (a_adv - amean) * arstd;
        const float s1 = an * ratio, s2 = an * fminf(1.f + clip, fmaxf(1.f - clip, ratio));
        const bool inside = ratio >= 1.f - clip && ratio <= 1.f + clip;
        const float g = (inside || s1 < s2) ? -an * ratio * m : 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          dout[k] = g * d[k] * ivar[k];
          gls[k] += g * (d[k] * d[k] * ivar[k] - 1.f);
          gb3[k] += dout[k];
        }
        st0 += -fminf(s1, s2) * m;
        st1 += ((ratio - 1.f) - lr) * m;
        st2 += (fabsf(ratio - 1.f) > clip ? 1.f : 0.f) * m;
      } else {
        const float diff = out[0] - a_ret;
        dout[0] = vf_coef * 2.f * diff * m;
        gb3[0] += dout[0];
        st0 += diff * diff * m;
      }
    }
#pragma unroll
    for (int fb = 0; fb < 8; fb++) store8(BH, BL, 8, tid, fb, h2 + 8 * fb);
    {
      const float dv[8] = {dout[0], dout[1], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store8(DH, DL, 1, tid, 0, dv);
    }
    publish_to_mma();
    if (tid == 0)   // M2: dW3 += H2' dOut   A = buffer B MN-major (next 8 units +128, next 8 samples +1024), B = dOut MN-major (next 8 samples +128)
    {
      umma3<8>(tmem + C_DW3, aBH, aBL, 2048, 1024, 128, aDH, aDL, 256, 128, 128, mk_idesc(64, 8, 1, 1), first_tile);
      umma_commit(bar);
    }
    // ---- P3: dz2 = (W3' dout)(1 - h2^2), over H2 once M2 has read it
#pragma unroll
    for (int j = 0; j < PH; j++) h2[j] = fmaf(W3[PH + j], dout[1], W3[j] * dout[0]) * (1.f - h2[j] * h2[j]);
#pragma unroll
    for (int k = 0; k < PIN; k++) x_next[k] = obs[i_next * PIN + k];      // next tile's observation: in flight across the M2 / M3 waits
    ok &= mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int fb = 0; fb < 8; fb++) store8(BH, BL, 8, tid, fb, h2 + 8 * fb);
    publish_to_mma();
    if (tid == 0) {
      // M3a: dH1 = dZ2 W2   A = buffer B K-major, B = W2 MN-major (N = inputs: next 8 +128; K = outputs: next 8 +1024)
      umma3<4>(tmem + C_DH1, aBH, aBL, 256, 128, 1024, aWH, aWL, 2048, 1024, 128, mk_idesc(128, 64, 0, 1), true);
      // M3b: dW2 | db2 += dZ2' [H1 | 1]   A = buffer B MN-major, B = buffer A MN-major (next 8 features +128, next 8 samples +1152)
      umma3<8>(tmem + C_DW2, aBH, aBL, 2048, 1024, 128, aAH, aAL, 2 * A_FB * 128, A_FB * 128, 128, mk_idesc(64, 72, 1, 1), first_tile);
      umma_commit(bar);
    }
    // ---- P4: dz1 = dh1 (1 - h1^2) in place over H1
    ok &= mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float *dh = h2;                                                               // same registers: dz2 is in shared memory now
#pragma unroll
      for (int c = 0; c < PH; c += 16) tmem_ld16_async(tlane + C_DH1 + c, dh + c);
      tmem_ld_wait();
#pragma unroll
      for (int fb = 0; fb < 8; fb++) {
        float h1[8];
        load8(AH, AL, A_FB, tid, fb, h1);
#pragma unroll
        for (int q = 0; q < 8; q++) h1[q] = dh[8 * fb + q] * (1.f - h1[q] * h1[q]);
        store8(AH, AL, A_FB, tid, fb, h1);
      }
    }
    publish_to_mma();
    if (tid == 0)   // M4: dW1 | db1 += dZ1' [X | 1]   A = buffer A MN-major, B = X MN-major
    {
      umma3<8>(tmem + C_DW1, aAH, aAL, 2 * A_FB * 128, A_FB * 128, 128, aXH, aXL, 256, 128, 128, mk_idesc(64, 8, 1, 1), first_tile);
      umma_commit(bar);
    }
    pending = true;
    first_tile = false;
  }
  if (pending) { ok &= mbar_wait(bar, parity); parity ^= 1u; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && fault) atomicExch(fault, 1);

  // ---- epilogue: M = 64 accumulators live in lanes (row % 16) + 32 (row / 16): the first 16 lanes of every warp
  if (!first_tile && ok) {
    const int j = 16 * warp + (tid & 15);
    const bool own = (tid & 31) < 16;
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      float v[16];
      tmem_ld16(tlane + C_DW2 + c, v);
      if (own) {
#pragma unroll
        for (int q = 0; q < 16; q++) atomicAdd(grad + OFF + PH * PIN + PH + j * PH + c + q, v[q]);
      }
    }
    float v8[8];
    tmem_ld8(tlane + C_DW2 + 64, v8);
    if (own) atomicAdd(grad + OFF + PH * PIN + PH + PH * PH + j, v8[0]);
    tmem_ld8(tlane + C_DW3, v8);
    if (own) {
#pragma unroll
      for (int o = 0; o < NOUT; o++) atomicAdd(grad + OFFW3 + o * PH + j, v8[o]);
    }
    tmem_ld8(tlane + C_DW1, v8);
    if (own) {
#pragma unroll
      for (int q = 0; q < PIN; q++) atomicAdd(grad + OFF + j * PIN + q, v8[q]);
      atomicAdd(grad + OFF + PH * PIN + j, v8[PIN]);
    }
  }
  const float q0 = warp_sum_f(st0), q1 = warp_sum_f(st1), q2 = warp_sum_f(st2), l0 = warp_sum_f(gls[0]), l1 = warp_sum_f(gls[1]);
  const float g0 = warp_sum_f(gb3[0]), g1 = warp_sum_f(gb3[1]);
  if ((tid & 31) == 0) {
    atomicAdd(grad + OFFB3, g0);
    if (ACTOR) {
      atomicAdd(grad + OFFB3 + 1, g1);
      atomicAdd(stats + 0, q0); atomicAdd(stats + 2, q1); atomicAdd(stats + 3, q2);
      atomicAdd(grad + OFF_LS, l0); atomicAdd(grad + OFF_LS + 1, l1);
    } else {
      atomicAdd(stats + 1, q0);
    }
  }
  if (ACTOR && blockIdx.x == 0 && tid < 2 && ent_coef != 0.f) atomicAdd(grad + OFF_LS + tid, -ent_coef);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace

// Launches both towers; *fault (device int, may be NULL) is set to 1 if a pipeline wait timed out (never observed; the waits are
// bounded so that a broken pipeline cannot hang the GPU).
extern "C" int brb_ppo_grad_tc_launch(const float *params, const float *obs, const float *actions, const float *old_logp, const float *adv,
                                      const float *returns, const int64_t *idx, int64_t mb, const float *adv_stats, float clip_range, float vf_coef,
                                      float ent_coef, float *grad, float *stats, int *fault, int sms, cudaStream_t s) {
  const long long tiles = (mb + 127) / 128, cap = 2LL * sms;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  if (cudaFuncSetAttribute(brb_ppo_grad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(brb_ppo_grad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return BRB_ECUDA;
  }
  brb_ppo_grad_tc_kernel<1><<<grid, 128, TC_SMEM, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats, clip_range,
                                                       vf_coef, ent_coef, grad, stats, fault);
  brb_ppo_grad_tc_kernel<0><<<grid, 128, TC_SMEM, s>>>(params, obs, actions, old_logp, adv, returns, (const long long *)idx, mb, adv_stats, clip_range,
                                                       vf_coef, ent_coef, grad, stats, fault);
  if (cudaGetLastError() != cudaSuccess) return BRB_ECUDA;
  return BRB_OK;
}
