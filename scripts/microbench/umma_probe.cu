// umma_probe.cu — de-risks the tcgen05 plumbing used by brb_ppo_grad_tc (csrc/brb_policy_tc.cu): TMEM alloc, no-swizzle
// shared-memory descriptors (K-major and MN-major over the SAME buffer), kind::f16 bf16 MMAs with M = 128 and M = 64,
// commit -> mbarrier, tcgen05.ld of both accumulator layouts.  Prints max errors against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor: start_address [0,14) (>>4), LBO [16,30) (>>4), SBO [32,46) (>>4), version [46,48) = 1, layout_type [61,64) = 0 (no swizzle)
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); spin++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;   // bounded: never hang the GPU
}
#define TMEM_LD16(taddr, r)                                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                          \
               : "r"(taddr))

// buffers: X [128 samples][72 features] and Y [128 samples][64 features] as 8x8 core matrices [sb][fb] (128 B each, sample-major rows
// of 16 B); W [64 out][64 in] as core matrices [ob][ib].
// GEMM1 (M=128,N=64,K=64): D1[s][o] = sum_i X[s][i] W[o][i]            A = X K-major, B = W K-major
// GEMM2 (M=128,N=64,K=64): D2[s][i] = sum_o Y[s][o] W[o][i]            A = Y K-major, B = W MN-major
// GEMM3 (M=64,N=72,K=128): D3[o][f] = sum_s Y[s][o] X[s][f]            A = Y MN-major, B = X MN-major
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *X, const __nv_bfloat16 *Y, const __nv_bfloat16 *W, float *D1, float *D2, float *D3, int *status) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __nv_bfloat16 *sX = (__nv_bfloat16 *)sm;                      // 16 x 9 x 128 B = 18432
  __nv_bfloat16 *sY = (__nv_bfloat16 *)(sm + 18432);            // 16 x 8 x 128 B = 16384
  __nv_bfloat16 *sW = (__nv_bfloat16 *)(sm + 18432 + 16384);    // 8 x 8 x 128 B = 8192
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 128 * 72; e += 128) { int s = e / 72, f = e % 72; sX[((s >> 3) * 9 + (f >> 3)) * 64 + (s & 7) * 8 + (f & 7)] = X[e]; }
  for (int e = tid; e < 128 * 64; e += 128) { int s = e / 64, f = e % 64; sY[((s >> 3) * 8 + (f >> 3)) * 64 + (s & 7) * 8 + (f & 7)] = Y[e]; }
  for (int e = tid; e < 64 * 64; e += 128) { int o = e / 64, i = e % 64; sW[((o >> 3) * 8 + (i >> 3)) * 64 + (o & 7) * 8 + (i & 7)] = W[e]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t aX = smem_u32(sX), aY = smem_u32(sY), aW = smem_u32(sW), b = smem_u32(&bar);
  if (tid == 0) {
    // GEMM1: A = X K-major (SBO = next 8 samples = 1152, LBO = next 8 features = 128), B = W K-major (SBO = next 8 outs = 1024, LBO = 128)
    for (int k = 0; k < 4; k++) umma(tmem + 0, make_desc(aX + k * 256, 128, 1152), make_desc(aW + k * 256, 128, 1024), make_idesc(128, 64, 0, 0), k > 0);
    // GEMM2: A = Y K-major (SBO 1024, LBO 128), B = W MN-major: N = in (SBO = next 8 ins = 128), K = out (LBO = next 8 outs = 1024)
    for (int k = 0; k < 4; k++) umma(tmem + 64, make_desc(aY + k * 256, 128, 1024), make_desc(aW + k * 2048, 1024, 128), make_idesc(128, 64, 0, 1), k > 0);
    // GEMM3: A = Y MN-major: M = feature (SBO 128), K = sample (LBO 1024); B = X MN-major: N = feature (SBO 128), K = sample (LBO 1152)
    for (int k = 0; k < 8; k++) umma(tmem + 128, make_desc(aY + k * 2048, 1024, 128), make_desc(aX + k * 2304, 1152, 128), make_idesc(64, 72, 1, 1), k > 0);
    umma_commit(b);
  }
  const bool ok = mbar_wait(b, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    uint32_t r[16];
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 16) {
      TMEM_LD16(lane_addr + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 16; q++) D1[tid * 64 + c + q] = __uint_as_float(r[q]);
      TMEM_LD16(lane_addr + 64 + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 16; q++) D2[tid * 64 + c + q] = __uint_as_float(r[q]);
    }
    // M = 64: row m lives in lane (m % 16) + 32 (m / 16); dump all 128 lanes x 80 columns so the host can check the layout
    for (int c = 0; c < 80; c += 16) {
      TMEM_LD16(lane_addr + 128 + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 16; q++) D3[tid * 80 + c + q] = __uint_as_float(r[q]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  const int NX = 128 * 72, NY = 128 * 64, NW = 64 * 64;
  __nv_bfloat16 *hX = new __nv_bfloat16[NX], *hY = new __nv_bfloat16[NY], *hW = new __nv_bfloat16[NW];
  float *fX = new float[NX], *fY = new float[NY], *fW = new float[NW];
  srand(1);
  auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.f; };
  for (int i = 0; i < NX; i++) { hX[i] = __float2bfloat16(rnd()); fX[i] = __bfloat162float(hX[i]); }
  for (int i = 0; i < NY; i++) { hY[i] = __float2bfloat16(rnd()); fY[i] = __bfloat162float(hY[i]); }
  for (int i = 0; i < NW; i++) { hW[i] = __float2bfloat16(rnd()); fW[i] = __bfloat162float(hW[i]); }
  __nv_bfloat16 *dX, *dY, *dW; float *d1, *d2, *d3; int *dst;
  cudaMalloc(&dX, NX * 2); cudaMalloc(&dY, NY * 2); cudaMalloc(&dW, NW * 2);
  cudaMalloc(&d1, 128 * 64 * 4); cudaMalloc(&d2, 128 * 64 * 4); cudaMalloc(&d3, 128 * 80 * 4); cudaMalloc(&dst, 4);
  cudaMemset(d3, 0, 128 * 80 * 4); cudaMemset(dst, 0, 4);
  cudaMemcpy(dX, hX, NX * 2, cudaMemcpyHostToDevice); cudaMemcpy(dY, hY, NY * 2, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, NW * 2, cudaMemcpyHostToDevice);
  const int smem = 18432 + 16384 + 8192;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dX, dY, dW, d1, d2, d3, dst);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  float *h1 = new float[128 * 64], *h2 = new float[128 * 64], *h3 = new float[128 * 80]; int st;
  cudaMemcpy(h1, d1, 128 * 64 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2, d2, 128 * 64 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3, d3, 128 * 80 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  printf("status (1 = mbarrier wait timed out): %d\n", st);
  double e1 = 0, e2 = 0, e3 = 0;
  for (int s = 0; s < 128; s++) for (int o = 0; o < 64; o++) {
    double a = 0, b2 = 0;
    for (int i = 0; i < 64; i++) { a += (double)fX[s * 72 + i] * fW[o * 64 + i]; b2 += (double)fY[s * 64 + i] * fW[i * 64 + o]; }
    e1 = fmax(e1, fabs(a - h1[s * 64 + o])); e2 = fmax(e2, fabs(b2 - h2[s * 64 + o]));
  }
  for (int o = 0; o < 64; o++) for (int f = 0; f < 72; f++) {
    double a = 0;
    for (int s = 0; s < 128; s++) a += (double)fY[s * 64 + o] * fX[s * 72 + f];
    const int lane = (o % 16) + 32 * (o / 16);
    e3 = fmax(e3, fabs(a - h3[lane * 80 + f]));
  }
  printf("max |err|: GEMM1 (K-major x K-major, M=128) %.3e   GEMM2 (K-major x MN-major, M=128) %.3e   GEMM3 (MN x MN, M=64, N=72) %.3e\n", e1, e2, e3);
  printf("%s\n", (e1 < 1e-3 && e2 < 1e-3 && e3 < 1e-3 && st == 0) ? "UMMA PROBE OK" : "UMMA PROBE FAILED");
  return 0;
}
