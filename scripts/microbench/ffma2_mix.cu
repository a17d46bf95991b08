// Does packed fma.rn.f32x2 relieve the ISSUE slot in a mixed FP32 / integer instruction stream at low occupancy (3 warps
// per scheduler, as in brb_step_kernel)?  Per inner iteration: 16 FMAs on 16 independent chains + NALU integer ops.
//   scalar: 16 FFMA (register operands) + NALU LOP3/IADD3       packed: 8 FFMA2 + NALU LOP3/IADD3
#include <cstdio>
#include <cuda_runtime.h>
template <int NALU, bool PACKED>
__global__ void __launch_bounds__(128) mix(float *out, int iters, const float *in) {
  float y[4], z[4];
  for (int i = 0; i < 4; i++) { y[i] = in[i]; z[i] = in[4 + i]; }
  unsigned u[8];
  for (int i = 0; i < 8; i++) u[i] = threadIdx.x * 2654435761u + i;
  float acc = 0.f;
  if (PACKED) {
    unsigned long long x[8], yy[2], zz[2];
    for (int i = 0; i < 8; i++) { float2 a = make_float2(threadIdx.x + i, i); x[i] = *reinterpret_cast<unsigned long long *>(&a); }
    for (int i = 0; i < 2; i++) { float2 b = make_float2(y[2 * i], y[2 * i + 1]), c = make_float2(z[2 * i], z[2 * i + 1]);
      yy[i] = *reinterpret_cast<unsigned long long *>(&b); zz[i] = *reinterpret_cast<unsigned long long *>(&c); }
    for (int k = 0; k < iters; k++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(yy[(i + j) & 1]), "l"(zz[(i + j + 1) & 1]));
#pragma unroll
        for (int i = 0; i < NALU; i++) u[i & 7] = (u[i & 7] ^ (u[(i + 1) & 7] >> 3)) + 0x9E3779B9u;
      }
    }
    for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2 *>(&x[i]); acc += v.x + v.y; }
  } else {
    float x[16];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int k = 0; k < iters; k++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], y[(i + j) & 3], z[(i + j + 1) & 3]);
#pragma unroll
        for (int i = 0; i < NALU; i++) u[i & 7] = (u[i & 7] ^ (u[(i + 1) & 7] >> 3)) + 0x9E3779B9u;
      }
    }
    for (int i = 0; i < 16; i++) acc += x[i];
  }
  unsigned s = 0; for (int i = 0; i < 8; i++) s += u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)s;
}
template <class F> double timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 1e30;
  for (int r = 0; r < 5; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 3, threads = 128, iters = 20000;    // 12 warps per SM = 3 per scheduler
  float *out, *in; cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&in, 64 * sizeof(float));
  float h[64]; for (int i = 0; i < 64; i++) h[i] = 0.999f + 1e-4f * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  const double fma = 16.0 * 8 * iters * (double)blocks * threads;
#define RUN(NALU) { double ts = timeit([&] { mix<NALU, false><<<blocks, threads>>>(out, iters, in); }), tp = timeit([&] { mix<NALU, true><<<blocks, threads>>>(out, iters, in); }); \
    printf("alu ops per 16 FMA = %2d: scalar %.3f ms (%.1f TFLOP/s), packed %.3f ms (%.1f TFLOP/s), packed/scalar time %.3f\n", NALU, ts, 2 * fma / ts / 1e9, tp, 2 * fma / tp / 1e9, tp / ts); }
  RUN(0) RUN(4) RUN(8) RUN(12) RUN(16)
  return 0;
}
