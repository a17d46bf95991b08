// FP32 pipe micro-benchmarks for the roofline denominator (sm_100a):
//   A: FFMA with constant-bank multiplier/addend (what brb_fp32_peak_flops measures)
//   B: FFMA with three distinct REGISTER operands (what most of the env-step kernel issues)
//   C: packed fma.rn.f32x2 (FFMA2) with register operands
//   D: FADD / FMUL register forms
#include <cstdio>
#include <cuda_runtime.h>

__global__ void kA(float *out, int iters, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int k = 0; k < iters; k++)
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kB(float *out, int iters, const float *in) {
  float x[8], y[8], z[8];
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = in[i] ; z[i] = in[8 + i]; }
  for (int k = 0; k < iters; k++)
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], y[(i + j) & 7], z[(i + 3 * j) & 7]);
  float s = 0; for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kC(float *out, int iters, const float *in) {
  unsigned long long x[8], y[8], z[8];
  for (int i = 0; i < 8; i++) {
    float2 a = make_float2(threadIdx.x + i, i), b = make_float2(in[i], in[i + 1]), c = make_float2(in[8 + i], in[9 + i]);
    x[i] = *reinterpret_cast<unsigned long long *>(&a); y[i] = *reinterpret_cast<unsigned long long *>(&b); z[i] = *reinterpret_cast<unsigned long long *>(&c);
  }
  for (int k = 0; k < iters; k++)
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y[(i + j) & 7]), "l"(z[(i + 3 * j) & 7]));
  float s = 0; for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2 *>(&x[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kD(float *out, int iters, const float *in) {
  float x[8], y[8];
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = in[i]; }
  for (int k = 0; k < iters; k++)
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = (j & 1) ? x[i] + y[(i + j) & 7] : x[i] * y[(i + j) & 7];
  float s = 0; for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> double timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 1e30;
  for (int r = 0; r < 5; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  float *out, *in; cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&in, 64 * sizeof(float));
  float h[64]; for (int i = 0; i < 64; i++) h[i] = 0.999f + 1e-4f * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  double n = 8.0 * 16 * iters * (double)blocks * threads;
  double tA = timeit([&] { kA<<<blocks, threads>>>(out, iters, 0.999999f, 1e-7f); });
  double tB = timeit([&] { kB<<<blocks, threads>>>(out, iters, in); });
  double tC = timeit([&] { kC<<<blocks, threads>>>(out, iters, in); });
  double tD = timeit([&] { kD<<<blocks, threads>>>(out, iters, in); });
  printf("{\"sms\": %d, \"ffma_const_tflops\": %.2f, \"ffma_reg_tflops\": %.2f, \"ffma2_reg_tflops\": %.2f, \"fadd_fmul_reg_tops\": %.2f}\n",
         p.multiProcessorCount, 2 * n / tA / 1e9, 2 * n / tB / 1e9, 4 * n / tC / 1e9, n / tD / 1e9);
  return 0;
}
