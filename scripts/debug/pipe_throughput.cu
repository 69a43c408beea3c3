// Micro-benchmark: per-SM throughput of the instructions the prefill softmax is built from, on
// sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_throughput pipe_throughput.cu
// Prints elements per clock per SM for each op (one CTA of 512 threads per SM, 8 independent chains
// per thread, clock64 around the loop).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cycles, float seed) {
  float a[CHAINS];
  uint32_t h[CHAINS];
  unsigned long long d[CHAINS / 2];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    a[i] = seed * (threadIdx.x + i + 1) * 1e-3f - 1.f;
    h[i] = 0x3c003c00u + i + threadIdx.x;  // two fp16 ~1.0
  }
#pragma unroll
  for (int i = 0; i < CHAINS / 2; i++)
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  unsigned long long sc;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(0.999f), "f"(0.999f));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(0.5f));
      if (OP == 6) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) % CHAINS]));
      if (OP == 7) {  // the softmax mix: scale-sub (fma) + ex2 + row-sum add, scalar
        float x;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x) : "f"(a[i]), "f"(seed), "f"(-0.25f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
      }
    }
#pragma unroll
    for (int i = 0; i < CHAINS / 2; i++) {
      if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(sc));
      if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(sc));
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc += a[i] + __uint_as_float(h[i]);
#pragma unroll
  for (int i = 0; i < CHAINS / 2; i++) acc += __uint_as_float((uint32_t)d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double elems_per_instr, int instrs_per_iter) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMalloc(&cyc, sms * sizeof(long long));
  k<OP><<<sms, 512>>>(out, cyc, 0.75f);
  k<OP><<<sms, 512>>>(out, cyc, 0.75f);
  cudaDeviceSynchronize();
  long long h[512];
  cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; i++) mean += h[i];
  mean /= sms;
  const double instr = (double)ITERS * instrs_per_iter * 512;  // thread-instructions per SM
  printf("%-28s %8.1f thread-instr/clk/SM  %8.1f elements/clk/SM  (%.0f cycles)\n", name, instr / mean,
         instr * elems_per_instr / mean, mean);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1, CHAINS);
  run<1>("ex2.approx.f16x2", 2, CHAINS);
  run<2>("ex2.approx.ftz.bf16x2", 2, CHAINS);
  run<3>("fma.rn.f32", 1, CHAINS);
  run<4>("fma.rn.f32x2", 2, CHAINS / 2);
  run<5>("add.rn.f32x2", 2, CHAINS / 2);
  run<6>("cvt.rn.bf16x2.f32", 2, CHAINS);
  run<7>("fma+ex2+add per element", 1.0 / 3, 3 * CHAINS);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
