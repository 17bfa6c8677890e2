// Micro-benchmark: do MUFU.EX2 and F2FP.F16.F32.PACK_AB share an issue pipe on sm_100a?  (nvcc -arch=sm_100a xu_pipe.cu)
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>   // 0: ex2 only, 1: pack only, 2: both, 3: ex2 + ffma, 4: ex2 + pack via integer ops
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned acc = 0;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float x = a[i], y = a[i + 1];
      if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
      }
      if (MODE == 1 || MODE == 2) {
        unsigned p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x), "f"(y));
        acc ^= p;
      }
      if (MODE == 3) { x = fmaf(x, 1.0001f, y); y = fmaf(y, 0.9999f, x); }
      if (MODE == 4) {
        unsigned ux = __float_as_uint(x) + 0x1000u, uy = __float_as_uint(y) + 0x1000u;
        acc ^= __byte_perm(ux >> 13, uy >> 13, 0x5410);
      }
      a[i] = x * 0.5f; a[i + 1] = y * 0.5f;
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}

template <int MODE>
void run(const char* name) {
  float* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(d, 16);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<148, 1024>>>(d, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double pairs = 148.0 * 1024 * iters * 4;   // (x,y) pairs processed
  printf("%-28s %8.3f ms  %7.2f pairs/clk/SM (at 1.9 GHz)\n", name, ms, pairs / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(d);
}

int main() {
  run<0>("ex2 x2 per pair");
  run<1>("f2fp pack per pair");
  run<2>("ex2 x2 + f2fp pack");
  run<3>("ex2 x2 + 2 ffma");
  run<4>("ex2 x2 + int pack");
  return 0;
}
