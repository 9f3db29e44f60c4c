// MUFU / conversion throughput microbenchmark (B200): warp-instructions per cycle per SM for a few XU-pipe ops.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = 0.001f * (threadIdx.x + j) + 0.5f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j])); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[j]));
      if (OP == 5) { unsigned h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h) : "f"(a[j])); a[j] += __uint_as_float(h); }
      if (OP == 6) { unsigned h; asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %1, %1; tanh.approx.f16x2 %0, t;}" : "=r"(h) : "f"(a[j])); a[j] += __uint_as_float(h); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int per_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP><<<148, 1024>>>(out, 16, cyc); cudaDeviceSynchronize();
  k<OP><<<148, 1024>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double winst = 32.0 * iters * 8 * per_iter;   // warp-instructions of the measured kind per SM (32 warps)
  printf("%-28s %8.3f warp-instr/cycle/SM  (%.1f lanes/clk/SM)\n", name, winst / c, 32.0 * winst / c);
}
int main() {
  run<0>("ex2.approx", 1); run<1>("rcp.approx", 1); run<2>("tanh.approx.f32", 1); run<3>("ex2+rcp (pairs)", 2);
  run<4>("fma (reference)", 1); run<5>("cvt.f16x2.f32 (+fadd)", 1); run<6>("cvt + tanh.f16x2 (+fadd)", 1);
  return 0;
}
