// Per-SM throughput of the instructions the attention softmax is made of (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each test: 148 CTAs x 256 threads (2 warps per SMSP), 8 independent chains per thread, long
// unrolled loop; reports results per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITERS 2048
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; p[i] = (unsigned long long)__float_as_uint(a[i]) | ((unsigned long long)__float_as_uint(a[i] + 0.5f) << 32); }
  const unsigned long long c2 = (unsigned long long)__float_as_uint(0.999f) | ((unsigned long long)__float_as_uint(0.998f) << 32);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 4) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r & 0x3fffffffu); }
      if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
      if (OP == 6) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r & 0x3fffffffu); }
      if (OP == 7) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 10) { uint32_t u = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      if (OP == 11) { uint32_t u = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      if (OP == 12) { uint32_t u = __float_as_uint(a[i]); float f; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(u)); a[i] = f + 1.0f; }
      if (OP == 13) { uint32_t u = __float_as_uint(a[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      if (OP == 14) { uint32_t u = __float_as_uint(a[i]); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u) : "r"(0x3c003c00u)); a[i] = __uint_as_float(u); }
      if (OP == 15) {   // the softmax inner step in fp16: pack two fp32 exponents (F2FP), one MUFU for both
        uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r & 0x3fffffffu); }
      if (OP == 16) {   // f16x2 -> two fp32 (for the fp32 row sum)
        uint32_t u = __float_as_uint(a[i]); float f0, f1;
        asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(f0), "=f"(f1) : "r"(u));
        a[i] = f0 + f1; }
      if (OP == 9) { uint32_t u = __float_as_uint(a[i]); asm volatile("shl.b32 %0, %0, 3;" : "+r"(u)); asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(seed))); a[i] = __uint_as_float(u); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_inst) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 256>>>(out, cyc, 0.5f);
  k<OP><<<148, 256>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double results = double(ITERS) * 8 * 256 * per_inst;
  printf("%-28s %8.1f results/clk/SM  (%.2f cycles per warp-inst per SMSP)\n", name, results / c, c / (double(ITERS) * 8 * 2));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<7>("tanh.approx.f32", 1);
  run<1>("fma.rn.f32", 1);
  run<2>("fma.rn.f32x2", 2);
  run<3>("add.rn.f32x2", 2);
  run<4>("cvt.rn.f16x2.f32", 2);
  run<6>("cvt.rn.bf16x2.f32", 2);
  run<5>("max.f32 (3-input)", 1);
  run<8>("max.f32 (2-input)", 1);
  run<9>("shl+add (u32)", 1);
  run<10>("ex2.approx.ftz.f16x2", 2);
  run<11>("ex2.approx.ftz.bf16x2", 2);
  run<12>("cvt.f32.f16 + add", 1);
  run<13>("tanh.approx.f16x2", 2);
  run<14>("fma.rn.f16x2", 2);
  run<15>("cvt.f16x2.f32 + ex2.f16x2", 2);
  run<16>("2x cvt.f32.f16 + add", 2);
  return 0;
}
