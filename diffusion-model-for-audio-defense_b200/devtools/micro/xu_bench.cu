// Development aid (not product): per-SM throughput of the transcendental (XU) pipe and of the conversions / packed fp32 math the
// k1_layer epilogue uses, at the epilogue's occupancy (8 warps per SM) and at 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_bench xu_bench.cu && ./xu_bench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP> __device__ __forceinline__ void body(float (&v)[8], unsigned (&u)[4]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[k]));
    if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
    if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
    if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[k]));
    if (OP == 4) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[k & 3]) : "f"(v[k]), "f"(v[(k + 1) & 7])); v[k] += __uint_as_float(u[k & 3]); }
    if (OP == 5) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[k & 3]) : "f"(v[k]), "f"(v[(k + 1) & 7])); v[k] += __uint_as_float(u[k & 3]); }
    if (OP == 6) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
    if (OP == 7) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
    if (OP == 9) asm volatile("add.rn.f32 %0, %0, %0;" : "+f"(v[k]));
    if (OP == 10) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[k & 3]));
    if (OP == 11) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[k & 3]));
    if (OP == 12) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(u[k & 3]));
    if (OP == 13) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[k & 3]));
    if (OP == 14) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(v[(k + 1) & 7]));
    if (OP == 15) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[k & 3]) : "f"(v[k]), "f"(v[(k + 1) & 7])); u[(k + 1) & 3] ^= u[k & 3]; }
    if (OP == 16) asm volatile("min.f32 %0, %0, %1;" : "+f"(v[k]) : "f"(v[(k + 1) & 7]));
  }
  if (OP == 8) {   // packed fp32x2 fma: 4 instructions = 8 elements
    unsigned long long* p = reinterpret_cast<unsigned long long*>(v);
#pragma unroll
    for (int k = 0; k < 4; ++k) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k]));
  }
}
template <int OP> __global__ void k(float* out, long long* clk) {
  float v[8];
  unsigned u[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i) + 0.5f;
  for (int i = 0; i < 4; ++i) u[i] = 0x38003400u + threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) body<OP>(v, u);
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + u[0] + u[1] + u[2] + u[3];
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int warps, int per_iter_instr, float* out, long long* clk) {
  k<OP><<<148, warps * 32>>>(out, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  const double winstr = double(ITERS) * per_iter_instr * warps;
  printf("%-22s warps/SM %2d: %8.0f clk, %6.3f clk per warp-instr per SM, %6.2f thread-results/clk/SM\n", name, warps, m, m / winstr,
         winstr * 32 * ((OP == 8 || (OP >= 10 && OP <= 13)) ? 2 : 1) / m);
}
int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    run<0>("tanh.approx", warps, 8, out, clk);
    run<1>("ex2.approx", warps, 8, out, clk);
    run<2>("rcp.approx", warps, 8, out, clk);
    run<6>("sqrt.approx", warps, 8, out, clk);
    run<7>("lg2.approx", warps, 8, out, clk);
    run<4>("cvt.bf16x2 (+fadd)", warps, 8, out, clk);
    run<5>("cvt.f16x2 (+fadd)", warps, 8, out, clk);
    run<3>("fma.f32", warps, 8, out, clk);
    run<9>("add.f32", warps, 8, out, clk);
    run<8>("fma.f32x2", warps, 4, out, clk);
    run<10>("tanh.approx.f16x2", warps, 8, out, clk);
    run<13>("tanh.approx.bf16x2", warps, 8, out, clk);
    run<11>("ex2.approx.f16x2", warps, 8, out, clk);
    run<12>("fma.f16x2", warps, 8, out, clk);
    run<14>("rcp.approx (indep)", warps, 8, out, clk);
    run<15>("cvt.bf16x2 (+xor)", warps, 8, out, clk);
    run<16>("min.f32", warps, 8, out, clk);
  }
  return 0;
}
