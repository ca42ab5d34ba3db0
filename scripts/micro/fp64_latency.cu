// Micro-benchmark: dependent-issue latency and throughput of DADD / DMUL chains on one SM sub-partition.
// usage: ./fp64_latency   (prints cycles per instruction for ILP = 1..8 and 1..8 warps per scheduler)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, double a, double b, int iters, long long* cycles) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = __dadd_rn(__dmul_rn(x[i], b), a);   // DMUL -> DADD, dependent
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
void run(int warps_per_sm) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<ILP><<<1, 32 * warps_per_sm>>>(out, 1.000001, 0.999999, iters, cyc);
  chain<ILP><<<1, 32 * warps_per_sm>>>(out, 1.000001, 0.999999, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr_per_warp = 2.0 * 8 * ILP * iters;
  const double warps_per_sched = warps_per_sm / 4.0;
  printf("ILP %d warps/SM %2d: %.2f cycles per fp64 instr per warp; per scheduler %.2f cycles per instr\n", ILP, warps_per_sm, h / instr_per_warp,
         h / (instr_per_warp * (warps_per_sched < 1 ? 1 : warps_per_sched)));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 24, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
