// tools/dmma_latency.cu -- dependent-issue latency and throughput of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) on this GPU:
// cycles per DMMA of ONE warp with 1..16 independent accumulator chains, and of W warps per SM with C chains each.
// Tells how many independent chains an SM needs in flight to keep the FP64 tensor pipe busy.  Not product code.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k(double* out, long long* cyc, int iters) {
  double c[ILP][2];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
  const int iters = 4000;
  k<ILP><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  k<ILP><<<148, warps * 32>>>(out, cyc, iters);
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double per_warp = (double)c / (iters * (double)ILP);
  printf("warps/SM %2d  chains/warp %2d  chains/SM %3d : %7.1f cycles per DMMA per warp, SM issues one DMMA every %6.2f cycles\n",
         warps, ILP, warps * ILP, per_warp, per_warp / warps);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc);
  }
  return 0;
}
