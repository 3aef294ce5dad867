// tools/store_pattern.cu -- write-bandwidth experiment (not product code): which store pattern reaches the HBM write
// ceiling?  Mimics the evaluation kernel's output streams without the arithmetic.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern store_pattern.cu && ./store_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NB = 64, NQ = 64;

__global__ void fill_kernel(double2* p, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2(1.0, 2.0);
}

// mode 0: warp per element, per point one 512 B N row (STG.64 x2) and one 1536 B dN row (STG.128 x3); elements interleaved
// mode 1: same, but per element first all N rows, then all dN rows
// mode 2: block (4 warps) per element, warp w handles points 16w..16w+15
// mode 3: warp per element, STG.128 for N as well (2 points per instruction pair)
template <int MODE>
__global__ void __launch_bounds__(128) pattern_kernel(double* N, double* dN, long long nelem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (MODE == 2) {
    for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
      for (int pt = 16 * warp; pt < 16 * warp + 16; ++pt) {
        double* n = N + (el * NQ + pt) * NB;
        n[lane] = 1.0; n[lane + 32] = 2.0;
        double2* d = reinterpret_cast<double2*>(dN + (el * NQ + pt) * NB * 3);
        d[lane] = make_double2(1, 2); d[lane + 32] = make_double2(1, 2); d[lane + 64] = make_double2(1, 2);
      }
    }
    return;
  }
  const long long warps = (long long)gridDim.x * 4, wid = (long long)blockIdx.x * 4 + warp;
  for (long long el = wid; el < nelem; el += warps) {
    if (MODE == 0 || MODE == 3) {
      for (int pt = 0; pt < NQ; ++pt) {
        double* n = N + (el * NQ + pt) * NB;
        if (MODE == 0) { n[lane] = 1.0; n[lane + 32] = 2.0; }
        else { reinterpret_cast<double2*>(n)[lane] = make_double2(1, 2); }
        double2* d = reinterpret_cast<double2*>(dN + (el * NQ + pt) * NB * 3);
        d[lane] = make_double2(1, 2); d[lane + 32] = make_double2(1, 2); d[lane + 64] = make_double2(1, 2);
      }
    } else {
      double2* n = reinterpret_cast<double2*>(N + el * NQ * NB);
      for (int j = lane; j < NQ * NB / 2; j += 32) n[j] = make_double2(1, 2);
      double2* d = reinterpret_cast<double2*>(dN + el * NQ * NB * 3);
      for (int j = lane; j < NQ * NB * 3 / 2; j += 32) d[j] = make_double2(1, 2);
    }
  }
}

template <class F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) f();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const long long nelem = 131072;
  const size_t nN = (size_t)nelem * NQ * NB, nD = nN * 3;
  double *N, *dN;
  CK(cudaMalloc(&N, nN * 8)); CK(cudaMalloc(&dN, nD * 8));
  const double gb = (nN + nD) * 8 / 1e9;
  float t = timeit([&] { fill_kernel<<<148 * 8, 256>>>((double2*)dN, nD / 2); fill_kernel<<<148 * 8, 256>>>((double2*)N, nN / 2); });
  printf("fill                 %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
  for (int bps = 1; bps <= 4; ++bps) {
    t = timeit([&] { pattern_kernel<0><<<148 * bps, 128>>>(N, dN, nelem); });
    printf("mode0 warp/elem bps=%d %.3f ms  %.0f GB/s\n", bps, t, gb / t * 1e3);
  }
  t = timeit([&] { pattern_kernel<3><<<148 * 3, 128>>>(N, dN, nelem); });
  printf("mode3 N as v2        %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
  t = timeit([&] { pattern_kernel<1><<<148 * 3, 128>>>(N, dN, nelem); });
  printf("mode1 N then dN      %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
  for (int bps = 1; bps <= 4; bps += 1) {
    t = timeit([&] { pattern_kernel<2><<<148 * bps, 128>>>(N, dN, nelem); });
    printf("mode2 block/elem bps=%d %.3f ms  %.0f GB/s\n", bps, t, gb / t * 1e3);
  }
  t = timeit([&] { pattern_kernel<0><<<148 * 16, 128>>>(N, dN, nelem); });
  printf("mode0 bps=16         %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
  CK(cudaDeviceSynchronize());
  return 0;
}
