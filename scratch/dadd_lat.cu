// microbenchmark: dependent DADD chain latency, register-fed and shared-memory-fed
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_reg(double* out, long long* cyc, int n) {
  double s = out[0], t = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) s = __dadd_rn(s, t);
  long long t1 = clock64();
  out[2] = s; cyc[0] = t1 - t0;
}
template <int U>
__global__ void k_smem(double* out, long long* cyc, int n) {
  __shared__ double buf[30 * 193 + 192];
  for (int i = threadIdx.x; i < 30 * 193 + 192; i += blockDim.x) buf[i] = out[0] * i;
  __syncthreads();
  const double* tp = buf + (threadIdx.x % 31) * 193;
  double s = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < n; ++rep) {
#pragma unroll U
    for (int k = 0; k < 192; ++k) s = __dadd_rn(s, tp[k]);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[3 + threadIdx.x] = s;
}
// software-pipelined: prefetch next group of 8 into registers while adding the current one
__global__ void k_smem_pipe(double* out, long long* cyc, int n) {
  __shared__ double buf[30 * 193 + 192];
  for (int i = threadIdx.x; i < 30 * 193 + 192; i += blockDim.x) buf[i] = out[0] * i;
  __syncthreads();
  const double* tp = buf + (threadIdx.x % 31) * 193;
  double s = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < n; ++rep) {
    double a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = tp[j];
    for (int k = 0; k < 192; k += 16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = tp[k + 8 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) s = __dadd_rn(s, a[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = tp[(k + 16 + j) % 192];
#pragma unroll
      for (int j = 0; j < 8; ++j) s = __dadd_rn(s, b[j]);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[3 + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  double h[2] = {1.0, 1e-9}; cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
  long long c;
  for (int rep = 0; rep < 2; ++rep) {
    k_reg<<<1, 32>>>(out, cyc, 100000); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("reg chain: %.2f cyc/add\n", c / 100000.0);
    k_smem<8><<<1, 32>>>(out, cyc, 100); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem unroll 8: %.2f cyc/add\n", c / (100.0 * 192));
    k_smem<16><<<1, 32>>>(out, cyc, 100); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem unroll 16: %.2f cyc/add\n", c / (100.0 * 192));
    k_smem<32><<<1, 32>>>(out, cyc, 100); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem unroll 32: %.2f cyc/add\n", c / (100.0 * 192));
    k_smem_pipe<<<1, 32>>>(out, cyc, 100); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem sw-pipelined: %.2f cyc/add\n", c / (100.0 * 192));
    k_smem<16><<<1, 1024>>>(out, cyc, 100); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem unroll 16, 32 warps busy: %.2f cyc/add\n", c / (100.0 * 192));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
