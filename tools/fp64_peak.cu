// Microbenchmark: sustained FP64 FMA throughput of the device as a function of resident warps and ILP.
// Used to set the FP64 roofline for the sum-factorisation kernels (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma_kernel(double *out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < ILP; ++i)
          x[i] = fma(x[i], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    s += x[i];
  if (s == 123.456)
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int threads_per_block, int blocks_per_sm, int n_sm, double *d_out)
{
  const int   iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fma_kernel<ILP><<<n_sm * blocks_per_sm, threads_per_block>>>(d_out, 10, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  fma_kernel<ILP><<<n_sm * blocks_per_sm, threads_per_block>>>(d_out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)n_sm * blocks_per_sm * threads_per_block * iters * 8.0 * ILP;
  printf("threads/SM %5d ILP %2d : %8.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", threads_per_block * blocks_per_sm, ILP,
         2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / n_sm / 1.965e9);
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  double *d;
  cudaMalloc(&d, 1 << 24);
  const int n = p.multiProcessorCount;
  for (int tpb : {128, 320, 640, 1024})
    {
      run<1>(tpb, 1, n, d);
      run<2>(tpb, 1, n, d);
      run<4>(tpb, 1, n, d);
      run<8>(tpb, 1, n, d);
      run<16>(tpb, 1, n, d);
    }
  run<8>(1024, 2, n, d);
  return 0;
}
