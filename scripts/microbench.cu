// fp64 / fp32 FMA issue-rate and latency microbenchmark for the second (ALU) ceiling in DESIGN.md
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, typename T>
__global__ void fma_kernel(T *out, T a, T b, int iters)
{
   T x[ILP];
   for (int i = 0; i < ILP; i++) x[i] = (T) (threadIdx.x + i);
   for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = x[i] * a + b;
   T s = 0;
   for (int i = 0; i < ILP; i++) s += x[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, typename T>
double run(int blocks, int threads, int iters)
{
   T *out;
   cudaMalloc(&out, sizeof(T) * blocks * threads);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   fma_kernel<ILP, T><<<blocks, threads>>>(out, (T) 1.0000001, (T) 1e-9, iters);
   cudaEventRecord(e0);
   fma_kernel<ILP, T><<<blocks, threads>>>(out, (T) 1.0000001, (T) 1e-9, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   cudaFree(out);
   return 2.0 * ILP * (double) iters * blocks * threads / (ms * 1e-3) / 1e12;
}
int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
   printf("fp64 TFLOP/s: 148x8 blocks x256 thr ILP8: %.2f\n", run<8, double>(148 * 8, 256, 20000));
   printf("fp64 TFLOP/s: 148x2 blocks x128 thr ILP1: %.2f\n", run<1, double>(148 * 2, 128, 100000));
   printf("fp64 TFLOP/s: 148x2 blocks x128 thr ILP2: %.2f\n", run<2, double>(148 * 2, 128, 100000));
   printf("fp64 TFLOP/s: 148x2 blocks x128 thr ILP4: %.2f\n", run<4, double>(148 * 2, 128, 50000));
   printf("fp64 TFLOP/s: 148x4 blocks x128 thr ILP4: %.2f\n", run<4, double>(148 * 4, 128, 50000));
   printf("fp64 TFLOP/s: 148x1 blocks x32 thr ILP1 (latency): %.4f\n", run<1, double>(148, 32, 200000));
   printf("fp32 TFLOP/s: 148x8 blocks x256 thr ILP8: %.2f\n", run<8, float>(148 * 8, 256, 40000));
   printf("fp32 TFLOP/s: 148x1 blocks x32 thr ILP1 (latency): %.4f\n", run<1, float>(148, 32, 400000));
   return 0;
}
