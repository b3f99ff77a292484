// pipe_probe.cu — which issue pipe do PRMT, IDP.4A (dp4a), FFMA, FMNMX, IMAD and HADD2.F32 use on this GPU?
// Each kernel runs 8 independent dependency chains per thread of one or two instruction kinds; if A+B interleaved takes
// max(time A, time B) the two kinds issue on different pipes, if it takes the sum they share one.  (Decides how the
// node test of traverse.cuh converts its 48 quantised plane bytes: the test is ALU-pipe bound, ncu r02c.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

#define ITER 4096
template <int A, int B>
__global__ void k(unsigned* out, unsigned seed, float fs)
{
  unsigned x[8];
  float    f[8];
  for(int i = 0; i < 8; i++)
  {
    x[i] = seed + threadIdx.x * 8 + i;
    f[i] = fs + i;
  }
  for(int it = 0; it < ITER; it++)
  {
#pragma unroll
    for(int i = 0; i < 8; i++)
    {
      auto op = [&](int kind) {
        if(kind == 1)
          asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(x[i]) : "r"(seed));
        if(kind == 2)
          asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(0x40u), "r"(seed));
        if(kind == 3)
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fs), "f"(1.0f));
        if(kind == 4)
          asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
        if(kind == 5)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(3u));
        if(kind == 6)
        {
          unsigned short h = (unsigned short)x[i];
          float          r;
          asm volatile("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
          f[i] += r;
        }
        if(kind == 7)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(seed), "r"(0x55u));
      };
      op(A);
      op(B);
    }
  }
  unsigned acc = 0;
  for(int i = 0; i < 8; i++)
    acc += x[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int A, int B>
float run(unsigned* d)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<A, B><<<148 * 4, 256>>>(d, 12345u, 1.0001f);
  cudaEventRecord(e0);
  k<A, B><<<148 * 4, 256>>>(d, 12345u, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main()
{
  unsigned* d;
  cudaMalloc(&d, 148 * 4 * 256 * 4);
  const char* name[8] = {"-", "PRMT", "DP4A", "FFMA", "FMNMX", "IMAD", "CVT.F32.F16", "LOP3"};
  printf("single kinds (8 chains x %d iterations, 592 blocks x 256 threads):\n", ITER);
#define ONE(A) printf("  %-12s %.3f ms\n", name[A], run<A, 0>(d));
  ONE(1) ONE(2) ONE(3) ONE(4) ONE(5) ONE(6) ONE(7)
  printf("pairs (sum of singles => same pipe, max => different pipes):\n");
#define TWO(A, B) printf("  %-12s + %-12s %.3f ms\n", name[A], name[B], run<A, B>(d));
  TWO(1, 3) TWO(2, 3) TWO(1, 2) TWO(2, 5) TWO(1, 4) TWO(2, 4) TWO(6, 3) TWO(6, 1) TWO(7, 2) TWO(7, 3)
  return 0;
}
