// Integer-issue peak of the box's GPU for the op classes the DP kernels are made of (SURVEY.md section 8d:
// "do not assume the lanes/clk figure -- measure it").  Dependent-free chains (8 per thread), full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu && ./int_peak > int_peak.json
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 16
#define ITERS 4096

template <int OP> __device__ __forceinline__ int step (int a, int b, int c) {
  if (OP == 0) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }	// IADD3 / VIADD
  if (OP == 1) { int r; asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }	// VIMNMX
  if (OP == 2) return __viaddmax_s32(a,b,c);					// VIADDMNMX: max(a+b,c)
  if (OP == 3) return (int) __vimax3_s16x2((unsigned) a,(unsigned) b,(unsigned) c);	// VIMNMX3.S16x2
  if (OP == 4) return (int) __viaddmax_s16x2((unsigned) a,(unsigned) b,(unsigned) c);	// VIADDMNMX.S16x2
  if (OP == 5) { int r; asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }	// LOP3
  if (OP == 6) return __shfl_up_sync(0xffffffffu,a,1);				// SHFL.UP
  if (OP == 7) return __vimax3_s32(a,b,c);					// VIMNMX3
  if (OP == 8) return (int) __vmaxs2((unsigned) a,(unsigned) b);			// VIMNMX.S16x2, two inputs
  if (OP == 9) {									// VIMNMX.S16x2 with both predicates + two predicated IMADs
    bool ph, pl; unsigned m = __vibmax_s16x2((unsigned) a,(unsigned) b,&ph,&pl);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, 16, %0;\n\t}" : "+r"(m) : "r"((unsigned) ph), "r"(c));
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, 1, %0;\n\t}" : "+r"(m) : "r"((unsigned) pl), "r"(c));
    return (int) m;
  }
  if (OP == 10) return (int) __vadd2((unsigned) a,(unsigned) b);			// VIADD.16x2
  if (OP == 11) { int r; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }	// PRMT
  if (OP == 12) { int r; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }	// IMAD
  if (OP == 13) {									// ISETP + predicated IMAD (the full fill's direction bits)
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %1, %2;\n\t@q mad.lo.u32 %0, %3, 16, %0;\n\t}" : "+r"(a) : "r"(b), "r"(a), "r"(c));
    return a;
  }
  if (OP == 14) { int r; asm volatile("shf.r.s32 %0, %1, 16, %1;" : "=r"(r) : "r"(a)); return r ^ b; }	// SHF + LOP
  return a;
}

template <int OP> __global__ void __launch_bounds__(256) k (int *out, int seed) {
  int v[CHAINS], b = seed + threadIdx.x, c = seed * 3 + blockIdx.x;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) v[i] = seed + i * 7 + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = step<OP>(v[i],b,c);
    if (OP != 0 && OP != 1 && OP != 5 && OP != 10) { b ^= it; c += it; }	// fresh operands every round: no collapsing of the chains
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s ^= v[i];
  if (s == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> double run (int *d, int blocks, const char *name, bool last) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; w++) k<OP><<<blocks,256>>>(d,w + 1);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int r = 0; r < reps; r++) k<OP><<<blocks,256>>>(d,r + 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double ops = (double) reps * blocks * 256.0 * ITERS * CHAINS;
  double tops = ops / (ms * 1e-3) / 1e12;
  printf("  \"%s\": %.3f%s\n",name,tops,last ? "" : ",");
  return tops;
}

int main () {
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0);
  int clk = 0; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
  int blocks = p.multiProcessorCount * 8;
  int *d; cudaMalloc(&d,(size_t) blocks * 256 * sizeof(int));
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d, \"unit\": \"Tera thread-ops/s (one op = one lane instruction; s16x2 ops carry two cells)\",\n",
	 p.name,p.multiProcessorCount,clk / 1000);
  run<0>(d,blocks,"iadd3",false);
  run<1>(d,blocks,"vimnmx_s32",false);
  run<7>(d,blocks,"vimnmx3_s32",false);
  run<2>(d,blocks,"viaddmnmx_s32",false);
  run<3>(d,blocks,"vimnmx3_s16x2",false);
  run<4>(d,blocks,"viaddmnmx_s16x2",false);
  run<5>(d,blocks,"lop3",false);
  run<8>(d,blocks,"vimnmx_s16x2",false);
  run<9>(d,blocks,"vimnmx_s16x2_2pred_plus_2_predicated_imad",false);
  run<10>(d,blocks,"viadd_16x2",false);
  run<11>(d,blocks,"prmt",false);
  run<12>(d,blocks,"imad",false);
  run<13>(d,blocks,"isetp_plus_predicated_imad",false);
  run<6>(d,blocks,"shfl_up",true);
  printf("}\n");
  return cudaGetLastError() != cudaSuccess;
}
