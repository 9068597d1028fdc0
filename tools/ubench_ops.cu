// Development aid (r2): what one instruction of each kind costs a sub-partition's issue port next to IMAD.WIDE.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_ab/ubench_ops tools/ubench_ops.cu
// Every trip of the timed loop is W IMAD.WIDE (8 independent accumulators) interleaved with A instructions of kind OP (8
// independent registers) and nothing else but the loop branch; 8 warps per sub-partition (1024 threads, one block per SM).
// Printed: clk per trip per warp slot, and the cost per OP instruction = (clk(W, A) - clk(W, 0)) / A.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

enum { OP_LOP3, OP_IADD, OP_ADDC, OP_SHF, OP_SEL, OP_IMADLO, OP_IMADIADD, OP_IMADHI, OP_IMADSHL, OP_LDS, OP_PRMT, OP_IADD3, OP_ISETP, OP_COUNT };
static const char* kOpName[] = {"LOP3", "IADD (add.u32)", "IADD3.X chain", "SHF", "LOP3 (maj)", "IMAD (lo)", "add via mad x1", "IMAD.HI", "LEA (shl+add)", "LDS.128", "PRMT", "IADD3 (3 in)", "ISETP+SEL"};

template <int W, int A, int OP>
__global__ void __launch_bounds__(1024, 1) k_op(u32* sink, long long* cycles, u32 seed, int trips) {
  __shared__ uint4 sm[1024];
  unsigned long long acc64[8];
  u32 m[8], y[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  sm[threadIdx.x] = make_uint4(t, seed, t ^ seed, 7u);
#pragma unroll
  for (int c = 0; c < 8; c++) { acc64[c] = ((unsigned long long)(seed + c) << 32) | (t + c); m[c] = seed * (c + 3) + t; y[c] = t ^ (c * 77u); }
  const u32 b = seed | 0x80000001u;
  const u32 sa = (u32)__cvta_generic_to_shared(&sm[threadIdx.x]);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
    constexpr int N = W > A ? W : A;
#pragma unroll
    for (int c = 0; c < N; c++) {
      if (c < W) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc64[c & 7]) : "r"((u32)acc64[(c + 3) & 7]), "r"(b));  // a multiplicand that changes: nothing to hoist
      if (c < A) {
        u32& r = y[c & 7];
        const u32 o = m[(c + 1) & 7];
        if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(o), "r"(b));
        else if (OP == OP_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(o));
        else if (OP == OP_ADDC) {
          if ((c & 7) == 0) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r) : "r"(o));
          else asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r) : "r"(o));
        }
        else if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(r) : "r"(o));
        else if (OP == OP_SEL) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r) : "r"(o), "r"(b));
        else if (OP == OP_IMADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(o), "r"(b));
        else if (OP == OP_IMADIADD) asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(r) : "r"(o));
        else if (OP == OP_IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r) : "r"(o));
        else if (OP == OP_IMADSHL) asm volatile("{ .reg .u32 q; shl.b32 q, %0, 3; add.u32 %0, q, %1; }" : "+r"(r) : "r"(o));
        else if (OP == OP_LDS) asm volatile("{ .reg .u32 q0, q1, q2; ld.volatile.shared.v4.u32 {%0,q0,q1,q2}, [%1]; }" : "=r"(r) : "r"(sa));
        else if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x1230;" : "+r"(r) : "r"(o));
        else if (OP == OP_IADD3) asm volatile("{ .reg .u32 q; add.u32 q, %0, %1; add.u32 %0, q, %2; }" : "+r"(r) : "r"(o), "r"(b));
        else if (OP == OP_ISETP) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(r) : "r"(o), "r"(b));
      }
    }
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= (u32)acc64[c] ^ (u32)(acc64[c] >> 32) ^ y[c];
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_sms;
static u32* g_sink;
static long long* g_cyc;

template <class K>
static double clk_per_trip(K kern, int trips) {  // per warp slot: 8 warps share a sub-partition, so block cycles / trips / 8
  kern<<<g_sms, 1024>>>(g_sink, g_cyc, 12345u, 10);
  CK(cudaDeviceSynchronize());
  kern<<<g_sms, 1024>>>(g_sink, g_cyc, 12345u, trips);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(g_sms);
  CK(cudaMemcpy(c.data(), g_cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < g_sms; i++) avg += (double)c[i];
  return avg / g_sms / trips / 8.0;
}

template <int OP>
static void row() {
  const double w32 = clk_per_trip(k_op<32, 0, OP>, 3000);
  const double a32 = clk_per_trip(k_op<0, 32, OP>, 3000);
  const double w32a16 = clk_per_trip(k_op<32, 16, OP>, 3000);
  const double w32a32 = clk_per_trip(k_op<32, 32, OP>, 3000);
  const double w16a32 = clk_per_trip(k_op<16, 32, OP>, 3000);
  const double w16 = clk_per_trip(k_op<16, 0, OP>, 3000);
  printf("%-16s alone %5.2f clk/op | next to 32 IMAD.WIDE (%.1f clk): +16 ops %5.2f clk/op, +32 ops %5.2f clk/op | next to 16 IMAD.WIDE (%.1f clk): +32 ops %5.2f clk/op\n",
         kOpName[OP], a32 / 32.0, w32, (w32a16 - w32) / 16.0, (w32a32 - w32) / 32.0, w16, (w16a32 - w16) / 32.0);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  CK(cudaMalloc(&g_sink, (size_t)g_sms * 1024 * 4));
  CK(cudaMalloc(&g_cyc, g_sms * sizeof(long long)));
  printf("SMs %d; clk per instruction per sub-partition (issue cost), 8 warps per sub-partition\n", g_sms);
  row<OP_LOP3>(); row<OP_IADD>(); row<OP_IADD3>(); row<OP_ADDC>(); row<OP_SHF>(); row<OP_SEL>(); row<OP_ISETP>(); row<OP_PRMT>();
  row<OP_IMADLO>(); row<OP_IMADIADD>(); row<OP_IMADSHL>(); row<OP_IMADHI>(); row<OP_LDS>();
  return 0;
}
