// ubench.cu -- development aid: issue-rate microbenchmarks for the integer pipes of sm_100a (B200).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_ab/ubench tools/ubench.cu && gpurun_ab/ubench
// Every kernel runs ONE block per SM with `warps` warps and reports warp-instructions per SMSP-cycle
// (from the SM clock read inside the kernel), so the numbers are independent of the clock frequency.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
typedef uint32_t u32;
typedef uint64_t u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// NW independent IMAD.WIDE.U32 (pairs), NA ALU instructions (LOP3 on independent registers), ND DFMA,
// NM "moves" that ptxas turns into IMAD.MOV (measured: do they cost multiply-pipe time?), NL LDS.128
template <int NW, int NA, int ND, int NM, int NL, bool CHAIN>
__global__ void __launch_bounds__(1024, 1) k_mix(u32* sink, long long* cycles, u32 seed, int trips) {
  __shared__ uint4 sm[1024];
  u32 lo[8], hi[8], x[8], y[8];
  double d[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { lo[c] = t + c; hi[c] = seed + c; x[c] = seed * (c + 3) + t; y[c] = t ^ (c * 77u); d[c] = 1.0 + 1e-9 * (t + c); }
  sm[threadIdx.x] = make_uint4(t, t + 1, t + 2, t + 3);
  __syncthreads();
  const u32 b = seed | 0x80000001u;
  const double db = 1.0 + 1e-12 * seed;
  u32 saddr = (u32)__cvta_generic_to_shared(&sm[threadIdx.x]);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      // fine-grained interleave in program order: after every IMAD.WIDE its share of the other instructions
#pragma unroll
      for (int c = 0; c < (NW > 0 ? NW : 1); c++) {
        if (NW > 0) {
          if (CHAIN) {  // one carry chain over the NW pairs (a row of a schoolbook product); the other instructions do not touch CC
            if (c == 0) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(x[0]), "r"(b));
            else asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c & 7]), "+r"(hi[c & 7]) : "r"(x[c & 7]), "r"(b));
          } else {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c & 7]), "+r"(hi[c & 7]) : "r"(hi[(c + 1) & 7]), "r"(b));
          }
        }
        constexpr int DIV = NW > 0 ? NW : 1;
#pragma unroll
        for (int q = 0; q < NA / DIV; q++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[(c * (NA / DIV) + q) & 7]) : "r"(y[(c * (NA / DIV) + q + 1) & 7]), "r"(b));
#pragma unroll
        for (int q = 0; q < ND / DIV; q++) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[(c * (ND / DIV) + q) & 7]) : "d"(db));
#pragma unroll
        for (int q = 0; q < NM / DIV; q++) asm volatile("mov.b32 %0, %1;" : "=r"(y[(c + 4) & 7]) : "r"(y[c & 3]));
        if (NL > 0 && (c % (DIV / (NL > 0 ? NL : 1) > 0 ? DIV / (NL > 0 ? NL : 1) : 1)) == 0) {
          u32 r0, r1, r2, r3;
          asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr) : "memory");
          x[c & 7] ^= r0 ^ r1 ^ r2 ^ r3;
        }
      }
    }
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c] ^ x[c] ^ y[c] ^ (u32)__double_as_longlong(d[c]);
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// dependent chains: LAT = 0: IMAD.WIDE.U32.X chain (carry-linked row, length 8, each row depends on the previous one through
// the accumulators); 1: IADD3.X chain of length 8
template <int KIND>
__global__ void __launch_bounds__(1024, 1) k_lat(u32* sink, long long* cycles, u32 seed, int trips) {
  u32 lo[8], hi[8], a[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { lo[c] = t + c; hi[c] = seed + c; a[c] = seed * (c + 3) + t; }
  const u32 b = seed | 0x80000001u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if (KIND == 0) {  // 4 pairs, one chain; the next row uses the same accumulators: dependent through data AND carry
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(b));
#pragma unroll
        for (int c = 1; c < 4; c++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(b));
      } else if (KIND == 1) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(a[0]));
#pragma unroll
        for (int c = 1; c < 8; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo[c]) : "r"(a[c]));
      } else if (KIND == 2) {  // fully dependent IMAD.WIDE: the product feeds the next multiplicand
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;\n\tmadc.hi.u32 %1, %1, %2, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(b));
      } else {  // fully dependent LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[0]) : "r"(a[1]), "r"(b));
      }
    }
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c];
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}


// The shape of a field multiplication in the ladder: a multiply phase (NROWS rows of 8 IMAD.WIDE in two carry chains over
// the same accumulators) followed by an ALU phase (NCH dependent carry chains of 8 additions).  MODE 0: phases one after
// the other (as the out-of-line routines are today); MODE 1: the ALU phase of the PREVIOUS operation's data interleaved
// chain by chain with the rows (software pipeline: one chain of additions after every row); MODE 2: as 0, but odd warps
// start half an operation later (does a stagger survive the arbiter?)
template <int MODE, int NROWS, int NCH>
__global__ void __launch_bounds__(1024, 1) k_phase(u32* sink, long long* cycles, u32 seed, int trips) {
  u32 lo[8], hi[8], a[8], x[8], y[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { lo[c] = t + c; hi[c] = seed + c; a[c] = seed * (c + 3) + t; x[c] = t * (c + 5); y[c] = seed ^ (c * 131u) ^ t; }
  auto row = [&](int r) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(x[r & 7]));
#pragma unroll
    for (int c = 1; c < 4; c++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(x[r & 7]));
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[4]), "+r"(hi[4]) : "r"(a[4]), "r"(x[r & 7]));
#pragma unroll
    for (int c = 5; c < 8; c++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(x[r & 7]));
  };
  auto chain = [&]() {
    // consumes the accumulator words of the rows issued so far (the next row overwrites them: ptxas keeps the order)
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(y[0]) : "r"(hi[0]));
#pragma unroll
    for (int c = 1; c < 8; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(hi[c]));
  };
  if (MODE == 2 && ((threadIdx.x >> 5) & 4)) {  // warps 4..7, 12..15, ...: every second warp of an SMSP
#pragma unroll
    for (int r = 0; r < NROWS / 2; r++) row(r);
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
    if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < NROWS; r++) {
        row(r);
        if (r < NCH) chain();
      }
#pragma unroll
      for (int r = NROWS; r < NCH; r++) chain();
    } else {
#pragma unroll
      for (int r = 0; r < NROWS; r++) row(r);
#pragma unroll
      for (int r = 0; r < NCH; r++) chain();
    }
    // the next operation's multiplier depends on this one's ALU phase and its multiplicand on the product (as in a chain of squarings)
    x[0] ^= y[3];
    a[0] ^= hi[7];
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c] ^ x[c] ^ y[c] ^ a[c];
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static u32* g_sink;
static long long* g_cyc;
static int g_sms;
template <class K> static double run(K kern, int warps, int trips, double ops_per_trip_per_warp);

// The issue-port model in its cleanest form: a loop body of W IMAD.WIDE.U32 (independent accumulators, loop-invariant
// multiplicands: no moves) and A LOP3 (independent registers), interleaved, NOTHING else but the loop branch; the table printed
// from it compares clk per trip with 4 W + A + 3 (the three loop-control instructions).
template <int W, int A>
__global__ void __launch_bounds__(1024, 1) k_model(u32* sink, long long* cycles, u32 seed, int trips) {
  u32 lo[8], hi[8], m[8], y[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { lo[c] = t + c; hi[c] = seed + c; m[c] = seed * (c + 3) + t; y[c] = t ^ (c * 77u); }
  const u32 b = seed | 0x80000001u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
    constexpr int N = W > A ? W : A;
#pragma unroll
    for (int c = 0; c < N; c++) {
      if (c < W) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c & 7]), "+r"(hi[c & 7]) : "r"(m[c & 7]), "r"(b));
      if (c < A) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c & 7]) : "r"(m[(c + 1) & 7]), "r"(b));
    }
  }
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c] ^ y[c];
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
#define MODEL(W, A)                                                                                              \
  do {                                                                                                           \
    const double per_clk = run(k_model<W, A>, 32, 4000, (double)((W) + (A)));                                    \
    const double clk = ((W) + (A)) / per_clk;                                                                    \
    printf("model W=%2d IMAD.WIDE + A=%2d LOP3 per trip: %7.1f clk per trip per sub-partition-warp slot; 4W + A + 3 = %3d; ratio %.3f\n", W, A, clk, 4 * (W) + (A) + 3, clk / (4 * (W) + (A) + 3)); \
  } while (0)


template <class K>
static double run(K kern, int warps, int trips, double ops_per_trip_per_warp) {
  kern<<<g_sms, warps * 32>>>(g_sink, g_cyc, 12345u, 10);
  CK(cudaDeviceSynchronize());
  kern<<<g_sms, warps * 32>>>(g_sink, g_cyc, 12345u, trips);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(g_sms);
  CK(cudaMemcpy(c.data(), g_cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < g_sms; i++) avg += (double)c[i];
  avg /= g_sms;
  // warp-instructions per SMSP-cycle: warps/4 per SMSP
  return ops_per_trip_per_warp * trips * (warps / 4.0) / avg;
}

#define MIX(NW, NA, ND, NM, NL, CH, W)                                                                                   \
  do {                                                                                                                   \
    const double total = 4.0 * ((NW) + (NA) + (ND) + (NM) + (NL));                                                      \
    const double r = run(k_mix<NW, NA, ND, NM, NL, CH>, W, 2000, total);                                                 \
    printf("mix wide=%d alu=%d dfma=%d mov=%d lds=%d chain=%d warps/SMSP=%d : %.3f instr/clk/SMSP  (wide %.3f, alu %.3f, dfma %.3f)\n", \
           NW, NA, ND, NM, NL, (int)CH, (W) / 4, r, r * (NW) / (total / 4), r * (NA) / (total / 4), r * (ND) / (total / 4));               \
  } while (0)


// Ping-pong: warps w and w+4 of a 256-thread block sit on the same SMSP; a pair passes a token through two named barriers
// so that one of them is in its multiply phase while the other is in its ALU phase.  PING 0: the same kernel without tokens.
template <int PING, int NROWS, int NCH>
__global__ void __launch_bounds__(256) k_pingpong(u32* sink, long long* cycles, u32 seed, int trips) {
  u32 lo[8], hi[8], a[8], x[8], y[8];
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { lo[c] = t + c; hi[c] = seed + c; a[c] = seed * (c + 3) + t; x[c] = t * (c + 5); y[c] = seed ^ (c * 131u) ^ t; }
  const u32 warp = threadIdx.x >> 5, q = warp & 3u, role = (warp >> 2) & 1u;
  const u32 id_mine = 1u + q * 2u + role, id_partner = 1u + q * 2u + (role ^ 1u);
  if (PING && role == 1u) asm volatile("bar.arrive %0, 64;" ::"r"(id_partner) : "memory");  // role 0 may start
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
    if (PING) asm volatile("bar.sync %0, 64;" ::"r"(id_mine) : "memory");
#pragma unroll
    for (int r = 0; r < NROWS; r++) {
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(x[r & 7]));
#pragma unroll
      for (int c = 1; c < 4; c++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(x[r & 7]));
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[4]), "+r"(hi[4]) : "r"(a[4]), "r"(x[r & 7]));
#pragma unroll
      for (int c = 5; c < 8; c++) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(x[r & 7]));
    }
    if (PING) asm volatile("bar.arrive %0, 64;" ::"r"(id_partner) : "memory");
#pragma unroll
    for (int r = 0; r < NCH; r++) {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(y[0]) : "r"(hi[0]));
#pragma unroll
      for (int c = 1; c < 8; c++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(hi[c]));
    }
    x[0] ^= y[3];
    a[0] ^= hi[7];
  }
  if (PING && role == 0u) asm volatile("bar.sync %0, 64;" ::"r"(id_mine) : "memory");  // consume the partner's last token
  const long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= lo[c] ^ hi[c] ^ x[c] ^ y[c] ^ a[c];
  sink[t] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <class K>
static double run_blocks(K kern, int blocks_per_sm, int trips, double ops_per_trip_per_warp) {
  const int nb = g_sms * blocks_per_sm;
  kern<<<nb, 256>>>(g_sink, g_cyc, 12345u, 10);
  CK(cudaDeviceSynchronize());
  kern<<<nb, 256>>>(g_sink, g_cyc, 12345u, trips);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(nb);
  CK(cudaMemcpy(c.data(), g_cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < nb; i++) avg += (double)c[i];
  avg /= nb;
  return ops_per_trip_per_warp * trips * (2.0 * blocks_per_sm) / avg;  // 8 warps per block = 2 per SMSP
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaMalloc(&g_sink, (size_t)g_sms * 8192 * 4));
  CK(cudaMalloc(&g_cyc, g_sms * 8 * sizeof(long long)));
  printf("SMs %d\n", g_sms);
  printf("== issue-port model, 8 warps per sub-partition (clk per trip = (W + A) / (instructions per clk per sub-partition) x 8 warps ... reported per warp slot)\n");
  MODEL(8, 0); MODEL(16, 0); MODEL(32, 0); MODEL(0, 32); MODEL(32, 8); MODEL(32, 16); MODEL(32, 32); MODEL(16, 32); MODEL(8, 32); MODEL(32, 64);
  // latencies (1 warp per SMSP)
  printf("lat IMAD.WIDE.X row of 4 (dependent rows): %.2f clk per instr\n", 1.0 / run(k_lat<0>, 4, 2000, 16.0));
  printf("lat IADD3.X chain of 8                   : %.2f clk per instr\n", 1.0 / run(k_lat<1>, 4, 2000, 32.0));
  printf("lat IMAD.WIDE fully dependent            : %.2f clk per instr\n", 1.0 / run(k_lat<2>, 4, 2000, 4.0));
  printf("lat LOP3 fully dependent                 : %.2f clk per instr\n", 1.0 / run(k_lat<3>, 4, 2000, 4.0));
  for (int w : {4, 8, 16, 32}) {
    printf("-- %d warps per SMSP\n", w / 4);
    if (w == 4) { MIX(8, 0, 0, 0, 0, false, 4); MIX(8, 0, 0, 0, 0, true, 4); MIX(0, 8, 0, 0, 0, false, 4); MIX(0, 0, 8, 0, 0, false, 4); MIX(8, 8, 0, 0, 0, false, 4); MIX(8, 16, 0, 0, 0, false, 4); MIX(8, 24, 0, 0, 0, false, 4);
                  MIX(8, 0, 8, 0, 0, false, 4); MIX(8, 0, 16, 0, 0, false, 4); MIX(8, 16, 8, 0, 0, false, 4); MIX(8, 0, 0, 8, 0, false, 4); MIX(8, 8, 0, 0, 4, false, 4); MIX(8, 16, 0, 0, 4, true, 4); }
    if (w == 8) { MIX(8, 0, 0, 0, 0, false, 8); MIX(8, 0, 0, 0, 0, true, 8); MIX(0, 8, 0, 0, 0, false, 8); MIX(0, 0, 8, 0, 0, false, 8); MIX(8, 8, 0, 0, 0, false, 8); MIX(8, 16, 0, 0, 0, false, 8); MIX(8, 24, 0, 0, 0, false, 8);
                  MIX(8, 0, 8, 0, 0, false, 8); MIX(8, 0, 16, 0, 0, false, 8); MIX(8, 16, 8, 0, 0, false, 8); MIX(8, 0, 0, 8, 0, false, 8); MIX(8, 8, 0, 0, 4, false, 8); MIX(8, 16, 0, 0, 4, true, 8); }
    if (w == 16) { MIX(8, 0, 0, 0, 0, false, 16); MIX(8, 0, 0, 0, 0, true, 16); MIX(0, 8, 0, 0, 0, false, 16); MIX(0, 0, 8, 0, 0, false, 16); MIX(8, 8, 0, 0, 0, false, 16); MIX(8, 16, 0, 0, 0, false, 16); MIX(8, 24, 0, 0, 0, false, 16);
                   MIX(8, 0, 8, 0, 0, false, 16); MIX(8, 0, 16, 0, 0, false, 16); MIX(8, 16, 8, 0, 0, false, 16); MIX(8, 0, 0, 8, 0, false, 16); MIX(8, 8, 0, 0, 4, false, 16); MIX(8, 16, 0, 0, 4, true, 16); }
    if (w == 32) { MIX(8, 0, 0, 0, 0, false, 32); MIX(8, 0, 0, 0, 0, true, 32); MIX(0, 8, 0, 0, 0, false, 32); MIX(0, 0, 8, 0, 0, false, 32); MIX(8, 16, 0, 0, 0, false, 32); MIX(8, 0, 16, 0, 0, false, 32); MIX(8, 16, 8, 0, 0, false, 32); MIX(8, 0, 0, 8, 0, false, 32); }
  }

  printf("== phase structure: 64 IMAD.WIDE then 56 dependent additions per operation; fraction of the IMAD.WIDE issue rate (1 per 4 clk per SMSP)\n");
  for (int w : {4, 8, 12, 16, 20, 24, 32}) {
    const double r0 = run(k_phase<0, 8, 7>, w, 2000, 64.0), r1 = run(k_phase<1, 8, 7>, w, 2000, 64.0), r2 = run(k_phase<2, 8, 7>, w, 2000, 64.0);
    printf("warps/SMSP=%d : phases in sequence %.3f | interleaved %.3f | sequence, staggered start %.3f   (x4 = share of the multiply pipe)\n", w / 4, r0 * 4, r1 * 4, r2 * 4);
  }

  printf("== ping-pong through named barriers (256-thread blocks, pairs w / w+4), share of the multiply pipe\n");
  for (int nb : {1, 2, 3, 4}) {
    const double r0 = run_blocks(k_pingpong<0, 8, 7>, nb, 2000, 64.0), r1 = run_blocks(k_pingpong<1, 8, 7>, nb, 2000, 64.0);
    const double r2 = run_blocks(k_pingpong<0, 8, 14>, nb, 2000, 64.0), r3 = run_blocks(k_pingpong<1, 8, 14>, nb, 2000, 64.0);
    printf("warps/SMSP=%d : 56 additions: free %.3f  ping-pong %.3f | 112 additions: free %.3f  ping-pong %.3f\n", 2 * nb, r0 * 4, r1 * 4, r2 * 4, r3 * 4);
  }
  return 0;
}
