// Integer-pipe micro-probes for sm_100a: issue cost of the IMAD forms a Montgomery multiplier can be built
// from.  Per thread: two independent accumulator sets x[2][8] (as 4 lo/hi pairs each... see modes), with
// loop-variant multiplicands y[8] so nothing is hoisted.  Throughput is measured at 8 warps per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_imad tools/probe_imad.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* sink, uint32_t seed, int iters) {
  uint32_t x[2][8], y[8];
  uint32_t b = seed * 3u + blockIdx.x;
#pragma unroll
  for (int k = 0; k < 8; k++) { x[0][k] = k + seed + threadIdx.x; x[1][k] = k * 7 + b; y[k] = (k ^ seed) * 2654435761u + threadIdx.x; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      uint32_t* a = x[s];
      if (MODE == 0) {  // 4 independent wide MACs, no carry flags (IMAD.WIDE.U32)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint64_t acc = ((uint64_t)a[2 * k + 1] << 32) | a[2 * k];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(y[4 * s + k]), "r"(b));
          a[2 * k] = (uint32_t)acc; a[2 * k + 1] = (uint32_t)(acc >> 32);
        }
      } else if (MODE == 1) {  // one carry chain over four (lo,hi) pairs (IMAD.WIDE.U32.X expected)
        asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1;"
                     "madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                     "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;"
                     "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                     : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                     : "r"(y[4 * s]), "r"(y[4 * s + 1]), "r"(y[4 * s + 2]), "r"(y[4 * s + 3]), "r"(b));
      } else if (MODE == 2) {  // 4 high-half MACs (IMAD.HI.U32)
#pragma unroll
        for (int k = 0; k < 4; k++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(y[4 * s + k]), "r"(b));
      } else if (MODE == 3) {  // 4 low-half MACs (IMAD)
#pragma unroll
        for (int k = 0; k < 4; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(y[4 * s + k]), "r"(b));
      } else if (MODE == 4) {  // add-with-carry chain of 8 (IADD3.X)
        asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
                     "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %10; addc.u32 %7, %7, %11;"
                     : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                     : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
      } else if (MODE == 5) {  // wide MAC with carry-OUT only, carry consumed by an add (4 pairs)
#pragma unroll
        for (int k = 0; k < 4; k++)
          asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                       : "+r"(a[2 * k]), "+r"(a[2 * k + 1]), "+r"(a[(2 * k + 3) % 8]) : "r"(y[4 * s + k]), "r"(b));
      } else if (MODE == 7 || MODE == 8) {
        // FP64 pipe: 4 independent fma.rz.f64 (DFMA) per set on doubles kept in the integer registers' bit patterns
        // (the 52-bit-limb multiplier of Emmart et al. forms its products this way); MODE 8 adds 2 wide integer
        // MACs per set to see whether the two pipes issue concurrently.
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double acc = __hiloint2double((int)(a[2 * k + 1] & 0x000fffffu) | 0x43300000, (int)a[2 * k]);
          double m1 = __hiloint2double(0x43300000 | (int)(y[4 * s + k] & 0xfffffu), (int)b);
          asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(m1), "d"(1.0000001));
          asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(m1), "d"(0.9999999));
          a[2 * k] = (uint32_t)__double2loint(acc); a[2 * k + 1] = (uint32_t)__double2hiint(acc);
        }
        if (MODE == 8) {
#pragma unroll
          for (int k = 0; k < 2; k++) {
            uint64_t acc = ((uint64_t)a[4 * k + 1] << 32) | a[4 * k];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(y[4 * s + k]), "r"(b));
            a[4 * k] = (uint32_t)acc; a[4 * k + 1] = (uint32_t)(acc >> 32);
          }
        }
      } else if (MODE == 6) {  // separate lo and hi MACs with carry chains (the CIOS-row formulation), 4 products
        asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.lo.cc.u32 %1, %9, %12, %1; madc.lo.cc.u32 %2, %10, %12, %2; madc.lo.cc.u32 %3, %11, %12, %3; addc.u32 %4, %4, 0;"
                     "mad.hi.cc.u32 %1, %8, %12, %1; madc.hi.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %10, %12, %3; madc.hi.cc.u32 %4, %11, %12, %4; addc.u32 %5, %5, 0;"
                     : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                     : "r"(y[4 * s]), "r"(y[4 * s + 1]), "r"(y[4 * s + 2]), "r"(y[4 * s + 3]), "r"(b));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) y[k] ^= x[k & 1][k];
    b = b * 3u + 1u;
  }
  uint32_t s = b;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= x[0][k] ^ x[1][k] ^ y[k];
  if (s == 0x12345u) sink[0] = s;
}

template <int MODE>
void run(const char* name, double mul32_per_iter) {
  uint32_t* sink; cudaMalloc(&sink, 16);
  const int blocks = 148 * 8, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<blocks, 256>>>(sink, 12345u, iters);
  probe<MODE><<<blocks, 256>>>(sink, 12345u, iters);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(sink, 12345u, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = (double)blocks * 8;
  double smsp_cycles = ms * 1e-3 * 1.965e9 * 148 * 4;
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"smsp_cycles_per_warp_iter\": %.2f, \"per_product_or_op\": %.2f}\n", name, ms,
         smsp_cycles / (warps * iters), smsp_cycles / (warps * iters) / mul32_per_iter);
  cudaFree(sink);
}

int main() {
  run<0>("mad.wide.u32 x8 (IMAD.WIDE.U32, no carry)", 8);
  run<1>("2 chains of 4 wide pairs with carry (IMAD.WIDE.U32.X)", 8);
  run<2>("mad.hi.u32 x8 (IMAD.HI.U32)", 8);
  run<3>("mad.lo.u32 x8 (IMAD)", 8);
  run<4>("2 addc chains of 8 (IADD3.X), per add", 16);
  run<5>("8 x (wide pair carry-out + addc)", 8);
  run<6>("2 x CIOS-row of 4 products (lo chain + hi chain)", 8);
  run<7>("fma.rz.f64 x16 (DFMA), per DFMA", 16);
  run<8>("fma.rz.f64 x16 + mad.wide.u32 x4 together, per DFMA", 16);
  return 0;
}
