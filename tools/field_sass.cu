// One Montgomery multiplication and one dedicated squaring per thread, nothing else: the SASS of this file
// (profiles/r2_field_mul_sqr.sass) is what DESIGN.md section 2.1 counts (IMAD.WIDE.U32 / IMAD.HI.U32 / IMAD).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o /tmp/field.cubin tools/field_sass.cu
//   cuobjdump -sass /tmp/field.cubin
#include <cuda_runtime.h>
#include "../vdf_b200/csrc/field.cuh"
using namespace vdf;
extern "C" __global__ void fp_mul_once(const fe* a, const fe* b, fe* out) {
  size_t i = blockIdx.x * blockDim.x + threadIdx.x;
  fe_store(out + i, Fp::mul(fe_load(a + i), fe_load(b + i)));
}
extern "C" __global__ void fp_sqr_once(const fe* a, fe* out) {
  size_t i = blockIdx.x * blockDim.x + threadIdx.x;
  fe_store(out + i, Fp::sqr(fe_load(a + i)));
}
