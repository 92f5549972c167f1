// Relaxed-R1CS kernels: sparse mat-vec (Az, Bz, Cz), fused folding cross-term T, and the fold update.
// They replace, inside nova-snark 0.8.0 (called from the reference at src/nova/proof.rs:342-349):
//   R1CSShape::multiply_vec      -> MultiplyVecFn       
//   R1CSShape::commit_T (T part) -> CrossTermFn         (row a6;  T = Az1.Bz2 + Az2.Bz1 - u1.Cz2 - u2.Cz1)
//   RelaxedR1CSWitness::fold     -> FoldFn              (row a7;  W <- W1 + r W2,  E <- E1 + r T)
//
// Layout in HBM: the three matrices are stored back to back as ONE CSR with 3*cons rows (A rows, then B,
// then C): row_ptr u32[3*cons+1], col u32[nnz], val 32-byte Montgomery field elements [nnz].  z is never
// materialised: column j < vars reads W[j], j == vars reads u, j > vars reads X[j-vars-1], so W stays
// resident across steps.  All kernels are bound by HBM bandwidth (36 B per non-zero + 32 B per row out).
#pragma once
#include "field.cuh"
#include "launch.cuh"

namespace vdf {

struct CsrView {
  const uint32_t* row_ptr;  // [3*cons + 1]
  const uint32_t* col;
  const fe* val;
  uint32_t cons, vars, io;
};

struct ZView {   // z = [W | u | X]
  const fe* W;
  const fe* u;   // one element
  const fe* X;
};

template <class F>
VDF_HD fe z_at(const ZView& z, uint32_t vars, uint32_t col) {
  const fe* p = col < vars ? z.W + col : (col == vars ? z.u : z.X + (col - vars - 1));
  return fe_load_gather(p);   // random 32-byte gather: do not pull the rest of the line
}

// acc += v * x.  Most R1CS coefficients are +1 or -1 (every A/B entry and all but one C entry of the MinRoot
// step circuit, src/nova/proof.rs:176-227): those cost an addition instead of a 245-instruction multiplication,
// which is what moves these kernels from the multiply pipe back to the HBM roofline.  Exact either way.
template <class F>
VDF_HD fe mul_acc(const fe& acc, const fe& v, const fe& x, const fe& one, const fe& minus_one) {
  if (F::eq(v, one)) return F::add(acc, x);
  if (F::eq(v, minus_one)) return F::sub(acc, x);
  return F::add(acc, F::mul(v, x));
}

template <class F>
VDF_HD fe csr_row_dot(const CsrView& m, uint32_t row, const ZView& z) {
  fe acc = F::zero();
  const fe one = F::one(), minus_one = F::neg(F::one());
  uint32_t lo = m.row_ptr[row], hi = m.row_ptr[row + 1];
  for (uint32_t k = lo; k < hi; k++) {
    fe v = fe_load(m.val + k);
    fe x = z_at<F>(z, m.vars, m.col[k]);
    acc = mul_acc<F>(acc, v, x, one, minus_one);
  }
  return acc;
}

// one thread per (matrix, row): out[mat][row]
template <class F>
struct MultiplyVecFn {
  CsrView m;
  ZView z;
  fe* Az; fe* Bz; fe* Cz;
  VDF_HD void operator()(size_t idx) const {
    uint32_t row = (uint32_t)idx;
    fe r = csr_row_dot<F>(m, row, z);
    uint32_t mat = row / m.cons, rr = row - mat * m.cons;
    fe* out = mat == 0 ? Az : (mat == 1 ? Bz : Cz);
    fe_store(out + rr, r);
  }
};

// one thread per constraint row: six dot products sharing the row's (col, val) loads, then
// T = Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1   (u2 = 1 for the fresh instance)
template <class F>
struct CrossTermFn {
  CsrView m;
  ZView z1, z2;
  fe* T;
  VDF_HD void operator()(size_t idx) const {
    uint32_t row = (uint32_t)idx;
    const fe one = F::one(), minus_one = F::neg(F::one());
    uint32_t ptr[3][2];
#pragma unroll
    for (uint32_t mat = 0; mat < 3; mat++) {   // six independent loads: one memory latency for all three matrices
      ptr[mat][0] = m.row_ptr[mat * m.cons + row];
      ptr[mat][1] = m.row_ptr[mat * m.cons + row + 1];
    }
    // (Measured and removed in round 2: prefetch.global.L2 of the later matrices' (col, val) entries and of the z
    // elements they select, issued here -- 0.245 -> 0.283 ms on the 2^21-constraint shape.  The kernel is not bound by
    // its chain of dependent loads: with full-size B coefficients a row costs 5 field multiplications for 252 bytes,
    // i.e. 0.14 ms of multiply pipe against 0.08 ms of HBM for this shape)
    // the A and B products are combined as soon as both are known, so at most four dot products are live at once
    fe a1, a2, t;
#pragma unroll
    for (uint32_t mat = 0; mat < 3; mat++) {
      fe s1 = F::zero(), s2 = F::zero();
      uint32_t lo = ptr[mat][0], hi = ptr[mat][1];
      for (uint32_t k = lo; k < hi; k++) {
        fe v = fe_load(m.val + k);
        uint32_t c = m.col[k];
        s1 = mul_acc<F>(s1, v, z_at<F>(z1, m.vars, c), one, minus_one);
        s2 = mul_acc<F>(s2, v, z_at<F>(z2, m.vars, c), one, minus_one);
      }
      if (mat == 0) {
        a1 = s1; a2 = s2;
      } else if (mat == 1) {
        t = F::add(F::mul(a1, s2), F::mul(a2, s1));          // Az1*Bz2 + Az2*Bz1
      } else {
        t = F::sub(F::sub(t, F::mul(fe_load(z1.u), s2)), s1);   // - u1*Cz2 - Cz1
      }
    }
    fe_store(T + row, t);
  }
};

// ---- transposed product for Spartan's inner sum-check ---------------------------------------
// M(y) = sum_x eq(x) * (rA A[x,y] + rB B[x,y] + rC C[x,y]) for every column y: nova-snark's
// compute_eval_table_sparse combined with the three challenges [R], reached from CompressedSNARK::prove
// (src/nova/proof.rs:363).  The scaled row table eq3[mat * cons + x] = r_mat * eq(x) first, then one THREAD per column
// over the column view (40 bytes per entry read in order + one gathered table element); the few heavy columns (the
// constant column of the step circuit holds one entry per round, with a full-size coefficient) are left to one BLOCK
// each: threads stride over the entries, shuffle trees and shared memory add them up.
struct CscView {
  const uint32_t* col_ptr;  // [ncols + 1]
  const uint32_t* srow;     // stacked row = mat * cons + row
  const fe* val;            // column order
  const uint32_t* heavy;    // columns with more than heavy_above entries
  uint32_t heavy_above;
};

template <class F>
struct ScaleRowsFn {
  const fe* eq;     // [cons]
  const fe* coef;   // [3]
  uint32_t cons;
  fe* eq3;          // [3 * cons]
  VDF_HD void operator()(size_t idx) const {
    const uint32_t mat = (uint32_t)(idx / cons), row = (uint32_t)(idx - (size_t)mat * cons);
    fe_store(eq3 + idx, F::mul(fe_load(coef + mat), fe_load(eq + row)));
  }
};

template <class F>
struct BindRowsFn {   // index = column
  CscView m;
  const fe* eq3;
  fe* out;          // [ncols]
  VDF_HD void operator()(size_t col) const {
    const uint32_t lo = m.col_ptr[col], hi = m.col_ptr[col + 1];
    if (hi - lo > m.heavy_above) return;   // BindHeavyFn's
    const fe one = F::one(), minus_one = F::neg(F::one());
    fe acc = F::zero();
    for (uint32_t k = lo; k < hi; k++)
      acc = mul_acc<F>(acc, fe_load(m.val + k), fe_load_gather(eq3 + m.srow[k]), one, minus_one);
    fe_store(out + col, acc);
  }
};

template <class F>
struct BindHeavyFn {   // index = (position in the heavy list) * THREADS + thread; launched with whole blocks of THREADS
  static constexpr uint32_t THREADS = 512;
  CscView m;
  const fe* eq3;
  fe* out;
  VDF_HD void operator()(size_t idx) const {
    const uint32_t col = m.heavy[idx / THREADS];
    const uint32_t lo = m.col_ptr[col], hi = m.col_ptr[col + 1];
    const fe one = F::one(), minus_one = F::neg(F::one());
    fe acc = F::zero();
#if defined(__CUDA_ARCH__)
    // the constant column of a t-round step circuit has ~3t entries with full-size coefficients: one multiplication
    // each, spread over the block; shuffle tree per warp, then the warp totals through shared memory
    __shared__ fe warp_tot[THREADS / 32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t k = lo + tid; k < hi; k += THREADS)
      acc = mul_acc<F>(acc, fe_load(m.val + k), fe_load_gather(eq3 + m.srow[k]), one, minus_one);
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
      fe o;
#pragma unroll
      for (int q = 0; q < 8; q++) o.v[q] = __shfl_down_sync(0xffffffffu, acc.v[q], d);
      acc = F::add(acc, o);
    }
    if (lane == 0) warp_tot[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      acc = lane < THREADS / 32 ? warp_tot[lane] : F::zero();
#pragma unroll 1
      for (int d = 8; d >= 1; d >>= 1) {
        fe o;
#pragma unroll
        for (int q = 0; q < 8; q++) o.v[q] = __shfl_down_sync(0xffffffffu, acc.v[q], d);
        acc = F::add(acc, o);
      }
      if (lane == 0) fe_store(out + col, acc);
    }
#else
    if (idx % THREADS) return;
    for (uint32_t k = lo; k < hi; k++) acc = mul_acc<F>(acc, fe_load(m.val + k), fe_load(eq3 + m.srow[k]), one, minus_one);
    fe_store(out + col, acc);
#endif
  }
};

// a[i] <- a[i] + r * b[i] over two vectors in one launch (W with W2, E with T)
template <class F>
struct FoldFn {
  fe* W1; const fe* W2; size_t nW;
  fe* E1; const fe* T; size_t nE;
  const fe* r;
  VDF_HD void operator()(size_t i) const {
    fe rr = fe_load(r);
    if (i < nW) {
      fe_store(W1 + i, F::add(fe_load(W1 + i), F::mul(rr, fe_load(W2 + i))));
    } else {
      size_t j = i - nW;
      fe_store(E1 + j, F::add(fe_load(E1 + j), F::mul(rr, fe_load(T + j))));
    }
  }
};

}  // namespace vdf
