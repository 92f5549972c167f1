// Sum-check building blocks over multilinear polynomials in evaluation form (tables of 2^k field elements):
// what CompressedSNARK::prove (reference src/nova/proof.rs:360-368) spends its non-MSM time in, inside nova-snark
// 0.8's spartan_with_ipa_pc.  [R]: that crate is not under /root/reference; the
// semantics below restate its sumcheck.rs / polynomial.rs from memory and are unique mathematical objects:
//
//   EqPolynomial::evals            eq[idx] = prod_j (bit_j(idx) ? r_j : 1 - r_j), r_0 <-> most significant bit
//   prove_cubic_with_additive_term one round: (e0, e2, e3) = sum_i comb(P(0), ..), comb(P(2), ..), comb(P(3), ..) with
//                                  comb(A, B, C, D) = A (B C - D), P(0) = lo, P(2) = 2 hi - lo, P(3) = 3 hi - 2 lo,
//                                  lo = P[i], hi = P[i + len/2]   (outer sum-check: A = eq(tau), B = Az, C = Bz, D = u Cz + E)
//   prove_quad                     one round: (e0, e2) with comb(A, B) = A B  (inner sum-check: ABC(r_x, .) and z)
//   bound_poly_var_top             P[i] <- lo + r (hi - lo), length halves
//   MultilinearPolynomial::evaluate  = <eq(r), P>
//
// All of it is streaming work over tables: 32 B per element read once per round -- HBM-bound at large sizes (the
// cubic round does 6 multiplications per 256 bytes: about even between the multiply pipe and HBM), launch-bound at
// Nova's own sizes (2^14 - 2^17 entries, L2-resident).  Reductions: per-thread strided partial sums, a shuffle tree
// per warp, shared memory across the warps of a block, one partial per block, a second one-block launch.
#pragma once
#include "curve.cuh"
#include "field.cuh"
#include "launch.cuh"

namespace vdf {

// ---- inner-product-argument building blocks (the polynomial-commitment opening of spartan_with_ipa_pc, [R]) --------
// One IPA round halves the vectors and the generators with the challenge r:
//   a' = r a_L + r^-1 a_R,  b' = r^-1 b_L + r b_R              (VecLinCombFn: out[i] = x a[i] + y b[i])
//   G'_i = r^-1 G_L,i + r G_R,i   (CommitGens::fold: n/2 independent two-term scalar multiplications;
//                                  nova computes each with vartime_multiscalar_mul on two points)
//   c_L = <a_L, b_R>, c_R = <a_R, b_L>                        (sc_dot_kernel below)
template <class F>
struct VecLinCombFn {
  const fe* a; const fe* b; const fe* xy;   // xy[0] = x, xy[1] = y (Montgomery)
  fe* out;
  VDF_HD void operator()(size_t i) const {
    fe_store(out + i, F::add(F::mul(fe_load(xy), fe_load(a + i)), F::mul(fe_load(xy + 1), fe_load(b + i))));
  }
};

// out[i] = w1 P[i] + w2 Q[i] on packed affine points, normalised to affine again.  Thread = CH consecutive outputs:
// each by interleaved double-and-add over the bits of both scalars (Shamir's trick: one doubling per bit, one addition
// of P, Q or P + Q), then ONE inversion for the CH results (Montgomery's trick).  C = curve, SF = its scalar field.
template <class C, class F, class SF>
struct PointLinCombFn {
  const affine_t* P; const affine_t* Q;
  const fe* w;            // w[0] = w1, w[1] = w2 (Montgomery, scalar field)
  affine_t* out;
  size_t n;
  static constexpr int CH = 4;
  VDF_HD void operator()(size_t t) const {
    const size_t lo = t * CH, hi = lo + CH < n ? lo + CH : n;
    const fe k1 = SF::from_mont(fe_load(w)), k2 = SF::from_mont(fe_load(w + 1));
    xyzz_t res[CH];
    fe pref[CH];
    fe run = F::one();
    for (size_t i = lo; i < hi; i++) {
      affine_t p, q;
      p.x = fe_load(&P[i].x); p.y = fe_load(&P[i].y);
      q.x = fe_load(&Q[i].x); q.y = fe_load(&Q[i].y);
      xyzz_t pq = C::from_affine(p);
      C::madd_signed(pq, q, false);                       // P + Q (identity operands and P = +-Q handled inside)
      xyzz_t acc = C::identity();
#pragma unroll 1
      for (int bit = 255; bit >= 0; bit--) {
        acc = C::dbl(acc);
        const uint32_t b1 = (k1.v[bit >> 5] >> (bit & 31)) & 1u, b2 = (k2.v[bit >> 5] >> (bit & 31)) & 1u;
        if (b1 & b2) C::add(acc, pq);
        else if (b1) C::madd_signed(acc, p, false);
        else if (b2) C::madd_signed(acc, q, false);
      }
      res[i - lo] = acc;
      pref[i - lo] = run;
      if (!C::is_inf(acc)) run = F::mul(run, F::mul(acc.ZZ, acc.ZZZ));
    }
    fe inv = F::inv(run);
    for (size_t i = hi; i > lo; i--) {
      const xyzz_t& r = res[i - 1 - lo];
      affine_t a;
      if (C::is_inf(r)) {
        a.x = F::zero(); a.y = F::zero();
      } else {
        const fe zi = F::mul(inv, pref[i - 1 - lo]);        // 1 / (ZZ * ZZZ)
        inv = F::mul(inv, F::mul(r.ZZ, r.ZZZ));
        a.x = F::mul(r.X, F::mul(zi, r.ZZZ));
        a.y = F::mul(r.Y, F::mul(zi, r.ZZ));
      }
      fe_store(&out[i - 1].x, a.x);
      fe_store(&out[i - 1].y, a.y);
    }
  }
};

constexpr int SC_BLOCK = 256;
constexpr int SC_MAX_POLYS = 4;

struct PolySet {
  fe* p[SC_MAX_POLYS];
};

// eq table over the bit range [first, first + nbits) of r (most significant first): out[idx], idx < 2^nbits
template <class F>
struct EqPartFn {
  const fe* r;       // [ell] Montgomery
  uint32_t first, nbits;
  fe* out;
  VDF_HD void operator()(size_t idx) const {
    fe acc = F::one();
    const fe one = F::one();
    for (uint32_t j = 0; j < nbits; j++) {
      const fe rj = fe_load(r + first + j);
      const bool bit = (idx >> (nbits - 1 - j)) & 1u;
      acc = F::mul(acc, bit ? rj : F::sub(one, rj));
    }
    fe_store(out + idx, acc);
  }
};

// out[idx] = hi[idx >> lo_bits] * lo[idx & (2^lo_bits - 1)]: one multiplication and one 32-byte store per entry
template <class F>
struct EqCombineFn {
  const fe* hi;
  const fe* lo;
  uint32_t lo_bits;
  fe* out;
  VDF_HD void operator()(size_t idx) const {
    fe_store(out + idx, F::mul(fe_load(hi + (idx >> lo_bits)), fe_load(lo + (idx & (((size_t)1 << lo_bits) - 1)))));
  }
};

// P[i] <- P[i] + r (P[i + half] - P[i]) for i < half, k polynomials in one launch (index = poly * half + i)
template <class F>
struct BindTopFn {
  PolySet polys;
  size_t half;
  const fe* r;
  VDF_HD void operator()(size_t idx) const {
    const size_t q = idx / half, i = idx - q * half;
    fe* P = polys.p[q];
    const fe lo = fe_load(P + i), hi = fe_load(P + half + i);
    fe_store(P + i, F::add(lo, F::mul(fe_load(r), F::sub(hi, lo))));
  }
};

// BindTopFn with the challenge as a kernel argument: the round loop has it on the host, one upload fewer per round
template <class F>
struct BindTopValFn {
  PolySet polys;
  size_t half;
  fe r;
  VDF_HD void operator()(size_t idx) const {
    const size_t q = idx / half, i = idx - q * half;
    fe* P = polys.p[q];
    const fe lo = fe_load(P + i), hi = fe_load(P + half + i);
    fe_store(P + i, F::add(lo, F::mul(r, F::sub(hi, lo))));
  }
};

#if defined(__CUDACC__)
// ---- reductions ------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ fe warp_sum(fe v) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    fe o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = __shfl_down_sync(0xffffffffu, v.v[k], d);
    v = F::add(v, o);
  }
  return v;
}

// block-wide sums of NACC per-thread accumulators -> out[blockIdx.x * NACC + a]
template <class F, int NACC>
__device__ __forceinline__ void block_sum_store(fe (&acc)[NACC], fe* out) {
  __shared__ fe sm[NACC][SC_BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; a++) {
    fe s = warp_sum<F>(acc[a]);
    if (lane == 0) sm[a][wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int a = 0; a < NACC; a++) {
      fe s = lane < SC_BLOCK / 32 ? sm[a][lane] : F::zero();
      s = warp_sum<F>(s);
      if (lane == 0) fe_store(out + (size_t)blockIdx.x * NACC + a, s);
    }
  }
}

// one round of the cubic sum-check with additive term: per-block partials of (e0, e2, e3)
template <class F>
__global__ void __launch_bounds__(SC_BLOCK) sc_cubic_round_kernel(const fe* A, const fe* B, const fe* C, const fe* D,
                                                                   size_t half, fe* partial) {
  fe acc[3] = {F::zero(), F::zero(), F::zero()};
  for (size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x; i < half; i += (size_t)gridDim.x * SC_BLOCK) {
    fe a = fe_load(A + i), b = fe_load(B + i), c = fe_load(C + i), d = fe_load(D + i);
    acc[0] = F::add(acc[0], F::mul(a, F::sub(F::mul(b, c), d)));
    const fe ah = fe_load(A + half + i), bh = fe_load(B + half + i), ch = fe_load(C + half + i), dh = fe_load(D + half + i);
    const fe da = F::sub(ah, a), db = F::sub(bh, b), dc = F::sub(ch, c), dd = F::sub(dh, d);
    a = F::add(ah, da); b = F::add(bh, db); c = F::add(ch, dc); d = F::add(dh, dd);          // P(2) = 2 hi - lo
    acc[1] = F::add(acc[1], F::mul(a, F::sub(F::mul(b, c), d)));
    a = F::add(a, da); b = F::add(b, db); c = F::add(c, dc); d = F::add(d, dd);              // P(3) = P(2) + hi - lo
    acc[2] = F::add(acc[2], F::mul(a, F::sub(F::mul(b, c), d)));
  }
  block_sum_store<F, 3>(acc, partial);
}

// one round of the quadratic sum-check: per-block partials of (e0, e2)
template <class F>
__global__ void __launch_bounds__(SC_BLOCK) sc_quad_round_kernel(const fe* A, const fe* B, size_t half, fe* partial) {
  fe acc[2] = {F::zero(), F::zero()};
  for (size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x; i < half; i += (size_t)gridDim.x * SC_BLOCK) {
    const fe a = fe_load(A + i), b = fe_load(B + i), ah = fe_load(A + half + i), bh = fe_load(B + half + i);
    acc[0] = F::add(acc[0], F::mul(a, b));
    acc[1] = F::add(acc[1], F::mul(F::add(ah, F::sub(ah, a)), F::add(bh, F::sub(bh, b))));
  }
  block_sum_store<F, 2>(acc, partial);
}

// <a, b> over n entries: per-block partials
template <class F>
__global__ void __launch_bounds__(SC_BLOCK) sc_dot_kernel(const fe* A, const fe* B, size_t n, fe* partial) {
  fe acc[1] = {F::zero()};
  for (size_t i = (size_t)blockIdx.x * SC_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * SC_BLOCK)
    acc[0] = F::add(acc[0], F::mul(fe_load(A + i), fe_load(B + i)));
  block_sum_store<F, 1>(acc, partial);
}

// second stage: out[a] = sum over blocks of partial[block * NACC + a]; one block
template <class F, int NACC>
__global__ void __launch_bounds__(SC_BLOCK) sc_final_kernel(const fe* partial, uint32_t nblocks, fe* out) {
  fe acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; a++) acc[a] = F::zero();
  for (uint32_t b = threadIdx.x; b < nblocks; b += SC_BLOCK)
#pragma unroll
    for (int a = 0; a < NACC; a++) acc[a] = F::add(acc[a], fe_load(partial + (size_t)b * NACC + a));
  block_sum_store<F, NACC>(acc, out);   // gridDim.x == 1: writes out[0 .. NACC)
}

// grid for a reduction over `items`: enough blocks to fill the GPU, never more partials than the final block likes
static inline uint32_t sc_grid(size_t items) {
  size_t blocks = (items + SC_BLOCK - 1) / SC_BLOCK;
  const size_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (uint32_t)blocks;
}
#endif  // __CUDACC__

}  // namespace vdf
