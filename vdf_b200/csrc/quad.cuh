// Point operations by a QUAD of lanes (lanes 4k..4k+3 of a warp) for the latency-bound stages of an MSM.
//
// The tail of a commitment (record levels, bucket reduction: msm.cuh stages 5-7) is a chain of dependent point
// operations executed by a handful of warps, one warp per scheduler: what it costs is the dependent latency of a field
// multiplication (~700 cycles for a lone warp, tools/probe_latency.py) times the multiplications of the chain -- 14 per
// XYZZ addition, 9 per doubling when one lane does them one after the other.  The group law has only 4 (addition) or
// 3 (doubling) LEVELS of mutually independent multiplications, so here every lane of a quad holds the same operands,
// computes ONE product of the level, and the quad exchanges the products by shuffles: 4 multiplication latencies per
// addition instead of 14.  Same formulas as curve.cuh (add-2008-s, madd-2008-s, dbl-2008-s-1 with a = 0), same special
// cases; the results are the same group elements in a possibly different XYZZ representation.
//
// Device only.  Blocks are one-dimensional with a multiple of 32 threads; all four lanes of a quad must call together
// with identical operands (shuffles use the quad's own mask, so different quads of a warp may diverge).
#pragma once
#include "curve.cuh"

#if defined(__CUDACC__)
namespace vdf {

#ifndef VDF_QUAD_MUL
#define VDF_QUAD_MUL mul_val   // one shared multiplier body per kernel (operands in registers): this code runs once per warp, cold
#endif

template <class F>
struct Quad {
  typedef Curve<F> C;

  static __device__ __forceinline__ unsigned mask() { return 0xFu << (threadIdx.x & 28u); }
  static __device__ __forceinline__ unsigned role() { return threadIdx.x & 3u; }

  // operand of lane `r` out of four
  static __device__ __forceinline__ fe pick(unsigned r, const fe& a0, const fe& a1, const fe& a2, const fe& a3) {
    fe o;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t lo = (r & 1u) ? a1.v[k] : a0.v[k];
      const uint32_t hi = (r & 1u) ? a3.v[k] : a2.v[k];
      o.v[k] = (r & 2u) ? hi : lo;
    }
    return o;
  }
  static __device__ __forceinline__ fe pick(unsigned r, const fe& a0, const fe& a1, const fe& a2) {
    fe o;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t lo = (r & 1u) ? a1.v[k] : a0.v[k];
      o.v[k] = (r & 2u) ? a2.v[k] : lo;
    }
    return o;
  }
  static __device__ __forceinline__ fe pick(unsigned r, const fe& a0, const fe& a1) {
    fe o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = (r & 1u) ? a1.v[k] : a0.v[k];
    return o;
  }

  // the value lane `src` of the quad holds
  static __device__ __forceinline__ fe from(unsigned m, const fe& v, int src) {
    fe o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = __shfl_sync(m, v.v[k], src, 4);
    return o;
  }

  // 2 * p  (three levels: {U^2, X^2}, {U*V, X*V, M^2, V*ZZ}, {M*(S - X3), W*Y, W*ZZZ})
  static __device__ __forceinline__ xyzz_t dbl(const xyzz_t& p) {
    const unsigned m = mask(), r = role();
    xyzz_t o;
    const fe U = F::dbl(p.Y);
    fe t = pick(r, U, p.X);
    t = F::VDF_QUAD_MUL(t, t);
    const fe V = from(m, t, 0), xx = from(m, t, 1);
    const fe M = F::add(F::dbl(xx), xx);
    t = F::VDF_QUAD_MUL(pick(r, U, p.X, M, V), pick(r, V, V, M, p.ZZ));
    const fe W = from(m, t, 0), S = from(m, t, 1), MM = from(m, t, 2);
    o.ZZ = from(m, t, 3);
    o.X = F::sub(MM, F::dbl(S));
    t = F::VDF_QUAD_MUL(pick(r, M, W, W), pick(r, F::sub(S, o.X), p.Y, p.ZZZ));
    o.Y = F::sub(from(m, t, 0), from(m, t, 1));
    o.ZZZ = from(m, t, 2);
    return o;
  }

  // acc += q, all special cases (four levels)
  static __device__ __forceinline__ void add(xyzz_t& acc, const xyzz_t& q) {
    if (C::is_inf(q)) return;
    if (C::is_inf(acc)) { acc = q; return; }
    const unsigned m = mask(), r = role();
    fe t = F::VDF_QUAD_MUL(pick(r, acc.X, q.X, acc.Y, q.Y), pick(r, q.ZZ, acc.ZZ, q.ZZZ, acc.ZZZ));
    const fe U1 = from(m, t, 0), S1 = from(m, t, 2);
    const fe P = F::sub(from(m, t, 1), U1), R = F::sub(from(m, t, 3), S1);
    if (F::is_zero(P)) {
      if (F::is_zero(R)) acc = dbl(acc);
      else acc = C::identity();
      return;
    }
    t = F::VDF_QUAD_MUL(pick(r, P, R, acc.ZZ, acc.ZZZ), pick(r, P, R, q.ZZ, q.ZZZ));
    const fe PP = from(m, t, 0), RR = from(m, t, 1), Z2 = from(m, t, 2), Z3 = from(m, t, 3);
    t = F::VDF_QUAD_MUL(pick(r, P, U1, Z2), PP);
    const fe PPP = from(m, t, 0), Q = from(m, t, 1);
    acc.ZZ = from(m, t, 2);
    acc.X = F::sub(F::sub(RR, PPP), F::dbl(Q));
    t = F::VDF_QUAD_MUL(pick(r, R, S1, Z3), pick(r, F::sub(Q, acc.X), PPP, PPP));
    acc.Y = F::sub(from(m, t, 0), from(m, t, 1));
    acc.ZZZ = from(m, t, 2);
  }

  // the point another quad of the warp holds: `delta` lanes up (full-warp shuffle: all 32 lanes call)
  static __device__ __forceinline__ xyzz_t down(const xyzz_t& v, int delta) {
    xyzz_t o;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      o.X.v[k] = __shfl_down_sync(0xffffffffu, v.X.v[k], delta);
      o.Y.v[k] = __shfl_down_sync(0xffffffffu, v.Y.v[k], delta);
      o.ZZ.v[k] = __shfl_down_sync(0xffffffffu, v.ZZ.v[k], delta);
      o.ZZZ.v[k] = __shfl_down_sync(0xffffffffu, v.ZZZ.v[k], delta);
    }
    return o;
  }
};

}  // namespace vdf
#endif
