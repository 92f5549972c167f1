// Short-Weierstrass y^2 = x^3 + 5 (Pallas over Fp, Vesta over Fq; a = 0) group law in XYZZ
// coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity <=> ZZ == 0.  Used by the MSM kernels that
// replace pasta-msm's CPU Pippenger.  Formulas: EFD "xyzz" for
// short Weierstrass (madd-2008-s 8M+2S, add-2008-s 12M+2S, dbl-2008-s-1 with a = 0).
// Everything here is __host__ __device__ so tests/emul can run the same code on the CPU.
#pragma once
#include "field.cuh"

namespace vdf {

struct affine_t {  // device-side packed affine point, 64 B; (0,0) (not on the curve) = identity
  fe x, y;
};

struct xyzz_t {
  fe X, Y, ZZ, ZZZ;
};

struct jac_t {  // pasta_curves Ep/Eq memory layout (X, Y, Z), identity <=> Z == 0
  fe X, Y, Z;
};

template <class F>
struct Curve {
  typedef F field;   // coordinate field
  static VDF_HD bool aff_is_inf(const affine_t& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
  static VDF_HD bool is_inf(const xyzz_t& p) { return F::is_zero(p.ZZ); }

  static VDF_HD xyzz_t identity() {
    xyzz_t r;
    r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero();
    return r;
  }

  static VDF_HD xyzz_t from_affine(const affine_t& p) {
    xyzz_t r;
    if (aff_is_inf(p)) return identity();
    r.X = p.x; r.Y = p.y; r.ZZ = F::one(); r.ZZZ = F::one();
    return r;
  }

  // 2 * (x, y) for an affine point (mdbl-2008-s-1, a = 0)
  static VDF_HD xyzz_t dbl_affine(const fe& x, const fe& y) {
    xyzz_t r;
    fe U = F::dbl(y);
    fe V = F::sqr(U);
    fe W = F::mul(U, V);
    fe S = F::mul(x, V);
    fe xx = F::sqr(x);
    fe M = F::add(F::dbl(xx), xx);
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
  }

  // dbl-2008-s-1, a = 0.  y == 0 cannot happen on a prime-order curve; identity maps to identity
  // because ZZ3 = V * ZZ1 = 0.
  static VDF_HD xyzz_t dbl(const xyzz_t& p) {
    xyzz_t r;
    fe U = F::dbl(p.Y);
    fe V = F::sqr(U);
    fe W = F::mul(U, V);
    fe S = F::mul(p.X, V);
    fe xx = F::sqr(p.X);
    fe M = F::add(F::dbl(xx), xx);
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, p.Y));
    r.ZZ = F::mul(V, p.ZZ);
    r.ZZZ = F::mul(W, p.ZZZ);
    return r;
  }

  // acc += (x2, y2) with the affine point NOT the identity (madd-2008-s); handles acc == identity,
  // acc == +-(x2, y2).
  static VDF_HD void madd(xyzz_t& acc, const fe& x2, const fe& y2) {
    if (is_inf(acc)) {
      acc.X = x2; acc.Y = y2; acc.ZZ = F::one(); acc.ZZZ = F::one();
      return;
    }
    fe U2 = F::mul(x2, acc.ZZ);
    fe S2 = F::mul(y2, acc.ZZZ);
    fe P = F::sub(U2, acc.X);
    fe R = F::sub(S2, acc.Y);
    if (F::is_zero(P)) {
      if (F::is_zero(R)) acc = dbl_affine(x2, y2);
      else acc = identity();
      return;
    }
    fe PP = F::sqr(P);
    fe PPP = F::mul(P, PP);
    fe Q = F::mul(acc.X, PP);
    fe X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    fe Y3 = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(acc.Y, PPP));
    acc.ZZ = F::mul(acc.ZZ, PP);
    acc.ZZZ = F::mul(acc.ZZZ, PPP);
    acc.X = X3;
    acc.Y = Y3;
  }

  // the doubling branch of a mixed addition is taken only when a bucket meets its own point again: keep it
  // out of the hot kernel's instruction stream
#if defined(__CUDACC__)
  static __device__ __noinline__ xyzz_t dbl_affine_cold(const fe& x, const fe& y) { return dbl_affine(x, y); }
#else
  static xyzz_t dbl_affine_cold(const fe& x, const fe& y) { return dbl_affine(x, y); }
#endif

#ifndef VDF_MADD_MUL
#define VDF_MADD_MUL mul_val   // by value: operands and result stay in registers across the call (mul_call: via the stack)
#define VDF_MADD_SQR sqr_val
#endif
  // madd with the multiplier called out of line (same arithmetic; used by the accumulate kernel)
  static VDF_HD void madd_call(xyzz_t& acc, const fe& x2, const fe& y2) {
    if (is_inf(acc)) {
      acc.X = x2; acc.Y = y2; acc.ZZ = F::one(); acc.ZZZ = F::one();
      return;
    }
    fe U2 = F::VDF_MADD_MUL(x2, acc.ZZ);
    fe S2 = F::VDF_MADD_MUL(y2, acc.ZZZ);
    fe P = F::sub(U2, acc.X);
    fe R = F::sub(S2, acc.Y);
    if (F::is_zero(P)) {
      if (F::is_zero(R)) acc = dbl_affine_cold(x2, y2);
      else acc = identity();
      return;
    }
    fe PP = F::VDF_MADD_SQR(P);
    fe PPP = F::VDF_MADD_MUL(P, PP);
    fe Q = F::VDF_MADD_MUL(acc.X, PP);
    fe X3 = F::sub(F::sub(F::VDF_MADD_SQR(R), PPP), F::dbl(Q));
    fe Y3 = F::sub(F::VDF_MADD_MUL(R, F::sub(Q, X3)), F::VDF_MADD_MUL(acc.Y, PPP));
    acc.ZZ = F::VDF_MADD_MUL(acc.ZZ, PP);
    acc.ZZZ = F::VDF_MADD_MUL(acc.ZZZ, PPP);
    acc.X = X3;
    acc.Y = Y3;
  }

  // acc += sign ? -(p) : p for a packed affine point (identity allowed)
  static VDF_HD void madd_signed(xyzz_t& acc, const affine_t& p, bool negate) {
    if (aff_is_inf(p)) return;
    fe y = negate ? F::neg(p.y) : p.y;
    madd(acc, p.x, y);
  }
  static VDF_HD void madd_signed_call(xyzz_t& acc, const affine_t& p, bool negate) {
    if (aff_is_inf(p)) return;
    fe y = negate ? F::neg(p.y) : p.y;
    madd_call(acc, p.x, y);
  }

  // acc += q (add-2008-s), all special cases handled
  static VDF_HD void add(xyzz_t& acc, const xyzz_t& q) {
    if (is_inf(q)) return;
    if (is_inf(acc)) { acc = q; return; }
    fe U1 = F::mul(acc.X, q.ZZ);
    fe U2 = F::mul(q.X, acc.ZZ);
    fe S1 = F::mul(acc.Y, q.ZZZ);
    fe S2 = F::mul(q.Y, acc.ZZZ);
    fe P = F::sub(U2, U1);
    fe R = F::sub(S2, S1);
    if (F::is_zero(P)) {
      if (F::is_zero(R)) acc = dbl(acc);
      else acc = identity();
      return;
    }
    fe PP = F::sqr(P);
    fe PPP = F::mul(P, PP);
    fe Q = F::mul(U1, PP);
    fe X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    fe Y3 = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(S1, PPP));
    acc.ZZ = F::mul(F::mul(acc.ZZ, q.ZZ), PP);
    acc.ZZZ = F::mul(F::mul(acc.ZZZ, q.ZZZ), PPP);
    acc.X = X3;
    acc.Y = Y3;
  }

  static VDF_HD xyzz_t neg(const xyzz_t& p) {
    xyzz_t r = p;
    r.Y = F::neg(p.Y);
    return r;
  }

  // canonical output: affine (x, y) with Z = 1, or (0, 0, 0) for the identity -- a valid
  // pasta_curves Ep/Eq value whose bytes are unique per group element.
  static VDF_HD jac_t to_jac_normalised(const xyzz_t& p) {
    jac_t r;
    if (is_inf(p)) {
      r.X = F::zero(); r.Y = F::zero(); r.Z = F::zero();
      return r;
    }
    // 1/ZZZ = i3;  1/ZZ = i3^2 * ZZ^2 ... cheaper: i = 1/(ZZ*ZZZ); 1/ZZ = i*ZZZ; 1/ZZZ = i*ZZ
    fe i = F::inv(F::mul(p.ZZ, p.ZZZ));
    r.X = F::mul(p.X, F::mul(i, p.ZZZ));
    r.Y = F::mul(p.Y, F::mul(i, p.ZZ));
    r.Z = F::one();
    return r;
  }

  // un-normalised Jacobian image of an XYZZ point: Z = ZZ (= z^2 of the underlying Jacobian z), so
  // X' = X*ZZ and Y' = Y*ZZZ satisfy X'/Z^2 = X/ZZ and Y'/Z^3 = Y/ZZZ.  Two multiplications, no inversion.
  static VDF_HD jac_t to_jac_raw(const xyzz_t& p) {
    jac_t r;
    if (is_inf(p)) {
      r.X = F::zero(); r.Y = F::zero(); r.Z = F::zero();
      return r;
    }
    r.X = F::mul(p.X, p.ZZ);
    r.Y = F::mul(p.Y, p.ZZZ);
    r.Z = p.ZZ;
    return r;
  }

  static VDF_HD xyzz_t from_jac(const jac_t& p) {
    xyzz_t r;
    if (F::is_zero(p.Z)) return identity();
    r.X = p.X; r.Y = p.Y;
    r.ZZ = F::sqr(p.Z);
    r.ZZZ = F::mul(r.ZZ, p.Z);
    return r;
  }

  // k * p, k an unsigned 32-bit integer (bucket-index weights in the reduction tree)
  static VDF_HD xyzz_t mul_u32(const xyzz_t& p, uint32_t k) {
    xyzz_t r = identity();
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
      r = dbl(r);
      if ((k >> bit) & 1u) add(r, p);
    }
    return r;
  }
};

typedef Curve<Fp> Pallas;  // coordinates in Fp
typedef Curve<Fq> Vesta;   // coordinates in Fq

}  // namespace vdf
