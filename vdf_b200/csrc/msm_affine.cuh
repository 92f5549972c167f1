// Batched-affine halving rounds between the sort and the XYZZ accumulation of the Pippenger MSM (msm.cuh).
//
// Every bucket of the sorted list is a plain sum of points, so one ROUND replaces each bucket segment
// e0 e1 e2 e3 ... by (e0+e1) (e2+e3) ... (an odd last element passes through).  All additions of a round are
// independent, which is what affine addition needs: lambda = (y2-y1)/(x2-x1) costs one inversion, and
// Montgomery's trick shares ONE inversion between all additions of a batch:
//   forward   p_j = d_0 d_1 ... d_(j-1) stored per addition, thread total P
//   invert    1/P, itself batched over the thread totals (BatchInvFn)
//   backward  1/d_j = inv * p_j; inv *= d_j; lambda, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1
// = 6 field multiplications per addition against 10 for the XYZZ mixed addition (curve.cuh).  After R rounds
// the list is 2^R times shorter and the range-based XYZZ accumulation finishes each bucket.
//
// Exceptional pairs (equal x: doubling or cancellation; an identity operand) are classified by pair_kind, the
// same way in both passes, and never poison a batch: their denominator is 2y (doubling) or they skip the batch.
#pragma once
#include "curve.cuh"
#include "launch.cuh"

namespace vdf {

VDF_HD uint32_t upper_bound_u32(const uint32_t* a, uint32_t n, uint32_t x) {
  // first index with a[idx] > x
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Where a round reads its operands: round 1 straight from the generator table through the sorted references
// (AoS, bit 31 of a reference = negate), later rounds from the previous round's output (SoA, so the forward
// pass, which needs only x, reads full cache lines).
struct PointSrc {
  const affine_t* aos;   // nullptr: SoA
  const fe* xs;
  const fe* ys;
  VDF_HD fe x(uint32_t ref) const {
    return aos ? fe_load_gather(&aos[ref & 0x7fffffffu].x) : fe_load(xs + ref);
  }
  template <class F>
  VDF_HD affine_t get(uint32_t ref) const {
    affine_t a;
    if (aos) {
      const affine_t* s = aos + (ref & 0x7fffffffu);
      a.x = fe_load_gather(&s->x);
      a.y = fe_load_gather(&s->y);
      if (ref >> 31) a.y = F::neg(a.y);   // -(0,0) = (0,0): the identity stays the identity
    } else {
      a.x = fe_load(xs + ref);
      a.y = fe_load(ys + ref);
    }
    return a;
  }
};

constexpr uint32_t PAIR_NONE = 0xffffffffu;

struct alignas(16) u32x4 {
  uint32_t a, b, c, d;
};

enum PairKind { PAIR_ADD = 0, PAIR_DBL, PAIR_TAKE1, PAIR_TAKE2, PAIR_ZERO };

template <class F>
VDF_HD int pair_kind(const affine_t& a, const affine_t& b) {
  if (F::is_zero(a.x) && F::is_zero(a.y)) return PAIR_TAKE2;
  if (F::is_zero(b.x) && F::is_zero(b.y)) return PAIR_TAKE1;
  if (!F::eq(a.x, b.x)) return PAIR_ADD;
  if (F::eq(a.y, b.y) && !F::is_zero(a.y)) return PAIR_DBL;
  return PAIR_ZERO;   // P + (-P); also y == 0, which no point of an odd-order curve has
}

struct HalfCountFn {
  const uint32_t* offs_in;
  uint32_t* cnt;
  VDF_HD void operator()(size_t b) const { cnt[b] = (offs_in[b + 1] - offs_in[b] + 1u) >> 1; }
};

// Operands of output o: ra[o], rb[o] = references of the two inputs (round 1: resolved through sref, so the
// arithmetic passes gather with one level of indirection; later: list positions); rb[o] = PAIR_NONE when the
// element passes through.
struct PairListFn {
  const uint32_t* offs_in;
  const uint32_t* offs_out;
  uint32_t NBK;
  const uint32_t* sref;   // nullptr after round 1
  uint32_t* ra;
  uint32_t* rb;
  static constexpr uint32_t CH = 16;   // consecutive outputs per thread: one binary search, then a walk
  VDF_HD void operator()(size_t t) const {
    const uint32_t M = offs_out[NBK];
    const uint64_t lo64 = (uint64_t)t * CH;
    if (lo64 >= M) return;
    const uint32_t lo = (uint32_t)lo64, hi = lo64 + CH < M ? (uint32_t)(lo64 + CH) : M;
    uint32_t b = upper_bound_u32(offs_out, NBK + 1, lo) - 1;
    uint32_t va[CH], vb[CH];
#pragma unroll
    for (uint32_t k = 0; k < CH; k++) {
      const uint32_t o = lo + k;
      if (o >= hi) { va[k] = 0; vb[k] = PAIR_NONE; continue; }
      while (o >= offs_out[b + 1]) b++;
      const uint32_t in0 = offs_in[b] + 2u * (o - offs_out[b]);
      const bool pair = in0 + 1 < offs_in[b + 1];
      va[k] = sref ? sref[in0] : in0;
      vb[k] = pair ? (sref ? sref[in0 + 1] : in0 + 1) : PAIR_NONE;
    }
    // the arrays are padded to a multiple of CH and 16-byte aligned: whole 16-byte stores, full sectors per thread
    u32x4* pa = reinterpret_cast<u32x4*>(ra + lo);
    u32x4* pb = reinterpret_cast<u32x4*>(rb + lo);
#pragma unroll
    for (uint32_t k = 0; k < CH; k += 4) {
      pa[k / 4] = u32x4{va[k], va[k + 1], va[k + 2], va[k + 3]};
      pb[k / 4] = u32x4{vb[k], vb[k + 1], vb[k + 2], vb[k + 3]};
    }
  }
};

// Thread t owns outputs t, t + T, t + 2T, ... (coalesced across the warp), at most K of them.
template <class F>
struct AffineFwdFn {
  PointSrc in;
  const uint32_t* ra;
  const uint32_t* rb;
  const uint32_t* m_out;   // -> number of outputs of this round
  uint32_t T, K;
  fe* prefix;              // [outputs]
  fe* ptot;                // [T]
  VDF_HD void operator()(size_t t) const {
    const uint32_t M = *m_out;
    fe run = F::one();
    // Software pipeline: the gathers of output j+1 are issued before the multiplication of output j, so two
    // iterations of random loads are in flight per thread (round 1 is bound by the latency of its gathers).
    uint32_t r1 = 0, r2 = PAIR_NONE;
    fe x1 = F::zero(), x2 = F::zero();
    bool live = t < M;
    if (live) load(t, r1, r2, x1, x2);
    for (uint32_t j = 0; live; j++) {
      const uint32_t o = (uint32_t)((uint64_t)j * T + t);
      const uint64_t on = (uint64_t)(j + 1) * T + t;
      const bool next = j + 1 < K && on < M;
      uint32_t n1 = 0, n2 = PAIR_NONE;
      fe nx1 = F::zero(), nx2 = F::zero();
      if (next) load((uint32_t)on, n1, n2, nx1, nx2);
      if (r2 != PAIR_NONE) {
        fe d = F::sub(x2, x1);
        bool use = true;
        if (F::is_zero(d) || F::is_zero(x1) || F::is_zero(x2)) {
          affine_t a = in.template get<F>(r1), b = in.template get<F>(r2);
          int kind = pair_kind<F>(a, b);
          if (kind == PAIR_DBL) d = F::dbl(a.y);
          else if (kind != PAIR_ADD) use = false;
        }
        if (use) {
          fe_store(prefix + o, run);
          run = F::mul(run, d);
        }
      }
      r1 = n1; r2 = n2; x1 = nx1; x2 = nx2;
      live = next;
    }
    fe_store(ptot + t, run);
  }
  VDF_HD void load(uint32_t o, uint32_t& r1, uint32_t& r2, fe& x1, fe& x2) const {
    r2 = rb[o];
    if (r2 == PAIR_NONE) return;
    r1 = ra[o];
    x1 = in.x(r1);
    x2 = in.x(r2);
  }
};

// v[i] <- 1 / v[i], J elements per thread with one Fermat inversion (no element is zero)
template <class F>
struct BatchInvFn {
  fe* v;
  uint32_t n;
#ifndef VDF_BATCHINV_J
#define VDF_BATCHINV_J 32
#endif
  static constexpr uint32_t J = VDF_BATCHINV_J;
  VDF_HD void operator()(size_t u) const {
    const uint32_t lo = (uint32_t)u * J, hi = lo + J < n ? lo + J : n;
    fe pre[J];
    fe run = F::one();
    for (uint32_t i = lo; i < hi; i++) {
      pre[i - lo] = run;
      run = F::mul(run, fe_load(v + i));
    }
    fe inv = F::inv(run);
    for (uint32_t i = hi; i > lo; i--) {
      fe x = fe_load(v + (i - 1));
      fe_store(v + (i - 1), F::mul(inv, pre[i - 1 - lo]));
      inv = F::mul(inv, x);
    }
  }
};

template <class F>
struct AffineBwdFn {
  PointSrc in;
  const uint32_t* ra;
  const uint32_t* rb;
  const uint32_t* m_out;
  uint32_t T, K;
  const fe* prefix;
  const fe* pinv;          // [T]: inverse of the thread total
  fe* out_x;               // [outputs]
  fe* out_y;
  VDF_HD void operator()(size_t t) const {
    const uint32_t M = *m_out;
    if (t >= M) return;
    uint32_t cnt = (uint32_t)(((uint64_t)M - t + T - 1) / T);   // outputs of this thread
    if (cnt > K) cnt = K;
    fe inv = fe_load(pinv + t);
    for (uint32_t j = cnt; j > 0; j--) {
      const uint32_t o = (uint32_t)((uint64_t)(j - 1) * T + t), r2 = rb[o];
      affine_t a = in.template get<F>(ra[o]);
      if (r2 == PAIR_NONE) {
        store(o, a);
        continue;
      }
      affine_t b = in.template get<F>(r2);
      fe d = F::sub(b.x, a.x);
      fe num;
      if (F::is_zero(d) || F::is_zero(a.x) || F::is_zero(b.x)) {
        int kind = pair_kind<F>(a, b);
        if (kind == PAIR_TAKE1) { store(o, a); continue; }
        if (kind == PAIR_TAKE2) { store(o, b); continue; }
        if (kind == PAIR_ZERO) { a.x = F::zero(); a.y = F::zero(); store(o, a); continue; }
        if (kind == PAIR_DBL) {
          d = F::dbl(a.y);
          fe xx = F::sqr(a.x);
          num = F::add(F::dbl(xx), xx);   // 3 x^2 (a = 0)
        } else {
          num = F::sub(b.y, a.y);
        }
      } else {
        num = F::sub(b.y, a.y);
      }
      fe dinv = F::mul(inv, fe_load(prefix + o));
      inv = F::mul(inv, d);
      fe lam = F::mul(num, dinv);
      affine_t r;
      r.x = F::sub(F::sub(F::sqr(lam), a.x), b.x);
      r.y = F::sub(F::mul(lam, F::sub(a.x, r.x)), a.y);
      store(o, r);
    }
  }
  VDF_HD void store(uint32_t o, const affine_t& p) const {
    fe_store(out_x + o, p.x);
    fe_store(out_y + o, p.y);
  }
};

// Items per thread close to `guess` such that the grid is a whole number of waves (`wave` = threads resident on
// the GPU at once): with only a few waves, a partly filled last wave costs as much as a full one.
static inline uint32_t fit_waves(size_t items, uint32_t guess, size_t wave) {
  if (guess < 1) guess = 1;
  size_t waves = (items + (size_t)guess * wave / 2) / ((size_t)guess * wave);   // nearest
  if (waves < 1) return guess;           // less than one wave of work: keep the parallelism
  if (waves > 16) return guess;          // many waves: the tail no longer matters
  return (uint32_t)((items + waves * wave - 1) / (waves * wave));
}

// upper bound of the list length after one round: sum ceil(s_b / 2) <= (sum s_b + buckets) / 2
static inline size_t affine_round_cap(size_t e_in, size_t nbk) { return (e_in + nbk + 1) / 2; }

#ifndef VDF_AFF_BWD_MINB
#define VDF_AFF_BWD_MINB 5
#endif
#ifndef VDF_AFF_FWD_MINB
#define VDF_AFF_FWD_MINB 7
#endif

// Runs `rounds` halving rounds.  In: the sorted references (sref, offs) over `pts`.  Out: *list_x / *list_y /
// *list_offs describe the reduced list (x and y arrays share ONE allocation, free *list_x and *list_offs).
// `e_cap` is the host-side bound of the list length, updated per round.
template <class L, class F>
void msm_affine_rounds(L& L_, uint32_t rounds, uint32_t K, uint32_t NBK, const uint32_t* sref, const affine_t* pts,
                       const uint32_t* offs, size_t& e_cap, fe** list_x, fe** list_y, uint32_t** list_offs) {
  PointSrc in{pts, nullptr, nullptr};
  const uint32_t* offs_in = offs;
  fe* cur = nullptr;
  uint32_t* cur_offs = nullptr;
  for (uint32_t r = 0; r < rounds; r++) {
    const size_t cap = affine_round_cap(e_cap, NBK);
    const uint32_t Kr = fit_waves(cap, K, (size_t)148 * VDF_AFF_BWD_MINB * 128);   // backward pass: whole waves
    const uint32_t T = (uint32_t)((cap + Kr - 1) / Kr);
    uint32_t* cnt = L_.template alloc<uint32_t>(NBK);
    uint32_t* offs_out = L_.template alloc<uint32_t>((size_t)NBK + 1);
    const size_t cap_pad = (cap + PairListFn::CH - 1) / PairListFn::CH * PairListFn::CH;
    uint32_t* refs = L_.template alloc<uint32_t>(2 * cap_pad);
    uint32_t *ra = refs, *rb = refs + cap_pad;
    fe* prefix = L_.template alloc<fe>(cap);
    fe* ptot = L_.template alloc<fe>(T);
    fe* out = L_.template alloc<fe>(2 * cap);
    L_.template run<256>(NBK, HalfCountFn{offs_in, cnt});
    L_.exclusive_scan(cnt, offs_out, NBK);
    L_.template run<128>((cap + PairListFn::CH - 1) / PairListFn::CH,
                         PairListFn{offs_in, offs_out, NBK, r == 0 ? sref : nullptr, ra, rb});
    L_.template run<128, VDF_AFF_FWD_MINB>(T, AffineFwdFn<F>{in, ra, rb, offs_out + NBK, T, Kr, prefix, ptot});
    L_.template run<64>((T + BatchInvFn<F>::J - 1) / BatchInvFn<F>::J, BatchInvFn<F>{ptot, T});
    L_.template run<128, VDF_AFF_BWD_MINB>(T, AffineBwdFn<F>{in, ra, rb, offs_out + NBK, T, Kr, prefix, ptot, out,
                                                             out + cap});
    L_.free(cnt); L_.free(refs); L_.free(prefix); L_.free(ptot);
    L_.free(cur); L_.free(cur_offs);
    cur = out;
    cur_offs = offs_out;
    in = PointSrc{nullptr, cur, cur + cap};
    offs_in = cur_offs;
    e_cap = cap;
  }
  *list_x = cur;
  *list_y = cur ? cur + e_cap : nullptr;
  *list_offs = cur_offs;
}

}  // namespace vdf
