// Batched MinRoot verification: one thread per independent (result, t, original) triple.
// Restates MinRootVDF::{inverse_step, inverse_round, inverse_eval, check} of the reference
// (src/minroot.rs:73-75 / :220-222, :338-344, :363-365, :369-371); Evaluation::verify / append
// (:424-438) are batches of the same check.  Per round: i' = i - 1; x' = y - i'; y' = x^5 - x' with
// x^5 = x * (x^2)^2 -- 2 squarings + 1 multiplication + 3 subtractions, 100 % integer-multiply pipe.
// States are the reference's State<F> in pasta_curves layout: x, y, i as 32-byte Montgomery elements.
#pragma once
#include "field.cuh"
#include "launch.cuh"

namespace vdf {

struct state_t {
  fe x, y, i;
};

template <class F>
VDF_HD state_t minroot_inverse_round(const state_t& s, const fe& one) {
  state_t r;
  r.i = F::sub(s.i, one);                      // minroot.rs:339
  r.x = F::sub(s.y, r.i);                      // minroot.rs:340
  fe x2 = F::sqr(s.x);
  fe x5 = F::mul(s.x, F::sqr(x2));             // minroot.rs:73-75
  r.y = F::sub(x5, r.x);                       // minroot.rs:341-342
  return r;
}

template <class F>
struct MinRootCheckFn {
  const state_t* results;
  const state_t* originals;
  const uint64_t* t_each;   // per-chain t, or nullptr
  uint64_t t_uniform;
  uint8_t* ok;
  VDF_HD void operator()(size_t idx) const {
    state_t s;
    s.x = fe_load(&results[idx].x);
    s.y = fe_load(&results[idx].y);
    s.i = fe_load(&results[idx].i);
    state_t o;
    o.x = fe_load(&originals[idx].x);
    o.y = fe_load(&originals[idx].y);
    o.i = fe_load(&originals[idx].i);
    // States come from outside: an encoding >= m is not a field element (pasta_curves' from_repr would have
    // refused it before check() ever ran), so the verdict is "no" rather than arithmetic on an unreduced value
    if (!(F::is_canonical(s.x) && F::is_canonical(s.y) && F::is_canonical(s.i) && F::is_canonical(o.x) &&
          F::is_canonical(o.y) && F::is_canonical(o.i))) {
      ok[idx] = 0;
      return;
    }
    const fe one = F::one();
    uint64_t t = t_each ? t_each[idx] : t_uniform;
#pragma unroll 1
    for (uint64_t k = 0; k < t; k++) s = minroot_inverse_round<F>(s, one);  // minroot.rs:363-365
    ok[idx] = (F::eq(s.x, o.x) && F::eq(s.y, o.y) && F::eq(s.i, o.i)) ? 1 : 0;  // minroot.rs:369-371
  }
};

// inverse_eval for many chains (used to produce witnesses/originals on device; same arithmetic)
template <class F>
struct MinRootInverseEvalFn {
  const state_t* results;
  uint64_t t;
  state_t* out;
  VDF_HD void operator()(size_t idx) const {
    state_t s;
    s.x = fe_load(&results[idx].x);
    s.y = fe_load(&results[idx].y);
    s.i = fe_load(&results[idx].i);
    const fe one = F::one();
#pragma unroll 1
    for (uint64_t k = 0; k < t; k++) s = minroot_inverse_round<F>(s, one);
    fe_store(&out[idx].x, s.x);
    fe_store(&out[idx].y, s.y);
    fe_store(&out[idx].i, s.i);
  }
};

// Step-circuit witness generation: the 4t+1 auxiliary values that
// InverseMinRootCircuit::synthesize allocates for one fold step -- per round new_x, tmp1 = x^2, tmp2 = tmp1^2,
// new_y = tmp2*x - new_x (src/nova/proof.rs:162-189), then final_i (:122-126) -- in allocation order, so the
// result is the step part of W and can feed commit(W) without passing through the host.  One thread per step:
// the circuits of a proof are independent once the VDF states are known (src/nova/proof.rs:284-296).
template <class F>
struct MinRootWitnessFn {
  const state_t* results;   // z_in = (x, y, i) of each step
  uint64_t t;
  fe* out;                  // [n][4t + 1]
  VDF_HD void operator()(size_t idx) const {
    fe x = fe_load(&results[idx].x), y = fe_load(&results[idx].y), i = fe_load(&results[idx].i);
    const fe one = F::one();
    fe* w = out + idx * (4 * t + 1);
#pragma unroll 1
    for (uint64_t k = 0; k < t; k++) {
      fe ni = F::sub(i, one);            // proof.rs:162-164
      fe nx = F::sub(y, ni);             // proof.rs:167-173
      fe t1 = F::sqr(x);                 // proof.rs:176
      fe t2 = F::sqr(t1);                // proof.rs:178
      fe ny = F::sub(F::mul(t2, x), nx); // proof.rs:181-189
      fe_store(w + 4 * k + 0, nx);
      fe_store(w + 4 * k + 1, t1);
      fe_store(w + 4 * k + 2, t2);
      fe_store(w + 4 * k + 3, ny);
      x = nx; y = ny; i = ni;
    }
    fe_store(w + 4 * t, i);              // final_i, proof.rs:122-126
  }
};

}  // namespace vdf
