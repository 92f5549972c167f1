// Pippenger multi-scalar multiplication  sum_i s_i * P_i  on Pallas / Vesta.
//
// Replaces the CPU MSM behind nova-snark's commit(W) / commit(T): Group::vartime_multiscalar_mul ->
// pasta_msm::{pallas,vesta} -> mult_pippenger_{pallas,vesta} (reached from
// the reference at src/nova/proof.rs:342-349).
//
// Pipeline (every stage is a functor over a 1-D index space, see launch.cuh):
//   1 digits     scalar (Montgomery) -> canonical -> signed c-bit digits; one key per (window, point);
//                histogram of bucket sizes with atomics; zero digits are dropped here.
//   2 scan       exclusive prefix sum of the histogram -> bucket offsets.
//   3 scatter    one-pass radix (counting) sort of point references by bucket id.  Order inside a bucket
//                is arbitrary: the group is commutative and the result is canonicalised at the end.
//   4 accumulate optional batched-affine halving rounds (msm_affine.cuh), then fixed-size ranges of the sorted list per thread (perfect balance whatever the digit
//                distribution), XYZZ mixed additions; bucket pieces that straddle a range boundary go
//                to a record list, complete buckets are stored directly.
//   5 fixup      segmented reduction of the record list (log depth: one warp per 32 records, shuffle sums),
//                then owners fold what is left.
//   6 reduce     sum_j j*B_j.  Small sets: bit decomposition (row b = sum of the buckets whose index has bit
//                b set, radix-4 tree sums, then 2^b applied per row in parallel).  Large sets: one level of
//                chunk-local running sums, then the same bit decomposition over the chunk totals.
//   7 final      tree sum of the partial results, Horner over windows (plain mode only), normalise to
//                (x, y, 1) and write a pasta_curves Jacobian point.
//
// Two generator layouts:
//   plain : points P_i only; window w has its own bucket set (NB = W sets).
//   table : levels 2^(c*w) * P_i precomputed once per generator set (Nova's generators are fixed for the
//           life of PublicParams, src/nova/proof.rs:232-237); all windows share ONE bucket set, so the
//           bucket reduction is paid once and there is no window combination.  Costs W x the point
//           memory -- what 180 GB of HBM is for.
#pragma once
#include "curve.cuh"
#include "launch.cuh"
#include "msm_affine.cuh"
#include "quad.cuh"

namespace vdf {

struct MsmPlan {
  uint32_t n = 0;          // points in this MSM
  uint32_t c = 0;          // window bits
  uint32_t W = 0;          // windows = ceil(256 / c)
  uint32_t B = 0;          // buckets per set = 2^(c-1), digit magnitudes 1..B
  uint32_t NB = 0;         // bucket sets: W (plain) or 1 (table)
  uint32_t table = 0;      // 1: points[w * level_stride + i] = 2^(c w) P_i
  uint64_t level_stride = 0;
  uint32_t S = 64;         // sorted entries per accumulate thread
  uint32_t G = 16;         // records per fixup-level thread
  uint32_t logm = 3;       // bucket-reduction chunk = 2^logm buckets
  uint32_t is_mont = 1;    // scalars arrive in Montgomery form
  uint32_t batch = 1;      // independent scalar vectors over the SAME points (one result each), <= MSM_MAX_BATCH
  uint32_t len[4] = {0, 0, 0, 0};  // length of each vector (<= n); shorter vectors are zero-padded
  uint32_t raw_jacobian = 0;       // 1: write (X*ZZ, Y*ZZZ, ZZ) without the final inversion (caller normalises)
  uint32_t rec_warp = 0;           // 1: record levels by warp-cooperative segmented sums (RecWarpLevelFn)
  uint32_t rec_bucket = 0;         // 1: small bucket sets reduce their boundary records per bucket (RecBucketFn; off: see there)
  uint32_t affine_rounds = 0;      // batched-affine halving rounds before the XYZZ accumulation (msm_affine.cuh)
  uint32_t affine_K = 64;          // additions per thread and round sharing one running product
};

constexpr uint32_t MSM_MAX_BATCH = 4;

constexpr uint32_t MSM_REC_BUCKET_MAX = 32768;   // bucket sets up to this size reduce their records per bucket (RecBucketFn)

#ifndef VDF_ACC_MINB
#define VDF_ACC_MINB 5   // resident blocks per SM the XYZZ accumulation kernel is compiled for (96 registers)
#endif

struct ScalarSet {
  const fe* v[MSM_MAX_BATCH];
};

enum MsmStage {
  MSM_STAGE_DIGITS = 0, MSM_STAGE_SCAN, MSM_STAGE_SCATTER, MSM_STAGE_ACCUMULATE, MSM_STAGE_RECORDS,
  MSM_STAGE_REDUCE, MSM_STAGE_FINAL, MSM_STAGE_END
};

constexpr uint32_t KEY_SKIP = 0xffffffffu;
constexpr uint32_t REC_NONE = 0xffffffffu;
constexpr uint32_t REC_FIRST = 1u, REC_LAST = 2u;

static inline uint32_t msm_windows(uint32_t c) { return (256 + c - 1) / c; }

// window size heuristic (bits): balances n*W additions against bucket-reduction work
static inline uint32_t msm_pick_c(size_t n, bool table) {
  uint32_t lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c;
  if (table) {
    c = (int)lg - 1;            // one shared bucket set of 2^(c-1) ~ n/4 buckets: reduction stays ~10 %
    // 2^15 .. 2^18 points: one bit less.  The bucket reduction is a latency chain that grows with the bucket count
    // while the accumulation is still small; c = 16 and c = 17 even have the same 16 windows.  Measured with
    // tools/sweep_window_bits.py (uniform scalars, us per commitment): 2^15: 410 -> 381 (c = 13), 2^17: 774 -> 758
    // (c = 15), 2^18: 1280 -> 1187 (c = 16); 2^16 lands on c = 15 either way (14 is skipped below).
    if (lg >= 15 && lg <= 18) c = (int)lg - 2;
  }
  else c = (int)lg - 6;         // W bucket sets of 2^(c-1) buckets each
  if (c < 4) c = 4;
  if (c > 20) c = 20;
  // Scalars are 254-bit values (both moduli are just above 2^254), so the TOP window holds 254 - (W-1)*c usable bits.
  // With c = 11, 12, 14 or 18 that is 1 or 2 bits: every scalar then drops one entry into one of at most three
  // buckets -- n/4 to n/2 entries each -- and those heavy buckets are cut into hundreds of boundary records whose
  // reduction is the longest dependent chain of a latency-regime MSM (measured: 13 904 points at c = 12 take 0.51 ms,
  // 16 384 points at c = 13 take 0.43 ms).  The next window size up has a healthy top window (7 or more bits, or none).
  while (c == 11 || c == 12 || c == 14 || c == 18) c++;
  return (uint32_t)c;
}

// ---- stage 1: digits + histogram -----------------------------------------------------------------
template <class SF>  // SF = scalar field
struct DigitsFn {
  ScalarSet scalars;
  uint32_t* keys;    // [batch][W][n]: bucket_global | sign << 31, or KEY_SKIP
  uint32_t* rank;    // [batch][W][n]: arrival number of the entry in its bucket (what the histogram atomic returns)
  uint32_t* count;   // [batch * NB * B]
  MsmPlan p;
  VDF_HD void operator()(size_t idx) const {
    const uint32_t bt = (uint32_t)(idx / p.n), i = (uint32_t)(idx - (size_t)bt * p.n);
    uint32_t* kb = keys + (size_t)bt * p.W * p.n;
    uint32_t* rb = rank + (size_t)bt * p.W * p.n;
    if (i >= p.len[bt]) {   // zero padding of a shorter vector
      for (uint32_t w = 0; w < p.W; w++) kb[(size_t)w * p.n + i] = KEY_SKIP;
      return;
    }
    fe s = fe_load(scalars.v[bt] + i);
    if (p.is_mont) s = SF::from_mont(s);
    uint32_t carry = 0;
    const uint32_t c = p.c, full = 1u << c;
    constexpr uint32_t U = 4;   // windows per group: their histogram atomics are in flight together
    for (uint32_t w0 = 0; w0 < p.W; w0 += U) {
      uint32_t key[U], rk[U];
#pragma unroll
      for (uint32_t k = 0; k < U; k++) {
        const uint32_t w = w0 + k;
        key[k] = KEY_SKIP;
        if (w >= p.W) continue;
        uint32_t bit = w * c, limb = bit >> 5, sh = bit & 31;
        uint64_t two = s.v[limb];
        if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
        uint32_t raw = (uint32_t)((two >> sh) & (full - 1)) + carry;
        if (raw > p.B) {
          uint32_t mag = full - raw;  // digit = raw - 2^c = -mag
          carry = 1;
          key[k] = mag ? ((mag - 1) | 0x80000000u) : KEY_SKIP;
        } else {
          carry = 0;
          key[k] = raw ? (raw - 1) : KEY_SKIP;
        }
        if (key[k] != KEY_SKIP) {
          uint32_t set = bt * p.NB + (p.table ? 0u : w);
          uint32_t gb = set * p.B + (key[k] & 0x7fffffffu);
          rk[k] = atomic_add_u32(count + gb, 1u);
          key[k] = gb | (key[k] & 0x80000000u);
        }
      }
#pragma unroll
      for (uint32_t k = 0; k < U; k++) {
        const uint32_t w = w0 + k;
        if (w >= p.W) continue;
        kb[(size_t)w * p.n + i] = key[k];
        if (key[k] != KEY_SKIP) rb[(size_t)w * p.n + i] = rk[k];
      }
    }
  }
};

// ---- stage 3: scatter (counting sort) --------------------------------------------------------------
// The position of an entry is its bucket's offset + its arrival rank from stage 1: no second round of atomics.
// Thread = one scalar, looping over its windows exactly like DigitsFn, so entries reach a bucket in (nearly) the
// order of their ranks: the 32-byte sectors of the sorted list fill up while they are still in L2.  (Scattering
// window by window instead spreads the writes to every sector over the whole kernel: 1.7 ms instead of 0.9 ms at
// 2^22 points, where the list no longer fits in L2.)
struct ScatterFn {
  const uint32_t* keys;
  const uint32_t* rank;
  const uint32_t* offs;
  uint32_t* sref;
  MsmPlan p;
  VDF_HD void operator()(size_t idx) const {
    const uint32_t bt = (uint32_t)(idx / p.n), i = (uint32_t)(idx - (size_t)bt * p.n);
    const size_t base = (size_t)bt * p.W * p.n + i;
    constexpr uint32_t U = 4;   // independent lookups in flight
    for (uint32_t w0 = 0; w0 < p.W; w0 += U) {
      uint32_t key[U], pos[U];
#pragma unroll
      for (uint32_t k = 0; k < U; k++) {
        const size_t e = base + (size_t)(w0 + k) * p.n;
        key[k] = (w0 + k < p.W) ? keys[e] : KEY_SKIP;
        pos[k] = key[k] != KEY_SKIP ? rank[e] : 0u;
      }
#pragma unroll
      for (uint32_t k = 0; k < U; k++)
        if (key[k] != KEY_SKIP) pos[k] += offs[key[k] & 0x7fffffffu];
#pragma unroll
      for (uint32_t k = 0; k < U; k++) {
        if (key[k] == KEY_SKIP) continue;
        const uint32_t ref = p.table ? (uint32_t)((w0 + k) * p.level_stride + i) : i;
        sref[pos[k]] = ref | (key[k] & 0x80000000u);
      }
    }
  }
};

// ---- stages 4/5: range-based segmented accumulation ---------------------------------------------------
// A "record" is a partial sum of a bucket that straddles a range boundary.
struct RecHdr {
  uint32_t bucket;  // REC_NONE if unused
  uint32_t flags;   // REC_FIRST: piece starts the bucket; REC_LAST: piece ends it
};

// Chunked MSMs: every chunk accumulates into its OWN zeroed bucket array with plain stores (an addition inside
// the divergent flush path would serialise the warp), and this fully converged kernel folds it into the
// running array afterwards: total[b] += chunk[b].
template <class C>
struct BucketMergeFn {
  xyzz_t* total;
  const xyzz_t* chunk;
  VDF_HD void operator()(size_t b) const {
    xyzz_t c = chunk[b];
    if (C::is_inf(c)) return;
    xyzz_t t = total[b];
    C::add(t, c);
    total[b] = t;
  }
};

template <class C, bool DIRECT = false>   // DIRECT: the list holds the points themselves (after affine rounds)
struct AccumulateFn {
  const uint32_t* offs;  // [NBK + 1]
  uint32_t NBK;
  const uint32_t* sref;
  const affine_t* pts;
  const fe* xs;          // DIRECT: x and y of list position pos
  const fe* ys;
  xyzz_t* buckets;       // [NBK], zero-initialised
  RecHdr* rec_hdr;       // [2 * threads]
  xyzz_t* rec_pt;
  uint32_t S;
  VDF_HD void operator()(size_t t) const {
    rec_hdr[2 * t].bucket = REC_NONE;
    rec_hdr[2 * t + 1].bucket = REC_NONE;
    const uint32_t M = offs[NBK];
    const uint64_t lo64 = (uint64_t)t * S;
    if (lo64 >= M) return;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = (lo64 + S < M) ? (uint32_t)(lo64 + S) : M;
    uint32_t b = upper_bound_u32(offs, NBK + 1, lo) - 1;
    uint32_t bs = offs[b], be = offs[b + 1];
    uint32_t seg_start = lo;
    xyzz_t acc = C::identity();
    // ONE flat loop over the range: the lanes of a warp stay converged on the mixed addition and only
    // diverge for the (short) flush when their own bucket ends.
    for (uint32_t pos = lo; pos < hi; pos++) {
      if (pos == be) {
        flush(t, lo, b, bs, be, seg_start, pos, acc);
        b++;
        while (offs[b + 1] == pos) b++;   // skip empty buckets (pos < hi <= M guarantees termination)
        bs = pos;
        be = offs[b + 1];
        seg_start = pos;
        acc = C::identity();
      }
#if defined(__CUDA_ARCH__) && defined(VDF_ACC_PREFETCH)
      // pull the point needed VDF_ACC_PREFETCH iterations ahead towards L1/L2 while this addition computes
      if (pos + VDF_ACC_PREFETCH < hi) {
        const affine_t* nxt = pts + (sref[pos + VDF_ACC_PREFETCH] & 0x7fffffffu);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
      }
#endif
      affine_t pt;
      uint32_t ref = 0;
      if (DIRECT) {
        pt.x = fe_load(xs + pos);
        pt.y = fe_load(ys + pos);
      } else {
        ref = sref[pos];
        const affine_t* src = pts + (ref & 0x7fffffffu);
        pt.x = fe_load_gather(&src->x);
        pt.y = fe_load_gather(&src->y);
      }
      C::madd_signed_call(acc, pt, (ref >> 31) != 0);
    }
    flush(t, lo, b, bs, be, seg_start, hi, acc);
  }

  // a piece [seg_start, seg_end) of bucket b = [bs, be): complete -> bucket store, else -> record
  VDF_HD void flush(size_t t, uint32_t lo, uint32_t b, uint32_t bs, uint32_t be, uint32_t seg_start,
                    uint32_t seg_end, const xyzz_t& acc) const {
    const bool first = seg_start == bs, last = seg_end == be;
    if (first && last) {
      buckets[b] = acc;
    } else {
      size_t slot = (seg_start == lo) ? 2 * t : 2 * t + 1;
      rec_hdr[slot].bucket = b;
      rec_hdr[slot].flags = (first ? REC_FIRST : 0u) | (last ? REC_LAST : 0u);
      rec_pt[slot] = acc;
    }
  }
};

// One level of the segmented reduction over the record list: thread g folds records
// [g*G, (g+1)*G) ; runs of equal bucket id collapse; a run that both starts with REC_FIRST and ends
// with REC_LAST is a finished bucket, otherwise it is re-emitted as a record of the next level.
template <class C>
struct RecLevelFn {
  const RecHdr* in_hdr;
  const xyzz_t* in_pt;
  size_t n_in;
  xyzz_t* buckets;
  RecHdr* out_hdr;  // [2 * threads]
  xyzz_t* out_pt;
  uint32_t G;
  VDF_HD void operator()(size_t g) const {
    out_hdr[2 * g].bucket = REC_NONE;
    out_hdr[2 * g + 1].bucket = REC_NONE;
    size_t lo = g * G, hi = lo + G < n_in ? lo + G : n_in;
    uint32_t cur = REC_NONE, flags = 0;
    xyzz_t acc = C::identity();
    bool have_head = false;  // first emitted run goes to slot 2g, any later one to 2g+1
    for (size_t r = lo; r <= hi; r++) {
      uint32_t b = (r < hi) ? in_hdr[r].bucket : REC_NONE;
      if (r < hi && b == REC_NONE) continue;
      if (b != cur || r == hi) {
        if (cur != REC_NONE) {
          if ((flags & REC_FIRST) && (flags & REC_LAST)) {
            buckets[cur] = acc;
          } else {
            size_t slot = have_head ? 2 * g + 1 : 2 * g;
            out_hdr[slot].bucket = cur;
            out_hdr[slot].flags = flags;
            out_pt[slot] = acc;
          }
          have_head = true;
        }
        if (r == hi) break;
        cur = b;
        flags = 0;
        acc = C::identity();
      }
      flags |= in_hdr[r].flags;
      C::add(acc, in_pt[r]);
    }
  }
};

// The same level with one WARP per 32 records: a segmented sum over the lanes by shuffles (5 steps, one point addition
// each) instead of a serial loop in one thread -- 32 records -> <= 2 with a chain of 5 additions, where RecLevelFn
// needs 8 additions to bring 8 down to 2.  Launched with a multiple of 32 threads; index = input record slot.
// (The CPU emulation runs the serial functor with G = 32: same outputs.)
template <class C>
struct RecWarpLevelFn {
  const RecHdr* in_hdr;
  const xyzz_t* in_pt;
  size_t n_in;
  xyzz_t* buckets;
  RecHdr* out_hdr;  // [2 * groups of 32]
  xyzz_t* out_pt;
  VDF_HD void operator()(size_t idx) const {
#if defined(__CUDA_ARCH__)
    const unsigned lane = threadIdx.x & 31u, full = 0xffffffffu;
    const size_t g = idx >> 5;
    if (lane == 0) {
      out_hdr[2 * g].bucket = REC_NONE;
      out_hdr[2 * g + 1].bucket = REC_NONE;
    }
    __syncwarp();
    uint32_t own = idx < n_in ? in_hdr[idx].bucket : REC_NONE;
    uint32_t flags = own != REC_NONE ? in_hdr[idx].flags : 0u;
    xyzz_t v = C::identity();
    if (own != REC_NONE) v = in_pt[idx];
    // empty slots join the run on their left (identity value), so that runs are contiguous in the warp
    uint32_t key = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t up = __shfl_up_sync(full, key, d);
      if (lane >= (unsigned)d && key == REC_NONE) key = up;
    }
    const uint32_t left = __shfl_up_sync(full, key, 1);
    const bool head = key != REC_NONE && (lane == 0 || left != key);
    // is there a real run before this lane's run?  (the first unfinished run goes to slot 2g, a later one to 2g+1)
    const unsigned heads = __ballot_sync(full, head);
    const bool first_run = (heads & ((1u << lane) - 1u)) == 0u;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t ok = __shfl_down_sync(full, key, d);
      const uint32_t of = __shfl_down_sync(full, flags, d);
      xyzz_t o;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        o.X.v[k] = __shfl_down_sync(full, v.X.v[k], d);
        o.Y.v[k] = __shfl_down_sync(full, v.Y.v[k], d);
        o.ZZ.v[k] = __shfl_down_sync(full, v.ZZ.v[k], d);
        o.ZZZ.v[k] = __shfl_down_sync(full, v.ZZZ.v[k], d);
      }
      if (lane + d < 32 && ok == key && key != REC_NONE) {
        C::add(v, o);
        flags |= of;
      }
    }
    if (head) {
      if ((flags & REC_FIRST) && (flags & REC_LAST)) {
        buckets[key] = v;
      } else {
        const size_t slot = first_run ? 2 * g : 2 * g + 1;
        out_hdr[slot].bucket = key;
        out_hdr[slot].flags = flags;
        out_pt[slot] = v;
      }
    }
#else
    if ((idx & 31) == 0) RecLevelFn<C>{in_hdr, in_pt, n_in, buckets, out_hdr, out_pt, 32u}(idx >> 5);
#endif
  }
};

// The record levels of the latency regime: one block takes 256 consecutive slots to <= 2 records.
//
// The records of a bucket are contiguous in the list, so after squeezing out the empty slots a bucket is a RUN of
// equal keys.  Every record knows its rank inside its run; in round j the records of rank = 0 mod 2^(j+1) add the
// record 2^j places to their right when it belongs to the same run.  All runs shrink together -- the chain is
// log2(longest run of the block) additions, 3-4 with ~9 pieces per bucket, however many buckets the block holds (a
// binary tree over the slots would add across one of its node boundaries at every one of its log2(256) levels, because
// every boundary cuts some bucket).  Each addition is done by a quad (quad.cuh).  The run heads then hold the run
// totals: a run with its bucket's first AND last piece is stored, the others -- at most the leftmost and the rightmost
// run of the block -- go to the next level as records 2g and 2g+1.  Launched with a multiple of 256 threads.
// (The CPU emulation runs the serial functor with G = 256: same outputs.)
template <class C>
struct RecRunsFn {
  static constexpr uint32_t NS = 256;
  const RecHdr* in_hdr;
  const xyzz_t* in_pt;
  size_t n_in;
  xyzz_t* buckets;
  RecHdr* out_hdr;  // [2 * blocks]
  xyzz_t* out_pt;
#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ fe* comp(xyzz_t* p, unsigned r) {
    return r == 0 ? &p->X : r == 1 ? &p->Y : r == 2 ? &p->ZZ : &p->ZZZ;
  }
#endif
  VDF_HD void operator()(size_t idx) const {
#if defined(__CUDA_ARCH__)
    typedef Quad<typename C::field> Q;
    __shared__ xyzz_t sp[NS];
    __shared__ uint32_t ckey[NS + 1], cflg[NS], rnk[NS];
    __shared__ uint32_t wtot[NS / 32], wmax[NS / 32], s_lo, s_hi, s_cnt;
    __shared__ uint16_t wl[NS];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, full = 0xffffffffu;
    const size_t g = idx >> 8, base = g * NS;
    // squeeze out the empty slots (slot tid -> compact index p)
    RecHdr h;
    h.bucket = REC_NONE; h.flags = 0;
    if (base + tid < n_in) h = in_hdr[base + tid];
    const bool real = h.bucket != REC_NONE;
    const unsigned bal = __ballot_sync(full, real);
    if (lane == 0) wtot[warp] = __popc(bal);
    if (tid == 0) { s_lo = 0xffffffffu; s_hi = 0u; }
    __syncthreads();
    unsigned off = 0, M = 0;
#pragma unroll
    for (unsigned w = 0; w < NS / 32; w++) {
      if (w < warp) off += wtot[w];
      M += wtot[w];
    }
    if (real) {
      const unsigned p = off + __popc(bal & ((1u << lane) - 1u));
      ckey[p] = h.bucket;
      cflg[p] = h.flags;
      sp[p] = in_pt[base + tid];
    }
    if (tid == 0) ckey[M] = REC_NONE;
    __syncthreads();
    // rank inside the run = p - (start of the run): running maximum of the head positions
    {
      const bool head = tid < M && (tid == 0 || ckey[tid - 1] != ckey[tid]);
      unsigned v = head ? tid : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(full, v, d);
        if (lane >= (unsigned)d && o > v) v = o;
      }
      if (lane == 31) wmax[warp] = v;
      __syncthreads();
      for (unsigned w = 0; w < warp; w++) v = wmax[w] > v ? wmax[w] : v;
      if (tid < M) rnk[tid] = tid - v;
    }
    __syncthreads();
    // pairing rounds: the additions of a round are listed first, so that every quad gets its share of REAL work (the
    // lanes of a warp would otherwise all wait through the additions of the quad that has the most candidates)
    const unsigned q = tid >> 2, r = tid & 3u;
#pragma unroll 1
    for (unsigned st = 1; st < NS; st <<= 1) {
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      if (tid + st < M && (rnk[tid] & (2 * st - 1)) == 0 && ckey[tid + st] == ckey[tid]) wl[atomicAdd(&s_cnt, 1u)] = tid;
      __syncthreads();
      const unsigned cnt = s_cnt;
      if (cnt == 0) break;
#pragma unroll 1
      for (unsigned i = q; i < cnt; i += NS / 4) {
        const unsigned p = wl[i];
        xyzz_t a = sp[p];
        const xyzz_t o = sp[p + st];
        __syncwarp(Q::mask());   // every lane of the quad has read both points before any lane overwrites one
        Q::add(a, o);
        fe_store(comp(&sp[p], r), Q::pick(r, a.X, a.Y, a.ZZ, a.ZZZ));
        if (r == 0) cflg[p] |= cflg[p + st];
      }
      __syncthreads();
    }
    // run heads: finished buckets are stored, the (<= 2) others are the block's output
    if (tid < M && rnk[tid] == 0) {
      if ((cflg[tid] & REC_FIRST) && (cflg[tid] & REC_LAST)) {
        buckets[ckey[tid]] = sp[tid];
      } else {
        atomicMin(&s_lo, tid);
        atomicMax(&s_hi, tid);
      }
    }
    __syncthreads();
    if (tid < 2) {
      const bool have = tid == 0 ? s_lo != 0xffffffffu : (s_lo != 0xffffffffu && s_hi != s_lo);
      const unsigned p = tid == 0 ? s_lo : s_hi;
      RecHdr o;
      o.bucket = have ? ckey[p] : REC_NONE;
      o.flags = have ? cflg[p] : 0u;
      out_hdr[2 * g + tid] = o;
      if (have) out_pt[2 * g + tid] = sp[p];
    }
#else
    if ((idx & 255) == 0) RecLevelFn<C>{in_hdr, in_pt, n_in, buckets, out_hdr, out_pt, NS}(idx >> 8);
#endif
  }
};

// last level: the owner (record flagged REC_FIRST) folds the following records of its bucket
template <class C>
struct RecOwnerFn {
  const RecHdr* hdr;
  const xyzz_t* pt;
  size_t n_rec;
  xyzz_t* buckets;
  VDF_HD void operator()(size_t r) const {
    uint32_t b = hdr[r].bucket;
    if (b == REC_NONE || !(hdr[r].flags & REC_FIRST)) return;
    xyzz_t acc = pt[r];
    if (!(hdr[r].flags & REC_LAST)) {
      for (size_t q = r + 1; q < n_rec; q++) {
        if (hdr[q].bucket == REC_NONE) continue;
        if (hdr[q].bucket != b) break;  // cannot happen: REC_LAST closes the bucket first
        C::add(acc, pt[q]);
        if (hdr[q].flags & REC_LAST) break;
      }
    }
    buckets[b] = acc;
  }
};

// Latency-regime alternative to the record levels: ONE launch, one warp per bucket.  The pieces of bucket b sit at
// known slots of the record array -- range thread t covers list positions [t*S, (t+1)*S), so b = [bs, be) was cut by
// threads t0 = bs/S .. t1 = (be-1)/S; t0's piece is in slot 2*t0 (+1 if b starts inside its range), every later
// thread's in slot 2*t -- so lane j adds pieces j, j+32, ..., a shuffle tree adds the lanes, lane 0 stores the bucket.
// A bucket that lies inside one range was stored by AccumulateFn itself.  Replaces 3-5 dependent launches
// (RecWarpLevelFn / RecLevelFn / RecOwnerFn) when there are few buckets.
// OFF BY DEFAULT (VDFGPU_MSM_RECBUCKET=1 enables it; measured in round 2, profiles/r2_experiments.md): with uniform
// digits it cuts the records stage from 0.095 to 0.041 ms at 2^14 points (c = 13) and 0.134 to 0.113 ms at 75 344
// (c = 15), but real inputs have HEAVY buckets -- the few digits of a narrow top window (c = 12, 14: three buckets of
// n/4 entries), the many 0/1 values of a Nova witness -- whose hundreds of pieces a group (or even a whole warp, one
// heavy bucket after the other) chains serially: 0.124 -> 0.19-0.26 ms at 13 904 points.  The record levels are
// log-depth whatever the distribution.
template <class C>
struct RecBucketFn {
  static constexpr uint32_t LANES = 8;   // lanes per bucket: a bucket of the latency regime has ~5-10 pieces
  static constexpr uint32_t HEAVY = 24;  // more pieces than this: the whole warp sums the bucket
  const uint32_t* offs;  // [NBK + 1] of the list AccumulateFn ran over
  uint32_t NBK, S;
  const xyzz_t* rec_pt;
  xyzz_t* buckets;
  VDF_HD void operator()(size_t idx) const {
    const uint32_t b = (uint32_t)(idx / LANES);
    uint32_t pieces = 0, t0 = 0, first_slot = 0;
    if (b < NBK) {
      const uint32_t bs = offs[b], be = offs[b + 1];
      if (be > bs) {
        t0 = bs / S;
        const uint32_t t1 = (be - 1) / S;
        if (t1 != t0) {
          pieces = t1 - t0 + 1;
          first_slot = 2 * t0 + (bs == t0 * S ? 0u : 1u);
        }
      }
    }
#if defined(__CUDA_ARCH__)
    const unsigned sub = (unsigned)idx & (LANES - 1), lane = (unsigned)idx & 31u, full = 0xffffffffu;
    // Heavy buckets (the top window's few digits, a witness's many 1s: hundreds of pieces) are summed by the WHOLE
    // warp, one after the other; the group's own 8 lanes would chain pieces / 8 additions.
    const bool heavy = pieces > HEAVY;
    unsigned heavy_groups = __ballot_sync(full, heavy && sub == 0);
    while (heavy_groups) {
      const int src = __ffs(heavy_groups) - 1;   // lane 0 of the heavy group
      heavy_groups &= heavy_groups - 1;
      const uint32_t hp = __shfl_sync(full, pieces, src), ht0 = __shfl_sync(full, t0, src), hfs = __shfl_sync(full, first_slot, src);
      const uint32_t hb = __shfl_sync(full, b, src);
      xyzz_t h = C::identity();
      for (uint32_t k = lane; k < hp; k += 32) C::add(h, rec_pt[k ? 2 * (size_t)(ht0 + k) : hfs]);
#pragma unroll 1
      for (int d = 16; d >= 1; d >>= 1) {
        xyzz_t o;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          o.X.v[q] = __shfl_down_sync(full, h.X.v[q], d);
          o.Y.v[q] = __shfl_down_sync(full, h.Y.v[q], d);
          o.ZZ.v[q] = __shfl_down_sync(full, h.ZZ.v[q], d);
          o.ZZZ.v[q] = __shfl_down_sync(full, h.ZZZ.v[q], d);
        }
        if (lane < (unsigned)d) C::add(h, o);
      }
      if (lane == 0) buckets[hb] = h;
    }
    const uint32_t light = heavy ? 0u : pieces;
    xyzz_t v = C::identity();
    for (uint32_t k = sub; k < light; k += LANES) C::add(v, rec_pt[k ? 2 * (size_t)(t0 + k) : first_slot]);
#pragma unroll 1
    for (int d = LANES / 2; d >= 1; d >>= 1) {
      // all 32 lanes shuffle (four buckets per warp); a group whose bucket has no piece above lane d adds nothing
      xyzz_t o;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        o.X.v[q] = __shfl_down_sync(full, v.X.v[q], d, LANES);
        o.Y.v[q] = __shfl_down_sync(full, v.Y.v[q], d, LANES);
        o.ZZ.v[q] = __shfl_down_sync(full, v.ZZ.v[q], d, LANES);
        o.ZZZ.v[q] = __shfl_down_sync(full, v.ZZZ.v[q], d, LANES);
      }
      if (sub < (unsigned)d && (uint32_t)d < light) C::add(v, o);
    }
    if (sub == 0 && light) buckets[b] = v;
#else
    if (idx % LANES || !pieces) return;
    xyzz_t v = C::identity();
    for (uint32_t k = 0; k < pieces; k++) C::add(v, rec_pt[k ? 2 * (size_t)(t0 + k) : first_slot]);
    buckets[b] = v;
#endif
  }
};

// ---- stage 6: bucket reduction ---------------------------------------------------------------------
// in: [NB][in_stride] points, first cnt of each row used, element j has weight j+1.
// Thread (set, t) handles chunk [t*m, min((t+1)*m, cnt)):
//   S_t = sum of the chunk                          -> next level input (weight t, index t-1)
//   A_t = sum (j - t*m + 1) * in[j], then * 2^shift -> acat (plain weight-1 sum later)
template <class C>
struct ReduceLevelFn {
  const xyzz_t* in;
  size_t in_stride;
  uint32_t cnt, T, logm;
  xyzz_t* next;        // [NB][T]   (next[set*T + t-1] = S_t for t >= 1)
  xyzz_t* acat;        // [NB][acat_stride], this level at column acat_off
  size_t acat_stride, acat_off;
  uint32_t shift;      // doublings applied to A_t: logm * (level - 1)
  VDF_HD void operator()(size_t idx) const {
    uint32_t set = (uint32_t)(idx / T), t = (uint32_t)(idx - (size_t)set * T);
    uint32_t m = 1u << logm;
    uint32_t lo = t * m, hi = lo + m < cnt ? lo + m : cnt;
    const xyzz_t* row = in + (size_t)set * in_stride;
    xyzz_t run = C::identity(), acc = C::identity();
    for (uint32_t j = hi; j > lo; j--) {
      C::add(run, row[j - 1]);
      C::add(acc, run);
    }
    for (uint32_t k = 0; k < shift; k++) acc = C::dbl(acc);
    acat[(size_t)set * acat_stride + acat_off + t] = acc;
    if (t >= 1) next[(size_t)set * T + (t - 1)] = run;
  }
};

// Small bucket sets (latency path): sum_j j*B_j = 2^(c-1)*B_top + sum_b 2^b * S_b with S_b the sum of the
// buckets j < B whose index has bit b set.  All S_b are plain tree sums (log depth, no serial running sums).
// First level: thread (set, b, q) adds the two buckets number 2q and 2q+1 among those with bit b set.
template <class C>
struct BitPairFn {
  const xyzz_t* buckets;  // [sets][B], element k holds digit value k+1
  uint32_t B, nbits;      // nbits = log2(B): weights < B use bits 0..nbits-1
  xyzz_t* out;            // [sets * nbits][B/4]
  VDF_HD static uint32_t insert_bit(uint32_t t, uint32_t b) {   // value with bit b set, other bits from t
    return ((t >> b) << (b + 1)) | (1u << b) | (t & ((1u << b) - 1u));
  }
  VDF_HD void operator()(size_t idx) const {
    const uint32_t q4 = B / 4;
    uint32_t row = (uint32_t)(idx / q4), q = (uint32_t)(idx - (size_t)row * q4);
    uint32_t set = row / nbits, b = row - set * nbits;
    const xyzz_t* bk = buckets + (size_t)set * B;
    xyzz_t acc = bk[insert_bit(2 * q, b) - 1];
    C::add(acc, bk[insert_bit(2 * q + 1, b) - 1]);
    out[(size_t)row * q4 + q] = acc;
  }
};

// BitPairFn with one QUAD per output (quad.cuh): index = output * 4 + lane of the quad
template <class C>
struct BitPairQuadFn {
  const xyzz_t* buckets;
  uint32_t B, nbits;
  xyzz_t* out;
  VDF_HD void operator()(size_t idx) const {
#if defined(__CUDA_ARCH__)
    const size_t o = idx >> 2;
    const uint32_t q4 = B / 4;
    uint32_t row = (uint32_t)(o / q4), q = (uint32_t)(o - (size_t)row * q4);
    uint32_t set = row / nbits, b = row - set * nbits;
    const xyzz_t* bk = buckets + (size_t)set * B;
    xyzz_t acc = bk[BitPairFn<C>::insert_bit(2 * q, b) - 1];
    Quad<typename C::field>::add(acc, bk[BitPairFn<C>::insert_bit(2 * q + 1, b) - 1]);
    if ((idx & 3) == 0) out[o] = acc;
#else
    if ((idx & 3) == 0) BitPairFn<C>{buckets, B, nbits, out}(idx >> 2);
#endif
  }
};

// thread (set, b), b = 0..nbits: term_b = 2^b * S_b (b doublings), the last one 2^nbits * arr[B-1]; the terms are
// summed by the caller with a short tree.  (A Horner loop over the bits would serialise nbits doublings AND
// nbits additions in one thread; here the longest chain is nbits doublings.)
template <class C>
struct BitScaleFn {
  const xyzz_t* bitsum;   // [sets * nbits][stride], element 0 of each row is S_b
  size_t stride;
  const xyzz_t* arr;      // [sets][B]
  uint32_t B, nbits;
  xyzz_t* out;            // [sets][nbits + 1]
  VDF_HD void operator()(size_t idx) const {
    const uint32_t per = nbits + 1;
    uint32_t set = (uint32_t)(idx / per), b = (uint32_t)(idx - (size_t)set * per);
    xyzz_t acc = b < nbits ? bitsum[((size_t)set * nbits + b) * stride] : arr[(size_t)set * B + (B - 1)];
    for (uint32_t k = 0; k < b; k++) acc = C::dbl(acc);
    out[idx] = acc;
  }
};

// BitScaleFn with one QUAD per term: the chain of b doublings costs 3 multiplication latencies each instead of 9
template <class C>
struct BitScaleQuadFn {
  const xyzz_t* bitsum;
  size_t stride;
  const xyzz_t* arr;
  uint32_t B, nbits;
  xyzz_t* out;            // [sets][nbits + 1]
  VDF_HD void operator()(size_t idx) const {
#if defined(__CUDA_ARCH__)
    const size_t o = idx >> 2;
    const uint32_t per = nbits + 1;
    uint32_t set = (uint32_t)(o / per), b = (uint32_t)(o - (size_t)set * per);
    xyzz_t acc = b < nbits ? bitsum[((size_t)set * nbits + b) * stride] : arr[(size_t)set * B + (B - 1)];
#pragma unroll 1
    for (uint32_t k = 0; k < b; k++) acc = Quad<typename C::field>::dbl(acc);
    if ((idx & 3) == 0) out[o] = acc;
#else
    if ((idx & 3) == 0) BitScaleFn<C>{bitsum, stride, arr, B, nbits, out}(idx >> 2);
#endif
  }
};

// BitScaleFn and the sum of its <= 32 terms in ONE launch: one warp per set, lane b scales its term by b doublings,
// a shuffle tree adds the lanes.  Index = set * 32 + lane; launched with whole warps.
template <class C>
struct BitScaleSumFn {
  const xyzz_t* bitsum;   // [sets * nbits][stride], element 0 of each row is S_b
  size_t stride;
  const xyzz_t* arr;      // [sets][B]
  uint32_t B, nbits;      // nbits + 1 <= 32
  xyzz_t* out;            // [sets]
  VDF_HD void operator()(size_t idx) const {
    const uint32_t set = (uint32_t)(idx >> 5);
#if defined(__CUDA_ARCH__)
    const unsigned lane = (unsigned)idx & 31u, full = 0xffffffffu;
    xyzz_t v = C::identity();
    if (lane <= nbits) {
      v = lane < nbits ? bitsum[((size_t)set * nbits + lane) * stride] : arr[(size_t)set * B + (B - 1)];
      for (uint32_t k = 0; k < lane; k++) v = C::dbl(v);
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
      xyzz_t o;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        o.X.v[q] = __shfl_down_sync(full, v.X.v[q], d);
        o.Y.v[q] = __shfl_down_sync(full, v.Y.v[q], d);
        o.ZZ.v[q] = __shfl_down_sync(full, v.ZZ.v[q], d);
        o.ZZZ.v[q] = __shfl_down_sync(full, v.ZZZ.v[q], d);
      }
      if (lane < (unsigned)d && (uint32_t)d <= nbits) C::add(v, o);
    }
    if (lane == 0) out[set] = v;
#else
    if (idx & 31) return;
    xyzz_t acc = C::identity();
    for (uint32_t b = 0; b <= nbits; b++) {
      xyzz_t t = b < nbits ? bitsum[((size_t)set * nbits + b) * stride] : arr[(size_t)set * B + (B - 1)];
      for (uint32_t k = 0; k < b; k++) t = C::dbl(t);
      C::add(acc, t);
    }
    out[set] = acc;
#endif
  }
};

// segmented tree sum: out[set][t] = sum of in[set][t*K .. (t+1)*K)
template <class C>
struct SumFn {
  const xyzz_t* in;
  size_t in_stride;
  uint32_t cnt, T, K;
  xyzz_t* out;
  size_t out_stride;
  VDF_HD void operator()(size_t idx) const {
    uint32_t set = (uint32_t)(idx / T), t = (uint32_t)(idx - (size_t)set * T);
    uint32_t lo = t * K, hi = lo + K < cnt ? lo + K : cnt;
    const xyzz_t* row = in + (size_t)set * in_stride;
    xyzz_t acc = C::identity();
    for (uint32_t j = lo; j < hi; j++) C::add(acc, row[j]);
    out[(size_t)set * out_stride + t] = acc;
  }
};

// Latency version of SumFn: one WARP per 32 consecutive elements of a row, summed by a shuffle tree (5 steps of one
// addition) -- 32x fewer elements per level where the radix-4 SumFn gives 4x for 3 additions.  It spends 5
// lane-additions per element, so it is used only when the level is small (msm_tree_levels).  Index = row * Tw * 32
// + warp * 32 + lane with Tw = ceil(cnt / 32); launched with whole warps.
template <class C>
struct SumWarpFn {
  const xyzz_t* in;
  size_t in_stride;
  uint32_t cnt, Tw;
  xyzz_t* out;
  size_t out_stride;
  VDF_HD void operator()(size_t idx) const {
    const size_t w = idx >> 5;
    const uint32_t row = (uint32_t)(w / Tw), t = (uint32_t)(w - (size_t)row * Tw);
    const xyzz_t* src = in + (size_t)row * in_stride;
#if defined(__CUDA_ARCH__)
    const unsigned lane = threadIdx.x & 31u, full = 0xffffffffu;
    const uint32_t j = t * 32u + lane;
    xyzz_t v = C::identity();
    if (j < cnt) v = src[j];
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
      xyzz_t o;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        o.X.v[k] = __shfl_down_sync(full, v.X.v[k], d);
        o.Y.v[k] = __shfl_down_sync(full, v.Y.v[k], d);
        o.ZZ.v[k] = __shfl_down_sync(full, v.ZZ.v[k], d);
        o.ZZZ.v[k] = __shfl_down_sync(full, v.ZZZ.v[k], d);
      }
      if (lane < (unsigned)d) C::add(v, o);
    }
    if (lane == 0) out[(size_t)row * out_stride + t] = v;
#else
    if (idx & 31) return;
    xyzz_t acc = C::identity();
    const uint32_t lo = t * 32u, hi = lo + 32u < cnt ? lo + 32u : cnt;
    for (uint32_t j = lo; j < hi; j++) C::add(acc, src[j]);
    out[(size_t)row * out_stride + t] = acc;
#endif
  }
};

// SumWarpFn with quads (quad.cuh): one WARP per 16 consecutive elements of a row; quad k adds elements 2k and 2k+1, then
// three shuffle steps add the quads -- 4 quad additions (4 multiplication latencies each) for 16x fewer elements, where
// the lane-per-element tree spends 5 lane additions (14 latencies each) for 32x.  Index = row * Tw * 32 + warp * 32 +
// lane with Tw = ceil(cnt / 16); launched with whole warps.
template <class C>
struct SumQuadFn {
  const xyzz_t* in;
  size_t in_stride;
  uint32_t cnt, Tw;
  xyzz_t* out;
  size_t out_stride;
  VDF_HD void operator()(size_t idx) const {
    const size_t w = idx >> 5;
    const uint32_t row = (uint32_t)(w / Tw), t = (uint32_t)(w - (size_t)row * Tw);
    const xyzz_t* src = in + (size_t)row * in_stride;
#if defined(__CUDA_ARCH__)
    typedef Quad<typename C::field> Q;
    const unsigned lane = threadIdx.x & 31u, qd = lane >> 2;
    const uint32_t j = t * 16u + 2u * qd;
    xyzz_t v = C::identity(), o = C::identity();
    if (j < cnt) v = src[j];
    if (j + 1 < cnt) o = src[j + 1];
    Q::add(v, o);
#pragma unroll 1
    for (int d = 4; d >= 1; d >>= 1) {
      o = Q::down(v, 4 * d);
      if (qd < (unsigned)d) Q::add(v, o);
    }
    if (lane == 0) out[(size_t)row * out_stride + t] = v;
#else
    if (idx & 31) return;
    xyzz_t acc = C::identity();
    const uint32_t lo = t * 16u, hi = lo + 16u < cnt ? lo + 16u : cnt;
    for (uint32_t j = lo; j < hi; j++) C::add(acc, src[j]);
    out[(size_t)row * out_stride + t] = acc;
#endif
  }
};

// rows x cnt  ->  rows x 1 by tree levels, ping-ponging between the two buffers (each at least rows * ceil(cnt / 4)
// elements after the first level); returns where the result lives and its row stride
template <class L, class C>
const xyzz_t* msm_tree_levels(L& L_, uint32_t rows, const xyzz_t* cur, size_t& cur_stride, uint32_t cnt, xyzz_t* buf_a,
                              xyzz_t* buf_b) {
  xyzz_t* dst = (cur == buf_a) ? buf_b : buf_a;
  while (cnt > 1) {
    // small levels are pure latency: the warp tree; big ones are throughput: radix 4 per thread
#ifdef VDF_NO_QUAD
    const bool warp = cnt > 4 && (size_t)rows * cnt <= 32768;
#else
#ifndef VDF_SUMQUAD_MAX
#define VDF_SUMQUAD_MAX 32768
#endif
    const bool warp = (size_t)rows * cnt <= VDF_SUMQUAD_MAX;
#endif
    uint32_t T;
    if (warp) {
#ifdef VDF_NO_QUAD
      T = (cnt + 31) / 32;
      L_.template run<128>((size_t)rows * T * 32, SumWarpFn<C>{cur, cur_stride, cnt, T, dst, T});
#else
      T = (cnt + 15) / 16;
      L_.template run<128>((size_t)rows * T * 32, SumQuadFn<C>{cur, cur_stride, cnt, T, dst, T});
#endif
    } else {
      T = (cnt + 3) / 4;
      L_.template run<128>((size_t)rows * T, SumFn<C>{cur, cur_stride, cnt, T, 4u, dst, T});
    }
    cur = dst;
    cur_stride = T;
    cnt = T;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
  return cur;
}

// ---- stage 7: final (one thread) -----------------------------------------------------------------------
template <class C>
struct FinalFn {
  const xyzz_t* in;   // [batch * NB][in_stride], cnt used per row
  size_t in_stride;
  uint32_t cnt, NB, c;
  jac_t* out;         // [batch]
  uint32_t raw_jacobian;
  VDF_HD void operator()(size_t bt) const {
    xyzz_t total = C::identity();
    for (uint32_t s = NB; s > 0; s--) {
      if (s != NB)
        for (uint32_t k = 0; k < c; k++) total = C::dbl(total);
      const xyzz_t* row = in + ((size_t)bt * NB + (s - 1)) * in_stride;
      for (uint32_t j = 0; j < cnt; j++) C::add(total, row[j]);
    }
    out[bt] = raw_jacobian ? C::to_jac_raw(total) : C::to_jac_normalised(total);
  }
};

// out[set] = sum_b 2^b S_b + 2^nbits arr[B-1] from the finished bit rows
template <class L, class C>
void msm_bit_combine(L& L_, uint32_t NBT, const xyzz_t* bitsum, size_t stride, const xyzz_t* arr, uint32_t B,
                     uint32_t nbits, xyzz_t* out) {
  const uint32_t per = nbits + 1;
#ifdef VDF_NO_QUAD
  if (per <= 32) {   // one launch: scale and sum in a warp per set
    L_.template run<32>((size_t)NBT * 32, BitScaleSumFn<C>{bitsum, stride, arr, B, nbits, out});
    return;
  }
#else
  if (per <= 256) {   // a quad per term for the doublings, then quad tree sums of the terms
    xyzz_t* terms = L_.template alloc<xyzz_t>((size_t)NBT * per);
    const uint32_t T1 = (per + 15) / 16;
    xyzz_t* t1 = L_.template alloc<xyzz_t>((size_t)NBT * T1);
    L_.template run<32>((size_t)NBT * per * 4, BitScaleQuadFn<C>{bitsum, stride, arr, B, nbits, terms});
    if (T1 == 1) {
      L_.template run<32>((size_t)NBT * 32, SumQuadFn<C>{terms, per, per, 1u, out, 1});
    } else {
      L_.template run<32>((size_t)NBT * T1 * 32, SumQuadFn<C>{terms, per, per, T1, t1, T1});
      L_.template run<32>((size_t)NBT * 32, SumQuadFn<C>{t1, T1, T1, 1u, out, 1});
    }
    L_.free(terms); L_.free(t1);
    return;
  }
#endif
  xyzz_t* terms = L_.template alloc<xyzz_t>((size_t)NBT * per);
  xyzz_t* t1 = L_.template alloc<xyzz_t>((size_t)NBT * ((per + 3) / 4));
  L_.template run<32>((size_t)NBT * per, BitScaleFn<C>{bitsum, stride, arr, B, nbits, terms});
  if (per > 4 && per <= 32) {
    // one warp per set sums its <= 32 terms in 5 shuffle steps
    L_.template run<32>((size_t)NBT * 32, SumWarpFn<C>{terms, per, per, 1u, out, 1});
  } else {
    const xyzz_t* cur = terms;
    size_t cur_stride = per;
    uint32_t cnt = per;
    xyzz_t* dst = t1;
    xyzz_t* spare = terms;
    while (cnt > 4) {
      uint32_t T = (cnt + 3) / 4;
      L_.template run<32>((size_t)NBT * T, SumFn<C>{cur, cur_stride, cnt, T, 4u, dst, T});
      cur = dst;
      cur_stride = T;
      cnt = T;
      xyzz_t* nx = (dst == t1) ? spare : t1;
      dst = nx;
    }
    L_.template run<32>(NBT, SumFn<C>{cur, cur_stride, cnt, 1u, cnt, out, 1});
  }
  L_.free(terms); L_.free(t1);
}

// per_set[s] = sum_{k=0}^{B-1} (k+1) * arr[s][k] for B a power of two, by bit decomposition (log depth):
//   = 2^(log2 B) * arr[B-1] + sum_b 2^b * S_b,  S_b = sum of the elements whose weight < B has bit b set.
template <class L, class C>
void msm_bit_weighted_sum(L& L_, uint32_t NBT, const xyzz_t* arr, uint32_t B, xyzz_t* per_set) {
  uint32_t nbits = 0;
  while ((1u << nbits) < B) nbits++;          // weights 1..B-1 use bits 0..nbits-1; weight B = 2^nbits
  const uint32_t rows = NBT * nbits;
  uint32_t cnt = B / 4;
  xyzz_t* bs_a = L_.template alloc<xyzz_t>((size_t)rows * cnt);
  xyzz_t* bs_b = L_.template alloc<xyzz_t>((size_t)rows * ((cnt + 3) / 4));
#ifdef VDF_NO_QUAD
  L_.template run<128>((size_t)rows * cnt, BitPairFn<C>{arr, B, nbits, bs_a});
#else
  if ((size_t)rows * cnt <= 65536) L_.template run<128>((size_t)rows * cnt * 4, BitPairQuadFn<C>{arr, B, nbits, bs_a});
  else L_.template run<128>((size_t)rows * cnt, BitPairFn<C>{arr, B, nbits, bs_a});
#endif
  size_t cur_stride = cnt;
  const xyzz_t* cur = msm_tree_levels<L, C>(L_, rows, bs_a, cur_stride, cnt, bs_a, bs_b);
  msm_bit_combine<L, C>(L_, NBT, cur, cur_stride, arr, B, nbits, per_set);
  L_.free(bs_a); L_.free(bs_b);
}

// per_set[s] = (four partial sums of A_t) + 2^logm * wsum[s]
template <class C>
struct CombineFn {
  const xyzz_t* sumA;   // [sets * 4][stride], element 0 of each row
  size_t stride;
  const xyzz_t* wsum;   // [sets]
  uint32_t logm;
  xyzz_t* out;
  VDF_HD void operator()(size_t set) const {
    xyzz_t acc = wsum[set];
    for (uint32_t k = 0; k < logm; k++) acc = C::dbl(acc);
    for (uint32_t k = 0; k < 4; k++) C::add(acc, sumA[(set * 4 + k) * stride]);
    out[set] = acc;
  }
};

// fallback of stage 6/7 for odd sizes: chunked running sums, recursively, then a tree sum and the final thread
template <class L, class C>
void msm_reduce_tree(L& L_, const MsmPlan& p, uint32_t NBT, const xyzz_t* buckets, jac_t* out) {
  const uint32_t m = 1u << p.logm;
  // level sizes
  uint32_t cnts[32], Ts[32];
  int levels = 0;
  size_t acat_len = 0;
  for (uint32_t cnt = p.B;;) {
    uint32_t T = (cnt + m - 1) / m;
    cnts[levels] = cnt; Ts[levels] = T;
    levels++;
    acat_len += T;
    if (T <= 1) break;
    cnt = T - 1;
  }
  xyzz_t* acat = L_.template alloc<xyzz_t>((size_t)NBT * acat_len);
  xyzz_t* lvl_a = L_.template alloc<xyzz_t>((size_t)NBT * Ts[0]);
  xyzz_t* lvl_b = L_.template alloc<xyzz_t>((size_t)NBT * (levels > 1 ? Ts[1] : 1));
  {
    const xyzz_t* in = buckets;
    size_t in_stride = p.B;
    size_t off = 0;
    xyzz_t* nxt = lvl_a;
    for (int l = 0; l < levels; l++) {
      L_.template run<128>((size_t)NBT * Ts[l],
                           ReduceLevelFn<C>{in, in_stride, cnts[l], Ts[l], p.logm, nxt, acat, acat_len, off,
                                            p.logm * (uint32_t)l});
      off += Ts[l];
      in = nxt;
      in_stride = Ts[l];
      nxt = (nxt == lvl_a) ? lvl_b : lvl_a;
    }
  }
  // tree sum of acat rows
  const uint32_t K = 16;
  xyzz_t* sum_a = L_.template alloc<xyzz_t>((size_t)NBT * ((acat_len + K - 1) / K));
  xyzz_t* sum_b = L_.template alloc<xyzz_t>((size_t)NBT * ((acat_len + (size_t)K * K - 1) / ((size_t)K * K)));
  const xyzz_t* cur = acat;
  size_t cur_stride = acat_len;
  uint32_t cur_cnt = (uint32_t)acat_len;
  xyzz_t* dst = sum_a;
  while (cur_cnt > 1) {
    uint32_t T = (cur_cnt + K - 1) / K;
    L_.template run<128>((size_t)NBT * T, SumFn<C>{cur, cur_stride, cur_cnt, T, K, dst, T});
    cur = dst;
    cur_stride = T;
    cur_cnt = T;
    dst = (dst == sum_a) ? sum_b : sum_a;
  }
  L_.mark(MSM_STAGE_FINAL);
  L_.template run<32>(p.batch, FinalFn<C>{cur, cur_stride, cur_cnt, p.NB, p.c, out, p.raw_jacobian});
  L_.mark(MSM_STAGE_END);

  L_.free(acat); L_.free(lvl_a); L_.free(lvl_b); L_.free(sum_a); L_.free(sum_b);
}

// ---- driver ------------------------------------------------------------------------------------------
// C = curve (coordinate field), SF = its scalar field.  All pointers are in the policy's memory space.
// Stages 1-5 for ONE chunk of the points: p.n = chunk length, pts / scalars point at the chunk.  The
// chunk's bucket sums are STORED into `buckets` ([batch*NB*B], zeroed by the caller).  Chunking lets the host
// API overlap the H2D copy of the next chunk of scalars with the accumulation of the current one while paying
// the bucket reduction only once (msm_merge_buckets folds the per-chunk arrays together).
template <class L, class C, class SF>
void msm_accumulate(L& L_, const MsmPlan& p, const affine_t* pts, const ScalarSet& scalars, xyzz_t* buckets) {
  const size_t n = p.n, E = n * p.W * p.batch;
  const uint32_t NBT = p.NB * p.batch;       // bucket sets in flight
  const uint32_t NBK = NBT * p.B;
  if (n == 0) return;
  uint32_t* keys = L_.template alloc<uint32_t>(E);
  uint32_t* rank = L_.template alloc<uint32_t>(E);
  uint32_t* count = L_.template alloc<uint32_t>(NBK);
  uint32_t* offs = L_.template alloc<uint32_t>((size_t)NBK + 1);
  uint32_t* sref = L_.template alloc<uint32_t>(E);
  L_.zero(count, (size_t)NBK * 4);

  L_.mark(MSM_STAGE_DIGITS);
  L_.template run<256>(n * p.batch, DigitsFn<SF>{scalars, keys, rank, count, p});
  L_.mark(MSM_STAGE_SCAN);
  L_.exclusive_scan(count, offs, NBK);
  L_.mark(MSM_STAGE_SCATTER);
  L_.template run<256>(n * p.batch, ScatterFn{keys, rank, offs, sref, p});

  // optional batched-affine halving rounds (msm_affine.cuh): the list shrinks 2^rounds times
  size_t e_cap = E;
  fe *alist_x = nullptr, *alist_y = nullptr;
  uint32_t* alist_offs = nullptr;
  uint32_t S = p.S;
  L_.mark(MSM_STAGE_ACCUMULATE);
  if (p.affine_rounds) {
    msm_affine_rounds<L, typename C::field>(L_, p.affine_rounds, p.affine_K, NBK, sref, pts, offs, e_cap, &alist_x,
                                            &alist_y, &alist_offs);
    size_t s2 = e_cap / (148 * 768);
    if (s2 < 32) s2 = 32;
    if (s2 < S) S = (uint32_t)s2;
  }
  S = fit_waves(e_cap, S, (size_t)148 * VDF_ACC_MINB * 128);   // a few waves of ranges: no partly filled last wave

  // accumulate over fixed-size ranges of the sorted list
  size_t T_acc = (e_cap + S - 1) / S;
  size_t n_rec = 2 * T_acc;
  RecHdr* hdr_a = L_.template alloc<RecHdr>(n_rec);
  xyzz_t* pt_a = L_.template alloc<xyzz_t>(n_rec);
  if (p.affine_rounds)
    L_.template run<128, VDF_ACC_MINB>(T_acc, AccumulateFn<C, true>{alist_offs, NBK, nullptr, nullptr, alist_x, alist_y, buckets,
                                                         hdr_a, pt_a, S});
  else
    L_.template run<128, VDF_ACC_MINB>(T_acc, AccumulateFn<C>{offs, NBK, sref, pts, nullptr, nullptr, buckets, hdr_a, pt_a, S});
  // segmented reduction of the records: log-depth levels, then owners
  RecHdr* hdr_b = nullptr;
  xyzz_t* pt_b = nullptr;
  L_.mark(MSM_STAGE_RECORDS);
  if (p.rec_warp && p.rec_bucket && NBK <= MSM_REC_BUCKET_MAX) {
    // latency regime: one warp per bucket gathers its pieces straight from the record array
    const size_t rb_threads = ((size_t)NBK * RecBucketFn<C>::LANES + 31) / 32 * 32;   // whole warps: every lane shuffles
    L_.template run<128>(rb_threads, RecBucketFn<C>{p.affine_rounds ? alist_offs : offs, NBK, S, pt_a, buckets});
    L_.free(alist_x); L_.free(alist_offs);
    L_.free(keys); L_.free(rank); L_.free(count); L_.free(offs); L_.free(sref);
    L_.free(hdr_a); L_.free(pt_a);
    return;
  }
  L_.free(alist_x); L_.free(alist_offs);
  // Each level maps G records to <= 2 (needs G > 2 to shrink).  Large problems: G = p.G (work-efficient),
  // stop at 4096 records.  Small problems are latency-bound (every level is a serial chain of <= G point
  // additions, and the owner pass is serial in the number of pieces of the heaviest bucket): G = 8, run
  // the levels down to 256 records.
  if (p.rec_warp) {
    const bool long_runs = n_rec > 2 * (size_t)NBK;
#ifdef VDF_NO_QUAD
    // (the lane-per-addition levels of round 2's second pass, kept for A/B builds)
    // one warp per 32 records (RecWarpLevelFn): 16x fewer records per level with a chain of 5 additions; down to 64
    // records so that the serial owner pass stays short even when one bucket owns every remaining piece
    // The warp level spends 5 lane-additions per record where the serial one spends 1: with many records in long
    // runs (several pieces per bucket) the first levels are throughput-bound, so they stay serial (G = 8).
    while (n_rec > 64) {
      const bool serial = long_runs && n_rec > 65536;
      const size_t per = serial ? 8 : 32;
      size_t groups = (n_rec + per - 1) / per;
      if (!hdr_b) {
        hdr_b = L_.template alloc<RecHdr>(2 * groups);
        pt_b = L_.template alloc<xyzz_t>(2 * groups);
      }
      if (serial) L_.template run<128>(groups, RecLevelFn<C>{hdr_a, pt_a, n_rec, buckets, hdr_b, pt_b, 8u});
      else L_.template run<128>(groups * 32, RecWarpLevelFn<C>{hdr_a, pt_a, n_rec, buckets, hdr_b, pt_b});
      RecHdr* th = hdr_a; hdr_a = hdr_b; hdr_b = th;
      xyzz_t* tp = pt_a; pt_a = pt_b; pt_b = tp;
      n_rec = 2 * groups;
    }
#else
    // serial levels (one lane per 8 records: work-efficient) while the list is long, then RecRunsFn down to one node
    while (n_rec > 2) {
#ifndef VDF_REC_SERIAL_ABOVE
#define VDF_REC_SERIAL_ABOVE (1u << 19)   // measured: Nova fold step 0.604 (2^17) -> 0.573 ms (2^19), profiles/r2_experiments.md
#endif
      const bool serial = long_runs && n_rec > (size_t)VDF_REC_SERIAL_ABOVE;
      const size_t per = serial ? 8 : (size_t)RecRunsFn<C>::NS;
      size_t groups = (n_rec + per - 1) / per;
      if (!hdr_b) {
        hdr_b = L_.template alloc<RecHdr>(2 * groups);
        pt_b = L_.template alloc<xyzz_t>(2 * groups);
      }
      if (serial) L_.template run<128>(groups, RecLevelFn<C>{hdr_a, pt_a, n_rec, buckets, hdr_b, pt_b, 8u});
      else L_.template run<256, 2>(groups * 256, RecRunsFn<C>{hdr_a, pt_a, n_rec, buckets, hdr_b, pt_b});
      RecHdr* th = hdr_a; hdr_a = hdr_b; hdr_b = th;
      xyzz_t* tp = pt_a; pt_a = pt_b; pt_b = tp;
      n_rec = 2 * groups;
    }
#endif
  } else {
  const bool small = n_rec <= (1u << 17);
  const uint32_t G = small ? 8u : (p.G < 4 ? 4u : p.G);
  const size_t rec_stop = small ? 256 : 4096;
  while (n_rec > rec_stop) {
    size_t groups = (n_rec + G - 1) / G;
    if (!hdr_b) {
      hdr_b = L_.template alloc<RecHdr>(2 * groups);
      pt_b = L_.template alloc<xyzz_t>(2 * groups);
    }
    L_.template run<128>(groups, RecLevelFn<C>{hdr_a, pt_a, n_rec, buckets, hdr_b, pt_b, G});
    RecHdr* th = hdr_a; hdr_a = hdr_b; hdr_b = th;
    xyzz_t* tp = pt_a; pt_a = pt_b; pt_b = tp;
    n_rec = 2 * groups;
  }
  }
  L_.template run<128>(n_rec, RecOwnerFn<C>{hdr_a, pt_a, n_rec, buckets});

  L_.free(keys); L_.free(rank); L_.free(count); L_.free(offs); L_.free(sref);
  L_.free(hdr_a); L_.free(pt_a); L_.free(hdr_b); L_.free(pt_b);
}

template <class L, class C>
void msm_merge_buckets(L& L_, const MsmPlan& p, xyzz_t* total, const xyzz_t* chunk) {
  L_.template run<128>((size_t)p.NB * p.batch * p.B, BucketMergeFn<C>{total, chunk});
}

// Stages 6-7: bucket sets -> p.batch normalised points
template <class L, class C>
void msm_finish(L& L_, const MsmPlan& p, const xyzz_t* buckets, jac_t* out) {
  const uint32_t NBT = p.NB * p.batch;
  L_.mark(MSM_STAGE_REDUCE);
  const uint32_t m = 1u << p.logm, T0 = p.B / m;
  if (p.B >= 8 && p.B <= 32768) {
    // latency path: bit-decomposition sums of the buckets themselves
    xyzz_t* per_set = L_.template alloc<xyzz_t>(NBT);
    msm_bit_weighted_sum<L, C>(L_, NBT, buckets, p.B, per_set);
    L_.mark(MSM_STAGE_FINAL);
    L_.template run<32>(p.batch, FinalFn<C>{per_set, 1, 1u, p.NB, p.c, out, p.raw_jacobian});
    L_.mark(MSM_STAGE_END);
    L_.free(per_set);
  } else if (p.B % m == 0 && (T0 & (T0 - 1)) == 0 && T0 >= 8 && T0 <= 65536) {
    // throughput path: ONE level of chunked running sums where parallelism is plentiful (B/m threads),
    // then the bit-decomposition on the B/m chunk totals and a plain sum of the chunk-local weighted sums:
    //   sum_j j*B_j = sum_t A_t + m * sum_t t*S_t
    // Both tree sums share their launches: the row buffer holds nbits bit rows per set (pairs already added,
    // length T0/4) followed by the A values viewed as four rows of length T0/4 per set.
    uint32_t nbits = 0;
    while ((1u << nbits) < T0) nbits++;
    const uint32_t q4 = T0 / 4, bit_rows = NBT * nbits, rows = bit_rows + NBT * 4;
    xyzz_t* S_arr = L_.template alloc<xyzz_t>((size_t)NBT * T0);
    xyzz_t* ra = L_.template alloc<xyzz_t>((size_t)rows * q4);
    xyzz_t* rb = L_.template alloc<xyzz_t>((size_t)rows * ((q4 + 3) / 4));
    xyzz_t* A_arr = ra + (size_t)bit_rows * q4;           // [NBT][T0] == [NBT*4][T0/4]
    L_.zero(S_arr, (size_t)NBT * T0 * sizeof(xyzz_t));   // element T0-1 (weight T0) stays the identity
    L_.template run<128>((size_t)NBT * T0, ReduceLevelFn<C>{buckets, p.B, p.B, T0, p.logm, S_arr, A_arr, T0, 0, 0u});
#ifdef VDF_NO_QUAD
    L_.template run<128>((size_t)bit_rows * q4, BitPairFn<C>{S_arr, T0, nbits, ra});
#else
    if ((size_t)bit_rows * q4 <= 65536) L_.template run<128>((size_t)bit_rows * q4 * 4, BitPairQuadFn<C>{S_arr, T0, nbits, ra});
    else L_.template run<128>((size_t)bit_rows * q4, BitPairFn<C>{S_arr, T0, nbits, ra});
#endif
    size_t cur_stride = q4;
    const xyzz_t* cur = msm_tree_levels<L, C>(L_, rows, ra, cur_stride, q4, ra, rb);
    xyzz_t* wsum = L_.template alloc<xyzz_t>(NBT);
    xyzz_t* per_set = L_.template alloc<xyzz_t>(NBT);
    msm_bit_combine<L, C>(L_, NBT, cur, cur_stride, S_arr, T0, nbits, wsum);
    L_.template run<32>(NBT, CombineFn<C>{cur + (size_t)bit_rows * cur_stride, cur_stride, wsum, p.logm, per_set});
    L_.mark(MSM_STAGE_FINAL);
    L_.template run<32>(p.batch, FinalFn<C>{per_set, 1, 1u, p.NB, p.c, out, p.raw_jacobian});
    L_.mark(MSM_STAGE_END);
    L_.free(S_arr); L_.free(ra); L_.free(rb); L_.free(wsum); L_.free(per_set);
  } else {
    msm_reduce_tree<L, C>(L_, p, NBT, buckets, out);
  }
}

// whole MSM in one chunk
template <class L, class C, class SF>
void msm_run(L& L_, const MsmPlan& p, const affine_t* pts, const ScalarSet& scalars, jac_t* out) {
  if (p.n == 0) {
    L_.zero(out, sizeof(jac_t) * p.batch);
    return;
  }
  const size_t NBK = (size_t)p.NB * p.batch * p.B;
  xyzz_t* buckets = L_.template alloc<xyzz_t>(NBK);
  L_.zero(buckets, NBK * sizeof(xyzz_t));
  msm_accumulate<L, C, SF>(L_, p, pts, scalars, buckets);
  msm_finish<L, C>(L_, p, buckets, out);
  L_.free(buckets);
}

// ---- generator-set construction ----------------------------------------------------------------------
// pasta_curves affine (72-byte stride: x, y, infinity:u8 + padding) -> packed 64-byte affine
struct RepackFn {
  const uint8_t* src;  // 72-byte stride
  affine_t* dst;
  VDF_HD void operator()(size_t i) const {
    const uint8_t* s = src + i * 72;
    affine_t a;
    // 72*i is 8-byte aligned only: read as 64-bit words
    const uint64_t* w = reinterpret_cast<const uint64_t*>(s);
    bool inf = s[64] != 0;
    for (int k = 0; k < 4; k++) {
      uint64_t xv = inf ? 0 : w[k], yv = inf ? 0 : w[4 + k];
      a.x.v[2 * k] = (uint32_t)xv; a.x.v[2 * k + 1] = (uint32_t)(xv >> 32);
      a.y.v[2 * k] = (uint32_t)yv; a.y.v[2 * k + 1] = (uint32_t)(yv >> 32);
    }
    fe_store(&dst[i].x, a.x);
    fe_store(&dst[i].y, a.y);
  }
};

struct UnpackFn {  // packed 64-byte affine -> 72-byte pasta_curves affine
  const affine_t* src;
  uint8_t* dst;
  VDF_HD void operator()(size_t i) const {
    uint8_t* d = dst + i * 72;
    uint64_t* w = reinterpret_cast<uint64_t*>(d);
    fe x = fe_load(&src[i].x), y = fe_load(&src[i].y);
    uint32_t o = 0;
    for (int k = 0; k < 8; k++) o |= x.v[k] | y.v[k];
    for (int k = 0; k < 4; k++) {
      w[k] = (uint64_t)x.v[2 * k] | ((uint64_t)x.v[2 * k + 1] << 32);
      w[4 + k] = (uint64_t)y.v[2 * k] | ((uint64_t)y.v[2 * k + 1] << 32);
    }
    w[8] = (o == 0) ? 1ull : 0ull;  // infinity flag byte + zero padding
  }
};

// Table levels: level[l][i] = 2^c * level[l-1][i], produced in XYZZ and normalised with a batched
// inversion (Montgomery's trick) over chunks of CH points per thread.
template <class C, class F>
struct TableLevelFn {
  const affine_t* prev;
  affine_t* next;
  size_t n;
  uint32_t c;
  static constexpr int CH = 8;
  VDF_HD void operator()(size_t t) const {
    size_t lo = t * CH, hi = lo + CH < n ? lo + CH : n;
    xyzz_t q[CH];
    fe pref[CH];
    fe run = F::one();
    for (size_t i = lo; i < hi; i++) {
      affine_t a;
      a.x = fe_load(&prev[i].x);
      a.y = fe_load(&prev[i].y);
      xyzz_t r = C::from_affine(a);
      for (uint32_t k = 0; k < c; k++) r = C::dbl(r);
      q[i - lo] = r;
      pref[i - lo] = run;
      if (!C::is_inf(r)) run = F::mul(run, F::mul(r.ZZ, r.ZZZ));
    }
    fe inv = F::inv(run);
    for (size_t i = hi; i > lo; i--) {
      const xyzz_t& r = q[i - 1 - lo];
      affine_t a;
      if (C::is_inf(r)) {
        a.x = F::zero(); a.y = F::zero();
      } else {
        fe zi = F::mul(inv, pref[i - 1 - lo]);          // 1 / (ZZ * ZZZ)
        inv = F::mul(inv, F::mul(r.ZZ, r.ZZZ));
        a.x = F::mul(r.X, F::mul(zi, r.ZZZ));
        a.y = F::mul(r.Y, F::mul(zi, r.ZZ));
      }
      fe_store(&next[i - 1].x, a.x);
      fe_store(&next[i - 1].y, a.y);
    }
  }
};

// Synthetic generator sets (bench / tests): P_i = (k0 + i*d) * G with G = (-1, 2), known discrete logs
// so an MSM of any size can be checked in O(n) scalar-field operations.
// Setup (one thread): K0 = k0*G, D = d*G (affine), Q = CH*D.  Thread t then emits points
// [t*CH, (t+1)*CH) starting from K0 + t*Q and stepping by D; chunks are normalised to affine with one
// batched inversion.
constexpr int PROG_CH = 16;

struct ProgSetup {
  xyzz_t K0, Q;
  fe dx, dy;   // affine D; (0,0) if d == 0
};

template <class C, class F>
struct ProgressionSetupFn {
  fe k0, d;     // canonical (non-Montgomery) 256-bit scalars
  ProgSetup* out;
  VDF_HD static xyzz_t mul_gen(const fe& k) {
    fe gx = F::neg(F::one()), gy = F::dbl(F::one());
    xyzz_t r = C::identity();
    for (int limb = 7; limb >= 0; limb--)
      for (int bit = 31; bit >= 0; bit--) {
        r = C::dbl(r);
        if ((k.v[limb] >> bit) & 1u) C::madd(r, gx, gy);
      }
    return r;
  }
  VDF_HD void operator()(size_t) const {
    ProgSetup s;
    s.K0 = mul_gen(k0);
    xyzz_t D = mul_gen(d);
    jac_t dj = C::to_jac_normalised(D);
    s.dx = dj.X;
    s.dy = dj.Y;
    if (C::is_inf(D)) { s.dx = F::zero(); s.dy = F::zero(); }
    xyzz_t q = D;
    for (int k = 1; k < PROG_CH; k <<= 1) q = C::dbl(q);
    s.Q = q;
    *out = s;
  }
};

template <class C, class F>
struct ProgressionFn {
  const ProgSetup* setup;
  affine_t* out;
  size_t n;
  static constexpr int CH = PROG_CH;
  VDF_HD void operator()(size_t t) const {
    size_t lo = t * CH, hi = lo + CH < n ? lo + CH : n;
    xyzz_t base = C::mul_u32(setup->Q, (uint32_t)t);
    C::add(base, setup->K0);
    affine_t D;
    D.x = setup->dx;
    D.y = setup->dy;
    xyzz_t q[CH];
    fe pref[CH];
    fe run = F::one();
    for (size_t i = lo; i < hi; i++) {
      q[i - lo] = base;
      pref[i - lo] = run;
      if (!C::is_inf(base)) run = F::mul(run, F::mul(base.ZZ, base.ZZZ));
      C::madd_signed(base, D, false);
    }
    fe inv = F::inv(run);
    for (size_t i = hi; i > lo; i--) {
      const xyzz_t& r = q[i - 1 - lo];
      affine_t a;
      if (C::is_inf(r)) {
        a.x = F::zero(); a.y = F::zero();
      } else {
        fe zi = F::mul(inv, pref[i - 1 - lo]);
        inv = F::mul(inv, F::mul(r.ZZ, r.ZZZ));
        a.x = F::mul(r.X, F::mul(zi, r.ZZZ));
        a.y = F::mul(r.Y, F::mul(zi, r.ZZ));
      }
      fe_store(&out[i - 1].x, a.x);
      fe_store(&out[i - 1].y, a.y);
    }
  }
};

// Sum of k Jacobian points in any representation (multi-GPU partial results arrive un-normalised): one warp,
// lane j folds points j, j+32, ..., a 5-step shuffle tree adds the lanes, lane 0 normalises ONCE.  Launched with 32
// threads.  (With k = 8 ranks: 3 dependent additions + one inversion instead of 8 additions + one inversion after
// every rank has already paid an inversion of its own.)
template <class C>
struct JacSumFn {
  const jac_t* in;
  uint32_t k;
  jac_t* out;
  VDF_HD void operator()(size_t idx) const {
#if defined(__CUDA_ARCH__)
    const unsigned lane = (unsigned)idx & 31u, full = 0xffffffffu;
    xyzz_t v = C::identity();
    for (uint32_t j = lane; j < k; j += 32) C::add(v, C::from_jac(in[j]));
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
      if (__all_sync(full, (uint32_t)d >= k)) continue;   // nothing above lane d
      xyzz_t o;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        o.X.v[q] = __shfl_down_sync(full, v.X.v[q], d);
        o.Y.v[q] = __shfl_down_sync(full, v.Y.v[q], d);
        o.ZZ.v[q] = __shfl_down_sync(full, v.ZZ.v[q], d);
        o.ZZZ.v[q] = __shfl_down_sync(full, v.ZZZ.v[q], d);
      }
      if (lane < (unsigned)d) C::add(v, o);
    }
    if (lane == 0) *out = C::to_jac_normalised(v);
#else
    if (idx) return;
    xyzz_t acc = C::identity();
    for (uint32_t j = 0; j < k; j++) C::add(acc, C::from_jac(in[j]));
    *out = C::to_jac_normalised(acc);
#endif
  }
};

}  // namespace vdf
