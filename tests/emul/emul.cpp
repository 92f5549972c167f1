// tests/emul: CPU execution of the library's kernel functors through vdf::HostLaunch.
//
// TEST TOOL ONLY.  libvdfgpu.so never contains or calls this code; the product has no CPU path.  The
// point is to let `pytest -m "not gpu"` check the pipeline *logic* that surrounds the PTX field
// arithmetic -- signed-digit recoding, the counting sort, range-based bucket accumulation with its record
// fix-up levels, the bucket-reduction tree, COO->CSR, the cross-term and fold formulas -- against the
// Python oracle in a container that has no GPU.  Field multiplication here is the portable 64-bit C path of
// field.cuh; the PTX path is checked on the GPU (tests/test_gpu_field.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../vdf_b200/csrc/minroot.cuh"
#include "../../vdf_b200/csrc/msm.cuh"
#include "../../vdf_b200/csrc/r1cs.cuh"
#include "../../vdf_b200/csrc/r1cs_host.hpp"

using namespace vdf;

namespace {

template <class F>
void field_op(int op, const fe* a, const fe* b, fe* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case 0: out[i] = F::mul(a[i], b[i]); break;
      case 1: out[i] = F::add(a[i], b[i]); break;
      case 2: out[i] = F::sub(a[i], b[i]); break;
      case 3: out[i] = F::sqr(a[i]); break;
      case 4: out[i] = F::inv(a[i]); break;
      case 5: out[i] = F::from_mont(a[i]); break;
      case 6: out[i] = F::neg(a[i]); break;
      case 7: out[i] = F::one(); break;
    }
  }
}

uint32_t g_affine_rounds = 0, g_affine_K = 8, g_rec_warp = 0;   // set by emul_set_affine

// HostLaunch whose temporaries come from ONE block sized beforehand by the sizing pass (PlanLaunch), exactly as the
// product's CudaLaunch does with its per-stream arena: a disagreement between the sizing pass and the run throws.
struct HostArenaLaunch : HostLaunch {
  Arena arena;
  std::vector<uint8_t> block;
  explicit HostArenaLaunch(size_t bytes) : block(bytes + 64) {
    arena.base = block.data() + (64 - reinterpret_cast<uintptr_t>(block.data()) % 64) % 64;
    arena.reset(bytes);
  }
  template <class T>
  T* alloc(size_t count) { return reinterpret_cast<T*>(arena.alloc(count * sizeof(T))); }
  void free(void* p) { arena.free(p); }
  void exclusive_scan(const uint32_t* in, uint32_t* out, size_t n) {
    size_t tiles = (n + 2047) / 2048;
    if (tiles == 0) tiles = 1;
    uint32_t* t = alloc<uint32_t>(tiles);   // mirrors CudaLaunch::exclusive_scan's temporary
    HostLaunch::exclusive_scan(in, out, n);
    free(t);
  }
};

// whole MSM through the arena: returns 0, or -7 when the sizing pass and the run disagree
template <class C, class SF>
int run_msm_arena(const MsmPlan& p, const affine_t* pts, const ScalarSet& ss, jac_t* out) {
  PlanLaunch PL;
  ScalarSet none{{nullptr, nullptr, nullptr, nullptr}};
  msm_run<PlanLaunch, C, SF>(PL, p, nullptr, none, nullptr);
  HostArenaLaunch L(PL.arena.high);
  try {
    msm_run<HostArenaLaunch, C, SF>(L, p, pts, ss, out);
  } catch (const std::exception&) {
    return -7;
  }
  if (L.arena.high != PL.arena.high || L.launches != PL.launches) return -7;
  return 0;
}

MsmPlan plan_for(size_t n, size_t stride, int table, uint32_t c, uint32_t S, uint32_t G, uint32_t logm, int is_mont) {
  MsmPlan p;
  p.n = (uint32_t)n;
  p.table = table ? 1u : 0u;
  p.c = c;
  p.W = msm_windows(c);
  p.B = 1u << (c - 1);
  p.NB = table ? 1u : p.W;
  p.level_stride = stride;
  p.S = S;
  p.G = G;
  p.logm = logm;
  p.is_mont = is_mont ? 1u : 0u;
  p.batch = 1;
  p.len[0] = (uint32_t)n;
  p.raw_jacobian = 0;
  p.affine_rounds = g_affine_rounds;
  p.affine_K = g_affine_K;
  p.rec_warp = g_rec_warp ? 1u : 0u;
  p.rec_bucket = g_rec_warp == 2 ? 1u : 0u;   // 2: per-bucket record reduction (RecBucketFn)
  return p;
}

}  // namespace

extern "C" {

// record levels in groups of 32 (the serial mirror of RecWarpLevelFn) for every following emul_msm* call
int emul_set_recwarp(uint32_t on) {
  g_rec_warp = on;
  return 0;
}

// batched-affine halving rounds (msm_affine.cuh) for every following emul_msm* call; 0 = off
int emul_set_affine(uint32_t rounds, uint32_t K) {
  g_affine_rounds = rounds;
  g_affine_K = K ? K : 8;
  return 0;
}

int emul_field_op(int field, int op, const void* a, const void* b, void* out, size_t n) {
  if (field == 0) field_op<Fp>(op, (const fe*)a, (const fe*)b, (fe*)out, n);
  else field_op<Fq>(op, (const fe*)a, (const fe*)b, (fe*)out, n);
  return 0;
}

// full MSM pipeline; points in pasta_curves 72-byte layout; table != 0 builds the 2^(c w) levels first
int emul_msm(int curve, int table, uint32_t c, uint32_t S, uint32_t G, uint32_t logm, const void* affine72, size_t n,
             const void* scalars, int is_mont, void* out96) {
  HostLaunch L;
  uint32_t W = table ? msm_windows(c) : 1;
  std::vector<affine_t> pts((size_t)W * (n ? n : 1));
  L.run(n, RepackFn{(const uint8_t*)affine72, pts.data()});
  if (table)
    for (uint32_t l = 1; l < W; l++) {
      size_t threads = (n + 7) / 8;
      if (curve == 0) L.run(threads, TableLevelFn<Pallas, Fp>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
      else L.run(threads, TableLevelFn<Vesta, Fq>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
    }
  MsmPlan p = plan_for(n, n, table, c, S, G, logm, is_mont);
  std::vector<fe> sc(n ? n : 1);
  std::memcpy(sc.data(), scalars, n * 32);
  jac_t out;
  ScalarSet ss{{sc.data(), nullptr, nullptr, nullptr}};
  int rc = curve == 0 ? run_msm_arena<Pallas, Fq>(p, pts.data(), ss, &out) : run_msm_arena<Vesta, Fp>(p, pts.data(), ss, &out);
  std::memcpy(out96, &out, 96);
  return rc;
}

// batched MSM: k scalar vectors (lengths lens[j] <= n, concatenated in `scalars`) over the same n points
int emul_msm_batch(int curve, int table, uint32_t c, uint32_t S, const void* affine72, size_t n, const void* scalars,
                   const uint32_t* lens, uint32_t k, void* out96k) {
  HostLaunch L;
  uint32_t W = table ? msm_windows(c) : 1;
  std::vector<affine_t> pts((size_t)W * (n ? n : 1));
  L.run(n, RepackFn{(const uint8_t*)affine72, pts.data()});
  if (table)
    for (uint32_t l = 1; l < W; l++) {
      size_t threads = (n + 7) / 8;
      if (curve == 0) L.run(threads, TableLevelFn<Pallas, Fp>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
      else L.run(threads, TableLevelFn<Vesta, Fq>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
    }
  MsmPlan p = plan_for(n, n, table, c, S, 4, 2, 1);
  p.batch = k;
  size_t total = 0;
  for (uint32_t j = 0; j < k && j < MSM_MAX_BATCH; j++) total += lens[j];
  std::vector<fe> sc(total + 1);
  const fe* src = static_cast<const fe*>(scalars);
  for (size_t q = 0; q < total; q++) sc[q] = src[q];
  ScalarSet ss{{nullptr, nullptr, nullptr, nullptr}};
  size_t off = 0;
  for (uint32_t j = 0; j < k; j++) { ss.v[j] = sc.data() + off; p.len[j] = lens[j]; off += lens[j]; }
  std::vector<jac_t> out(k);
  int rc = curve == 0 ? run_msm_arena<Pallas, Fq>(p, pts.data(), ss, out.data())
                      : run_msm_arena<Vesta, Fp>(p, pts.data(), ss, out.data());
  std::memcpy(out96k, out.data(), 96 * (size_t)k);
  return rc;
}

// the host API's chunked schedule: stages 1-5 per point-range chunk into one bucket array, stages 6-7 once
int emul_msm_chunked(int curve, int table, uint32_t c, uint32_t S, const void* affine72, size_t n, const void* scalars,
                     uint32_t chunks, void* out96) {
  HostLaunch L;
  uint32_t W = table ? msm_windows(c) : 1;
  std::vector<affine_t> pts((size_t)W * (n ? n : 1));
  L.run(n, RepackFn{(const uint8_t*)affine72, pts.data()});
  if (table)
    for (uint32_t l = 1; l < W; l++) {
      size_t threads = (n + 7) / 8;
      if (curve == 0) L.run(threads, TableLevelFn<Pallas, Fp>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
      else L.run(threads, TableLevelFn<Vesta, Fq>{pts.data() + (size_t)(l - 1) * n, pts.data() + (size_t)l * n, n, c});
    }
  MsmPlan full = plan_for(n, n, table, c, S, 8, 3, 1);
  std::vector<fe> sc(n + 1);
  std::memcpy(sc.data(), scalars, n * 32);
  std::vector<xyzz_t> buckets((size_t)full.NB * full.B), chunk_buckets((size_t)full.NB * full.B);
  std::memset(buckets.data(), 0, buckets.size() * sizeof(xyzz_t));
  size_t per = (n + chunks - 1) / chunks;
  for (uint32_t k = 0; k < chunks; k++) {
    size_t lo = k * per, len = lo < n ? (lo + per <= n ? per : n - lo) : 0;
    if (!len) continue;
    MsmPlan p = full;
    p.n = (uint32_t)len;
    p.len[0] = (uint32_t)len;
    ScalarSet ss{{sc.data() + lo, nullptr, nullptr, nullptr}};
    xyzz_t* dst = k ? chunk_buckets.data() : buckets.data();
    if (k) std::memset(dst, 0, chunk_buckets.size() * sizeof(xyzz_t));
    if (curve == 0) {
      msm_accumulate<HostLaunch, Pallas, Fq>(L, p, pts.data() + lo, ss, dst);
      if (k) msm_merge_buckets<HostLaunch, Pallas>(L, full, buckets.data(), chunk_buckets.data());
    } else {
      msm_accumulate<HostLaunch, Vesta, Fp>(L, p, pts.data() + lo, ss, dst);
      if (k) msm_merge_buckets<HostLaunch, Vesta>(L, full, buckets.data(), chunk_buckets.data());
    }
  }
  jac_t out;
  if (curve == 0) msm_finish<HostLaunch, Pallas>(L, full, buckets.data(), &out);
  else msm_finish<HostLaunch, Vesta>(L, full, buckets.data(), &out);
  std::memcpy(out96, &out, 96);
  return 0;
}

int emul_progression(int curve, const void* k0, const void* d, size_t n, void* out72) {
  HostLaunch L;
  fe fk0, fd;
  std::memcpy(fk0.v, k0, 32);
  std::memcpy(fd.v, d, 32);
  ProgSetup setup;
  std::vector<affine_t> pts(n ? n : 1);
  size_t threads = (n + PROG_CH - 1) / PROG_CH;
  if (curve == 0) {
    L.run(1, ProgressionSetupFn<Pallas, Fp>{fk0, fd, &setup});
    L.run(threads, ProgressionFn<Pallas, Fp>{&setup, pts.data(), n});
  } else {
    L.run(1, ProgressionSetupFn<Vesta, Fq>{fk0, fd, &setup});
    L.run(threads, ProgressionFn<Vesta, Fq>{&setup, pts.data(), n});
  }
  L.run(n, UnpackFn{pts.data(), (uint8_t*)out72});
  return 0;
}

int emul_point_sum(int curve, const void* pts96, size_t k, void* out96) {
  HostLaunch L;
  jac_t out;
  if (curve == 0) L.run(1, JacSumFn<Pallas>{(const jac_t*)pts96, (uint32_t)k, &out});
  else L.run(1, JacSumFn<Vesta>{(const jac_t*)pts96, (uint32_t)k, &out});
  std::memcpy(out96, &out, 96);
  return 0;
}

int emul_minroot_check(int field, const void* results, const void* originals, const uint64_t* t_each,
                       uint64_t t_uniform, size_t n, uint8_t* ok) {
  HostLaunch L;
  if (field == 0) L.run(n, MinRootCheckFn<Fp>{(const state_t*)results, (const state_t*)originals, t_each, t_uniform, ok});
  else L.run(n, MinRootCheckFn<Fq>{(const state_t*)results, (const state_t*)originals, t_each, t_uniform, ok});
  return 0;
}

int emul_minroot_inverse_eval(int field, const void* results, uint64_t t, size_t n, void* out) {
  HostLaunch L;
  if (field == 0) L.run(n, MinRootInverseEvalFn<Fp>{(const state_t*)results, t, (state_t*)out});
  else L.run(n, MinRootInverseEvalFn<Fq>{(const state_t*)results, t, (state_t*)out});
  return 0;
}

int emul_minroot_witness(int field, const void* results, uint64_t t, size_t n, void* out) {
  HostLaunch L;
  if (field == 0) L.run(n, MinRootWitnessFn<Fp>{(const state_t*)results, t, (fe*)out});
  else L.run(n, MinRootWitnessFn<Fq>{(const state_t*)results, t, (fe*)out});
  return 0;
}

// R1CS: mode 0 = multiply_vec (out = Az|Bz|Cz, 3*cons elements, z1 only), mode 1 = cross-term T (cons elements),
// mode 2 = bind_rows (W1 = eq table of cons elements, u1 = the three challenges; out = vars+1+io elements)
int emul_r1cs(int field, int mode, size_t num_cons, size_t num_vars, size_t num_io, const uint64_t* a_rows,
              const uint64_t* a_cols, const void* a_vals, size_t a_nnz, const uint64_t* b_rows, const uint64_t* b_cols,
              const void* b_vals, size_t b_nnz, const uint64_t* c_rows, const uint64_t* c_cols, const void* c_vals,
              size_t c_nnz, const void* W1, const void* u1, const void* X1, const void* W2, const void* X2, void* out) {
  const uint64_t* rows[3] = {a_rows, b_rows, c_rows};
  const uint64_t* cols[3] = {a_cols, b_cols, c_cols};
  const uint8_t* vals[3] = {(const uint8_t*)a_vals, (const uint8_t*)b_vals, (const uint8_t*)c_vals};
  size_t nnzs[3] = {a_nnz, b_nnz, c_nnz};
  HostCsr csr;
  try {
    csr = coo_to_csr(num_cons, num_vars + 1 + num_io, rows, cols, vals, nnzs);
  } catch (const std::exception&) {
    return -1;
  }
  std::vector<fe> val(csr.val.size() / 32 + 1);
  std::memcpy(val.data(), csr.val.data(), csr.val.size());
  CsrView m{csr.row_ptr.data(), csr.col.data(), val.data(), (uint32_t)num_cons, (uint32_t)num_vars, (uint32_t)num_io};
  HostLaunch L;
  if (mode == 2) {
    const size_t ncols = num_vars + 1 + num_io;
    const HostCsc csc = csr_to_csc(csr, ncols);
    std::vector<fe> cval(csc.val.size() / 32 + 1), eq(num_cons), coef(3), eq3(3 * num_cons);
    std::memcpy(cval.data(), csc.val.data(), csc.val.size());
    std::memcpy(eq.data(), W1, num_cons * 32);
    std::memcpy(coef.data(), u1, 96);
    std::vector<uint32_t> heavy(csc.heavy);
    heavy.push_back(0);
    CscView v{csc.col_ptr.data(), csc.srow.data(), cval.data(), heavy.data(), HostCsc::HEAVY};
    fe* o = (fe*)out;
    const size_t nh = csc.heavy.size();
    if (field == 0) {
      L.run(3 * num_cons, ScaleRowsFn<Fp>{eq.data(), coef.data(), (uint32_t)num_cons, eq3.data()});
      L.run(ncols, BindRowsFn<Fp>{v, eq3.data(), o});
      L.run(nh * BindHeavyFn<Fp>::THREADS, BindHeavyFn<Fp>{v, eq3.data(), o});
    } else {
      L.run(3 * num_cons, ScaleRowsFn<Fq>{eq.data(), coef.data(), (uint32_t)num_cons, eq3.data()});
      L.run(ncols, BindRowsFn<Fq>{v, eq3.data(), o});
      L.run(nh * BindHeavyFn<Fq>::THREADS, BindHeavyFn<Fq>{v, eq3.data(), o});
    }
    return (int)nh;   // number of heavy columns (the test wants the warp path exercised)
  }
  std::vector<fe> w1(num_vars + 1), w2(num_vars + 1), ux1(1 + num_io), ux2(1 + num_io);
  std::memcpy(w1.data(), W1, num_vars * 32);
  std::memcpy(ux1.data(), u1, 32);
  std::memcpy(ux1.data() + 1, X1, num_io * 32);
  ZView z1{w1.data(), ux1.data(), ux1.data() + 1};
  fe* o = (fe*)out;
  if (mode == 0) {
    if (field == 0) L.run(3 * num_cons, MultiplyVecFn<Fp>{m, z1, o, o + num_cons, o + 2 * num_cons});
    else L.run(3 * num_cons, MultiplyVecFn<Fq>{m, z1, o, o + num_cons, o + 2 * num_cons});
    return 0;
  }
  std::memcpy(w2.data(), W2, num_vars * 32);
  ux2[0] = field == 0 ? Fp::one() : Fq::one();
  std::memcpy(ux2.data() + 1, X2, num_io * 32);
  ZView z2{w2.data(), ux2.data(), ux2.data() + 1};
  if (field == 0) L.run(num_cons, CrossTermFn<Fp>{m, z1, z2, o});
  else L.run(num_cons, CrossTermFn<Fq>{m, z1, z2, o});
  return 0;
}

int emul_fold(int field, void* W1, const void* W2, size_t nW, void* E1, const void* T, size_t nE, const void* r) {
  HostLaunch L;
  fe rr;
  std::memcpy(&rr, r, 32);
  if (field == 0) L.run(nW + nE, FoldFn<Fp>{(fe*)W1, (const fe*)W2, nW, (fe*)E1, (const fe*)T, nE, &rr});
  else L.run(nW + nE, FoldFn<Fq>{(fe*)W1, (const fe*)W2, nW, (fe*)E1, (const fe*)T, nE, &rr});
  return 0;
}

}  // extern "C"
