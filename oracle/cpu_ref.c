/* oracle/cpu_ref.c -- C restatement of the reference's CPU algorithms for the hot path.
 *
 * TEST INFRASTRUCTURE + CPU BASELINE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (libvdfgpu.so) never links or calls this file.
 *
 * PARITY UNPINNED against the Rust crates (no Rust toolchain, dependencies not vendored, no golden values
 * in the reference's tests: SURVEY.md 0.3/0.4).  It is pinned against oracle/pasta.py (Python integers)
 * by tests/test_cpu_ref.py, and through it against the reference's own constants and addition chains.
 *
 * What it restates (kind = "port" in bench.py):
 *   fe_*            pasta_curves 0.4 Fp/Fq: four u64 limbs, Montgomery R = 2^256 (Cargo.toml:17)
 *   ref_msm         pasta-msm 0.1 / sppark CPU Pippenger: signed windows, XYZZ buckets, (window x point
 *                   tile) tasks over a thread pool (Cargo.toml:18; reached from src/nova/proof.rs:342)
 *   ref_multiply_vec / ref_cross_term / ref_fold   nova-snark 0.8 R1CSShape::multiply_vec (three COO
 *                   products run concurrently), commit_T's T, RelaxedR1CSWitness::fold (Cargo.toml:15)
 *   ref_minroot_check   MinRootVDF::check over independent chains, src/minroot.rs:338-371 verbatim
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { fe m, one, r2; uint64_t inv; } field_t;   /* inv = -m^-1 mod 2^64 */
typedef struct { fe x, y; } aff_t;                           /* (0,0) = identity */
typedef struct { fe X, Y, ZZ, ZZZ; } xyzz_t;

/* SURVEY.md section 8 / Appendix A constants */
static const field_t FP = {
  {{0x992d30ed00000001ull, 0x224698fc094cf91bull, 0x0000000000000000ull, 0x4000000000000000ull}},
  {{0x34786d38fffffffdull, 0x992c350be41914adull, 0xffffffffffffffffull, 0x3fffffffffffffffull}},
  {{0, 0, 0, 0}}, 0x992d30ecffffffffull};
static const field_t FQ = {
  {{0x8c46eb2100000001ull, 0x224698fc0994a8ddull, 0x0000000000000000ull, 0x4000000000000000ull}},
  {{0x5b2b3e9cfffffffdull, 0x992c350be3420567ull, 0xffffffffffffffffull, 0x3fffffffffffffffull}},
  {{0, 0, 0, 0}}, 0x8c46eb20ffffffffull};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}

static inline void fe_cond_sub(fe* r, uint64_t hi, const field_t* f) {
  fe t; u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)r->l[i] - f->m.l[i] - (uint64_t)b;
    t.l[i] = (uint64_t)d; b = (d >> 64) & 1;
  }
  if (hi || !b) *r = t;
}

static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* f) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
  fe_cond_sub(r, (uint64_t)c, f);
}

static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* f) {
  u128 bw = 0; fe t;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)bw;
    t.l[i] = (uint64_t)d; bw = (d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)t.l[i] + f->m.l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
  }
  *r = t;
}

static inline void fe_neg(fe* r, const fe* a, const field_t* f) { fe z = {{0, 0, 0, 0}}; fe_sub(r, &z, a, f); }

/* Montgomery multiplication, 4 x 64-bit limbs, R = 2^256 (pasta_curves' representation).
 *
 * x86-64 with BMI2 + ADX: CIOS with the two carry chains of MULX / ADCX / ADOX, the scheme of the assembly that
 * pasta-msm's field layer (semolina) and blst use; valid because the top limb of both moduli is 2^62 (no carry out
 * of the fifth limb).  Measured on the build host against the portable u128 loop below: see DESIGN.md section 6.
 * Other targets: the portable CIOS loop. */
#if defined(__x86_64__) && defined(__BMI2__) && defined(__ADX__) && !defined(VDF_REF_PORTABLE)
#define VDF_REF_ADX 1
#define MONT_ROW(boff)                                                                          \
  "movq " #boff "(%[b]), %%rdx\n\t"                                                             \
  "xorq %[lo], %[lo]\n\t"                                                                       \
  "mulx 0(%[a]), %[lo], %[hi]\n\t"  "adox %[lo], %[t0]\n\t"                                     \
  "adcx %[hi], %[t1]\n\t" "mulx 8(%[a]), %[lo], %[hi]\n\t"  "adox %[lo], %[t1]\n\t"             \
  "adcx %[hi], %[t2]\n\t" "mulx 16(%[a]), %[lo], %[hi]\n\t" "adox %[lo], %[t2]\n\t"             \
  "adcx %[hi], %[t3]\n\t" "mulx 24(%[a]), %[lo], %[A]\n\t"  "adox %[lo], %[t3]\n\t"             \
  "movl $0, %k[lo]\n\t"   "adcx %[lo], %[A]\n\t"            "adox %[lo], %[A]\n\t"              \
  "movq %[inv], %%rdx\n\t" "imulq %[t0], %%rdx\n\t"                                             \
  "xorq %[lo], %[lo]\n\t"                                                                       \
  "mulx 0(%[m]), %[lo], %[hi]\n\t"  "adcx %[t0], %[lo]\n\t" "movq %[hi], %[t0]\n\t"             \
  "adcx %[t1], %[t0]\n\t" "mulx 8(%[m]), %[lo], %[t1]\n\t"  "adox %[lo], %[t0]\n\t"             \
  "adcx %[t2], %[t1]\n\t" "mulx 16(%[m]), %[lo], %[t2]\n\t" "adox %[lo], %[t1]\n\t"             \
  "adcx %[t3], %[t2]\n\t" "mulx 24(%[m]), %[lo], %[t3]\n\t" "adox %[lo], %[t2]\n\t"             \
  "movl $0, %k[lo]\n\t"   "adcx %[lo], %[t3]\n\t"           "adox %[A], %[t3]\n\t"

static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* f) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, A, lo, hi;
  __asm__(MONT_ROW(0) MONT_ROW(8) MONT_ROW(16) MONT_ROW(24)
          : [t0] "+&r"(t0), [t1] "+&r"(t1), [t2] "+&r"(t2), [t3] "+&r"(t3), [A] "=&r"(A), [lo] "=&r"(lo), [hi] "=&r"(hi)
          : [a] "r"(a->l), [b] "r"(b->l), [m] "r"(f->m.l), [inv] "rm"(f->inv), "m"(*a), "m"(*b), "m"(f->m)
          : "rdx", "cc");
  fe o = {{t0, t1, t2, t3}};
  fe_cond_sub(&o, 0, f);
  *r = o;
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* f) { fe_mul(r, a, a, f); }
#else
/* portable CIOS, u128 accumulators */
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* f) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * f->inv;
    c = ((u128)q * f->m.l[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { c += (u128)q * f->m.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe o = {{t[0], t[1], t[2], t[3]}};
  fe_cond_sub(&o, t[4], f);
  *r = o;
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* f) { fe_mul(r, a, a, f); }
#endif

static void fe_from_mont(fe* r, const fe* a, const field_t* f) { fe one = {{1, 0, 0, 0}}; fe_mul(r, a, &one, f); }

static void fe_inv(fe* r, const fe* a, const field_t* f) {  /* a^(m-2) */
  fe e = f->m; e.l[0] -= 2;                                  /* m0 ends ...0001: borrow needed */
  /* m.l[0] = ....00000001, so subtracting 2 borrows from higher bits of the same limb (nonzero) */
  fe acc = f->one;
  for (int i = 255; i >= 0; i--) {
    fe_sqr(&acc, &acc, f);
    if ((e.l[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a, f);
  }
  *r = acc;
}

/* ---- XYZZ group law on y^2 = x^3 + 5 (EFD madd-2008-s / add-2008-s / dbl-2008-s-1, a = 0) ---- */
static inline int xyzz_is_inf(const xyzz_t* p) { return fe_is_zero(&p->ZZ); }
static inline int aff_is_inf(const aff_t* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }

static void xyzz_dbl(xyzz_t* r, const xyzz_t* p, const field_t* f) {
  if (xyzz_is_inf(p)) { memset(r, 0, sizeof *r); return; }
  fe U, V, W, S, M, t, X3, Y3;
  fe_add(&U, &p->Y, &p->Y, f); fe_sqr(&V, &U, f); fe_mul(&W, &U, &V, f); fe_mul(&S, &p->X, &V, f);
  fe_sqr(&t, &p->X, f); fe_add(&M, &t, &t, f); fe_add(&M, &M, &t, f);
  fe_sqr(&X3, &M, f); fe_sub(&X3, &X3, &S, f); fe_sub(&X3, &X3, &S, f);
  fe_sub(&t, &S, &X3, f); fe_mul(&Y3, &M, &t, f); fe_mul(&t, &W, &p->Y, f); fe_sub(&Y3, &Y3, &t, f);
  fe zz, zzz; fe_mul(&zz, &V, &p->ZZ, f); fe_mul(&zzz, &W, &p->ZZZ, f);
  r->X = X3; r->Y = Y3; r->ZZ = zz; r->ZZZ = zzz;
}

static void xyzz_madd(xyzz_t* acc, const fe* x2, const fe* y2, const field_t* f) {
  if (xyzz_is_inf(acc)) { acc->X = *x2; acc->Y = *y2; acc->ZZ = f->one; acc->ZZZ = f->one; return; }
  fe U2, S2, P, R, PP, PPP, Q, t, X3, Y3;
  fe_mul(&U2, x2, &acc->ZZ, f); fe_mul(&S2, y2, &acc->ZZZ, f);
  fe_sub(&P, &U2, &acc->X, f); fe_sub(&R, &S2, &acc->Y, f);
  if (fe_is_zero(&P)) {
    if (fe_is_zero(&R)) { xyzz_t a = {*x2, *y2, f->one, f->one}; xyzz_dbl(acc, &a, f); }
    else memset(acc, 0, sizeof *acc);
    return;
  }
  fe_sqr(&PP, &P, f); fe_mul(&PPP, &P, &PP, f); fe_mul(&Q, &acc->X, &PP, f);
  fe_sqr(&X3, &R, f); fe_sub(&X3, &X3, &PPP, f); fe_sub(&X3, &X3, &Q, f); fe_sub(&X3, &X3, &Q, f);
  fe_sub(&t, &Q, &X3, f); fe_mul(&Y3, &R, &t, f); fe_mul(&t, &acc->Y, &PPP, f); fe_sub(&Y3, &Y3, &t, f);
  fe_mul(&acc->ZZ, &acc->ZZ, &PP, f); fe_mul(&acc->ZZZ, &acc->ZZZ, &PPP, f);
  acc->X = X3; acc->Y = Y3;
}

static void xyzz_add(xyzz_t* acc, const xyzz_t* q, const field_t* f) {
  if (xyzz_is_inf(q)) return;
  if (xyzz_is_inf(acc)) { *acc = *q; return; }
  fe U1, U2, S1, S2, P, R, PP, PPP, Q, t, X3, Y3;
  fe_mul(&U1, &acc->X, &q->ZZ, f); fe_mul(&U2, &q->X, &acc->ZZ, f);
  fe_mul(&S1, &acc->Y, &q->ZZZ, f); fe_mul(&S2, &q->Y, &acc->ZZZ, f);
  fe_sub(&P, &U2, &U1, f); fe_sub(&R, &S2, &S1, f);
  if (fe_is_zero(&P)) {
    if (fe_is_zero(&R)) { xyzz_t a = *acc; xyzz_dbl(acc, &a, f); }
    else memset(acc, 0, sizeof *acc);
    return;
  }
  fe_sqr(&PP, &P, f); fe_mul(&PPP, &P, &PP, f); fe_mul(&Q, &U1, &PP, f);
  fe_sqr(&X3, &R, f); fe_sub(&X3, &X3, &PPP, f); fe_sub(&X3, &X3, &Q, f); fe_sub(&X3, &X3, &Q, f);
  fe_sub(&t, &Q, &X3, f); fe_mul(&Y3, &R, &t, f); fe_mul(&t, &S1, &PPP, f); fe_sub(&Y3, &Y3, &t, f);
  fe_mul(&t, &acc->ZZ, &q->ZZ, f); fe_mul(&acc->ZZ, &t, &PP, f);
  fe_mul(&t, &acc->ZZZ, &q->ZZZ, f); fe_mul(&acc->ZZZ, &t, &PPP, f);
  acc->X = X3; acc->Y = Y3;
}

static void xyzz_to_jac96(uint8_t* out, const xyzz_t* p, const field_t* f) {
  if (xyzz_is_inf(p)) { memset(out, 0, 96); return; }
  fe t, i, x, y;
  fe_mul(&t, &p->ZZ, &p->ZZZ, f); fe_inv(&i, &t, f);
  fe_mul(&t, &i, &p->ZZZ, f); fe_mul(&x, &p->X, &t, f);
  fe_mul(&t, &i, &p->ZZ, f); fe_mul(&y, &p->Y, &t, f);
  memcpy(out, &x, 32); memcpy(out + 32, &y, 32); memcpy(out + 64, &f->one, 32);
}

static const field_t* base_field(int curve) { return curve == 0 ? &FP : &FQ; }
static const field_t* scalar_field(int curve) { return curve == 0 ? &FQ : &FP; }
static const field_t* field_by_id(int id) { return id == 0 ? &FP : &FQ; }

/* ---- persistent thread pool: run fn(arg, task) for task in [0, ntasks) on up to nthreads threads -------------
 * Workers are created once and sleep on a condition variable between jobs (pasta-msm / sppark keep a thread pool
 * too); the calling thread works alongside them.  One job at a time (callers are serialised by job_mu). */
#define POOL_MAX 256
static struct {
  pthread_mutex_t job_mu, mu;
  pthread_cond_t wake, done;
  pthread_t th[POOL_MAX];
  int nworkers;
  unsigned long generation;
  void (*fn)(void*, size_t);
  void* arg;
  size_t ntasks, next, finished;
  int active_limit, active;      /* workers allowed to join this job / workers that did */
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
            {0}, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static void pool_drain(void) {   /* called with g_pool.mu held; returns with it held */
  while (g_pool.next < g_pool.ntasks) {
    size_t t = g_pool.next++;
    pthread_mutex_unlock(&g_pool.mu);
    g_pool.fn(g_pool.arg, t);
    pthread_mutex_lock(&g_pool.mu);
    if (++g_pool.finished == g_pool.ntasks) pthread_cond_broadcast(&g_pool.done);
  }
}

static void* pool_worker(void* unused) {
  (void)unused;
  unsigned long seen = 0;
  pthread_mutex_lock(&g_pool.mu);
  for (;;) {
    while (g_pool.generation == seen) pthread_cond_wait(&g_pool.wake, &g_pool.mu);
    seen = g_pool.generation;
    if (g_pool.active < g_pool.active_limit) {
      g_pool.active++;
      pool_drain();
    }
  }
  return NULL;
}

static void parallel_for(size_t ntasks, int nthreads, void (*fn)(void*, size_t), void* arg) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > POOL_MAX) nthreads = POOL_MAX;
  if ((size_t)nthreads > ntasks) nthreads = (int)ntasks;
  if (nthreads <= 1) { for (size_t t = 0; t < ntasks; t++) fn(arg, t); return; }
  pthread_mutex_lock(&g_pool.job_mu);
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.nworkers < nthreads - 1) {
    if (pthread_create(&g_pool.th[g_pool.nworkers], NULL, pool_worker, NULL) != 0) break;
    pthread_detach(g_pool.th[g_pool.nworkers]);
    g_pool.nworkers++;
  }
  g_pool.fn = fn; g_pool.arg = arg; g_pool.ntasks = ntasks; g_pool.next = 0; g_pool.finished = 0;
  g_pool.active_limit = nthreads - 1; g_pool.active = 0;
  g_pool.generation++;
  pthread_cond_broadcast(&g_pool.wake);
  pool_drain();
  while (g_pool.finished < g_pool.ntasks) pthread_cond_wait(&g_pool.done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  pthread_mutex_unlock(&g_pool.job_mu);
}

/* ---- MSM ---- */
typedef struct {
  const field_t* f; const aff_t* pts; const fe* sc; size_t n;
  unsigned c, W, slices; size_t slice_len; xyzz_t* partial;   /* [W][slices] */
} msm_job_t;

static inline unsigned get_bits(const fe* s, unsigned bit, unsigned c) {
  unsigned limb = bit >> 6, sh = bit & 63;
  uint64_t v = s->l[limb] >> sh;
  if (sh + c > 64 && limb + 1 < 4) v |= s->l[limb + 1] << (64 - sh);
  return (unsigned)(v & ((1ull << c) - 1));
}

/* signed digit of window w: booth recoding on (c+1) bits, as in sppark's pippenger */
static inline int booth_digit(const fe* s, unsigned w, unsigned c) {
  unsigned bit = w * c;
  unsigned raw;
  if (bit == 0) raw = get_bits(s, 0, c) << 1;
  else {
    unsigned avail = 256 - (bit - 1);
    unsigned take = c + 1 < avail ? c + 1 : avail;
    raw = get_bits(s, bit - 1, take);
  }
  /* raw has c+1 bits: value = -(top bit)*2^c + low c bits ... booth: d = ((raw + 1) >> 1) - (top ? 2^c : 0) */
  int top = (raw >> c) & 1;
  int d = (int)((raw + 1) >> 1);
  if (top) d -= (1 << c);
  return d;
}

static void msm_task(void* arg, size_t task) {
  msm_job_t* j = (msm_job_t*)arg;
  unsigned w = (unsigned)(task / j->slices), sl = (unsigned)(task % j->slices);
  size_t lo = sl * j->slice_len, hi = lo + j->slice_len < j->n ? lo + j->slice_len : j->n;
  size_t nb = (size_t)1 << (j->c - 1);
  xyzz_t* buckets = (xyzz_t*)calloc(nb + 1, sizeof(xyzz_t));
  for (size_t i = lo; i < hi; i++) {
    if (aff_is_inf(&j->pts[i])) continue;
    int d = booth_digit(&j->sc[i], w, j->c);
    if (d > 0) xyzz_madd(&buckets[d], &j->pts[i].x, &j->pts[i].y, j->f);
    else if (d < 0) { fe ny; fe_neg(&ny, &j->pts[i].y, j->f); xyzz_madd(&buckets[-d], &j->pts[i].x, &ny, j->f); }
  }
  xyzz_t run, acc; memset(&run, 0, sizeof run); memset(&acc, 0, sizeof acc);
  for (size_t b = nb; b >= 1; b--) { xyzz_add(&run, &buckets[b], j->f); xyzz_add(&acc, &run, j->f); }
  j->partial[task] = acc;
  free(buckets);
}

static unsigned pick_window(size_t n) {
  unsigned lg = 0; while (((size_t)1 << (lg + 1)) <= n) lg++;
  unsigned c = lg > 4 ? lg - 3 : 2;   /* sppark/pasta-msm use ~log2(n) - 2..3 on the CPU */
  if (c > 18) c = 18;
  return c;
}

static void unpack72(aff_t* dst, const uint8_t* src, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (src[72 * i + 64]) memset(&dst[i], 0, sizeof(aff_t));
    else memcpy(&dst[i], src + 72 * i, 64);
  }
}

int ref_msm(int curve, const uint8_t* affine72, size_t n, const uint8_t* scalars, int is_mont, int nthreads,
            uint8_t* out96) {
  const field_t* f = base_field(curve);
  const field_t* sf = scalar_field(curve);
  if (n == 0) { memset(out96, 0, 96); return 0; }
  aff_t* pts = (aff_t*)malloc(n * sizeof(aff_t));
  fe* sc = (fe*)malloc(n * sizeof(fe));
  unpack72(pts, affine72, n);
  memcpy(sc, scalars, n * 32);
  if (is_mont) for (size_t i = 0; i < n; i++) fe_from_mont(&sc[i], &sc[i], sf);
  msm_job_t j; j.f = f; j.pts = pts; j.sc = sc; j.n = n;
  j.c = pick_window(n);
  j.W = (255 + j.c) / j.c;              /* booth needs one extra bit: windows cover 256 bits */
  if (j.W * j.c < 256) j.W++;
  j.slices = (2 * (unsigned)nthreads + j.W - 1) / j.W; if (j.slices < 1) j.slices = 1;
  if (n < 4096) j.slices = 1;
  j.slice_len = (n + j.slices - 1) / j.slices;
  j.partial = (xyzz_t*)calloc((size_t)j.W * j.slices, sizeof(xyzz_t));
  parallel_for((size_t)j.W * j.slices, nthreads, msm_task, &j);
  xyzz_t total; memset(&total, 0, sizeof total);
  for (unsigned w = j.W; w-- > 0;) {
    for (unsigned k = 0; k < j.c; k++) { xyzz_t t = total; xyzz_dbl(&total, &t, f); }
    for (unsigned s = 0; s < j.slices; s++) xyzz_add(&total, &j.partial[(size_t)w * j.slices + s], f);
  }
  xyzz_to_jac96(out96, &total, f);
  free(pts); free(sc); free(j.partial);
  return 0;
}

/* known-dlog progression P_i = (k0 + i d) G, G = (-1, 2): synthetic generator sets for the CPU baseline.
 * Point ranges run in parallel (range start = K0 + lo * D by double-and-add over the 64-bit lo), each range is
 * normalised to affine with batched inversions (Montgomery's trick, 256 points per inversion). */
typedef struct { const field_t* f; xyzz_t K0, D; fe dx, dy; int d_inf; size_t n, chunk; uint8_t* out72; } prog_job_t;

static void prog_task(void* arg, size_t task) {
  prog_job_t* j = (prog_job_t*)arg;
  const field_t* f = j->f;
  size_t lo = task * j->chunk, hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
  xyzz_t cur; memset(&cur, 0, sizeof cur);
  for (int bit = 63; bit >= 0; bit--) {               /* cur = lo * D */
    xyzz_t t = cur; xyzz_dbl(&cur, &t, f);
    if (((uint64_t)lo >> bit) & 1) xyzz_add(&cur, &j->D, f);
  }
  xyzz_add(&cur, &j->K0, f);
  enum { B = 256 };
  xyzz_t q[B]; fe pre[B];
  for (size_t base = lo; base < hi; base += B) {
    size_t m = base + B < hi ? B : hi - base;
    fe run = f->one;
    for (size_t k = 0; k < m; k++) {
      q[k] = cur; pre[k] = run;
      if (!xyzz_is_inf(&cur)) { fe t; fe_mul(&t, &cur.ZZ, &cur.ZZZ, f); fe_mul(&run, &run, &t, f); }
      if (!j->d_inf) xyzz_madd(&cur, &j->dx, &j->dy, f);
    }
    fe inv; fe_inv(&inv, &run, f);
    for (size_t k = m; k-- > 0;) {
      uint8_t* o = j->out72 + 72 * (base + k);
      memset(o, 0, 72);
      if (xyzz_is_inf(&q[k])) { o[64] = 1; continue; }
      fe zi, t, x, y;
      fe_mul(&zi, &inv, &pre[k], f);                                  /* 1 / (ZZ * ZZZ) */
      fe_mul(&t, &q[k].ZZ, &q[k].ZZZ, f); fe_mul(&inv, &inv, &t, f);
      fe_mul(&t, &zi, &q[k].ZZZ, f); fe_mul(&x, &q[k].X, &t, f);
      fe_mul(&t, &zi, &q[k].ZZ, f); fe_mul(&y, &q[k].Y, &t, f);
      memcpy(o, &x, 32); memcpy(o + 32, &y, 32);
    }
  }
}

int ref_progression_mt(int curve, const uint8_t* k0_le32, const uint8_t* d_le32, size_t n, int nthreads, uint8_t* out72) {
  const field_t* f = base_field(curve);
  fe gx, gy; fe_neg(&gx, &f->one, f); fe_add(&gy, &f->one, &f->one, f);
  fe k0, d; memcpy(&k0, k0_le32, 32); memcpy(&d, d_le32, 32);
  prog_job_t j; j.f = f; j.n = n; j.out72 = out72;
  memset(&j.K0, 0, sizeof j.K0); memset(&j.D, 0, sizeof j.D);
  for (int i = 255; i >= 0; i--) {
    xyzz_t t = j.K0; xyzz_dbl(&j.K0, &t, f); t = j.D; xyzz_dbl(&j.D, &t, f);
    if ((k0.l[i >> 6] >> (i & 63)) & 1) xyzz_madd(&j.K0, &gx, &gy, f);
    if ((d.l[i >> 6] >> (i & 63)) & 1) xyzz_madd(&j.D, &gx, &gy, f);
  }
  uint8_t dj[96]; xyzz_to_jac96(dj, &j.D, f);
  memcpy(&j.dx, dj, 32); memcpy(&j.dy, dj + 32, 32);
  j.d_inf = xyzz_is_inf(&j.D);
  if (nthreads < 1) nthreads = 1;
  j.chunk = (n + (size_t)nthreads * 4 - 1) / ((size_t)nthreads * 4);
  if (j.chunk < 256) j.chunk = 256;
  if (n) parallel_for((n + j.chunk - 1) / j.chunk, nthreads, prog_task, &j);
  return 0;
}

int ref_progression(int curve, const uint8_t* k0_le32, const uint8_t* d_le32, size_t n, uint8_t* out72) {
  return ref_progression_mt(curve, k0_le32, d_le32, n, 1, out72);
}

/* ---- MinRoot check (src/minroot.rs:338-371) ---- */
typedef struct { const field_t* f; const uint8_t* res; const uint8_t* orig; const uint64_t* t_each; uint64_t t_uniform;
                 size_t n; uint8_t* ok; size_t chunk; } mr_job_t;
static void mr_task(void* arg, size_t task) {
  mr_job_t* j = (mr_job_t*)arg;
  size_t lo = task * j->chunk, hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
  const field_t* f = j->f;
  for (size_t k = lo; k < hi; k++) {
    fe x, y, i; memcpy(&x, j->res + 96 * k, 32); memcpy(&y, j->res + 96 * k + 32, 32); memcpy(&i, j->res + 96 * k + 64, 32);
    uint64_t t = j->t_each ? j->t_each[k] : j->t_uniform;
    for (uint64_t r = 0; r < t; r++) {
      fe ni, nx, x2, x4, x5, ny;
      fe_sub(&ni, &i, &f->one, f);            /* minroot.rs:339 */
      fe_sub(&nx, &y, &ni, f);                /* :340 */
      fe_sqr(&x2, &x, f); fe_sqr(&x4, &x2, f); fe_mul(&x5, &x, &x4, f);  /* :73-75 */
      fe_sub(&ny, &x5, &nx, f);               /* :341-342 */
      x = nx; y = ny; i = ni;
    }
    fe ox, oy, oi; memcpy(&ox, j->orig + 96 * k, 32); memcpy(&oy, j->orig + 96 * k + 32, 32); memcpy(&oi, j->orig + 96 * k + 64, 32);
    j->ok[k] = fe_eq(&x, &ox) && fe_eq(&y, &oy) && fe_eq(&i, &oi);
  }
}
int ref_minroot_check(int field, const uint8_t* results, const uint8_t* originals, const uint64_t* t_each,
                      uint64_t t_uniform, size_t n, int nthreads, uint8_t* ok) {
  mr_job_t j = {field_by_id(field), results, originals, t_each, t_uniform, n, ok, 0};
  j.chunk = (n + (size_t)nthreads * 8 - 1) / ((size_t)nthreads * 8); if (j.chunk < 1) j.chunk = 1;
  parallel_for((n + j.chunk - 1) / j.chunk, nthreads, mr_task, &j);
  return 0;
}

/* ---- R1CS (nova-snark r1cs.rs semantics) ---- */
typedef struct { const field_t* f; size_t cons, vars; const uint64_t* rows[3]; const uint64_t* cols[3]; const uint8_t* vals[3];
                 size_t nnz[3]; const fe* W; const fe* u; const fe* X; fe* out[3]; } mv_job_t;
static void mv_task(void* arg, size_t m) {
  mv_job_t* j = (mv_job_t*)arg;
  memset(j->out[m], 0, j->cons * sizeof(fe));
  for (size_t k = 0; k < j->nnz[m]; k++) {
    uint64_t c = j->cols[m][k];
    const fe* z = c < j->vars ? &j->W[c] : (c == j->vars ? j->u : &j->X[c - j->vars - 1]);
    fe v, p; memcpy(&v, j->vals[m] + 32 * k, 32);
    fe_mul(&p, &v, z, j->f);
    fe_add(&j->out[m][j->rows[m][k]], &j->out[m][j->rows[m][k]], &p, j->f);
  }
}
int ref_multiply_vec(int field, size_t cons, size_t vars, size_t io,
                     const uint64_t* ar, const uint64_t* ac, const uint8_t* av, size_t an,
                     const uint64_t* br, const uint64_t* bc, const uint8_t* bv, size_t bn,
                     const uint64_t* cr, const uint64_t* cc, const uint8_t* cv, size_t cn,
                     const uint8_t* W, const uint8_t* u, const uint8_t* X, int nthreads, uint8_t* AzBzCz) {
  (void)io;
  mv_job_t j; j.f = field_by_id(field); j.cons = cons; j.vars = vars;
  j.rows[0] = ar; j.rows[1] = br; j.rows[2] = cr; j.cols[0] = ac; j.cols[1] = bc; j.cols[2] = cc;
  j.vals[0] = av; j.vals[1] = bv; j.vals[2] = cv; j.nnz[0] = an; j.nnz[1] = bn; j.nnz[2] = cn;
  j.W = (const fe*)W; j.u = (const fe*)u; j.X = (const fe*)X;
  for (int m = 0; m < 3; m++) j.out[m] = (fe*)(AzBzCz + (size_t)m * cons * 32);
  parallel_for(3, nthreads, mv_task, &j);   /* nova: rayon::join over the three products */
  return 0;
}

/* T = Az1.Bz2 + Az2.Bz1 - u1.Cz2 - Cz1 from the six products (u2 = 1) */
int ref_cross_term(int field, size_t cons, const uint8_t* ABC1, const uint8_t* ABC2, const uint8_t* u1, uint8_t* T) {
  const field_t* f = field_by_id(field);
  const fe* a1 = (const fe*)ABC1; const fe* b1 = a1 + cons; const fe* c1 = b1 + cons;
  const fe* a2 = (const fe*)ABC2; const fe* b2 = a2 + cons; const fe* c2 = b2 + cons;
  fe uu; memcpy(&uu, u1, 32);
  for (size_t i = 0; i < cons; i++) {
    fe t, s;
    fe_mul(&t, &a1[i], &b2[i], f); fe_mul(&s, &a2[i], &b1[i], f); fe_add(&t, &t, &s, f);
    fe_mul(&s, &uu, &c2[i], f); fe_sub(&t, &t, &s, f); fe_sub(&t, &t, &c1[i], f);
    memcpy(T + 32 * i, &t, 32);
  }
  return 0;
}

int ref_fold(int field, uint8_t* a, const uint8_t* b, size_t n, const uint8_t* r) {   /* a <- a + r*b */
  const field_t* f = field_by_id(field);
  fe rr; memcpy(&rr, r, 32);
  for (size_t i = 0; i < n; i++) {
    fe x, y, p; memcpy(&x, a + 32 * i, 32); memcpy(&y, b + 32 * i, 32);
    fe_mul(&p, &rr, &y, f); fe_add(&x, &x, &p, f); memcpy(a + 32 * i, &x, 32);
  }
  return 0;
}

int ref_field_mul(int field, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  const field_t* f = field_by_id(field);
  for (size_t i = 0; i < n; i++) {
    fe x, y, p; memcpy(&x, a + 32 * i, 32); memcpy(&y, b + 32 * i, 32);
    fe_mul(&p, &x, &y, f); memcpy(out + 32 * i, &p, 32);
  }
  return 0;
}
