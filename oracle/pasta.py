"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the protocol/vdf Nova/MinRoot hot path.

PARITY UNPINNED: the reference (/root/reference) is pure Rust whose arithmetic lives in
third-party crates that are not vendored (nova-snark ^0.8.0, pasta-msm ^0.1.1,
pasta_curves ^0.4.0; Cargo.toml:15,17,18), no Rust toolchain exists in this image and the
reference's tests hold no golden values.  What *is* pinned here:

  * the reference's own constants and addition chains (src/minroot.rs:88-127, :223-261,
    :273-285) are restated below and must satisfy  chain(x)**5 == x  under our moduli --
    this pins the two moduli and the field multiplication;
  * the reference's round-trip tests (src/minroot.rs:449-542) are restated in tests/;
  * everything else is exact arithmetic in a prime field / prime-order group with a
    canonical encoding, so "bit-exact with the reference" == "mathematically correct and
    same encoding" (4x u64 little-endian Montgomery limbs, R = 2**256, pasta_curves repr-c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Plain Python integers everywhere: slow, obviously correct, independent of all limb code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Fields (SURVEY.md section 8 notation).  Fp = Pallas base = Vesta scalar; Fq = Pallas scalar
# = Vesta base.  pasta_curves stores value * 2**256 mod m as four little-endian u64 limbs.
# --------------------------------------------------------------------------------------
P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
R_BITS = 256
R = 1 << R_BITS

FIELD_FP = 0  # modulus P
FIELD_FQ = 1  # modulus Q
MODULUS = {FIELD_FP: P, FIELD_FQ: Q}

CURVE_PALLAS = 0  # coordinates in Fp, scalars in Fq, group order Q
CURVE_VESTA = 1   # coordinates in Fq, scalars in Fp, group order P

# src/minroot.rs:273-285 (little-endian u64 limbs)
FP_RESCUE_INVALPHA = [0xE0F0F3F0CCCCCCCD, 0x4E9EE0C9A10A60E2, 0x3333333333333333, 0x3333333333333333]
FQ_RESCUE_INVALPHA = [0xD69F2280CCCCCCCD, 0x4E9EE0C9A143BA4A, 0x3333333333333333, 0x3333333333333333]


def limbs_to_int(limbs: Sequence[int]) -> int:
    return sum(int(l) << (64 * k) for k, l in enumerate(limbs))


def to_mont(v: int, m: int) -> int:
    return (v % m) * R % m


def from_mont(v: int, m: int) -> int:
    return v * pow(R, -1, m) % m


def fe_to_bytes(v: int, m: int) -> bytes:
    """canonical integer -> 32 bytes, pasta_curves in-memory layout (Montgomery, LE limbs)."""
    return to_mont(v, m).to_bytes(32, "little")


def fe_from_bytes(b: bytes, m: int) -> int:
    raw = int.from_bytes(b[:32], "little")
    assert raw < m, "non-canonical Montgomery limb value"
    return from_mont(raw, m)


def fes_to_bytes(vs: Iterable[int], m: int) -> bytes:
    rm = R % m
    return b"".join(((v % m) * rm % m).to_bytes(32, "little") for v in vs)


def fes_from_bytes(b: bytes, m: int) -> List[int]:
    rinv = pow(R, -1, m)
    return [int.from_bytes(b[k:k + 32], "little") * rinv % m for k in range(0, len(b), 32)]


# --------------------------------------------------------------------------------------
# Synthetic-input generator of the reference's tests: XorShiftRng::from_seed([42;16]) and
# pasta_curves Field::random (512-bit LE integer reduced mod m).  SURVEY.md section 8(c).
# --------------------------------------------------------------------------------------
TEST_SEED = bytes([42] * 16)  # src/lib.rs:4


class XorShiftRng:
    def __init__(self, seed: bytes = TEST_SEED):
        assert len(seed) == 16
        w = [int.from_bytes(seed[4 * k:4 * k + 4], "little") for k in range(4)]
        if not any(w):
            w = [0xBAD5EED, 0xBAD5EED, 0xBAD5EED, 0xBAD5EED]
        self.x, self.y, self.z, self.w = w

    def next_u32(self) -> int:
        x = self.x
        t = (x ^ (x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        w = self.w
        self.w = (w ^ (w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo


def field_random(rng: XorShiftRng, m: int) -> int:
    wide = 0
    for k in range(8):
        wide |= rng.next_u64() << (64 * k)
    return wide % m


# --------------------------------------------------------------------------------------
# Curves  y^2 = x^3 + 5  (both).  Affine points are (x, y) tuples or None for the identity.
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Curve:
    cid: int
    name: str
    base: int    # coordinate field modulus
    order: int   # scalar field modulus == group order
    b: int = 5

    @property
    def gen(self) -> Tuple[int, int]:
        return (self.base - 1, 2)  # (-1, 2): 4 == -1 + 5

    def on_curve(self, pt) -> bool:
        if pt is None:
            return True
        x, y = pt
        return (y * y - x * x * x - self.b) % self.base == 0

    def neg(self, pt):
        if pt is None:
            return None
        return (pt[0], (-pt[1]) % self.base)

    def add(self, a, b):
        m = self.base
        if a is None:
            return b
        if b is None:
            return a
        x1, y1 = a
        x2, y2 = b
        if x1 == x2:
            if (y1 + y2) % m == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, m) % m
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, m) % m
        x3 = (lam * lam - x1 - x2) % m
        y3 = (lam * (x1 - x3) - y1) % m
        return (x3, y3)

    # Jacobian (X, Y, Z), identity Z == 0; used for bulk work (no inversions)
    def jac_double(self, pt):
        m = self.base
        X, Y, Z = pt
        if Z == 0 or Y == 0:
            return (0, 1, 0)
        A = X * X % m
        B = Y * Y % m
        C = B * B % m
        D = 2 * ((X + B) * (X + B) - A - C) % m
        E = 3 * A % m
        F = E * E % m
        X3 = (F - 2 * D) % m
        Y3 = (E * (D - X3) - 8 * C) % m
        Z3 = 2 * Y * Z % m
        return (X3, Y3, Z3)

    def jac_add_affine(self, pt, q):
        m = self.base
        if q is None:
            return pt
        X1, Y1, Z1 = pt
        x2, y2 = q
        if Z1 == 0:
            return (x2, y2, 1)
        Z1Z1 = Z1 * Z1 % m
        U2 = x2 * Z1Z1 % m
        S2 = y2 * Z1 * Z1Z1 % m
        H = (U2 - X1) % m
        r = (S2 - Y1) % m
        if H == 0:
            if r == 0:
                return self.jac_double(pt)
            return (0, 1, 0)
        HH = H * H % m
        HHH = H * HH % m
        V = X1 * HH % m
        X3 = (r * r - HHH - 2 * V) % m
        Y3 = (r * (V - X3) - Y1 * HHH) % m
        Z3 = Z1 * H % m
        return (X3, Y3, Z3)

    def jac_to_affine(self, pt):
        m = self.base
        X, Y, Z = pt
        if Z == 0:
            return None
        zi = pow(Z, -1, m)
        zi2 = zi * zi % m
        return (X * zi2 % m, Y * zi2 * zi % m)

    def mul(self, k: int, pt):
        k %= self.order
        acc = (0, 1, 0)
        if pt is None or k == 0:
            return None
        for bit in bin(k)[2:]:
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add_affine(acc, pt)
        return self.jac_to_affine(acc)

    def msm_naive(self, scalars: Sequence[int], points: Sequence) -> Optional[Tuple[int, int]]:
        """sum_i s_i * P_i by independent double-and-add (definition of a4)."""
        acc = None
        for s, pt in zip(scalars, points):
            acc = self.add(acc, self.mul(s, pt))
        return acc

    def msm(self, scalars: Sequence[int], points: Sequence, c: int = 8):
        """Unsigned-window Pippenger in Python ints (algorithmically unlike the CUDA path:
        unsigned digits, Jacobian buckets, per-window running sum)."""
        assert len(scalars) == len(points)
        nwin = (255 + c - 1) // c
        total = (0, 1, 0)
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jac_double(total)
            buckets = [(0, 1, 0)] * (1 << c)
            for s, pt in zip(scalars, points):
                d = ((s % self.order) >> (w * c)) & ((1 << c) - 1)
                if d:
                    buckets[d] = self.jac_add_affine(buckets[d], pt)
            run = None
            acc = None
            for d in range((1 << c) - 1, 0, -1):
                run = self.add(run, self.jac_to_affine(buckets[d]))
                acc = self.add(acc, run)
            total = self._jac_add_aff_any(total, acc)
        return self.jac_to_affine(total)

    def _jac_add_aff_any(self, pt, q):
        return self.jac_add_affine(pt, q)

    def progression(self, k0: int, d: int, n: int) -> List:
        """Known-dlog points P_i = (k0 + i*d) * G  (SURVEY 8c/8d, config C2)."""
        step = self.mul(d, self.gen)
        cur = self.mul(k0, self.gen)
        out = []
        for _ in range(n):
            out.append(cur)
            cur = self.add(cur, step)
        return out

    def msm_known_dlog(self, scalars: Sequence[int], k0: int, d: int):
        """O(N) check value for an MSM over progression(k0, d, n)."""
        acc = 0
        for i, s in enumerate(scalars):
            acc += s * (k0 + i * d)
        return self.mul(acc % self.order, self.gen)


PALLAS = Curve(CURVE_PALLAS, "pallas", P, Q)
VESTA = Curve(CURVE_VESTA, "vesta", Q, P)
CURVES = {CURVE_PALLAS: PALLAS, CURVE_VESTA: VESTA}


# ---- pasta_curves repr-c byte layouts ---------------------------
AFFINE_STRIDE = 72  # x:32, y:32, infinity:u8, 7 bytes padding
JAC_STRIDE = 96     # X, Y, Z


def affine_to_bytes(curve: Curve, pt) -> bytes:
    if pt is None:
        return bytes(64) + b"\x01" + bytes(7)
    return fe_to_bytes(pt[0], curve.base) + fe_to_bytes(pt[1], curve.base) + bytes(8)


def affines_to_bytes(curve: Curve, pts: Sequence) -> bytes:
    return b"".join(affine_to_bytes(curve, p) for p in pts)


def affine_from_bytes(curve: Curve, b: bytes):
    if b[64] != 0:
        return None
    return (fe_from_bytes(b[0:32], curve.base), fe_from_bytes(b[32:64], curve.base))


def jac_from_bytes(curve: Curve, b: bytes):
    """decode a 96-byte Jacobian point to canonical affine (or None)."""
    X = fe_from_bytes(b[0:32], curve.base)
    Y = fe_from_bytes(b[32:64], curve.base)
    Z = fe_from_bytes(b[64:96], curve.base)
    return curve.jac_to_affine((X, Y, Z))


def jac_to_bytes(curve: Curve, pt) -> bytes:
    """canonical Jacobian encoding used at the C ABI: (x, y, 1) or (0, 0, 0)."""
    if pt is None:
        return bytes(96)
    m = curve.base
    return fe_to_bytes(pt[0], m) + fe_to_bytes(pt[1], m) + fe_to_bytes(1, m)


# --------------------------------------------------------------------------------------
# MinRoot (src/minroot.rs).  Field id selects the modulus: PallasVDF works in Fq
# (minroot.rs:38), VestaVDF in Fp (minroot.rs:199).
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class State:  # src/minroot.rs:267-272
    x: int
    y: int
    i: int


class MinRootVDF:
    """Restatement of trait MinRootVDF (src/minroot.rs:287-374) over Python ints."""

    def __init__(self, field_id: int):
        self.field_id = field_id
        self.m = MODULUS[field_id]
        self.exponent = limbs_to_int(FQ_RESCUE_INVALPHA if field_id == FIELD_FQ else FP_RESCUE_INVALPHA)

    inverse_exponent = 5  # minroot.rs:68-70, :215-217

    def element(self, n: int) -> int:  # minroot.rs:60-62
        return n % self.m

    def forward_step(self, x: int) -> int:  # minroot.rs:312-314 (pow_vartime)
        return pow(x, self.exponent, self.m)

    def forward_step_addition_chain(self, x: int) -> int:
        """minroot.rs:88-127 (Fq) and :223-261 (Fp), restated to pin modulus + multiplication."""
        m = self.m

        def sqr(v, n):
            for _ in range(n):
                v = v * v % m
            return v

        def sqr_mul(v, n, y):
            return y * sqr(v, n) % m

        a1 = x
        a10 = sqr(a1, 1)
        a11 = a10 * a1 % m
        a101 = a10 * a11 % m
        a110 = sqr(a11, 1)
        a111 = a110 * a1 % m
        a1001 = a111 * a10 % m
        a1111 = a1001 * a110 % m
        r2 = sqr_mul(a110, 3, a11)
        r4 = sqr_mul(r2, 8, r2)
        r8 = sqr_mul(r4, 16, r4)
        r16 = sqr_mul(r8, 32, r8)
        r32 = sqr_mul(r16, 64, r16)
        if self.field_id == FIELD_FQ:
            tail = [(5, a1001), (8, a111), (4, a1), (2, r4), (7, a11), (6, a1001), (3, a101), (7, a101),
                    (7, a111), (4, a111), (5, a1001), (5, a101), (3, a11), (4, a101), (3, a101),
                    (6, a1111), (4, a1001), (6, a101), (37, r8), (2, a1)]
        else:
            tail = [(5, a1001), (8, a111), (4, a1), (2, r4), (7, a11), (6, a1001), (3, a101), (5, a1),
                    (7, a101), (4, a11), (8, a111), (4, a1), (4, a111), (9, a1111), (8, a1111),
                    (6, a1111), (2, a11), (34, r8), (2, a1)]
        acc = r32
        for n, y in tail:
            acc = sqr_mul(acc, n, y)
        return acc

    def inverse_step(self, x: int) -> int:  # minroot.rs:73-75, :220-222: x * (x^2)^2
        m = self.m
        x2 = x * x % m
        return x * (x2 * x2 % m) % m

    def round(self, s: State) -> State:  # minroot.rs:329-335
        m = self.m
        return State(self.forward_step((s.x + s.y) % m), (s.x + s.i) % m, (s.i + 1) % m)

    def inverse_round(self, s: State) -> State:  # minroot.rs:338-344
        m = self.m
        i = (s.i - 1) % m
        x = (s.y - i) % m
        y = (self.inverse_step(s.x) - x) % m
        return State(x, y, i)

    def eval(self, s: State, t: int) -> State:  # minroot.rs:352-359
        for _ in range(t):
            s = self.round(s)
        return s

    def inverse_eval(self, s: State, t: int) -> State:  # minroot.rs:363-365
        for _ in range(t):
            s = self.inverse_round(s)
        return s

    def check(self, result: State, t: int, original: State) -> bool:  # minroot.rs:369-371
        return original == self.inverse_eval(result, t)


def PallasVDF() -> MinRootVDF:
    return MinRootVDF(FIELD_FQ)


def VestaVDF() -> MinRootVDF:
    return MinRootVDF(FIELD_FP)


@dataclass
class Evaluation:  # src/minroot.rs:376-439
    vdf: MinRootVDF
    result: State
    t: int

    @classmethod
    def eval(cls, vdf: MinRootVDF, x: State, t: int):
        result = vdf.eval(x, t)
        return [result.x, result.y, result.i], cls(vdf, result, t)

    def verify(self, original: State) -> bool:  # :424-426
        return self.vdf.check(self.result, self.t, original)

    def append(self, other: "Evaluation") -> Optional["Evaluation"]:  # :428-438
        if other.verify(self.result):
            return Evaluation(self.vdf, other.result, self.t + other.t)
        return None


def state_to_bytes(s: State, m: int) -> bytes:
    return fe_to_bytes(s.x, m) + fe_to_bytes(s.y, m) + fe_to_bytes(s.i, m)


# --------------------------------------------------------------------------------------
# R1CS (nova-snark 0.8.0 r1cs.rs semantics).
# z = [W | u | X]; column j < num_vars -> W[j]; j == num_vars -> u ("one"); else X.
# --------------------------------------------------------------------------------------
@dataclass
class R1CSShape:
    m: int                 # scalar field modulus
    num_cons: int
    num_vars: int
    num_io: int
    A: List[Tuple[int, int, int]]  # COO (row, col, val)
    B: List[Tuple[int, int, int]]
    C: List[Tuple[int, int, int]]

    def multiply_vec(self, z: Sequence[int]):
        assert len(z) == self.num_vars + 1 + self.num_io
        out = []
        for M in (self.A, self.B, self.C):
            v = [0] * self.num_cons
            for (r, c, val) in M:
                v[r] = (v[r] + val * z[c]) % self.m
            out.append(v)
        return out

    def bind_rows(self, eq_rows: Sequence[int], r_abc: Sequence[int]) -> List[int]:
        """out[y] = sum_x eq_rows[x] (rA A[x,y] + rB B[x,y] + rC C[x,y]): the table of Spartan's inner sum-check
        (nova-snark 0.8 spartan_with_ipa_pc, compute_eval_table_sparse combined with r_A, r_B, r_C [R]; reached from
        CompressedSNARK::prove, src/nova/proof.rs:363).  A unique mathematical object: restated from the definition."""
        assert len(eq_rows) == self.num_cons and len(r_abc) == 3
        out = [0] * (self.num_vars + 1 + self.num_io)
        for M, rm in zip((self.A, self.B, self.C), r_abc):
            for (r, c, val) in M:
                out[c] = (out[c] + rm * val % self.m * eq_rows[r]) % self.m
        return out

    def z_of(self, W, u, X):
        return list(W) + [u] + list(X)

    def is_sat_relaxed(self, W, E, u, X) -> bool:
        Az, Bz, Cz = self.multiply_vec(self.z_of(W, u, X))
        return all((a * b - u * c - e) % self.m == 0 for a, b, c, e in zip(Az, Bz, Cz, E))

    def cross_term(self, W1, u1, X1, W2, X2):
        """T = Az1.Bz2 + Az2.Bz1 - u1.Cz2 - u2.Cz1 with u2 = 1 (commit_T)."""
        Az1, Bz1, Cz1 = self.multiply_vec(self.z_of(W1, u1, X1))
        Az2, Bz2, Cz2 = self.multiply_vec(self.z_of(W2, 1, X2))
        m = self.m
        return [(a1 * b2 + a2 * b1 - u1 * c2 - c1) % m
                for a1, b1, c1, a2, b2, c2 in zip(Az1, Bz1, Cz1, Az2, Bz2, Cz2)]


def fold_vec(v1: Sequence[int], v2: Sequence[int], r: int, m: int) -> List[int]:
    """v1 + r * v2  (RelaxedR1CSWitness::fold)."""
    return [(a + r * b) % m for a, b in zip(v1, v2)]


class _ShapeBuilder:
    def __init__(self, m: int, num_io: int):
        self.m = m
        self.num_io = num_io
        self.rows: List[Tuple[dict, dict, dict]] = []
        self.values: List[int] = []  # aux variable values (witness W)

    ONE = -1  # symbolic column for the constant

    def alloc(self, value: int) -> int:
        self.values.append(value % self.m)
        return len(self.values) - 1

    def enforce(self, a: dict, b: dict, c: dict):
        self.rows.append((a, b, c))

    def finish(self) -> Tuple[R1CSShape, List[int]]:
        nv = len(self.values)

        def col(v):
            return nv if v == self.ONE else (v if v >= 0 else nv + (-v - 1))

        mats = ([], [], [])
        for r, row in enumerate(self.rows):
            for k in range(3):
                for v, coeff in row[k].items():
                    coeff %= self.m
                    if coeff:
                        mats[k].append((r, col(v), coeff))
        shape = R1CSShape(self.m, len(self.rows), nv, self.num_io, *mats)
        return shape, list(self.values)


def synth_inverse_minroot(sb: _ShapeBuilder, vdf: MinRootVDF, z_in: Tuple[int, int, int], t: int):
    """InverseMinRootCircuit::synthesize + inverse_round gadget, src/nova/proof.rs:87-140,
    :155-230.  z_in are already-allocated variable ids for (x, y, i).  Allocation order per
    round: new_x, tmp1, tmp2, new_y; then final_i.  Constraint order per round: tmp1, tmp2,
    round relation; then the final_i epilogue.  3t+1 constraints, 4t+1 variables."""
    m = sb.m
    ONE = sb.ONE
    x, y, i0 = z_in
    xv, yv, iv = (sb.values[x], sb.values[y], sb.values[i0])
    i_lc = {i0: 1}  # Num<F> linear combination; value iv
    for j in range(t):
        # proof.rs:162-164  new_i = i - 1 (pure linear combination, no variable)
        new_i_lc = dict(i_lc)
        new_i_lc[ONE] = (new_i_lc.get(ONE, 0) - 1) % m
        new_iv = (iv - 1) % m
        new_x = sb.alloc(yv - new_iv)                     # proof.rs:167-173
        tmp1 = sb.alloc(xv * xv)                          # proof.rs:176
        sb.enforce({x: 1}, {x: 1}, {tmp1: 1})
        t1 = sb.values[tmp1]
        tmp2 = sb.alloc(t1 * t1)                          # proof.rs:178
        sb.enforce({tmp1: 1}, {tmp1: 1}, {tmp2: 1})
        t2 = sb.values[tmp2]
        new_y = sb.alloc(t2 * xv - sb.values[new_x])      # proof.rs:181-189
        c = {new_y: 1, y: 1}                              # proof.rs:219-227
        for v, coeff in i_lc.items():
            c[v] = (c.get(v, 0) - coeff) % m
        c[ONE] = (c.get(ONE, 0) + 1) % m
        sb.enforce({tmp2: 1}, {x: 1}, c)
        x, y = new_x, new_y
        xv, yv, iv = sb.values[new_x], sb.values[new_y], new_iv
        i_lc = new_i_lc
    final_i = sb.alloc(iv)                                # proof.rs:122-126
    sb.enforce({final_i: 1}, {ONE: 1}, dict(i_lc))        # proof.rs:128-133
    return x, y, final_i


def synth_augmented_block(sb: _ShapeBuilder, rng: XorShiftRng, n_cons: int):
    """SYNTHETIC stand-in for the NovaAugmentedCircuit part of the primary shape (SURVEY 8d
    C3: ~9.8k constraints, about half boolean witness values).  Always satisfiable."""
    m = sb.m
    ONE = sb.ONE
    pool = [sb.alloc(field_random(rng, m)) for _ in range(4)]
    while len(sb.rows) < n_cons:
        if rng.next_u32() & 1:
            b = sb.alloc(rng.next_u32() & 1)
            sb.enforce({b: 1}, {ONE: 1, b: m - 1}, {})     # b * (1 - b) = 0
            pool.append(b)
        else:
            k = 1 + rng.next_u32() % 3
            a_lc, b_lc = {}, {}
            for lc in (a_lc, b_lc):
                for _ in range(k):
                    v = pool[rng.next_u32() % len(pool)]
                    coeff = [1, m - 1, 2, field_random(rng, m)][rng.next_u32() % 4]
                    lc[v] = (lc.get(v, 0) + coeff) % m
            av = sum(sb.values[v] * c for v, c in a_lc.items()) % m
            bv = sum(sb.values[v] * c for v, c in b_lc.items()) % m
            o = sb.alloc(av * bv)
            sb.enforce(a_lc, b_lc, {o: 1})
            pool.append(o)
            if len(pool) > 64:
                pool = pool[-64:]


def make_step_instance(field_id: int, t: int, result: State, aug_cons: int = 0,
                       seed: bytes = TEST_SEED):
    """Shape + satisfying (W, X) for one Nova step over the inverse-MinRoot circuit.
    io = 2 as in nova's augmented circuit; X is synthetic (two random scalars that no
    constraint touches, as the hash outputs are only bound inside the real augmented part)."""
    vdf = MinRootVDF(field_id)
    m = vdf.m
    rng = XorShiftRng(seed)
    sb = _ShapeBuilder(m, 2)
    if aug_cons:
        synth_augmented_block(sb, rng, aug_cons)
    zin = (sb.alloc(result.x), sb.alloc(result.y), sb.alloc(result.i))
    out = synth_inverse_minroot(sb, vdf, zin, t)
    shape, W = sb.finish()
    X = [field_random(rng, m), field_random(rng, m)]
    outs = State(W[out[0]], W[out[1]], W[out[2]])
    return shape, W, X, outs


def shape_to_coo_bytes(shape: R1CSShape):
    """(rows u64[], cols u64[], vals 32B[]) x 3, the layout vdfgpu_r1cs_create takes."""
    import struct
    out = []
    for M in (shape.A, shape.B, shape.C):
        rows = struct.pack("<%dQ" % len(M), *[e[0] for e in M])
        cols = struct.pack("<%dQ" % len(M), *[e[1] for e in M])
        vals = fes_to_bytes([e[2] for e in M], shape.m)
        out.append((rows, cols, vals, len(M)))
    return out


# ---- sum-check building blocks -------------------------------------------------------------
# What CompressedSNARK::prove (src/nova/proof.rs:360-368) runs inside nova-snark 0.8's spartan_with_ipa_pc
# (sumcheck.rs, polynomial.rs).  [R]: that crate is NOT under /root/reference; these are the mathematical
# definitions of its EqPolynomial::evals, prove_cubic_with_additive_term, prove_quad, bound_poly_var_top and
# MultilinearPolynomial::evaluate, restated from memory -- PARITY UNPINNED like the rest of this file.
def eq_evals(r: Sequence[int], m: int) -> List[int]:
    """eq[idx] = prod_j (bit_j(idx) ? r_j : 1 - r_j), r[0] <-> most significant bit of idx."""
    out = [1]
    for rj in r:
        out = [v for e in out for v in (e * (1 - rj) % m, e * rj % m)]
    return out


def bind_top(P: Sequence[int], r: int, m: int) -> List[int]:
    half = len(P) // 2
    return [(P[i] + r * (P[half + i] - P[i])) % m for i in range(half)]


def sumcheck_cubic_round(A, B, C, D, m: int):
    """(e0, e2, e3) of one round with comb(A, B, C, D) = A (B C - D)."""
    half = len(A) // 2
    e = [0, 0, 0]
    for i in range(half):
        lo = (A[i], B[i], C[i], D[i])
        hi = (A[half + i], B[half + i], C[half + i], D[half + i])
        for k, t in enumerate((0, 2, 3)):
            a, b, c, d = ((l + t * (h - l)) % m for l, h in zip(lo, hi))
            e[k] = (e[k] + a * (b * c - d)) % m
    return tuple(e)


def sumcheck_quad_round(A, B, m: int):
    half = len(A) // 2
    e0 = sum(A[i] * B[i] for i in range(half)) % m
    e2 = sum((2 * A[half + i] - A[i]) * (2 * B[half + i] - B[i]) for i in range(half)) % m
    return e0, e2


def sumcheck_prove(tables: Sequence[Sequence[int]], m: int, challenge):
    """All rounds; `challenge(round, evals) -> r`.  Returns (per-round evals, challenges, final evaluations)."""
    tabs = [list(t) for t in tables]
    evals, rs = [], []
    rnd = 0
    while len(tabs[0]) > 1:
        e = sumcheck_cubic_round(*tabs, m) if len(tabs) == 4 else sumcheck_quad_round(*tabs, m)
        r = challenge(rnd, e)
        evals.append(e)
        rs.append(r)
        tabs = [bind_top(t, r, m) for t in tabs]
        rnd += 1
    return evals, rs, [t[0] for t in tabs]


def poly_evaluate(P: Sequence[int], r: Sequence[int], m: int) -> int:
    return sum(a * b for a, b in zip(eq_evals(r, m), P)) % m
