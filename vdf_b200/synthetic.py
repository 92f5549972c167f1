"""Synthetic inputs for benchmarks and demos (host side, Python integers): the R1CS shape of one Nova step over
the inverse-MinRoot circuit with a satisfying witness.

The step part follows InverseMinRootCircuit::synthesize and the inverse_round gadget of the reference exactly
(src/nova/proof.rs:87-140, :155-230): per round the variables new_x, tmp1, tmp2, new_y and the constraints
tmp1 = x*x, tmp2 = tmp1*tmp1, tmp2*x = new_y + y - i + (j+1); then final_i.  The Nova augmented circuit
(bellperson, out of scope here) is replaced by a SYNTHETIC block of boolean and product constraints of about the
same size.  Nothing here is on a measured path; it only produces inputs.
"""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

from .encoding import MODULUS

Coo = List[Tuple[int, int, int]]
ONE = -1  # symbolic column of the constant 1 (z = [W | u | X])


class ShapeBuilder:
    def __init__(self, m: int, num_io: int):
        self.m, self.num_io = m, num_io
        self.rows: List[Tuple[Dict[int, int], Dict[int, int], Dict[int, int]]] = []
        self.values: List[int] = []

    def alloc(self, v: int) -> int:
        self.values.append(v % self.m)
        return len(self.values) - 1

    def enforce(self, a, b, c) -> None:
        self.rows.append((a, b, c))

    def finish(self):
        nv = len(self.values)
        mats: Tuple[Coo, Coo, Coo] = ([], [], [])
        for r, row in enumerate(self.rows):
            for k in range(3):
                for v, coeff in row[k].items():
                    coeff %= self.m
                    if coeff:
                        mats[k].append((r, nv if v == ONE else v, coeff))
        return len(self.rows), nv, self.num_io, mats[0], mats[1], mats[2], list(self.values)


def synth_augmented_block(sb: ShapeBuilder, rng: random.Random, n_cons: int) -> None:
    m = sb.m
    pool = [sb.alloc(rng.randrange(m)) for _ in range(4)]
    while len(sb.rows) < n_cons:
        if rng.random() < 0.5:
            b = sb.alloc(rng.randrange(2))
            sb.enforce({b: 1}, {ONE: 1, b: m - 1}, {})          # b * (1 - b) = 0
            pool.append(b)
        else:
            a_lc: Dict[int, int] = {}
            b_lc: Dict[int, int] = {}
            for lc in (a_lc, b_lc):
                for _ in range(1 + rng.randrange(3)):
                    v = pool[rng.randrange(len(pool))]
                    coeff = [1, m - 1, 2, rng.randrange(m)][rng.randrange(4)]
                    lc[v] = (lc.get(v, 0) + coeff) % m
            av = sum(sb.values[v] * c for v, c in a_lc.items()) % m
            bv = sum(sb.values[v] * c for v, c in b_lc.items()) % m
            o = sb.alloc(av * bv)
            sb.enforce(a_lc, b_lc, {o: 1})
            pool.append(o)
            pool = pool[-64:]


def synth_inverse_minroot(sb: ShapeBuilder, z_in: Tuple[int, int, int], t: int) -> None:
    m = sb.m
    x, y, i0 = z_in
    xv, yv, iv = sb.values[x], sb.values[y], sb.values[i0]
    i_lc: Dict[int, int] = {i0: 1}
    for _ in range(t):
        new_i_lc = dict(i_lc)
        new_i_lc[ONE] = (new_i_lc.get(ONE, 0) - 1) % m          # proof.rs:162-164
        new_iv = (iv - 1) % m
        new_x = sb.alloc(yv - new_iv)                           # proof.rs:167-173
        tmp1 = sb.alloc(xv * xv)                                # proof.rs:176
        sb.enforce({x: 1}, {x: 1}, {tmp1: 1})
        tmp2 = sb.alloc(sb.values[tmp1] ** 2)                   # proof.rs:178
        sb.enforce({tmp1: 1}, {tmp1: 1}, {tmp2: 1})
        new_y = sb.alloc(sb.values[tmp2] * xv - sb.values[new_x])  # proof.rs:181-189
        c = {new_y: 1, y: 1}                                    # proof.rs:219-227
        for v, coeff in i_lc.items():
            c[v] = (c.get(v, 0) - coeff) % m
        c[ONE] = (c.get(ONE, 0) + 1) % m
        sb.enforce({tmp2: 1}, {x: 1}, c)
        x, y = new_x, new_y
        xv, yv, iv = sb.values[new_x], sb.values[new_y], new_iv
        i_lc = new_i_lc
    final_i = sb.alloc(iv)                                      # proof.rs:122-126
    sb.enforce({final_i: 1}, {ONE: 1}, dict(i_lc))              # proof.rs:128-133


def step_instance(field_id: int, t: int, aug_cons: int, seed: int = 42):
    """(num_cons, num_vars, num_io, A, B, C, W, X): shape and satisfying (W, X) of one synthetic Nova step."""
    m = MODULUS[field_id]
    rng = random.Random(seed)
    sb = ShapeBuilder(m, 2)
    if aug_cons:
        synth_augmented_block(sb, rng, aug_cons)
    zin = (sb.alloc(rng.randrange(m)), sb.alloc(rng.randrange(m)), sb.alloc(t + 5))
    synth_inverse_minroot(sb, zin, t)
    cons, nv, io, A, B, C, W = sb.finish()
    X = [rng.randrange(m), rng.randrange(m)]
    return cons, nv, io, A, B, C, W, X
