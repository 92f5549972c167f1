"""Shared helpers for the parity tests: seeded inputs and byte marshalling (oracle side)."""
import ctypes
import random

import numpy as np

from oracle import pasta as O


def aligned(b: bytes) -> np.ndarray:
    a = np.frombuffer(bytes(b) if len(b) else bytes(16), dtype=np.uint8).copy()
    assert a.ctypes.data % 16 == 0
    return a


def ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def rand_scalars(rng: O.XorShiftRng, m: int, n: int):
    return [O.field_random(rng, m) for _ in range(n)]


def nova_like_scalars(pyrng: random.Random, rng: O.XorShiftRng, m: int, n: int):
    """Witness-like distribution: about half bits / tiny values, the rest uniform."""
    out = []
    for _ in range(n):
        k = pyrng.random()
        if k < 0.25:
            out.append(0)
        elif k < 0.5:
            out.append(1)
        elif k < 0.6:
            out.append(pyrng.randrange(1 << 16))
        else:
            out.append(O.field_random(rng, m))
    return out


def edge_field_values(m: int):
    return [0, 1, 2, m - 1, m - 2, (1 << 254), (1 << 254) - 1, (1 << 128), (1 << 128) - 1, 0xFFFFFFFF, 1 << 32,
            (m - 1) // 2, (m + 1) // 2]


from vdf_b200.encoding import known_dlog_scalar  # noqa: E402,F401  (O(n) numpy side of the known-dlog identity)
