"""CPU, gloo, world_size 2: the host-side logic of the point-range-sharded MSM: shard ranges,
the byte all-gather and the combine.  The per-rank MSM and the point sum are played by the oracle here
(no GPU in this container); on GPUs they are Generators.commit_bytes and msm.point_sum."""
import os
import random
import socket

import pytest

from oracle import pasta as O
from vdf_b200 import dist as D


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cv = O.PALLAS
        k0, d = 31, 7
        py = random.Random(11)
        sc = [py.randrange(cv.order) for _ in range(n)]          # same global scalar vector on every rank
        first, count = D.shard_range(n, rank, world)
        pts = cv.progression(k0 + first * d, d, count)           # this rank's generator slice

        def commit_shard():
            return O.jac_to_bytes(cv, cv.msm(sc[first:first + count], pts))

        def combine(parts):
            acc = None
            for p in parts:
                acc = cv.add(acc, O.jac_from_bytes(cv, p))
            return O.jac_to_bytes(cv, acc)

        total = D.sharded_commit(commit_shard, combine)
        want = O.jac_to_bytes(cv, cv.msm_known_dlog(sc, k0, d))
        gathered = D.all_gather_bytes(bytes([rank]) * 5)
        # sharded MinRoot verification: 7 chains (ragged 4 + 3), chain 5 corrupted; the oracle plays the checker
        vdf = O.MinRootVDF(O.FIELD_FQ)
        origs = [O.State(3 + k, 5 * k + 1, 0) for k in range(7)]
        results = [vdf.eval(o, 6) for o in origs]
        results[5] = O.State(results[5].x, (results[5].y + 1) % vdf.m, results[5].i)

        def check_shard(lo, cnt):
            return bytes(1 if vdf.check(results[k], 6, origs[k]) else 0 for k in range(lo, lo + cnt))

        verdicts = D.sharded_verify(check_shard, 7)
        q.put((rank, total == want, gathered == [bytes([r]) * 5 for r in range(world)] and verdicts == bytes([1, 1, 1, 1, 1, 0, 1])))
    finally:
        dist.destroy_process_group()


def test_sharded_msm_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, n = 2, 97            # ragged split: 49 + 48
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
