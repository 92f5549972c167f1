#!/usr/bin/env python
"""Multi-GPU parity on real NCCL, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/gpu/dist_check.py
(1) point-range-sharded MSM: every rank commits its contiguous slice of 2^18 seeded scalars over its slice of the
    generators (k0 + i d) G, the 96-byte partials are all-gathered and summed on the GPU; the result must equal
    (sum s_i (k0 + i d)) G from Python integers on every rank;
(2) sharded batched MinRoot verification: 4 096 chains of t = 64 rounds, every 97th corrupted, sliced by chain index;
    the gathered verdicts must equal the expected pattern on every rank.
Rank 0 prints one JSON line.  oracle/ is used as the checker only (this is a test tool)."""
import json
import os
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import pasta as O  # noqa: E402
from vdf_b200 import _lib, dist as D, minroot as MR, msm as G  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.check(_lib.load().vdfgpu_init(local))

# (1) sharded MSM
cv = O.PALLAS
n, k0, d = 1 << 18, 0x1234567, 0x89ABCDEF01
py = random.Random(77)
sc = [py.randrange(cv.order) for _ in range(n)]                     # the same global vector on every rank
first, count = D.shard_range(n, rank, world)
gens = G.Generators.progression(cv.cid, k0 + first * d, d, count, table=True)
part_scalars = O.fes_to_bytes(sc[first:first + count], cv.order)
total = D.sharded_commit(lambda: gens.commit_bytes(part_scalars), lambda ps: G.point_sum(cv.cid, b"".join(ps)))
want = O.jac_to_bytes(cv, cv.mul(sum(s * (k0 + i * d) for i, s in enumerate(sc)) % cv.order, cv.gen))
msm_ok = total == want

# (2) sharded verification
vdf = MR.PallasVDF()
chains, t = 4096, 64
ovdf = O.MinRootVDF(O.FIELD_FQ)
# 64 chains through the slow direction on the host (the sequential fifth-root chain, oracle), the others the cheap
# way round: arbitrary RESULTS whose originals come from the GPU's inverse direction
head_origs = [O.State(py.randrange(vdf.m), py.randrange(vdf.m), 0) for _ in range(64)]
head_res = [ovdf.eval(o, t) for o in head_origs]
rest = [MR.State(py.randrange(vdf.m), py.randrange(vdf.m), t) for _ in range(chains - 64)]
rest_origs = vdf.inverse_eval_batch(rest, t)
results = [MR.State(r.x, r.y, r.i) for r in head_res] + rest
originals = [MR.State(o.x, o.y, o.i) for o in head_origs] + rest_origs
expect = bytearray([1] * chains)
for k in range(0, chains, 97):
    results[k] = MR.State(results[k].x, (results[k].y + 1) % vdf.m, results[k].i)
    expect[k] = 0


def check_shard(lo, cnt):
    return bytes(1 if ok else 0 for ok in vdf.check_batch(results[lo:lo + cnt], t, originals[lo:lo + cnt]))


verdicts = D.sharded_verify(check_shard, chains)
verify_ok = verdicts == bytes(expect)

flag = torch.tensor([int(msm_ok), int(verify_ok)], device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"dist_check": "ok" if int(flag.min()) == 1 else "FAILED", "world": world,
                      "sharded_msm_equals_python_integers_on_all_ranks": bool(flag[0].item()),
                      "sharded_verify_pattern_on_all_ranks": bool(flag[1].item()), "msm_points": n, "chains": chains}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if int(flag.min()) == 1 else 1)
