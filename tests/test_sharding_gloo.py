"""The N > 1 path on CPU: two gloo ranks shard a sweep, solve their pairs (the CPU oracle stands in for the GPU
kernel -- this test is about the host-side plumbing), gather the result records and agree with a single rank."""
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r})
import numpy as np, torch, torch.distributed as dist
import riptrm_b200 as rb
from riptrm_b200 import sharding
from oracle.c import binding as detc
rank, world, _ = sharding.env_rank_world()
dist.init_process_group("gloo")
TOTAL, B = 11, 4
opt = {{"maxiter": 5, "tolresid": 0}}
# ragged sharding of a fixed sweep (strong scaling)
lo, hi = sharding.shard_range(TOTAL, rank, world)
Z, x0, y0 = rb.datagen.nonnegpca_batch(lo, hi - lo, 50)
_, _, sm = detc.solve_many(Z, x0, y0, opt)
allsm = sharding.gather_ragged_records(torch.from_numpy(sm), world).numpy()
# weak scaling: B pairs per rank, ids rank*B..
lo2, hi2 = sharding.weak_range(B, rank)
Z, x0, y0 = rb.datagen.nonnegpca_batch(lo2, B, 50)
_, _, sm2 = detc.solve_many(Z, x0, y0, opt)
allsm2 = sharding.gather_records(torch.from_numpy(sm2), world).numpy()
tmax = sharding.max_over_ranks(10.0 + rank, world, torch.device("cpu"))
if rank == 0:
    np.save({out!r} + "_ragged.npy", allsm)
    np.save({out!r} + "_weak.npy", allsm2)
    open({out!r} + "_tmax.txt", "w").write(str(tmax))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gloo_ranks_shard_and_gather(tmp_path):
    out = str(tmp_path / "res")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, env=env, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-3000:]
    sys.path.insert(0, REPO)
    import riptrm_b200 as rb
    from oracle.c import binding as detc
    opt = {"maxiter": 5, "tolresid": 0}
    Z, x0, y0 = rb.datagen.nonnegpca_batch(0, 11, 50)
    _, _, sm = detc.solve_many(Z, x0, y0, opt)
    assert np.array_equal(np.load(out + "_ragged.npy"), sm)          # same records, same (global id) order
    Z, x0, y0 = rb.datagen.nonnegpca_batch(0, 8, 50)
    _, _, sm = detc.solve_many(Z, x0, y0, opt)
    assert np.array_equal(np.load(out + "_weak.npy"), sm)
    assert float(open(out + "_tmax.txt").read()) == 11.0              # max over ranks


def test_shard_range_covers_everything():
    from riptrm_b200 import sharding
    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
