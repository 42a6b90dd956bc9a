"""Multi-rank host logic on CPU: world_size-2 gloo processes shard a batch by instance index, each solves its block
(with the oracle standing in for the GPU solve -- tests may use it), rank 0 gathers; the result must equal the
single-process solve of the whole batch.  No collective sits on the data path, only the final gather."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import skeelberzins_jl_b200 as sb
    from skeelberzins_jl_b200 import examples as ex
    from skeelberzins_jl_b200.sharding import shard_range, solve_sharded
    from oracle import sb_oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = ex.baseline_case("2", B=B, Nx=64)
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:4]
    u0 = c.initial()
    seen = {}

    def solver(prm, lo, hi):
        seen["range"] = (lo, hi)
        assert (lo, hi) == shard_range(B, world, rank)
        return sb_oracle.pdepe_transient(c.functor, c.m, c.xmesh, prm, u0[lo:hi], tg, None, nthreads=1)["u"]

    res = solve_sharded(solver, c.params, dist=dist)
    if rank == 0:
        np.save(out_path, res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_two_rank_sharded_solve_matches_single_process(tmp_path, oracle, B):
    import torch.multiprocessing as mp
    import skeelberzins_jl_b200 as sb
    from skeelberzins_jl_b200 import examples as ex
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = np.load(out)
    c = ex.baseline_case("2", B=B, Nx=64)
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    ref = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), tg[:4], None)["u"]
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_shard_ranges_partition_the_batch():
    from skeelberzins_jl_b200.sharding import shard_range
    for B in (1, 5, 8, 4096, 16384, 4097):
        for W in (1, 2, 4, 8):
            r = [shard_range(B, W, k) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) == -(-B // W)


def test_strided_shards_partition_the_batch_and_balance_a_sweep():
    """shard_indices: every instance belongs to exactly one rank, and a monotone per-instance cost (a parameter sweep) is
    spread evenly -- which contiguous blocks do not do."""
    import numpy as np
    from skeelberzins_jl_b200.sharding import shard_indices
    for B in (1, 7, 64, 4096, 4099):
        for W in (1, 2, 3, 8):
            for mode in ("strided", "contiguous"):
                idx = np.concatenate([shard_indices(B, W, k, mode) for k in range(W)])
                assert sorted(idx.tolist()) == list(range(B)), (B, W, mode)
    cost = 500.0 + 100.0 * np.arange(4096) / 4096          # cfg2: 503..604 Newton iterations across the sweep
    load = lambda mode: np.array([cost[shard_indices(4096, 8, k, mode)].sum() for k in range(8)])
    assert load("strided").max() / load("strided").mean() < 1.001
    assert load("contiguous").max() / load("contiguous").mean() > 1.05
