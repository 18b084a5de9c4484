"""CPU, world_size 2 over gloo: the host-side logic of the sharded ADMM round -- block partition of agents and the
ONE all-gather per round (scvx_b200.batch.shard_bounds / allgather_shards).  The numerical kernels are not involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scvx_b200.batch import allgather_shards, exchange_states, shard_bounds


def test_shard_bounds_cover_all_agents():
    for N in (1, 5, 16, 255, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                per, i0, i1 = shard_bounds(N, world, r)
                assert 0 <= i1 - i0 <= per
                seen += list(range(i0, i1))
            assert seen == list(range(N))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, N, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = torch.as_tensor(rng.normal(size=(N, 3, K))); U = torch.as_tensor(rng.normal(size=(N, 2, K)))
        per, i0, i1 = shard_bounds(N, world, rank)
        for _round in range(2):
            Xl = X[i0:i1] + _round; Ul = U[i0:i1] - _round
            Xa, Ua = allgather_shards(Xl, Ul, N, per, dist, None)
            ok = torch.equal(Xa, X + _round) and torch.equal(Ua, U - _round) and Xa.shape == (N, 3, K)
            # the zero-copy exchange of BatchedADMM: padded send buffer in, all shards in agent order, in place
            send = torch.zeros((per, 3, K), dtype=X.dtype); send[:i1 - i0] = Xl
            X_all = torch.full((world * per, 3, K), float("nan"), dtype=X.dtype)
            out = exchange_states(send, X_all, dist, None)
            ok = ok and out is X_all and torch.equal(X_all[:N], X + _round)
            q.put((rank, _round, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [5, 8, 1])
def test_allgather_shards_gloo_world2(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(4)]
    assert all(ok for _, _, ok in res) and len(res) == 4


def test_single_process_is_identity():
    X = torch.zeros((3, 3, 4)); U = torch.ones((3, 2, 4))
    Xa, Ua = allgather_shards(X, U, 3, 3, None, None)
    assert Xa is X and Ua is U
    X_all = torch.empty((3, 3, 4))
    assert exchange_states(X + 2.0, X_all) is X_all and torch.equal(X_all, X + 2.0)


def test_lane_bounds_of_the_pipelined_driver():
    """Lanes of PipelinedSCvx: contiguous, cover every agent once, ragged tail, never more lanes than agents."""
    import pytest
    from scvx_b200.batch import lane_bounds
    assert lane_bounds(1024, 4) == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    assert lane_bounds(7, 3) == [(0, 3), (3, 6), (6, 7)]
    assert lane_bounds(2, 8) == [(0, 1), (1, 2)]
    assert lane_bounds(5, 1) == [(0, 5)] and lane_bounds(5, 0) == [(0, 5)]
    for n in (1, 9, 100, 1000):
        for lanes in (1, 2, 3, 4, 7, 16):
            b = lane_bounds(n, lanes)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:])) and len(b) <= max(1, min(lanes, n))
    with pytest.raises(ValueError):
        lane_bounds(0, 4)
