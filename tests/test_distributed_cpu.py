"""N > 1 host logic on CPU: two processes over the gloo backend (no GPU): the NCCL-id hand-off, the
contiguous shard partition, and the sharded error-norm arithmetic (sum of shard partials -> allreduce ->
sqrt((1/N) S)) against the unsharded CPU oracle, including the accept/reject decision it drives."""
import os
import socket

import numpy as np
import pytest

import oracle as O
from numericalnim_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unique-id hand-off: only rank 0's generator is used
        uid = D.exchange_unique_id(lambda: bytes([7 + rank]) * 128)
        # 2. shard partition + sharded error norm of one DOPRI54 attempt
        n = 10007
        rng = np.random.default_rng(5)
        y = rng.uniform(-1, 1, n)
        ks = [rng.uniform(-2, 2, n) for _ in range(7)]
        off, ln = D.shard_range(n, rank, world)
        sl = slice(off, off + ln)
        yn_loc, _, S_loc, _ = O.pair_finish("dopri54", 0.02, 1e-6, 1e-6, y[sl], [k[sl] for k in ks])
        err = D.global_error_norm(S_loc, n)
        q.put((rank, uid, off, ln, yn_loc, err))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_norm():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == bytes([7]) * 128
    # shards tile [0, n) in rank order
    assert res[0][2] == 0 and res[0][2] + res[0][3] == res[1][2] and res[1][2] + res[1][3] == 10007
    # element-wise results are shard-independent (bitwise); the norm agrees within the stated rtol
    n = 10007
    rng = np.random.default_rng(5)
    y = rng.uniform(-1, 1, n)
    ks = [rng.uniform(-2, 2, n) for _ in range(7)]
    yn, _, S, E = O.pair_finish("dopri54", 0.02, 1e-6, 1e-6, y, ks)
    assert np.array_equal(np.concatenate([res[0][4], res[1][4]]).view(np.uint64), yn.view(np.uint64))
    assert res[0][5] == res[1][5]  # every rank takes the identical controller decision
    assert abs(res[0][5] - E) <= 1e-13 * E
    assert (res[0][5] <= 1) == (E <= 1)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1000, 10007, 1 << 23, (1 << 27) + 1])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_ranges_tile_the_vector(n, world):
    pos = 0
    for r in range(world):
        off, ln = D.shard_range(n, r, world)
        assert off == pos and off % 4 == 0 or ln == 0
        pos = off + ln
    assert pos == n
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)
