"""World-size-2 gloo test of the shard / gather host logic (the N>1 path of bench.py and distributed.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from temporal_inverse_kinematics_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeModel:
    """Stands in for the CUDA model on CPU: poses = a deterministic function of each clip."""

    def __call__(self, x):
        return {"poses": x.reshape(x.shape[0], -1)[:, :6].reshape(x.shape[0], 1, 6) * 2.0 + 1.0}


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_total, 4, 17, 3, generator=g)
        got = D.solve_sharded(_FakeModel(), x)
        want = _FakeModel()(x)["poses"]
        q.put((rank, bool(torch.equal(got, want)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 65536):
        for w in (1, 2, 3, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_two_rank_gloo_gather_even_and_ragged():
    for n_total in (8, 7):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(ok for _, ok, _ in res) and all(shape == (n_total, 1, 6) for _, _, shape in res)
