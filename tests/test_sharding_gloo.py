"""N>1 path on CPU: world_size-2 `gloo` process group.  Each rank takes the frame shard
`k % world == rank` (the rule pa_exec.shard_rank/shard_count implements on the GPU), computes its partial
int64 histogram (with the oracle standing in for the CUDA kernels: this is a test), and the partials are
combined with one reduce -- the same collective bench.py issues over NCCL.  The sum must be bit-identical
to the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fixtures


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_lib as orc
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    accs = []
    for jn, refnum, rq, cnt in fx.jobs()[:3]:
        job = orc.job_setup(traj, rq)
        h, w, o = orc.histogram(traj, job, nthreads=1, shard=(rank, world))
        accs.append(np.concatenate([h, [w, o]]))          # accumulator layout of pa_session: [hist | within | oob]
    acc = torch.from_numpy(np.concatenate(accs))
    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reduce_equals_unsharded(world, oracle, tmp_path):
    name = "re4"
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    got = np.load(out)
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    exp = []
    for jn, refnum, rq, cnt in fx.jobs()[:3]:
        job = oracle.job_setup(traj, rq)
        h, w, o = oracle.histogram(traj, job)
        assert (w, o) == cnt
        exp.append(np.concatenate([h, [w, o]]))
    assert np.array_equal(got, np.concatenate(exp))
