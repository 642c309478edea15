"""CPU suite: the N > 1 path of bench.py with two processes over gloo (no GPU). The hot path shards by utterance with no collective
on the data path, so what N > 1 adds is (a) the partition of the prompt lists over the ranks - weak
scaling for the configs[2] batch, strong scaling for the configs[4] sweep - and (b) the MAX-over-ranks reduction of the timings.
Both are bench.py's own functions, run here exactly as under torchrun."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, sweep_batch, per_rank_batch, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from __graft_entry__ import load_package
    pkg = load_package()
    # (a) weak-scaling batch: every rank takes its own per_rank_batch prompts of ONE global, seeded list
    lo, hi = bench.rank_slice(rank, per_rank_batch)
    mine = pkg.synth.make_prompts(per_rank_batch * world, seed=123)[lo:hi]
    # (b) strong-scaling sweep
    sb, nb = bench.sweep_plan(total, world, sweep_batch)
    slo, shi = bench.rank_slice(rank, nb * sb)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi, slo, shi, sb, nb, mine))
    # (c) the job is as slow as its slowest rank
    slow = bench.reduce_max(10.0 + rank, world, "cpu")
    dist.barrier()
    if rank == 0:
        np.save(Path(out_dir) / "ok.npy", np.array([slow]))
        import pickle
        (Path(out_dir) / "spans.pkl").write_bytes(pickle.dumps(spans))
    dist.destroy_process_group()


@pytest.mark.parametrize("total,sweep_batch", [(8192, 1024), (1000, 256)])
def test_two_rank_sharding_and_max_reduce(tmp_path, total, sweep_batch):
    import pickle
    import torch.multiprocessing as mp
    world, per_rank_batch = 2, 16
    mp.spawn(_worker, args=(world, _free_port(), total, sweep_batch, per_rank_batch, str(tmp_path)), nprocs=world, join=True)
    assert float(np.load(tmp_path / "ok.npy")[0]) == 11.0            # MAX over ranks, not rank 0's own value
    spans = pickle.loads((tmp_path / "spans.pkl").read_bytes())
    # weak-scaling batch: disjoint, contiguous, different prompts on every rank
    assert [s[:2] for s in spans] == [(0, 16), (16, 32)]
    assert not set(spans[0][6]) & set(spans[1][6])
    # sweep: every rank the same call size and count, slices tile [0, world * nb * sb) without gaps or overlap
    sb, nb = spans[0][4], spans[0][5]
    assert all(s[4] == sb and s[5] == nb for s in spans) and sb <= sweep_batch and sb * nb * world <= total
    assert [s[2:4] for s in spans] == [(0, nb * sb), (nb * sb, 2 * nb * sb)]
