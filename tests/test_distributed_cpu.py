"""Multi-rank plumbing on CPU (gloo, world_size 2): the sample-split scheme of SURVEY.md §8e. Rank r renders batches
r, r+N, ... with the unmodified seed formula into a local SUM image; one reduce to rank 0 gives the same image as a
single process accumulating all batches. The renderer used here is the CPU oracle (this is host-logic coverage; the
CUDA path runs the same partition in bench.py under torchrun)."""
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


def _worker(rank, world, port, total_batches, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    rb = ol.rb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = rb.configs.cornell(32, 24, samples_per_pixel=1, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    acc = np.zeros((24, 32, 4), np.float32)
    mine = list(range(rank, total_batches, world))
    for b in mine:
        acc, _ = sc.render_batch(32, 24, rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM, wl.push_constants(b), acc, threads=1)
    t = torch.from_numpy(acc)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    counts = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        np.save(out_path, t.numpy() / float(counts.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_split_equals_single_process(tmp_path, ol, rb):
    import torch.multiprocessing as mp
    total = 6
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = np.load(out)
    wl = rb.configs.cornell(32, 24, samples_per_pixel=1, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    ref = np.zeros((24, 32, 4), np.float32)
    for b in range(total):
        ref, _ = sc.render_batch(32, 24, rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM, wl.push_constants(b), ref, threads=1)
    ref = ref / float(total)
    # identical sample set, only the fp32 summation order differs (two partial sums instead of one running sum)
    assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-7)
    # and the union of the two ranks' batches is exactly the single-process batch set
    assert sorted(list(range(0, total, 2)) + list(range(1, total, 2))) == list(range(total))


def test_bench_batch_partition_is_a_partition():
    """bench.py: rank r of N takes batches r + i*N — disjoint, and their union over ranks is 0..K*N-1."""
    for world in (1, 2, 4, 8):
        k = 5
        seen = sorted(r + i * world for r in range(world) for i in range(k))
        assert seen == list(range(k * world))
