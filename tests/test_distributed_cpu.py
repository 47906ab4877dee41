"""Multi-rank plumbing on CPU (gloo, world_size 2): the sample-split and interleaved-tile schemes of SURVEY.md §8e. Rank r renders batches
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


def _tile_worker(rank, world, port, batches, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    rb = ol.rb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = rb.configs.cornell(48, 36, samples_per_pixel=2, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    img = np.zeros((36, 48, 4), np.float32)
    for b in range(batches):          # every rank renders EVERY batch, of its own tiles only (running average)
        img, _ = sc.render_batch(48, 36, rb.RB200_FLAG_NEE, wl.push_constants(b), img, threads=1, tiles=(rank, world, 8))
    t = torch.from_numpy(img)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)          # the one collective: images are zero outside the own tiles
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_interleaved_tiles_reduce_to_the_single_process_image_bit_for_bit(tmp_path, ol, rb):
    """Latency mode: 8x8 tiles round-robin over 2 ranks, full sample count each, one reduce.
    Every pixel is accumulated by exactly one rank in batch order, so the reduced image is not merely close to the
    single-process one — it is the same bits."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "tiles.npy")
    mp.spawn(_tile_worker, args=(2, _free_port(), 3, out), nprocs=2, join=True)
    got = np.load(out)
    wl = rb.configs.cornell(48, 36, samples_per_pixel=2, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    ref = np.zeros((36, 48, 4), np.float32)
    for b in range(3):
        ref, _ = sc.render_batch(48, 36, rb.RB200_FLAG_NEE, wl.push_constants(b), ref, threads=1)
    assert (got.view(np.uint32) == ref.view(np.uint32)).all()


def test_tile_owner_rule_is_a_partition():
    """(y // tile) * tilesX + x // tile, modulo the rank count: every pixel has exactly one owner and the owners are
    spread evenly (the rule in k_generate and in the oracle)."""
    W, H, tile = 100, 70, 16
    tx = (W + tile - 1) // tile
    y, x = np.mgrid[0:H, 0:W]
    for world in (1, 2, 3, 8):
        owner = ((y // tile) * tx + x // tile) % world
        counts = np.bincount(owner.ravel(), minlength=world)
        assert counts.sum() == W * H and counts.min() > 0.6 * W * H / world


def test_bench_batch_partition_is_a_partition():
    """bench.py: rank r of N takes batches r + i*N — disjoint, and their union over ranks is 0..K*N-1."""
    for world in (1, 2, 4, 8):
        k = 5
        seen = sorted(r + i * world for r in range(world) for i in range(k))
        assert seen == list(range(k * world))
