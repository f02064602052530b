"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding covers every batch
exactly once, and the MAX / SUM merge of the collector state is exact and order-independent.
The CUDA kernels are not involved: the per-rank state tensors are filled by hand."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TESTS = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(TESTS), "pytorch-quantity_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common.quantity import DistributionCollector
    from tools.pytorch_quantizer import Quantity, _dist_info
    assert _dist_info() == (rank, world)

    # --- sharding: batches 0..MAX round-robin, nothing else touched
    q = object.__new__(Quantity)
    q._max_img_num, q.rank, q.world_size = 6, rank, world
    mine = [i for i, _ in q._my_batches(["b%d" % i for i in range(10)])]
    np.save(os.path.join(out_dir, "mine%d.npy" % rank), np.array(mine))

    # --- merge: every rank holds partial maxima (as fp32 bit patterns) and partial counts
    names = ["a", "b", "c"]
    col = DistributionCollector(names, device="cpu")
    rng = np.random.default_rng(100 + rank)
    maxima = rng.random(3).astype(np.float32) * (10.0 ** rng.integers(-3, 3, size=3)).astype(np.float32)
    if rank == 1:
        maxima[2] = 0.0                                  # a tensor this rank never saw
    col._max_bits = torch.from_numpy(maxima.view(np.int32).copy())
    col._hist = torch.from_numpy(rng.integers(0, 2 ** 40, size=(3, 2048), dtype=np.int64))
    np.save(os.path.join(out_dir, "max%d.npy" % rank), maxima)
    np.save(os.path.join(out_dir, "hist%d.npy" % rank), col._hist.numpy().copy())
    col.all_reduce_max()
    col.all_reduce_hist()
    mv = col.max_vals
    np.save(os.path.join(out_dir, "merged_max%d.npy" % rank), np.array([float(mv[n]) for n in names]))
    np.save(os.path.join(out_dir, "merged_hist%d.npy" % rank), col._hist.numpy())
    assert col.distributions["a"].dtype == np.int64      # counts above 2^31 are not narrowed
    # --- per-channel maxima (extension): one MAX all-reduce per tensor on the bit patterns
    chan = rng.random((2, 5)).astype(np.float32)
    col._chan_bits = {"a": torch.from_numpy(chan[0].view(np.int32).copy()),
                      "c": torch.from_numpy(chan[1].view(np.int32).copy())}
    np.save(os.path.join(out_dir, "chan%d.npy" % rank), chan)
    col.all_reduce_channel_max()
    np.save(os.path.join(out_dir, "merged_chan%d.npy" % rank),
            np.stack([col.channel_max_vals["a"], col.channel_max_vals["c"]]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mine = [np.load(tmp_path / ("mine%d.npy" % r)).tolist() for r in range(world)]
    assert sorted(mine[0] + mine[1]) == list(range(7))            # batches 0..MAX_CALI_IMG_NUM, once each
    assert mine[0] == [0, 2, 4, 6] and mine[1] == [1, 3, 5]
    maxima = np.stack([np.load(tmp_path / ("max%d.npy" % r)) for r in range(world)])
    hists = np.stack([np.load(tmp_path / ("hist%d.npy" % r)) for r in range(world)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("merged_max%d.npy" % r)),
                              maxima.max(axis=0).astype(np.float64))        # integer MAX on bit patterns == float max
        assert np.array_equal(np.load(tmp_path / ("merged_hist%d.npy" % r)), hists.sum(axis=0))
    chans = np.stack([np.load(tmp_path / ("chan%d.npy" % r)) for r in range(world)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("merged_chan%d.npy" % r)), chans.max(axis=0))
