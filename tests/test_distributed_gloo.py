"""The N>1 host logic on CPU: world_size-2 gloo run of the shard/gather plumbing bench.py uses.

The device call is replaced by the oracle (this is a test of the sharding contract -- shot offsets,
packed-row gather, rank-0 assembly -- not of the kernel, which `-m gpu` covers)."""

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tsim_b200.noise import ChannelSampler
from tsim_b200.shard import gather_packed_rows, pack_bool_rows, shard_range
from tsim_b200.synthetic import noise_probs, synthetic_program

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prog = synthetic_program("cfg3p_rank1")
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 5e-3), seed=9).sample(B)  # same on all ranks
    lo, hi = shard_range(B, rank, world)
    bits = oracle.sample_program(prog, f[lo:hi], (4, 2), shot_offset=lo, check_norm=False)
    rows = torch.from_numpy(pack_bool_rows(bits).view(np.int64))
    full = gather_packed_rows(rows, B, rank, world)
    if rank == 0:
        np.save(out_path, full.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_process(tmp_path):
    B, world = 1001, 2  # ragged split
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), B, out), nprocs=world, join=True)
    got = np.load(out)
    prog = synthetic_program("cfg3p_rank1")
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 5e-3), seed=9).sample(B)
    want = oracle.sample_program(prog, f, (4, 2), check_norm=False)
    assert np.array_equal(got, pack_bool_rows(want))


def test_shard_range_covers_batch():
    for B in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
