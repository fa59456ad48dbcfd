"""Run under torchrun on N GPUs: the sharded sampler must reproduce the single-GPU bits.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tsim_b200.backend import DeviceProgram, split_key  # noqa: E402
from tsim_b200.distributed import ShardedDetectorSampler  # noqa: E402
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler  # noqa: E402
from tsim_b200.synthetic import noise_probs, synthetic_program  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    prog = synthetic_program("cfg2_distill35")
    tables = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 1e-3), seed=0)._sparse_data
    shots, batch = 300_001, 120_000  # ragged batches and ragged shards
    for mode in ("fast", "sliced"):
        s = ShardedDetectorSampler(prog, tables, prog.infer_num_f(), seed=11, mode=mode)
        got = s.sample_packed(shots, batch_size=batch).cpu().numpy().view(np.uint64)
        # single-GPU replay of the same schedule on this rank's own GPU
        dp = DeviceProgram(prog, device=local, mode=mode)
        noise = DeviceChannelSampler(tables, prog.infer_num_f(), seed=11, device=local)
        key, rows = (0, 11), []
        for _ in range(-(-shots // batch)):
            key, sub = split_key(key)
            rows.append(dp.sample_noisy(noise, batch, sub, packed_out=True)[0].copy())
        want = np.concatenate(rows)[:shots]
        ok = np.array_equal(got, want)
        flags = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[multi_gpu_check] world={world} mode={mode} shots={shots}: {'OK' if flags.item() else 'MISMATCH'}"
                  f" (ones fraction {np.unpackbits(got.view(np.uint8), axis=1).mean():.4f})", flush=True)
        if not flags.item():
            dist.destroy_process_group()
            sys.exit(1)
    # peer-memory gather (copy-engine pushes into symmetric memory) vs NCCL all_gather on the sampler's own output rows
    try:
        from tsim_b200.distributed import PeerGather

        n, wo = 50_000, dp.info["words_out64"]
        pg = PeerGather(n, wo, local)
        st = torch.cuda.current_stream().cuda_stream
        key = (7, 7)
        ok = True
        for i in range(5):
            key, sub = split_key(key)
            d_f = torch.from_numpy(noise.sample_packed(n, shot_offset=rank * n, call=100 + i).view(np.int64)).cuda()
            rows = pg.local_rows(i)
            dp.sample_device(d_f.data_ptr(), n, sub, rows.data_ptr(), shot_offset=rank * n, stream=st)
            done = pg.push(i)
            ref = torch.empty((world * n, wo), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(ref, rows.clone())
            torch.cuda.current_stream().wait_event(done)
            ok = ok and bool(torch.equal(pg.gathered(i), ref))
            pg.mark_consumed(i)
        flags = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[multi_gpu_check] world={world} peer-memory gather == NCCL all_gather over 5 steps: {'OK' if flags.item() else 'MISMATCH'}", flush=True)
        if not flags.item():
            dist.destroy_process_group()
            sys.exit(1)
    except Exception as exc:
        if rank == 0:
            print(f"[multi_gpu_check] peer-memory gather unavailable on this box: {exc!r}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
