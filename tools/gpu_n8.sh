#!/bin/bash
# run with: gpurun --gpus N -- bash tools/gpu_n8.sh N   (short: N x the box time is charged)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout -s KILL 300 $TR tests/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
timeout -s KILL 300 $TR bench.py --gpus $N --steps 20 --warmup 3 $Q > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err; echo "peer rc=$?"; tail -3 gpurun_out/bench_n${N}_peer.err
TSIM_B200_GATHER=nccl timeout -s KILL 300 $TR bench.py --gpus $N --steps 20 --warmup 3 $Q > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "nccl rc=$?"
python - <<PY
import json
for k in ("peer","nccl"):
    try:
        d=json.load(open(f"gpurun_out/bench_n${N}_{k}.json"))
        print(k, "value %.4g ms/step %.4f memo %.4g e2e %.4g |"%(d["value"], d["ms_per_step"], d["value_memoised"], d["e2e"]["value"]), d["config"]["parallelism"])
    except Exception as e: print(k, "failed", e)
PY
