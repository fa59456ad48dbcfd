#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: shots/sec on the 35-qubit distillation
program at p = 1e-3), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shots S]

A *step* is one ``sample_program`` pass over one batch of S (default 10^6) shots per GPU.
``value``   : device-resident throughput -- packed f rows already in HBM, CUDA events on the launch stream.
``e2e``     : the same batch through the reference-facing call ``tsim_b200.sampler.sample_program`` with
              host buffers (uint8[B, num_f] in pinned memory in, bool[B, n_out] out), copies in the timed region.
``roofline``: algorithmic bytes of the dominant kernel / its measured duration vs the measured HBM peak.
``cpu_baseline`` / ``--impl reference``: the NumPy oracle (restatement of the reference's JAX path; jax is
              not installable here) on all host cores over a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cfg2_distill35"
METRIC = "shots/sec, 35-qubit distillation-shaped program, p=1e-3"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    BAD = {
        0x8: "hw_slowdown",
        0x40: "hw_thermal_slowdown",
        0x20: "sw_thermal_slowdown",
    }
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover
            self._nv = None
            log(f"[bench] pynvml unavailable ({exc}); clocks not sampled")

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_workload(shots: int, seed: int = 12345):
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    prog = synthetic_program(os.environ.get("TSIM_B200_BENCH_WORKLOAD", WORKLOAD))
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 1e-3), seed=seed)
    return prog, cs


def cpu_reference_run(prog, f_sample, key, steps: int, warmup: int):
    """CPU restatement of the reference path on all host cores; returns (shots/s, cores, seconds per step, what ran).

    Preferred: the C restatement (oracle/c/oracle.c, validated bit for bit against the NumPy oracle) on a thread pool;
    fallback when gcc is missing: the NumPy oracle on a process pool."""
    cores = os.cpu_count() or 1
    try:
        from oracle import cport

        cport.load()
        for _ in range(max(1, warmup)):
            cport.sample_program(prog, f_sample[: max(cores * 64, 512)], key, threads=cores)
        t0 = time.perf_counter()
        for _ in range(steps):
            cport.sample_program(prog, f_sample, key, threads=cores)
        dt = time.perf_counter() - t0
        return steps * f_sample.shape[0] / dt, cores, dt / steps, "C restatement of the reference path (oracle/c/oracle.c), thread pool"
    except Exception as exc:  # pragma: no cover - depends on the box
        log(f"[bench] C oracle unavailable ({exc}); timing the NumPy oracle")
    from oracle.parallel import OraclePool

    pool = OraclePool(prog, key, cores)
    try:
        for _ in range(warmup):
            pool.sample(f_sample[: max(cores * 256, 1024)], slice_rows=256)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.sample(f_sample, slice_rows=max(256, f_sample.shape[0] // (cores * 4)))
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    return steps * f_sample.shape[0] / dt, cores, dt / steps, "NumPy restatement of the reference path (oracle/), process pool"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_shots
    prog, cs = make_workload(sample)
    f = cs.sample(sample)
    key = (0, 42)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded: the C restatement does about 1e4 shots/s/core; 2^20 shots per step on 16 cores is about 8 s
    steps = min(steps, 5)
    warmup = min(warmup, 1)
    value, cores, sec, what = cpu_reference_run(prog, f, key, steps, warmup)
    desc = f"{sample} shots per step of the {WORKLOAD} workload (same program, same noise model), {steps} steps; {what}"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "shots/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": sec * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "shots_per_step": sample, "note": "CPU restatement of the reference path (jax not installable here), all host cores: " + what},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from tsim_b200.backend import DeviceProgram, PinnedArray, split_key
    from tsim_b200.build import build
    import tsim_b200.sampler as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # The sampling kernel is one wave of CTAs that fill their SMs (896 threads x 72 registers, 227 KB of shared
        # memory); the launch plan leaves 8 SMs free.  Eight NCCL channels make the all-gather fit into those SMs so that
        # it overlaps the next step's kernel instead of queueing behind it (N = 8: 7.0e9 -> 7.7e9 shots/s).
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")
        # NCCL prints its version banner on stdout when the first communicator comes up: send fd 1 to stderr until the
        # warm-up is over, so that rank 0's stdout carries the one JSON line and nothing else
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()

    shots = args.shots
    prog, cs = make_workload(shots)
    dp = DeviceProgram(prog, device=local, mode=args.mode)
    info = dp.info
    n_out, wf, wo = info["num_outputs"], info["words_f64"], info["words_out64"]
    dev = torch.device("cuda", local)

    # ---- inputs: rotate over enough distinct buffers that the working set exceeds the 126 MB L2
    per_step_bytes = shots * 8 * (wf + wo)
    n_buf = max(2, int(np.ceil(160e6 / per_step_bytes)))
    n_buf = min(n_buf, 64)
    rng_cs = cs
    f_host = [rng_cs.sample_packed(shots) for _ in range(min(n_buf, 3))]
    d_f = [torch.from_numpy(f_host[i % len(f_host)].view(np.int64)).to(dev) for i in range(n_buf)]
    d_out = [torch.empty((shots, wo), dtype=torch.int64, device=dev) for _ in range(n_buf)]
    d_all = [torch.empty((world * shots, wo), dtype=torch.int64, device=dev) for _ in range(2)] if world > 1 else None
    pending = []  # in-flight gathers: step i's gather overlaps step i+1's kernel
    stream = torch.cuda.current_stream().cuda_stream
    key = (0, 42)
    shot_offset = rank * shots  # weak scaling: the global batch is world * shots, this rank owns one slice

    def step(i, key):
        key, sub = split_key(key)
        b = i % n_buf
        dp.sample_device(d_f[b].data_ptr(), shots, sub, d_out[b].data_ptr(), shot_offset=shot_offset, stream=stream)
        if world > 1:
            if len(pending) >= 2:
                pending.pop(0).wait()  # frees the gather buffer about to be reused
            # the single gather of output bitstrings (NCCL), asynchronous w.r.t. the next step's kernel
            pending.append(dist.all_gather_into_tensor(d_all[i % 2], d_out[b], async_op=True))
        return key

    def drain():
        while pending:
            pending.pop(0).wait()

    for i in range(args.warmup):
        key = step(i, key)
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, lib_ms = [], []
    with ClockSampler(local) as clocks:
        ev0.record()
        for i in range(args.steps):
            key = step(args.warmup + i, key)
        drain()
        ev1.record()
        torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    total_ms = float(t.item())
    value = world * shots * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel duration, measured live: a few isolated launches with events around each
    for i in range(5):
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        key, sub = split_key(key)
        b = i % n_buf
        ka.record()
        dp.sample_device(d_f[b].data_ptr(), shots, sub, d_out[b].data_ptr(), shot_offset=shot_offset, stream=stream)
        kb.record()
        torch.cuda.synchronize()
        kernel_ms.append(ka.elapsed_time(kb))
        lib_ms.append(dp.last_kernel_ms())
    step_ms_isolated = float(np.mean(kernel_ms))
    # the library's own events: around the sampling kernel alone for sliced programs (K1s), around
    # derive_subkeys + sample_kernel for per-row programs
    k_ms = float(np.mean([m for m, _ in lib_ms])) if all(m > 0 for m, _ in lib_ms) else step_ms_isolated
    launches_per_step = int(lib_ms[-1][1]) if lib_ms else 2

    # ---- e2e through the reference-facing call with host buffers (rank-local batch)
    e2e = None
    f_pin = PinnedArray((shots, info["num_f"]), np.uint8)
    f_pin.array[...] = cs.sample(shots)
    S.check_norm_deviations = lambda devs: None  # synthetic program: not a probability tree
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        key, sub = split_key(key)
        out_bits = S.sample_program(dp, f_pin.array, sub)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        key, sub = split_key(key)
        out_bits = S.sample_program(dp, f_pin.array, sub)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    assert out_bits.shape == (shots, n_out)
    e2e = {
        "value": world * shots * e2e_steps / e2e_s,
        "unit": "shots/s",
        "h2d_bytes_per_step": int(shots * info["num_f"]),
        "d2h_bytes_per_step": int(shots * n_out),
        "steps": e2e_steps,
        "call": "tsim_b200.sampler.sample_program(program, uint8[B,num_f] pinned host, key) -> bool[B,n_out] host",
    }

    # ---- extras (N = 1 only; never the headline): optional pattern cache, and the full sample() API with
    #      host noise (reference ChannelSampler stream) vs noise generated on the device (K5)
    extras = None
    if world == 1 and not args.no_extras:
        from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
        from tsim_b200.synthetic import noise_probs

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        extras = {}
        q = noise_probs(info["num_f"], 1e-3)
        det_host = S.CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=1), seed=2)
        det_dev = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1, device=local), seed=2)
        extras["sample_api_host_noise_shots_per_s"] = shots / timed(lambda: det_host.sample(shots, batch_size=shots, bit_packed=True), 2)
        extras["sample_api_device_noise_shots_per_s"] = shots / timed(lambda: det_dev.sample(shots, batch_size=shots, bit_packed=True))
        for wmax in ((0, 1, 2) if info["mode"] != 2 else ()):
            entries = dp.set_pattern_cache(wmax)
            ev = []
            for i in range(5):
                ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                key, sub = split_key(key)
                b = i % n_buf
                ka.record()
                dp.sample_device(d_f[b].data_ptr(), shots, sub, d_out[b].data_ptr(), shot_offset=shot_offset, stream=stream)
                kb.record()
                torch.cuda.synchronize()
                ev.append(ka.elapsed_time(kb))
            key, sub = split_key(key)
            t = timed(lambda: S.sample_program(dp, f_pin.array, sub))
            extras[f"pattern_cache_w{wmax}"] = {
                "table_entries": entries,
                "device_shots_per_s": shots / (float(np.mean(ev[1:])) * 1e-3),
                "e2e_shots_per_s": shots / t,
            }
        if info["mode"] != 2:
            det_dev._device_program.set_pattern_cache(2)
            extras["sample_api_device_noise_cache_w2_shots_per_s"] = shots / timed(lambda: det_dev.sample(shots, batch_size=shots, bit_packed=True))
            dp.set_pattern_cache(None)
        extras["note"] = ("bit-identical speed-ups outside the headline: pattern_cache tabulates the probability trees of light "
                          "f patterns; sample_api = CompiledDetectorSampler.sample(shots, bit_packed=True) including noise sampling")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1)
    peak, peak_kind = measured_peaks()
    alg_bytes = dp.packed.g_bytes + shots * 8 * (wf + wo)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, ncu_pipes = None, None
    kernel_name = "sample_sliced_kernel" if info["mode"] == 2 else "sample_kernel"
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath) and args.workload == WORKLOAD and shots == 1_000_000:
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload
            entry = json.load(open(tpath)).get(kernel_name, {})
            traffic = entry.get("dram_bytes_per_launch")
            ncu_pipes = entry.get("ncu_pipes")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "peak_kind": peak_kind,
        "kernel": kernel_name,
        "kernel_ms": k_ms,
        "step_ms_isolated": step_ms_isolated,
        "algorithmic_bytes": int(alg_bytes),
        # the ceilings that actually bind this kernel (issue slots, shared-memory wavefronts), from the committed ncu capture
        "ncu_pipes": ncu_pipes,
        "note": "shared-memory / issue bound by construction (about 2e4 instructions and 150 shared-memory word reads per shot vs 16 B of mandatory HBM traffic); see DESIGN.md",
    }

    # ---- CPU baseline beside it (N = 1 only): oracle on all host cores over a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        fs = cs.sample(args.cpu_shots)
        v, cores, sec, what = cpu_reference_run(prog, fs, (0, 42), 1, 1)
        cpu = {
            "value": v,
            "unit": "shots/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{args.cpu_shots} shots of the same workload, 1 warm-up + 1 timed pass ({sec:.1f} s); {what}",
        }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "shots/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": {
            "workload": args.workload,
            "shots_per_gpu_per_step": shots,
            "batch_size": shots,
            "num_f": info["num_f"],
            "num_outputs": n_out,
            "stabiliser_terms": int(sum(lv.num_graphs for c in prog.components for lv in c.compiled_scalar_graphs)),
            "kernel_mode": ("faithful", "fast", "sliced")[info["mode"]],
            "g_resident_in_smem": bool(info["resident"]),
            "l2": f"inputs/outputs rotate over {n_buf} buffer pairs ({n_buf * per_step_bytes / 1e6:.0f} MB > 126 MB L2)",
            "parallelism": f"shots sharded over {world} GPU(s), one all-gather of packed outputs per step" if world > 1 else "single GPU",
            "synthetic_program": "shape-matched random g (tsim's compile stages need stim/pyzx_param, absent here)",
        },
        "clocks": clocks.summary(),
        "e2e": e2e,
        # per step: derive_subkeys + sample_kernel (per-row) or derive_subkeys + transpose_in + sample_sliced + assemble_out
        # + norm_check on a side stream (sliced)
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if extras is not None:
        line["extras"] = extras
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--cpu-shots", type=int, default=1 << 22)
    ap.add_argument("--mode", default="auto", choices=["auto", "fast", "faithful", "sliced"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--workload", default=WORKLOAD, help="synthetic configuration (default: the headline cfg2_distill35)")
    args = ap.parse_args()
    os.environ["TSIM_B200_BENCH_WORKLOAD"] = args.workload
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
