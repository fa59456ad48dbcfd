#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: shots/sec on the 35-qubit distillation
program at p = 1e-3), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shots S] [--program x.npz]

A *step* is one ``sample_program`` pass over one batch of S (default 10^6) shots per GPU.
``value``          : device-resident throughput of the full evaluation (every shot through the sampling kernel K1s, which
                     also transposes its f rows and assembles its output rows; pattern cache off) -- packed f rows already in HBM, CUDA events on the launch stream.
``value_memoised`` : the same batches on the library's default path (pattern cache on: light f patterns walk tabulated
                     probability trees, the rest take the full evaluation; bit-identical outputs).
``e2e``            : the same batch through the reference-facing call ``tsim_b200.sampler.sample_program`` with host
                     buffers (uint8[B, num_f] in pinned memory in, bool[B, n_out] out), copies in the timed region, full
                     evaluation; ``e2e.memoised`` / ``e2e.pageable`` are the default path and pageable input.
``roofline``       : algorithmic bytes of one step / its device time vs the measured HBM peak, plus the ceilings that bind.
``configs``        : the other BASELINE.json configurations (cfg3, cfg4, cfg5), device and end-to-end.
``parity``         : the step's bits against the C oracle on the full batch + float32 margin census (N = 1).
``cpu_baseline`` / ``--impl reference``: the C restatement of the reference's path (jax is not installable here) on all
                     host cores over a bounded sample of the same workload.
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
    """Samples SM clock and throttle reasons during the timed region (pynvml), about once per millisecond."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index: int, period_s: float = 0.001):
        self.index = index
        self.period = period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.power_w = []
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
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        i = 0
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(reasons_fn(self._h))
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
                if i % 16 == 0:
                    self.power_w.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            i += 1
            self._stop.wait(self.period)

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
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
            "sm_mhz_min": int(min(self.samples)),
            "power_w_max": float(max(self.power_w)) if self.power_w else None,
        }


def make_workload(args, seed: int = 12345):
    """-> (program, host channel sampler, sparse noise tables, workload name)."""
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    if args.program:
        from tsim_b200.program import load_npz, load_npz_noise

        prog = load_npz(args.program)
        noise = load_npz_noise(args.program)
        name = os.path.splitext(os.path.basename(args.program))[0]
        if noise is not None:
            cs = ChannelSampler.from_sparse(noise[0], noise[1], seed=seed)
            return prog, cs, name
        cs = ChannelSampler.from_bit_probs(np.full(prog.infer_num_f(), 1e-3), seed=seed)
        return prog, cs, name
    prog = synthetic_program(args.workload)
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 1e-3), seed=seed)
    return prog, cs, args.workload


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
    prog, cs, name = make_workload(args)
    f = cs.sample(sample)
    key = (0, 42)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded: the C restatement does about 3e4 shots/s/core; 2^22 shots per step on 16 cores is about 10 s
    steps = min(steps, 5)
    warmup = min(warmup, 1)
    value, cores, sec, what = cpu_reference_run(prog, f, key, steps, warmup)
    desc = f"{sample} shots per step of the {name} workload (same program, same noise model), {steps} steps; {what}"
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
        "data": "synthetic" if not args.program else "program file",
        "config": {"workload": name, "shots_per_step": sample, "note": "CPU restatement of the reference path (jax not installable here), all host cores: " + what},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


class DeviceLoop:
    """Device-resident stepping of one program: rotating input/output buffers (> L2), optional all-gather of the outputs."""

    def __init__(self, cx: Ctx, dp, cs, shots: int, gather: bool = True):
        import torch

        from tsim_b200.backend import split_key

        self.cx, self.dp, self.shots, self.split_key = cx, dp, shots, split_key
        info = dp.info
        self.wf, self.wo = info["words_f64"], info["words_out64"]
        per_step_bytes = shots * 8 * (self.wf + self.wo)
        self.per_step_bytes = per_step_bytes
        self.n_buf = min(64, max(2, int(np.ceil(160e6 / max(1, per_step_bytes)))))
        f_host = [cs.sample_packed(shots) for _ in range(min(self.n_buf, 3))]
        self.d_f = [torch.from_numpy(f_host[i % len(f_host)].view(np.int64)).to(cx.dev) for i in range(self.n_buf)]
        self.d_out = [torch.empty((shots, self.wo), dtype=torch.int64, device=cx.dev) for _ in range(self.n_buf)]
        self.f_host0 = f_host[0]
        self.gather = gather and cx.world > 1
        self.d_all, self.pg, self.gather_kind = None, None, "none"
        if self.gather:
            kind = os.environ.get("TSIM_B200_GATHER", "peer")
            if kind == "peer":
                try:
                    from tsim_b200.distributed import PeerGather

                    self.pg = PeerGather(shots, self.wo, cx.local)
                    self.gather_kind = "peer-memory copy-engine push + signal-pad barrier (no SM)"
                except Exception as exc:  # pragma: no cover - depends on the box
                    log(f"[bench] peer-memory gather unavailable ({exc!r}); using NCCL all_gather")
            if self.pg is None:
                self.d_all = [torch.empty((cx.world * shots, self.wo), dtype=torch.int64, device=cx.dev) for _ in range(2)]
                self.gather_kind = "NCCL all_gather_into_tensor"
        self.pending = []
        self.stream = torch.cuda.current_stream().cuda_stream
        self.key = (0, 42)
        self.shot_offset = cx.rank * shots  # weak scaling: the global batch is world * shots, this rank owns one slice
        self.i = 0

    def step(self):
        self.key, sub = self.split_key(self.key)
        b = self.i % self.n_buf
        if self.pg is not None:
            # the pipeline writes this rank's rows straight into its slice of the receive buffer; the push of step i runs on
            # the copy engines while step i + 1 samples
            if len(self.pending) >= 2:
                self.cx.torch.cuda.current_stream().wait_event(self.pending.pop(0))
            rows = self.pg.local_rows(self.i)
            self.dp.sample_device(self.d_f[b].data_ptr(), self.shots, sub, rows.data_ptr(), shot_offset=self.shot_offset, stream=self.stream)
            self.pending.append(self.pg.push(self.i))
            self.i += 1
            return sub, b
        self.dp.sample_device(self.d_f[b].data_ptr(), self.shots, sub, self.d_out[b].data_ptr(), shot_offset=self.shot_offset, stream=self.stream)
        if self.gather:
            if len(self.pending) >= 2:
                self.pending.pop(0).wait()  # frees the gather buffer about to be reused
            # the single gather of output bitstrings (NCCL), asynchronous w.r.t. the next step's kernel
            self.pending.append(self.cx.dist.all_gather_into_tensor(self.d_all[self.i % 2], self.d_out[b], async_op=True))
        self.i += 1
        return sub, b

    def drain(self):
        while self.pending:
            w = self.pending.pop(0)
            if self.pg is not None:
                self.cx.torch.cuda.current_stream().wait_event(w)
            else:
                w.wait()

    def timed(self, steps: int, warmup: int, clocks: ClockSampler | None = None) -> float:
        """-> total ms of `steps` steps (max over ranks), CUDA events on the launch stream."""
        torch = self.cx.torch
        for _ in range(warmup):
            self.step()
        self.drain()
        self.cx.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if clocks is not None:
            clocks.__enter__()
        ev0.record()
        for _ in range(steps):
            self.step()
        self.drain()
        ev1.record()
        torch.cuda.synchronize()
        if clocks is not None:
            clocks.__exit__(None, None, None)
        ms = self.cx.max_over_ranks(ev0.elapsed_time(ev1))
        if self.cx.world > 1:
            self.cx.dist.barrier()
        return ms

    def isolated(self, n: int = 5):
        """-> (mean ms of one call alone, mean ms of the library's own kernel events, launches per call)."""
        torch = self.cx.torch
        call_ms, lib = [], []
        for _ in range(n):
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka.record()
            self.step()
            kb.record()
            self.drain()
            torch.cuda.synchronize()
            call_ms.append(ka.elapsed_time(kb))
            lib.append(self.dp.last_kernel_ms())
        k = float(np.mean([m for m, _ in lib])) if all(m > 0 for m, _ in lib) else float(np.mean(call_ms))
        return float(np.mean(call_ms)), k, int(lib[-1][1]) if lib else 0


def time_e2e(cx: Ctx, dp, f_host: np.ndarray, steps: int) -> float:
    """Seconds per call of tsim_b200.sampler.sample_program(dp, f_host, key) (max over ranks)."""
    import tsim_b200.sampler as S
    from tsim_b200.backend import split_key

    key = (5, 5)
    for _ in range(2):
        key, sub = split_key(key)
        out_bits = S.sample_program(dp, f_host, sub)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        key, sub = split_key(key)
        out_bits = S.sample_program(dp, f_host, sub)
    dt = cx.max_over_ranks(time.perf_counter() - t0)
    assert out_bits.shape == (f_host.shape[0], dp.num_outputs)
    return dt / steps


def time_sample_api(cx: Ctx, prog, shots: int, steps: int):
    """Seconds per call (max over ranks) and D2H bytes per call of CompiledDetectorSampler.sample(shots, append_observables=True,
    bit_packed=True) with the noise generated on this rank's GPU (K5): noise, sampling, column layout and bit packing on the
    device, only the packed result crosses PCIe.  Every rank samples its own `shots` (weak scaling, rank-specific seeds)."""
    import tsim_b200.sampler as S
    from tsim_b200.noise import DeviceChannelSampler
    from tsim_b200.synthetic import noise_probs

    nz = DeviceChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=1000 + cx.rank, device=cx.local)
    smp = S.CompiledDetectorSampler(prog, nz, seed=7 + cx.rank, device=cx.local)
    kw = dict(append_observables=True, bit_packed=True)
    for _ in range(3):
        out = smp.sample(shots, **kw)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = smp.sample(shots, **kw)
    dt = cx.max_over_ranks(time.perf_counter() - t0)
    assert out.shape == (shots, (prog.num_outputs + 7) // 8)
    return dt / steps, int(out.nbytes)


def bench_config(cx: Ctx, name: str, total_shots: int, steps: int):
    """Another BASELINE.json configuration: `total_shots` per step sharded over the ranks.  -> dict for the JSON line."""
    import tsim_b200.sampler as S
    from tsim_b200.backend import DeviceProgram, PinnedArray
    from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    prog = synthetic_program(name)
    nf = prog.infer_num_f()
    q = noise_probs(nf, 1e-3)
    shots = max(1, total_shots // cx.world)
    out = {"workload": name, "shots_per_step": shots * cx.world, "shots_per_gpu": shots, "num_f": nf, "num_outputs": prog.num_outputs,
           "stabiliser_terms": int(sum(lv.num_graphs for c in prog.components for lv in c.compiled_scalar_graphs))}
    if not prog.components:
        # rank-1 (Clifford) program: every output is a direct f bit, the device work is noise sampling + gather + packing.
        # Measured through CompiledDetectorSampler.sample(bit_packed=True) with the device channel sampler (K5 -> direct gather -> column layout -> D2H).
        det = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1 + cx.rank, device=cx.local), seed=2, device=cx.local)
        for _ in range(3):  # the pinned result pool settles once two results have been alive at the same time
            res = det.sample(shots, bit_packed=True)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = det.sample(shots, bit_packed=True)
        dt = cx.max_over_ranks(time.perf_counter() - t0) / steps
        out.update({"kernel_mode": "direct (K5 noise + direct gather + column layout, all on the device)", "e2e_shots_per_s": cx.world * shots / dt,
                    "e2e_call": "CompiledDetectorSampler.sample(shots, bit_packed=True), device channel sampler",
                    "d2h_bytes_per_step": int(res.nbytes), "device_shots_per_s": None})
        return out
    cs = ChannelSampler.from_bit_probs(q, seed=12345 + cx.rank)
    dp = DeviceProgram(prog, device=cx.local, mode="auto", pattern_cache=None)
    out["kernel_mode"] = ("faithful", "fast", "sliced")[dp.info["mode"]]
    loop = DeviceLoop(cx, dp, cs, shots, gather=True)
    ms = loop.timed(steps, 3)
    out["device_shots_per_s"] = cx.world * shots * steps / (ms * 1e-3)
    out["device_ms_per_step"] = ms / steps
    n = dp.set_pattern_cache(3)
    loop2 = DeviceLoop(cx, dp, cs, shots, gather=True)
    ms = loop2.timed(steps, 3)
    out["device_memoised_shots_per_s"] = cx.world * shots * steps / (ms * 1e-3)
    out["pattern_cache_entries"] = n
    dp.set_pattern_cache(None)
    f_pin = PinnedArray((shots, nf), np.uint8)
    f_pin.array[...] = cs.sample(shots)
    out["e2e_shots_per_s"] = cx.world * shots / time_e2e(cx, dp, f_pin.array, max(1, min(steps, 3)))
    out["e2e_call"] = "tsim_b200.sampler.sample_program(program, uint8[B,num_f] pinned host, key)"
    return out


def parity_block(prog, dp, cs, shots: int):
    """One batch of the headline workload: device bits vs the C oracle on every shot, and the float32 margin census."""
    from oracle import cport

    f = cs.sample(shots)
    key = (0, 4242)
    t0 = time.perf_counter()
    got, dev = dp.sample(f, key)
    want, want_dev = cport.sample_program(prog, f, key, return_deviations=True, threads=cport.max_threads())
    c = cport.census(prog, f, key)
    return {
        "shots": int(shots),
        "draws": c["draws"],
        "bits_equal_c_oracle": bool(np.array_equal(got, want)),
        "differing_shots": int(np.count_nonzero((np.asarray(got) != want).any(axis=1))),
        "norm_dev_equal": bool(np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32))),
        "margin_tol": c["tol"],
        "margin_draws": c["margin_draws"],
        "outside_margin_flips_f64": c["outside_margin_flips_f64"],
        "margin_draws_wide": c["margin_draws_wide"],
        "margin_tol_wide": c["tol_wide"],
        "flips_alt_f32_lowering": c["flips_alt_f32_lowering"],
        "max_rel_dev_f32_vs_f64": c["max_rel_dev_f32_vs_f64"],
        "seconds": time.perf_counter() - t0,
        "note": "margin_draws = Bernoulli draws with |u - p1/prev| <= tol * p1/prev: the only draws whose bit can depend on how XLA orders or "
                "contracts the float32 tail of the approximate branch (evaluate.py:56-59); every draw outside the band has the same bit under a "
                "float64 evaluation of the same amplitudes (outside_margin_flips_f64 = 0)",
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from tsim_b200.backend import DeviceProgram, PinnedArray
    from tsim_b200.build import build
    import tsim_b200.sampler as S

    cx = Ctx(args)
    rank, world, local = cx.rank, cx.world, cx.local
    if world != args.gpus:
        log(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        from tsim_b200.distributed import nccl_defaults

        nccl_defaults()  # channel cap: the gather's CTAs fit into the SMs the sampling kernel's launch plan leaves free
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
    prog, cs, wl_name = make_workload(args)
    S.check_norm_deviations = lambda devs: None  # synthetic programs are not probability trees
    dp = DeviceProgram(prog, device=local, mode=args.mode, pattern_cache=None)
    info = dp.info
    n_out, wf, wo = info["num_outputs"], info["words_f64"], info["words_out64"]

    # ---- headline: full evaluation of every shot (pattern cache off)
    loop = DeviceLoop(cx, dp, cs, shots)
    loop.timed(0, args.warmup)  # warm-up, also brings NCCL up
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    with ClockSampler(local) as clocks_short:
        total_ms = loop.timed(args.steps, 0)
    value = world * shots * args.steps / (total_ms * 1e-3)
    step_ms_isolated, k_ms, launches_per_step = loop.isolated(5)

    # ---- sustained sub-run (>= 1.5 s of back-to-back steps): clocks under load, same kernel
    sustain_steps = args.steps if args.no_sustain else int(min(20000, max(args.steps, np.ceil(1500.0 / max(1e-3, total_ms / args.steps)))))
    clocks = ClockSampler(local)
    sustained_ms = loop.timed(sustain_steps, 0, clocks)
    sustained_value = world * shots * sustain_steps / (sustained_ms * 1e-3)

    # ---- default path: pattern cache on (same batches, bit-identical outputs)
    memo_entries = dp.set_pattern_cache(3) if info["n_draws"] > 0 else 0
    loop_m = DeviceLoop(cx, dp, cs, shots)
    memo_ms = loop_m.timed(args.steps, args.warmup)
    value_memoised = world * shots * args.steps / (memo_ms * 1e-3)
    memo_iso_ms, _, memo_launches = loop_m.isolated(5)

    # ---- e2e through the reference-facing call with host buffers (rank-local batch), four variants
    f_bytes = cs.sample(shots)
    f_pin = PinnedArray((shots, info["num_f"]), np.uint8)
    f_pin.array[...] = f_bytes
    e2e_steps = max(1, min(args.steps, 5))
    e2e_memo_pin = time_e2e(cx, dp, f_pin.array, e2e_steps)
    e2e_memo_page = time_e2e(cx, dp, f_bytes, e2e_steps)
    dp.set_pattern_cache(None)
    e2e_pin = time_e2e(cx, dp, f_pin.array, e2e_steps)
    e2e_page = time_e2e(cx, dp, f_bytes, e2e_steps)
    e2e = {
        "value": world * shots / e2e_pin,
        "unit": "shots/s",
        "h2d_bytes_per_step": int(shots * info["num_f"]),
        "d2h_bytes_per_step": int(shots * n_out),
        "steps": e2e_steps,
        "call": "tsim_b200.sampler.sample_program(program, uint8[B,num_f] pinned host, key) -> bool[B,n_out] host; full evaluation (pattern cache off)",
        "pageable": world * shots / e2e_page,
        "memoised": world * shots / e2e_memo_pin,
        "memoised_pageable": world * shots / e2e_memo_page,
        "note": "pageable = the same call on an ordinary NumPy array (what tsim's ChannelSampler hands over); memoised = library default (pattern cache on)",
    }
    try:
        api_s, api_bytes = time_sample_api(cx, prog, shots, e2e_steps)
        e2e["device_noise"] = {
            "value": world * shots / api_s,
            "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": api_bytes,
            "call": "CompiledDetectorSampler.sample(shots, append_observables=True, bit_packed=True) per rank, device channel sampler (K5), "
                    "library defaults (pattern cache on): noise, sampling, column layout and bit packing on the GPU",
        }
    except Exception as exc:  # pragma: no cover - keep the headline line alive
        e2e["device_noise"] = {"error": repr(exc)}

    # ---- the other BASELINE.json configurations
    configs = None
    if not args.no_configs and not args.program and args.workload == WORKLOAD:
        configs = []
        for name, total, st in (("cfg3_surface_d5", 10_000_000, 3), ("cfg4_cultivation_d3", 1_000_000, 3), ("cfg5_distill85", 100_000, 10)):
            try:
                configs.append(bench_config(cx, name, total, st))
            except Exception as exc:  # pragma: no cover - keep the headline line alive
                configs.append({"workload": name, "error": repr(exc)})

    # ---- extras (N = 1 only; never the headline): the full sample() API with host noise (reference ChannelSampler
    #      stream) vs noise generated on the device (K5)
    extras = None
    if world == 1 and not args.no_extras and not args.program:
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
        det_host = S.CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=1), seed=2, device=local)
        det_dev = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1, device=local), seed=2, device=local)
        extras["sample_api_host_noise_shots_per_s"] = shots / timed(lambda: det_host.sample(shots, batch_size=shots, bit_packed=True), 2)
        extras["sample_api_device_noise_shots_per_s"] = shots / timed(lambda: det_dev.sample(shots, batch_size=shots, bit_packed=True))
        extras["note"] = ("sample_api = CompiledDetectorSampler.sample(shots, bit_packed=True) including noise sampling, library defaults "
                          "(pattern cache on); host noise = the reference's exact PCG64 stream, device noise = K5 (statistical parity)")

    # ---- parity on the step's workload (N = 1, rank 0): all bits vs the C oracle + margin census
    parity = None
    if world == 1 and not args.no_cpu and not args.no_parity:
        try:
            parity = parity_block(prog, dp, cs, min(shots, 1_000_000))
        except Exception as exc:  # pragma: no cover
            parity = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline.  Algorithmic bytes of one step: g + B * 8 * (words_f + words_out), over the device time of
    #      the whole step (subkeys + K1s, which reads the f rows and writes the output rows itself), isolated launches.
    peak, peak_kind = measured_peaks()
    alg_bytes = dp.packed.g_bytes + shots * 8 * (wf + wo)
    achieved = alg_bytes / (step_ms_isolated * 1e-3) / 1e9
    traffic, ncu_pipes = None, None
    kernel_name = "sample_sliced_kernel" if info["mode"] == 2 else "sample_kernel"
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath) and wl_name == WORKLOAD and shots == 1_000_000:
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload
            entry = json.load(open(tpath)).get(kernel_name, {})
            traffic = entry.get("dram_bytes_per_launch")
            ncu_pipes = entry.get("ncu_pipes")
        except Exception:
            traffic = None
    csum = clocks.summary()
    live = None
    if ncu_pipes and csum.get("sm_mhz"):
        # live ceiling fractions: instruction / wavefront counts of the committed capture (properties of the program and the
        # batch, not of the run) over this run's kernel time and SM clock
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        cycles = k_ms * 1e-3 * csum["sm_mhz"] * 1e6
        live = {
            "issue_slot_frac": ncu_pipes["warp_instructions"] / (cycles * 4 * sms),
            "lsu_shared_wavefront_frac": ncu_pipes["shared_wavefronts"] / (cycles * sms),
            "sm_mhz": csum["sm_mhz"],
            "sms": sms,
            "how": "warp_instructions / (kernel_ms * sm_clock * 4 schedulers * SMs); shared wavefronts / (kernel_ms * sm_clock * SMs); counts from the committed ncu capture",
        }
    roofline = {
        "bound": "hbm",
        "bound_actual": "issue/lsu",
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
        "bytes_over": "one whole step (derive_subkeys + K1s with its fused input transpose and output assembly): f rows in, output rows out, g once; divided by step_ms_isolated",
        "memoised": {"achieved": alg_bytes / (memo_iso_ms * 1e-3) / 1e9, "frac": alg_bytes / (memo_iso_ms * 1e-3) / 1e9 / peak, "step_ms_isolated": memo_iso_ms},
        "live": live,
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

    cshort = clocks_short.summary()
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
        "data": "synthetic" if not args.program else "program file",
        "config": {
            "workload": wl_name,
            "shots_per_gpu_per_step": shots,
            "batch_size": shots,
            "num_f": info["num_f"],
            "num_outputs": n_out,
            "stabiliser_terms": int(sum(lv.num_graphs for c in prog.components for lv in c.compiled_scalar_graphs)),
            "kernel_mode": ("faithful", "fast", "sliced")[info["mode"]],
            "pattern_cache": "off for value / e2e.value (full evaluation of every shot); on (library default) for value_memoised / e2e.memoised",
            "g_resident_in_smem": bool(info["resident"]),
            "l2": f"inputs/outputs rotate over {loop.n_buf} buffer pairs ({loop.n_buf * loop.per_step_bytes / 1e6:.0f} MB > 126 MB L2)",
            "parallelism": f"shots sharded over {world} GPU(s), one gather of packed outputs per step: {loop.gather_kind}" if world > 1 else "single GPU",
            "synthetic_program": "shape-matched random g (tsim's compile stages need stim/pyzx_param, absent here)" if not args.program else args.program,
        },
        "clocks": {**csum, "region": f"sustained sub-run: {sustain_steps} back-to-back steps, {sustained_ms:.0f} ms", "timed_region_samples": cshort.get("samples", 0),
                   "timed_region_sm_mhz": cshort.get("sm_mhz")},
        "sustained": {"value": sustained_value, "steps": sustain_steps, "ms_per_step": sustained_ms / sustain_steps},
        "value_memoised": value_memoised,
        "memoised": {"ms_per_step": memo_ms / args.steps, "pattern_cache_entries": memo_entries, "max_weight": 3, "gpu_launches_per_step": memo_launches,
                     "note": "light f patterns (weight <= 3 per component) walk tabulated probability trees, the rest take the full evaluation; bit-identical"},
        "e2e": e2e,
        # per step: derive_subkeys + sample_sliced (input transpose and output assembly fused) + norm_check on a side stream (sliced), or
        # derive_subkeys + sample_kernel (per-row)
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
    }
    if configs is not None:
        line["configs"] = configs
    if parity is not None:
        line["parity"] = parity
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
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sustain", action="store_true", help="profiling runs: skip the >= 1.5 s sustained sub-run")
    ap.add_argument("--workload", default=WORKLOAD, help="synthetic configuration (default: the headline cfg2_distill35)")
    ap.add_argument("--program", default=None, help="a program dumped by tools/dump_tsim_programs.py (.npz) instead of a synthetic one")
    args = ap.parse_args()
    os.environ["TSIM_B200_BENCH_WORKLOAD"] = args.workload
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
