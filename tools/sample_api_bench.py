"""Time CompiledDetectorSampler.sample(shots, ...) with a device channel sampler: layout on the device vs host NumPy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tsim_b200.sampler as S
from tsim_b200.noise import DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

S.check_norm_deviations = lambda devs: None
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_distill35"
shots = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
prog = synthetic_program(name)
for dev_layout in (False, True):
    s = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=3), seed=1)
    s.DEVICE_LAYOUT = dev_layout
    for kw in (dict(bit_packed=True), dict(), dict(separate_observables=True, bit_packed=True), dict(append_observables=True, use_observable_reference_sample=True)):
        for _ in range(3):
            s.sample(shots, **kw)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            r = s.sample(shots, **kw)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        print(f"{name} device_layout={dev_layout} {kw}: {1e3*t:.3f} ms -> {shots/t:.3e} shots/s")
