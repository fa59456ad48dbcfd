"""Device-resident time of one sample_device call vs batch size, sliced vs per-row records (cfg2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsim_b200.backend import DeviceProgram
from tsim_b200.synthetic import synthetic_program

prog = synthetic_program("cfg2_distill35")
dps = {m: DeviceProgram(prog, mode=m) for m in ("sliced-direct", "sliced", "fast")}
for B in (1024, 8192, 32768, 65536, 131072, 200000, 262144, 524288):
    f = torch.zeros((B, 1), dtype=torch.int64, device="cuda")
    out = torch.zeros((B, 1), dtype=torch.int64, device="cuda")
    row = [f"B={B:7d}"]
    for m, dp in dps.items():
        ts = []
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dp.sample_device(f.data_ptr(), B, (1, i), out.data_ptr())
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        row.append(f"{m} {np.median(ts[1:]):.3f} ms (K {dp.last_kernel_ms()[0]:.3f})")
    print("  ".join(row), flush=True)
