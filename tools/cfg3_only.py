import sys, time; sys.path.insert(0,'/root/repo')
import bench, argparse
import tsim_b200.sampler as S
S.check_norm_deviations = lambda devs: None
args = argparse.Namespace(gpus=1)
cx = bench.Ctx(args)
import torch; torch.cuda.set_device(0)
print(bench.bench_config(cx, "cfg3_surface_d5", 10_000_000, 3))
from tsim_b200.noise import DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program
prog = synthetic_program("cfg3_surface_d5")
det = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(noise_probs(121), seed=1, device=0), seed=2)
for i in range(8):
    t0=time.perf_counter(); r=det.sample(10_000_000, bit_packed=True); print(i, round(1e3*(time.perf_counter()-t0),2), r.shape)
