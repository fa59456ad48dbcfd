import sys; sys.path.insert(0,'/root/repo')
import tsim_b200.sampler as S
from tsim_b200.noise import DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program
S.check_norm_deviations = lambda d: None
name=sys.argv[1]; shots=int(sys.argv[2])
prog=synthetic_program(name)
s=S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=3), seed=1)
for _ in range(3): s.sample(shots, bit_packed=True)
