#!/usr/bin/env python
"""Dump the reference's benchmark circuits as ``.npz`` programs for this backend.  RUNS WHERE tsim IS INSTALLED
(jax, stim, pyzx_param, equinox): none of them exist in this repo's build container, so nothing here imports this file.

    python tools/dump_tsim_programs.py --out programs/ [--p 1e-3] [--which distill35,distill85,surface_d5,readme]
    python bench.py --program programs/distill35_p1e-3.npz

What is built (reference file:line):
  * ``distill35``  -- the 35-qubit 5-to-1 distillation circuit: ``SteaneEncoder`` (``src/tsim/utils/encoder.py:176-207``)
    around the notebook's logical circuit (``docs/demos/magic_state_distillation.ipynb`` cell 20, with ``SteaneEncoder`` in
    place of ``ColorEncoder5``; BASELINE.json config 2), basis Z, ``p_prep = p``, gate noise ``p / 5`` (cell 23).
  * ``distill85``  -- the same with ``ColorEncoder5`` (``encoder.py:210-260``; cell 20 verbatim; config 5).
  * ``surface_d5`` -- ``stim.Circuit.generated("surface_code:rotated_memory_z", distance=5, rounds=5, ...)`` at p (config 3).
  * ``readme``     -- the README's 2-qubit T + CNOT + DEPOLARIZE2 circuit (``README.md:40-55``; config 1).
  * ``--stim-file F`` adds any other circuit (e.g. the d=3 cultivation circuit of config 4, which the reference ships only
    as a benchmark figure) under the file's stem.
Each program goes through ``Circuit.compile_detector_sampler(seed=0)`` (``compile/pipeline.py:26-102``); the dump holds the
``CompiledProgram`` (``core/types.py:80-107``) and the channel sampler's sparse tables (``noise/channels.py:578-622``), i.e.
everything ``tsim_b200`` needs to draw f vectors and sample.  A small golden batch (f vectors, key, tsim's own bits) is stored
beside it for a true end-to-end parity check: ``tests/test_program_io.py::test_golden_batch`` picks up ``*.golden.npz``.
"""

from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

THETA = float(np.arccos(np.sqrt(1 / 3)) / np.pi)  # the notebook's magic angle in units of pi

LOGICAL = """
SQRT_X 0 1 4
DEPOLARIZE1({noise}) 0 1 4
CZ 0 1 2 3
DEPOLARIZE2({noise}) 0 1 2 3
SQRT_Y 0 3
DEPOLARIZE1({noise}) 0 3
CZ 0 2 3 4
DEPOLARIZE2({noise}) 0 2 3 4
TICK
SQRT_X_DAG 0
DEPOLARIZE1({noise}) 0
CZ 0 4
DEPOLARIZE2({noise}) 0 4
TICK
CZ 1 3
DEPOLARIZE2({noise}) 1 3
TICK
SQRT_X_DAG 0 1 2 3 4
DEPOLARIZE1({noise}) 0 1 2 3 4
M 0 1 2 3 4
DETECTOR rec[-5]
DETECTOR rec[-4]
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
OBSERVABLE_INCLUDE(0) rec[-5]
OBSERVABLE_INCLUDE(1) rec[-4]
OBSERVABLE_INCLUDE(2) rec[-3]
OBSERVABLE_INCLUDE(3) rec[-2]
OBSERVABLE_INCLUDE(4) rec[-1]
"""


def distillation(encoder_cls, p: float):
    enc = encoder_cls()
    enc.initialize(f"""
        R 0 1 2 3 4
        R_X({THETA}) 0 1 2 3 4
        T_DAG 0 1 2 3 4
        DEPOLARIZE1({p}) 0 1 2 3 4
        """)
    enc.encode_transversally(LOGICAL.format(noise=p / 5))
    return enc.circuit


def build(which: str, p: float):
    import tsim

    if which == "distill35":
        from tsim.utils.encoder import SteaneEncoder

        return distillation(SteaneEncoder, p)
    if which == "distill85":
        from tsim.utils.encoder import ColorEncoder5

        return distillation(ColorEncoder5, p)
    if which == "surface_d5":
        import stim

        c = stim.Circuit.generated(
            "surface_code:rotated_memory_z", distance=5, rounds=5, after_clifford_depolarization=p,
            before_measure_flip_probability=p, after_reset_flip_probability=p,
        )
        return tsim.Circuit(str(c))
    if which == "readme":
        return tsim.Circuit("""
            RX 0
            R 1
            T 0
            PAULI_CHANNEL_1(0.1, 0.1, 0.2) 0 1
            H 0
            CNOT 0 1
            DEPOLARIZE2(0.01) 0 1
            M 0 1
            DETECTOR rec[-1] rec[-2]
            """)
    raise SystemExit(f"unknown circuit {which!r}")


def dump(circuit, name: str, out_dir: str, golden_shots: int) -> str:
    import jax
    import tsim
    import tsim.sampler as TS

    from tsim_b200.program import from_tsim, save_npz

    sampler = circuit.compile_detector_sampler(seed=0)
    prog = from_tsim(sampler._program, num_f=int(sampler._channel_sampler.signature_matrix.shape[1]))
    path = os.path.join(out_dir, name + ".npz")
    save_npz(path, prog, noise=sampler._channel_sampler,
             meta={"circuit": name, "tsim": getattr(tsim, "__version__", "?"), "repr": repr(sampler)})
    # golden batch through tsim's own sample_program (reference sampler.py:117-167): f vectors, key, bits
    f = sampler._channel_sampler.sample(golden_shots)
    key = jax.random.key(20260101)
    bits = np.asarray(TS.sample_program(sampler._program, jax.numpy.asarray(f), key))
    np.savez_compressed(os.path.join(out_dir, name + ".golden.npz"), f=np.packbits(f, axis=1, bitorder="little"),
                        num_f=np.array([f.shape[1]]), key=np.asarray(jax.random.key_data(key)), bits=np.packbits(bits, axis=1, bitorder="little"),
                        n_out=np.array([bits.shape[1]]))
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="programs")
    ap.add_argument("--p", type=float, default=1e-3)
    ap.add_argument("--which", default="distill35,distill85,surface_d5,readme")
    ap.add_argument("--stim-file", action="append", default=[])
    ap.add_argument("--golden-shots", type=int, default=4096)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    tag = f"_p{args.p:g}"
    for which in [w for w in args.which.split(",") if w]:
        print(dump(build(which, args.p), which + ("" if which == "readme" else tag), args.out, args.golden_shots))
    for path in args.stim_file:
        import tsim

        print(dump(tsim.Circuit(open(path).read()), os.path.splitext(os.path.basename(path))[0], args.out, args.golden_shots))


if __name__ == "__main__":
    main()
