"""tsim_b200 -- B200-native backend for tsim's compiled autoregressive sampler.

Importing the package does not load CUDA; the first use of :class:`DeviceProgram` (or of the sampler
classes) loads ``libtsim_b200.so`` and fails loudly if it is missing -- there is no CPU fallback.
"""

from .program import (  # noqa: F401
    CompiledComponent,
    CompiledProgram,
    CompiledScalarGraphs,
    from_tsim,
    load_npz,
    make_program,
    make_scalar_graphs,
    save_npz,
)
from .noise import ChannelSampler, pack_f_rows  # noqa: F401
from .pack import pack_program  # noqa: F401


def __getattr__(name):
    if name in ("DeviceProgram", "PinnedArray", "split_key", "key_words"):
        from . import backend

        return getattr(backend, name)
    if name in (
        "sample_program",
        "install",
        "CompiledDetectorSampler",
        "CompiledMeasurementSampler",
        "CompiledStateProbs",
    ):
        from . import sampler

        return getattr(sampler, name)
    raise AttributeError(name)
