"""huygens_b200 — B200-native (sm_100a) filter-ensemble hot path of amcerbu/huygens.

Scope: the per-sample tick loops of Filterbank<T>, Oscbank<T,N> and RMSbank<T,N>
(/root/reference/src/{filterbank,oscbank,rmsbank}.h) behind the reference's own class surface,
with a block-process entry point.  Kernels live in huygens_b200/csrc (hand-written CUDA, C ABI in
include/huygens_cuda.h); C++ drop-in templates in include/huygens/; this package is the Python
host mirror used by the tests and bench.py.
"""
from . import design  # noqa: F401
from ._lib import HuygensError, LIB_PATH, build, lib  # noqa: F401
from .banks import Comm, Filterbank, Latchbank, Oscbank, RMSbank, Slidebank, Stickbank, chain_latency, fma_peak  # noqa: F401
