"""mcdiscrete_b200 -- B200-native (sm_100a) per-electron trajectory loop of MC-discrete.

The compute path is libmcd.so (hand-written CUDA behind the C ABI in include/mcd.h); this package is
the Python host side that mirrors the reference's worker interface.  There is no CPU fallback."""
from .material import material, scatter_params            # noqa: F401
from .fileTools import readInput, thingsToSave, zipDict    # noqa: F401

__all__ = ['material', 'scatter_params', 'readInput', 'thingsToSave', 'zipDict', 'Engine', 'genTables',
           'multiTraj_DS', 'multiTraj_cont', 'scatterMultiEl_DS', 'scatterMultiEl_cont']


def __getattr__(name):
    # the CUDA-backed pieces are imported lazily so that `import mcdiscrete_b200` works for host-only use
    # (parsers, materials); touching any compute entry point without libmcd.so raises ImportError.
    if name == 'Engine':
        from .engine import Engine
        return Engine
    if name == 'genTables':
        from .probTables import genTables
        return genTables
    if name in ('multiTraj_DS', 'multiTraj_cont', 'scatterMultiEl_DS', 'scatterMultiEl_cont'):
        from . import multiScatter
        return getattr(multiScatter, name)
    raise AttributeError(name)
