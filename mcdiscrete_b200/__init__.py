"""mcdiscrete_b200 -- B200-native (sm_100a) per-electron trajectory loop of MC-discrete.

The compute path is libmcd.so (hand-written CUDA behind the C ABI in include/mcd.h); this package is
the Python host side that mirrors the reference's worker interface.  There is no CPU fallback."""
from .material import material, scatter_params            # noqa: F401
from .fileTools import readInput, thingsToSave, zipDict    # noqa: F401

__all__ = ['material', 'scatter_params', 'readInput', 'thingsToSave', 'zipDict', 'Engine', 'genTables', 'run',
           'multiTraj_DS', 'multiTraj_cont', 'scatterMultiEl_DS', 'scatterMultiEl_cont']


def run(inputPar, *, seed=None, device=None, count=None, first_id=0, ordered='columns', tables=None):
    """One call from input parameters to columnar results: builds the context on `device`, builds
    (or takes) the inverse-CDF tables, simulates `count` electrons (default inputPar['num_el']) with ids from
    `first_id` and returns a RunResult -- `els` structured array, `scats` columns in the reference's row order,
    counters.  Under torchrun every rank simulates its own id range (mcdiscrete_b200.distributed.rank_range).
    The context is closed before returning; use Engine directly to run several batches on one set of tables."""
    from .engine import Engine
    from .probTables import genTables
    from .distributed import init_from_env, rank_range
    n = int(inputPar['num_el'] if count is None else count)
    if device is None:
        rank, local, world = init_from_env()
        lo, n = rank_range(n, rank, world)
        first_id, device = first_id + lo, local
    eng = Engine(inputPar, device=device, seed=seed)
    try:
        if 'DS' in inputPar['mode']:
            if tables is None:
                genTables(inputPar, engine=eng)
            else:
                eng.table_upload(0, tables['Moller'].Es, tables['Moller'].Ws, tables['Moller'].table)
                eng.table_upload(1, tables['Gryz'][0].Es, tables['Gryz'][0].Ws, tables['Gryz'][0].table)
                eng.tables = tables
        res = eng.run(count=n, first_id=first_id, ordered=ordered)
        # detach the rows from buffers the context owns
        res.els = res.els.copy()
        return res
    finally:
        eng.close()


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
