import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def _gpu_usable():
    """libmcd.so built and at least one CUDA device visible (the library itself has no CPU fallback)."""
    import ctypes
    so = os.path.join(ROOT, 'mcdiscrete_b200', 'libmcd.so')
    if not os.path.exists(so):
        return False
    try:
        n = ctypes.c_int(0)
        return ctypes.CDLL(so).mcd_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a host without a GPU: the gpu-marked tests are skipped (not failed) so that a regression in
    the CPU suites is not buried under MCD_ERR_NO_DEVICE failures.  An explicit `-m gpu` run still fails loudly."""
    if 'gpu' in (config.getoption('-m') or '') and 'not gpu' not in (config.getoption('-m') or ''):
        return
    gpu_items = [it for it in items if it.get_closest_marker('gpu')]
    if gpu_items and not _gpu_usable():
        skip = pytest.mark.skip(reason='needs a CUDA device (libmcd has no CPU fallback)')
        for it in gpu_items:
            it.add_marker(skip)
