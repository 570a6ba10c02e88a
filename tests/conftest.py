import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # the two in-tree libraries are build products (git-ignored): build them when a fresh checkout has none (nvcc + g++, a few minutes)
    pkg = os.path.join(ROOT, 'cacto_b200')
    if not (os.path.exists(os.path.join(pkg, 'libcacto_b200.so')) and os.path.exists(os.path.join(pkg, 'libcacto_b200_torch.so'))):
        sys.path.insert(0, pkg)
        import importlib.util
        spec = importlib.util.spec_from_file_location('_cacto_b200_build', os.path.join(pkg, 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
