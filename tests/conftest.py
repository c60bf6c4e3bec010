import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    # the C-ABI library is a build artefact (git-ignored): build it in-tree when a fresh checkout has none
    # (nvcc cross-compiles sm_100a without a GPU; cached objects make this a no-op afterwards)
    so = os.path.join(ROOT, 'led-net_b200', 'libledb200.so')
    if not os.path.isfile(so):
        import importlib.util
        spec = importlib.util.spec_from_file_location('ledb200_build', os.path.join(ROOT, 'led-net_b200', 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
