import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def built_lib():
    """The in-tree libssdcodec.so (built on demand; nvcc cross-compiles without a GPU)."""
    from jpeg_detection_resnet_ssd_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def ctx(built_lib):
    import jpeg_detection_resnet_ssd_b200 as pkg
    return pkg.get_context()
