import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """one dfgpu context per test session; fails loudly when the CUDA library / device is missing"""
    from datafusion_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def task_ctx(gpu_ctx):
    from datafusion_b200.exec import SessionConfig, TaskContext
    return TaskContext(SessionConfig(), gpu_ctx)
