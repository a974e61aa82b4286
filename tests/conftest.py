import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
  """Make sure libwavenet_b200.so exists (nvcc cross-compiles on CPU-only boxes)."""
  from wavenets_b200 import build
  build.build()
  yield
