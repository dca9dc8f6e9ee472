import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library, (re)built when stale.  No GPU needed to build or load."""
    from pytorch_news_recommender_b200 import _build, _lib
    _build.build()
    return _lib.load()
