import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # No test of this suite needs more than a couple of minutes: a hang (a deadlocked worker thread, a scheduler
    # that never reaches its fixpoint) must fail, not eat the GPU box's time.  pytest-timeout is in the image.
    if config.pluginmanager.hasplugin("timeout") and not getattr(config.option, "timeout", None):
        config.option.timeout = 300


@pytest.fixture(scope="session")
def built():
    """Native pieces are built once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def product(built):
    import leansdr_b200 as P
    P.load()
    return P


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
