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


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (`-m "not gpu"`) is host work only -- oracle runs, g++ builds of the emulated kernels,
    ThreadSanitizer binaries -- and every test is independent of the others: spread it over worker processes when
    pytest-xdist is in the image and the caller did not choose (`-n ...`, `-p no:xdist`, LDVB_TESTS_SERIAL=1).
    The GPU suite is left as the caller runs it (one process unless `-n` is given)."""
    opt = config.option
    if (os.environ.get("PYTEST_XDIST_WORKER") or os.environ.get("LDVB_TESTS_SERIAL")
            or not config.pluginmanager.hasplugin("xdist") or getattr(opt, "numprocesses", None) is not None
            or getattr(opt, "collectonly", False) or getattr(opt, "usepdb", False)
            or "not gpu" not in (getattr(opt, "markexpr", "") or "")):
        return None
    n = min(6, max(1, (os.cpu_count() or 1) - 1))
    if n > 1:
        opt.numprocesses = n
        opt.dist = "load"
    return None


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
