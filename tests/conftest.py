import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_sessionstart(session):
    """The engine library is a build artefact (git-ignored): compile it when it is missing or older
    than its sources, so that a fresh checkout can run the suite (host-side entry points such as the
    PDB parser are used by CPU tests too).  A failure here is reported by the tests that need it."""
    try:
        from pyjess_b200 import _build
        _build.build()
    except Exception as err:  # pragma: no cover
        print(f"[conftest] could not build libjess_b200.so: {err}", file=sys.stderr)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine context on cuda:0 (gpu tests only)."""
    from pyjess_b200 import _capi
    return _capi.default_context()


@pytest.fixture(autouse=True)
def _device_memory_log(request):
    """JB_MEMLOG=<file>: free device memory after every test (leak hunting on the GPU box)."""
    yield
    path = os.environ.get("JB_MEMLOG")
    if path:
        import torch
        free, total = torch.cuda.mem_get_info()
        with open(path, "a") as f:
            f.write(f"{(total - free) / 2**20:10.0f} MiB used after {request.node.nodeid}\n")
