import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dna-sequences-pg-extension_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref():
    """The CPU oracle (oracle/ref_cpu.c) -- the checker."""
    from oracle import ref_cpu
    ref_cpu.lib()
    return ref_cpu


@pytest.fixture(scope="session")
def gpu():
    """One libdnagpu context on cuda:0.  Fails (not skips) when the library or GPU is missing."""
    import dnagpu
    ctx = dnagpu.Context(0)
    yield ctx
    ctx.close()
