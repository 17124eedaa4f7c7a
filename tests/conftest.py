import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what is buildable here: the oracle always, the reference when its sources are
    mounted, the product library when the .so is missing (nvcc cross-compiles without a GPU)."""
    from oracle import oracle as O
    O.build()
    import gmerlin_avdecoder_b200 as g
    if not os.path.exists(g.LIB_PATH) or not os.path.exists(g.PLUGIN_PATH):
        g.build_library()
    yield


@pytest.fixture(autouse=True)
def _gpu_tests_compare_with_the_compiled_reference(request):
    """Every -m gpu test takes its expected output from oracle/_ref (the unmodified lib/RTjpeg.c compiled by
    oracle/Makefile; it travels to the GPU box with the repository).  Without it tests/streams.py would fall back to
    the restatement -- still pinned through the golden fixtures, but not what the parity claim says: fail loudly."""
    if request.node.get_closest_marker("gpu"):
        from oracle import oracle as O
        assert O.have_ref(), "oracle/_ref/librtjref.so is missing: the GPU parity tests compare with the compiled reference itself"
    yield
