import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `pytest -m gpu` on a B200)")
    # oracle/_ref: the pieces of the reference that compile from where they lie (this container only).  Built before collection, because the
    # pinning tests are skipped when the libraries are absent (as they are on the GPU box, where /root/reference does not exist).
    if os.path.isdir("/root/reference/include/wt") and not all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)) for f in ("libref_distributions.so", "libref_traverse.so", "libref_mueller.so")):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"], check=False)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    # the product library and the oracle are built in-tree (they travel to the GPU box with the snapshot)
    from wave_tracer_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    import _oracle
    _oracle.lib()

