import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATAFILES = os.path.join(GOLDEN, "datafiles")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The oracle's C part is compiled on demand (gcc only)."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    # the CUDA library too (nvcc cross-compiles without a GPU): the C-ABI symbol test loads it on CPU
    from igdetective_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()


@pytest.fixture(autouse=True)
def _gpu_tests_need_a_device(request):
    """A test marked `gpu` on a machine without any CUDA device (plain `pytest tests` in the CPU container) is skipped,
    not failed; on a GPU box the count is >= 1 and nothing is skipped."""
    if request.node.get_closest_marker("gpu") is not None:
        from igdetective_b200 import _lib
        if _lib.load().igd_device_count() == 0:
            pytest.skip("no CUDA device visible: GPU tests need a B200")


@pytest.fixture(scope="session")
def motifs():
    from oracle import igd_oracle
    return igd_oracle.load_motifs_text(os.path.join(DATAFILES, "motif_sets.txt"))


@pytest.fixture(scope="session")
def engine(motifs):
    """One Engine per session; creating it fails loudly when the CUDA library or a B200 is missing."""
    from igdetective_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):        # a fresh checkout: the test harness builds, the product never does
        from igdetective_b200 import build
        build.build()
    from igdetective_b200.engine import Engine
    if _lib.load().igd_device_count() == 0:
        # a machine without any CUDA device (the CPU container): `pytest tests` reports skips, not errors.  On a GPU box
        # the count is >= 1 and everything below fails loudly if the library cannot drive the device.
        pytest.skip("no CUDA device visible: GPU tests need a B200 (run with -m 'not gpu' here)")
    e = Engine(0, motifs)
    yield e
    e.close()


def golden_input(name):
    return os.path.join(GOLDEN, "inputs", name + ".fa.gz")


def canonical(gene_dir, locus="IGH"):
    from oracle import igd_oracle
    return {g: igd_oracle.fasta_dict(os.path.join(DATAFILES, gene_dir, "%s%s.fa.gz" % (locus, g)), upper=True)
            for g in ("V", "J")}
