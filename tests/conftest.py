import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def small_case():
    """3 genomes x 20 kb (one a strain copy), 30 simulated reads + the hand-shaped edge reads."""
    from monica_b200 import synth
    names, seqs = synth.make_genomes(11, 3, 60000, strain_frac=0.34)
    reads, truth = synth.simulate_reads(12, seqs, 30, 2500, 0.10, junk_frac=0.05)
    reads = synth.edge_reads(13, seqs) + reads
    return names, seqs, reads


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "monica", "genomes"))
