import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def example_reads():
    """The reference's only fixture, committed as tests/golden/sample.fasta (Example/sample.fasta)."""
    path = os.path.join(ROOT, "tests", "golden", "sample.fasta")
    out, name, buf = [], None, []
    for line in open(path, "rb"):
        line = line.strip()
        if line.startswith(b">"):
            if name is not None:
                out.append((name.decode(), b"".join(buf)))
            name, buf = line[1:], []
        elif line:
            buf.append(line)
    out.append((name.decode(), b"".join(buf)))
    return out
