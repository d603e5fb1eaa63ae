import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imomd-rrtstar_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref")
    config.addinivalue_line("markers", "refsrc: also needs the reference's OSM files (/root/reference)")


def pytest_collection_modifyitems(config, items):
    import oraclelib

    have_ref = oraclelib.have_ref()
    have_src = os.path.isdir(os.path.join(REFERENCE_DIR, "osm_data"))
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="oracle/_ref not built"))
        if "refsrc" in item.keywords and not (have_ref and have_src):
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
