"""The reference's own test files (G-Research/ahocorasick_rs tests/test_ac.py and tests/test_ac_bytes.py, copied
byte for byte -- they are the behaviour contract of the drop-in API, not code of this repo), run UNMODIFIED against
this repo through the `ahocorasick_rs` shim package.  Every search in them goes through the CUDA library, so the whole
directory is `gpu`-marked."""
import pytest


def pytest_collection_modifyitems(config, items):
    for item in items:
        if "ref_tests" in str(item.fspath):
            item.add_marker(pytest.mark.gpu)
