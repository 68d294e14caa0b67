"""The driver's round-end smoke entry point must keep working: run it as a test."""
import pytest

import __graft_entry__ as graft


@pytest.mark.gpu
def test_graft_smoke():
    graft.build()
    graft.smoke()
