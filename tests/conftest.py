import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """The C ABI library and the oracle must exist (built by __graft_entry__.build())."""
    from meng_zhang_b200 import capi
    if not os.path.isfile(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.lib()


@pytest.fixture(scope="session")
def fe_pot_file(tmp_path_factory, built):
    import util
    return util.write_fe_potential(tmp_path_factory.mktemp("pot") / "fe_annp_potential.ann")


@pytest.fixture(scope="session")
def ni_pot_file(tmp_path_factory, built):
    import util
    return util.write_ni_potential(tmp_path_factory.mktemp("pot") / "ni_annp_potential.ann")


@pytest.fixture(scope="session")
def anna_pot_file(tmp_path_factory, built):
    import util
    return util.write_anna_fe_potential(tmp_path_factory.mktemp("pot") / "fe_adp_potential.anna")
