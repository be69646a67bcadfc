import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tp():
    import trajectory_planner_b200 as tp_
    if not os.path.exists(tp_._capi.LIB_PATH):
        tp_.build()
    return tp_


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.lib()  # builds liborc.so if missing
    return O


@pytest.fixture(scope="session")
def sq_map(tp):
    return tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "square_static.tpm"))


@pytest.fixture(scope="session")
def sq_omap(orc, sq_map):
    """Oracle map built from the raw occupied cells; the oracle does its own inflation."""
    from helpers import oracle_map_from
    return oracle_map_from(orc, sq_map)


@pytest.fixture(scope="session")
def engine(tp, sq_map):
    e = tp.Engine(0)
    e.set_map(sq_map)
    yield e
    e.close()


@pytest.fixture(scope="session")
def problems(tp, sq_map, sq_omap):
    from helpers import make_problems
    return make_problems(tp, sq_map, sq_omap, 96, seed=20261018)
