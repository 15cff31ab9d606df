import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine_10g_r50():
    """SCRFD-10G + ArcFace R50 engine (R50 keeps the CPU oracle side of the tests fast)."""
    from person_capture_b200.engine import Engine
    eng = Engine(0, scrfd="scrfd_10g_bnkps", arcface="arcface_r50")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_25g_r50():
    from person_capture_b200.engine import Engine
    eng = Engine(0, scrfd="scrfd_2.5g_bnkps", arcface="arcface_r50")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_10g_r100():
    """SCRFD-10G + ArcFace R100: the models BASELINE configs 2-5 (and bench.py) are quoted on."""
    from person_capture_b200.engine import Engine
    eng = Engine(0, scrfd="scrfd_10g_bnkps", arcface="arcface_r100")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engines_val():
    """Engines on libpcb200_val.so (product sources + the CUDA-core validation convolution, conv impl 1), built lazily per
    (scrfd, arcface) pair: the product library itself does not contain that kernel."""
    from person_capture_b200 import _lib
    from person_capture_b200.engine import Engine
    made = {}

    def get(scrfd, arcface):
        key = (scrfd, arcface)
        if key not in made:
            made[key] = Engine(0, scrfd=scrfd, arcface=arcface, lib_path=_lib.VAL_LIB_PATH)
        return made[key]

    yield get
    for eng in made.values():
        eng.close()
