import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libwipa.so and the C oracle exist (compiles them when absent; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from whisper_ipa_b200 import _lib
    if not (os.path.exists(_lib.LIB_PATH) and os.path.exists(_lib.LIB_PATH_BF16)):
        g.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def tiny_sd():
    from oracle import hf_reference as hf
    return hf.state_dict_f32(hf.build_hf_model("tiny", seed=0))


@pytest.fixture(scope="session")
def tiny_gain_sd():
    from oracle import hf_reference as hf
    return hf.state_dict_f32(hf.build_hf_model("tiny", seed=0, init_gain=3.0))
