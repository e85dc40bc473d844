import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly (not skip) when the CUDA library or device is missing.
    pass


@pytest.fixture(scope="session")
def tiny_lm():
    """Tiny config weights shared by the LM tests (oracle finishes in seconds)."""
    from fish_speech_rs_b200 import synth
    cfg, tok = dict(synth.TINY), dict(synth.TINY_TOKENS)
    return cfg, tok, synth.make_lm_weights(cfg, seed=1234)


@pytest.fixture(scope="session")
def codec_weights():
    from fish_speech_rs_b200 import synth
    return synth.make_codec_weights(seed=4321, with_encoder=True)
