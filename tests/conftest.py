import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    from minsdtf_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def unet_sd():
    from minsdtf_b200 import synth
    return synth.make_state_dict("unet")


@pytest.fixture(scope="session")
def vae_sd():
    from minsdtf_b200 import synth
    return synth.make_vae_state_dict()


@pytest.fixture(scope="session")
def cnet_sd():
    from minsdtf_b200 import synth
    return synth.make_controlnet_state_dict()


@pytest.fixture(scope="session")
def engine_unet(engine, unet_sd):
    if "unet" not in engine.loaded:
        engine.load_state_dict(unet_sd, "unet")
    return engine


@pytest.fixture(scope="session")
def engine_vae(engine, vae_sd):
    for comp in ("vae_decoder", "vae_encoder"):
        if comp not in engine.loaded:
            engine.load_state_dict(vae_sd, comp)
    return engine


@pytest.fixture(scope="session")
def engine_cnet(engine, cnet_sd):
    if "controlnet" not in engine.loaded:
        engine.load_state_dict(cnet_sd, "controlnet")
    return engine
