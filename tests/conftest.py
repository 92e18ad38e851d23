import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tiny_cfg():
    from tools.q3cfg import DecoderConfig
    return DecoderConfig.tiny()


@pytest.fixture(scope="session")
def full_cfg():
    from tools.q3cfg import DecoderConfig
    return DecoderConfig()


@pytest.fixture(scope="session")
def tiny_dir(tiny_cfg):
    from tools.fixtures import checkpoint_dir
    return os.path.join(checkpoint_dir(tiny_cfg, seed=7), "speech_tokenizer")


@pytest.fixture(scope="session")
def full_dir(full_cfg):
    from tools.fixtures import checkpoint_dir
    return os.path.join(checkpoint_dir(full_cfg), "speech_tokenizer")


@pytest.fixture(scope="session")
def tiny_oracle(tiny_dir):
    import torch
    from oracle import weights, decoder
    cfg, w = weights.load_decoder(tiny_dir)
    return cfg.decoder_config, w, decoder.OracleDecoder(cfg.decoder_config, w, torch.float64)


@pytest.fixture(scope="session")
def full_oracle(full_dir):
    import torch
    from oracle import weights, decoder
    cfg, w = weights.load_decoder(full_dir)
    return cfg.decoder_config, w, decoder.OracleDecoder(cfg.decoder_config, w, torch.float32)
