"""Host side of the library under AddressSanitizer + UBSan: the safetensors / config.json reader with the
sanitize rules (checkpoint.cc) and the host-only C ABI (api_host.cc), fed valid checkpoints and deliberately corrupt ones.
Every input must end in a status code; any sanitizer report fails the test."""
import json
import os
import shutil
import struct
import subprocess

import pytest

from tools.fixtures import checkpoint_dir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "swift-qwen3-tts_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    r = subprocess.run(["make", "-C", CSRC, "asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(ROOT, "swift-qwen3-tts_b200", "build", "sanitize_harness")


def test_host_library_is_clean_under_asan_and_ubsan(harness, tiny_cfg, tmp_path):
    good = [os.path.join(checkpoint_dir(tiny_cfg, seed=7, **kw), "speech_tokenizer")
            for kw in (dict(), dict(dtype="float16"), dict(dtype="bfloat16"), dict(with_encoder_stub=True), dict(mlx_layout=True))]
    from tools.q3cfg import EncoderConfig
    good.append(os.path.join(checkpoint_dir(tiny_cfg, seed=7, encoder_cfg=EncoderConfig.tiny()), "speech_tokenizer"))   # a real encoder shard
    bad = []

    def variant(name):
        d = tmp_path / name
        shutil.copytree(good[0], d)
        bad.append(str(d))
        return d

    d = variant("huge_header")
    with open(d / "model.safetensors", "r+b") as f:
        f.write(struct.pack("<Q", 1 << 40))
    d = variant("truncated")
    size = os.path.getsize(d / "model.safetensors")
    with open(d / "model.safetensors", "r+b") as f:
        f.truncate(size // 2)
    d = variant("tiny_file")
    with open(d / "model.safetensors", "wb") as f:
        f.write(b"\x03\x00\x00")
    d = variant("garbage_json_header")
    with open(d / "model.safetensors", "r+b") as f:
        f.seek(8)
        f.write(b'{"a":[[[[[[' * 4)
    d = variant("offsets_past_the_end")
    with open(d / "model.safetensors", "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        hdr = json.loads(f.read(n))
        rest = f.read()
    k = next(k for k in hdr if k != "__metadata__")
    hdr[k]["data_offsets"] = [hdr[k]["data_offsets"][0], hdr[k]["data_offsets"][1] + (1 << 33)]
    blob = json.dumps(hdr).encode()
    with open(d / "model.safetensors", "wb") as f:
        f.write(struct.pack("<Q", len(blob)) + blob + rest)
    d = variant("bad_config")
    with open(d / "config.json", "w") as f:
        f.write('{"decoder_config": {"upsample_rates": [8, 5, "x"], "latent_dim": 1e99')
    d = variant("negative_dims_config")
    with open(d / "config.json", "w") as f:
        json.dump({"decoder_config": {"latent_dim": -4, "num_quantizers": 100000, "upsample_rates": [0] * 40}}, f)
    d = tmp_path / "encoder_with_a_truncated_shard"
    shutil.copytree(good[-1], d)
    bad_enc = str(d)
    with open(d / "model-encoder.safetensors", "r+b") as f:
        f.truncate(os.path.getsize(d / "model-encoder.safetensors") // 3)
    bad.append(str(tmp_path / "does_not_exist"))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1",
               Q3TTS_HARNESS_WAV=str(tmp_path / "x.wav"))
    r = subprocess.run([harness] + good + bad + [bad_enc], capture_output=True, text=True, env=env, timeout=600)
    out = r.stdout + r.stderr
    assert "AddressSanitizer" not in out and "runtime error" not in out and "LeakSanitizer" not in out, out[-4000:]
    assert r.returncode == 0, out[-4000:]
    assert f"harness: {len(good) + 1} parsed, {len(bad)} rejected" in out, out[-2000:]   # the decoder half of bad_enc is intact
    assert out.count("encoder OK") == 1 and "encoder_with_a_truncated_shard: encoder:" in out, out[-3000:]
    assert os.path.getsize(tmp_path / "x.wav") == 44 + 12
