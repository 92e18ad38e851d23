"""CPU tests of the C-ABI library: it loads, exports every symbol include/*.h declares, and its
host-only entry points (checkpoint reader, scheduler, PCM post-processing) behave like the
reference's caller-side code.  No compute call is made here (no GPU)."""
import json
import os
import re
import shutil
import struct

import numpy as np
import pytest

import qwen3tts_cuda as q
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = q.lib()
    hdr = open(os.path.join(ROOT, "include", "qwen3tts_cuda.h")).read()
    declared = set(re.findall(r"\b(q3tts_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert L.q3tts_abi_version() == 1
    assert set(L._q3_symbols) == declared


def test_inspect_full_checkpoint(full_dir):
    cfg = q.checkpoint_inspect(full_dir)
    assert (cfg.latent_dim, cfg.codebook_dim, cfg.decoder_dim, cfg.hidden_size) == (1024, 512, 1536, 512)
    assert cfg.total_upsample == 1920 and cfg.decode_upsample_rate == 1920 and cfg.output_sample_rate == 24000
    assert list(cfg.upsample_rates[:4]) == [8, 5, 4, 3] and list(cfg.upsampling_ratios[:2]) == [2, 2]
    assert cfg.num_decoder_tensors == 271                  # docs/paper.tex:218
    # decode-path parameters after codebook folding, without the unused input_proj (114.3 M published, docs/paper.tex:554)
    assert abs(cfg.num_parameters / 1e6 - 114.55) < 0.05
    assert cfg.has_encoder_config == 0 and cfg.sliding_window == 72


def test_config_defaults_when_keys_absent(tmp_path, tiny_dir):
    # Cfg.swift:388-408: every key is decodeIfPresent ?? default
    d = tmp_path / "st"
    shutil.copytree(tiny_dir, d)
    with open(d / "config.json", "w") as f:
        json.dump({"decoder_config": {}}, f)
    with pytest.raises(q.AudioDecodingFailed) as e:   # defaults describe the FULL model: tiny tensors mismatch
        q.checkpoint_inspect(str(d))
    assert e.value.status == 3 and "shape" in str(e.value)


def test_missing_decoder_config_is_an_error(tmp_path, tiny_dir):
    d = tmp_path / "st"
    shutil.copytree(tiny_dir, d)
    with open(d / "config.json", "w") as f:
        json.dump({"output_sample_rate": 24000}, f)
    with pytest.raises(q.AudioDecodingFailed) as e:   # ST.swift:801-805 fatalError("Decoder config is required")
        q.checkpoint_inspect(str(d))
    assert e.value.status == 3 and "decoder_config" in str(e.value)


def test_io_and_format_errors(tmp_path, tiny_dir):
    with pytest.raises(q.AudioDecodingFailed) as e:
        q.checkpoint_inspect(str(tmp_path / "nope"))
    assert e.value.status == 2
    d = tmp_path / "st"
    shutil.copytree(tiny_dir, d)
    with open(d / "model.safetensors", "r+b") as f:
        f.write(struct.pack("<Q", 1 << 40))           # absurd header length
    with pytest.raises(q.AudioDecodingFailed) as e:
        q.checkpoint_inspect(str(d))
    assert e.value.status == 3


def test_missing_tensor_is_strict(tmp_path, tiny_cfg):
    # the reference's update(verify: []) would silently accept this (Q3.swift:1486); we do not
    from safetensors.torch import load_file, save_file
    src = os.path.join(checkpoint_dir(tiny_cfg, seed=7), "speech_tokenizer")
    d = tmp_path / "st"
    shutil.copytree(src, d)
    t = load_file(str(d / "model.safetensors"))
    del t["decoder.decoder.3.block.2.conv1.conv.bias"]
    save_file(t, str(d / "model.safetensors"))
    with pytest.raises(q.AudioDecodingFailed) as e:
        q.checkpoint_inspect(str(d))
    assert e.value.status == 3 and "block2.res1.conv1.conv.bias" in str(e.value)


@pytest.mark.parametrize("kw", [dict(dtype="float16"), dict(dtype="bfloat16"), dict(with_encoder_stub=True),
                                dict(mlx_layout=True)])
def test_checkpoint_variants_parse(tiny_cfg, kw):
    # fp16-on-disk == the 'lite' variant; encoder.* tensors are ignored; already-MLX layouts
    st = os.path.join(checkpoint_dir(tiny_cfg, seed=7, **kw), "speech_tokenizer")
    cfg = q.checkpoint_inspect(st)
    assert cfg.num_decoder_tensors == 271 - 0 if tiny_cfg.num_hidden_layers == 8 else cfg.num_decoder_tensors > 0
    assert cfg.has_encoder_config == (1 if kw.get("with_encoder_stub") else 0)
    assert cfg.total_upsample == tiny_cfg.total_upsample


def test_model_load_without_gpu_fails_loudly(tiny_dir):
    if q.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(q.AudioDecodingFailed) as e:
        q.Qwen3TTSSpeechTokenizer(tiny_dir)
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)


def test_partition_lpt():
    frames = [750, 25, 400, 400, 30, 700, 100, 90]
    part = q.partition_lpt(frames, 2)
    loads = [sum(f for f, p in zip(frames, part) if p == k) for k in range(2)]
    assert sorted(set(part.tolist())) == [0, 1]
    assert abs(loads[0] - loads[1]) <= 60 and sum(loads) == sum(frames)
    assert np.array_equal(part, q.partition_lpt(frames, 2))            # deterministic
    assert q.partition_lpt([], 4).shape == (0,)
    assert q.partition_lpt([5, 5, 5], 1).tolist() == [0, 0, 0]
    one = q.partition_lpt([10, 20], 8)
    assert len(set(one.tolist())) == 2
    with pytest.raises(q.AudioDecodingFailed):
        q.partition_lpt([1, 2], 0)


def test_pcm_post_processing(tmp_path):
    # Q3.swift:746-752, 1196-1199; main.swift:134-165
    assert q.trim_length(100, 40) == 40 and q.trim_length(100, 0) == 100 and q.trim_length(100, 100) == 100
    assert q.voice_clone_cut(1, 4, 100) == 25 and q.voice_clone_cut(0, 4, 100) == 0 and q.voice_clone_cut(4, 4, 100) == 0
    pcm = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -3.0], dtype=np.float32)
    i16 = q.pcm_to_int16(pcm)
    assert i16.tolist() == [0, 16383, -16383, 32767, -32767, 32767, -32767]
    path = str(tmp_path / "a.wav")
    q.write_wav(path, pcm, 24000)
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and len(raw) == 44 + 2 * len(pcm)
    assert struct.unpack("<I", raw[24:28])[0] == 24000 and struct.unpack("<H", raw[34:36])[0] == 16
    assert np.frombuffer(raw[44:], dtype="<i2").tolist() == i16.tolist()


# ---- row N3: the encoder's checkpoint reader runs before any CUDA call, so its verdicts are visible without a GPU ----
def _encoder_status(st_dir):
    """(status, message) of q3tts_encoder_load on this machine: 4 (ECUDA, "no CUDA device") means the checkpoint was accepted."""
    try:
        q.Qwen3TTSSpeechTokenizerEncoder(st_dir).close()
        return 0, ""
    except q.AudioDecodingFailed as e:
        return e.status, str(e)


def test_encoder_checkpoint_is_read_strictly(tmp_path, tiny_cfg):
    from safetensors.torch import load_file, save_file
    from tools.q3cfg import EncoderConfig
    good = os.path.join(checkpoint_dir(tiny_cfg, seed=7, encoder_cfg=EncoderConfig.tiny()), "speech_tokenizer")
    st, msg = _encoder_status(good)
    assert st in (0, 4), msg                                     # accepted (then: no GPU here, or a working encoder)
    # a decode-only ("lite") checkpoint has no encoder: the reference throws encoderNotAvailable (SpeechTokenizer.swift:842-844)
    st, msg = _encoder_status(os.path.join(checkpoint_dir(tiny_cfg, seed=7), "speech_tokenizer"))
    assert st == 3 and "encoder_config" in msg
    # a missing / misshapen tensor is an error, not a silently random layer (the reference's update(verify: []) accepts it)
    for victim, mutate in (("encoder.encoder.layers.4.block.3.conv.bias", None),
                           ("encoder.quantizer.acoustic_residual_vector_quantizer.layers.14.codebook.embed_sum", None),
                           ("encoder.downsample.conv.weight", lambda t: t[:, :, :2].contiguous())):
        d = tmp_path / ("enc_" + victim.split(".")[-3] + "_" + victim.split(".")[-1])
        shutil.copytree(good, d)
        t = load_file(str(d / "model-encoder.safetensors"))
        if mutate is None:
            del t[victim]
        else:
            t[victim] = mutate(t[victim])
        save_file(t, str(d / "model-encoder.safetensors"))
        st, msg = _encoder_status(str(d))
        assert st == 3 and "encoder" in msg, (victim, st, msg)
    # tensors past encoder_valid_num_quantizers are never read: dropping one is fine
    d = tmp_path / "enc_deep_codebook_dropped"
    shutil.copytree(good, d)
    t = load_file(str(d / "model-encoder.safetensors"))
    del t["encoder.quantizer.acoustic_residual_vector_quantizer.layers.17.codebook.embed_sum"]
    save_file(t, str(d / "model-encoder.safetensors"))
    st, msg = _encoder_status(str(d))
    assert st in (0, 4), msg
