"""Host-side mirror of the reference's speech-tokenizer API over the C ABI (ctypes).

The reference is Swift (no toolchain in this image), so the host layer above
``libqwen3tts_cuda.so`` is mirrored here in Python with the reference's names and argument
meaning, so that parity tests read like the reference's own test
(Tests/Qwen3TTSTests/Qwen3TTSTests.swift:25-283):

  Qwen3TTSSpeechTokenizer.decode(audio_codes [B,T,16]) -> (audio [B,1920*T], audio_lengths [B])
      -- Sources/Qwen3TTS/Models/SpeechTokenizer.swift:823-836
  Qwen3TTSSpeechTokenizerDecoder.__call__(codes [B,16,T]) -> [B,1,1920*T]
      -- SpeechTokenizer.swift:754-784
  .has_encoder -- SpeechTokenizer.swift:816

There is NO CPU fallback: if the shared library is missing or no sm_100 GPU is usable, loading a
model raises.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

__all__ = ["Qwen3TTSSpeechTokenizer", "Qwen3TTSDecoderPool", "DecodeStream", "Qwen3TTSSpeechTokenizerDecoder", "AudioDecodingFailed", "lib",
           "partition_lpt", "PREC_FP32", "PREC_FP16", "PREC_BF16", "ATTN_REFERENCE", "ATTN_CAUSAL_SW",
           "library_path", "checkpoint_inspect", "CodecEmbedder", "pcm_to_int16", "write_wav", "trim_length",
           "voice_clone_cut", "device_count"]

PREC_FP32, PREC_FP16, PREC_BF16 = 0, 1, 2
ATTN_REFERENCE, ATTN_CAUSAL_SW = 0, 1
CODES_BQT, CODES_BTQ = 0, 1
_STATUS = {0: "OK", 1: "EINVAL", 2: "EIO", 3: "EFORMAT", 4: "ECUDA", 5: "ENOMEM", 6: "ESTATE"}


class AudioDecodingFailed(RuntimeError):
    """Maps a non-zero C status, like the Swift shim maps it to
    AudioGenerationError.audioDecodingFailed(String) (Core/GenerationTypes.swift:67)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[{_STATUS.get(status, status)}] {message}")
        self.status = status


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("precision", C.c_int32),
                ("attn_mode", C.c_int32), ("workspace_bytes", C.c_uint64),
                ("max_frames_per_launch", C.c_int32), ("reserved", C.c_int32)]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "latent_dim", "codebook_dim", "codebook_size", "decoder_dim", "hidden_size", "intermediate_size",
        "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "head_dim", "sliding_window",
        "num_quantizers", "num_semantic_quantizers", "semantic_codebook_size", "num_upsample_rates")] + [
        ("upsample_rates", C.c_int32 * 8), ("num_upsampling_ratios", C.c_int32),
        ("upsampling_ratios", C.c_int32 * 8), ("total_upsample", C.c_int32),
        ("decode_upsample_rate", C.c_int32), ("output_sample_rate", C.c_int32),
        ("has_encoder_config", C.c_int32), ("rms_norm_eps", C.c_float), ("rope_theta", C.c_float),
        ("layer_scale_initial_scale", C.c_float), ("num_decoder_tensors", C.c_int64),
        ("num_parameters", C.c_int64)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("launches", C.c_int32),
                ("flops", C.c_double), ("bytes", C.c_double)]


class StageTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_float), ("launches", C.c_int32),
                ("flops", C.c_double), ("bytes", C.c_double)]


def library_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    return os.environ.get("QWEN3TTS_CUDA_LIB") or os.path.normpath(
        os.path.join(here, "..", "..", "lib", "libqwen3tts_cuda.so"))


_lib = None


def lib() -> C.CDLL:
    """Load libqwen3tts_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise OSError(f"{path} not found: build it with __graft_entry__.build() "
                      f"(make -C swift-qwen3-tts_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p
    sigs = {
        "q3tts_abi_version": (C.c_int, []),
        "q3tts_last_error": (cp, []),
        "q3tts_device_count": (C.c_int, []),
        "q3tts_options_default": (None, [C.POINTER(Options)]),
        "q3tts_model_load": (C.c_int, [cp, C.POINTER(Options), C.POINTER(vp)]),
        "q3tts_model_free": (None, [vp]),
        "q3tts_model_config": (C.c_int, [vp, C.POINTER(Config)]),
        "q3tts_checkpoint_inspect": (C.c_int, [cp, C.POINTER(Config)]),
        "q3tts_output_samples": (i64, [vp, i64]),
        "q3tts_decode": (C.c_int, [vp, vp, i32, i32, i32, vp, vp]),
        "q3tts_decode_varlen": (C.c_int, [vp, vp, vp, i32, vp, vp]),
        "q3tts_decode_int16": (C.c_int, [vp, vp, i32, i32, i32, vp, vp]),
        "q3tts_decode_varlen_int16": (C.c_int, [vp, vp, vp, i32, vp, vp]),
        "q3tts_decode_device": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp]),
        "q3tts_decode_varlen_device": (C.c_int, [vp, vp, vp, i32, vp, vp, vp]),
        "q3tts_codec_embedder_load": (C.c_int, [cp, i32, C.POINTER(vp)]),
        "q3tts_codec_embedder_free": (None, [vp]),
        "q3tts_codec_embedder_info": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp]),
        "q3tts_codec_embed_sum": (C.c_int, [vp, vp, i64, vp]),
        "q3tts_codec_embed_sum_device": (C.c_int, [vp, vp, i64, vp, vp]),
        "q3tts_sync": (C.c_int, [vp, vp]),
        "q3tts_set_taps": (C.c_int, [vp, i32]),
        "q3tts_set_graphs": (C.c_int, [vp, i32]),
        "q3tts_stage_tap_shape": (C.c_int, [vp, cp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "q3tts_stage_tap": (C.c_int, [vp, cp, vp, i64]),
        "q3tts_weight_shape": (C.c_int, [vp, cp, C.POINTER(i32), C.POINTER(i64 * 4)]),
        "q3tts_stream_open": (C.c_int, [vp, C.POINTER(vp)]),
        "q3tts_stream_push": (C.c_int, [vp, vp, i32, vp]),
        "q3tts_stream_push_batch": (C.c_int, [vp, i32, vp, vp, vp]),
        "q3tts_stream_frames": (i64, [vp]),
        "q3tts_stream_close": (None, [vp]),
        "q3tts_partition_lpt": (C.c_int, [vp, i32, i32, vp]),
        "q3tts_trim_length": (i64, [i64, i64]),
        "q3tts_voice_clone_cut": (i64, [i64, i64, i64]),
        "q3tts_pcm_to_int16": (C.c_int, [vp, i64, vp]),
        "q3tts_write_wav": (C.c_int, [cp, vp, i64, i32]),
        "q3tts_profile_enable": (C.c_int, [vp, i32]),
        "q3tts_profile_get": (C.c_int, [vp, C.POINTER(StageTime), i32]),
        "q3tts_profile_kernels": (C.c_int, [vp, C.POINTER(KernelTime), i32]),
        "q3tts_launch_count": (i64, [vp]),
        "q3tts_debug_fused_unit": (C.c_int, [i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             C.POINTER(C.c_float)]),
        "q3tts_debug_resunit": (C.c_int, [i32, i32, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "q3tts_debug_conv_gemm": (C.c_int, [i32, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "q3tts_debug_attention": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp]),
        "q3tts_pool_open": (C.c_int, [cp, C.POINTER(Options), vp, i32, C.POINTER(vp)]),
        "q3tts_pool_close": (None, [vp]),
        "q3tts_pool_size": (i32, [vp]),
        "q3tts_pool_decode_varlen": (C.c_int, [vp, vp, vp, i32, vp, vp]),
        "q3tts_pool_decode_varlen_int16": (C.c_int, [vp, vp, vp, i32, vp, vp]),
        "q3tts_pool_decode": (C.c_int, [vp, vp, i32, i32, i32, vp, vp]),
        "q3tts_pool_last_stats": (C.c_int, [vp, vp, vp, i32]),
        "q3tts_encoder_load": (C.c_int, [cp, C.POINTER(Options), C.POINTER(vp)]),
        "q3tts_encoder_free": (None, [vp]),
        "q3tts_encoder_info": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
        "q3tts_encode_frames": (i64, [vp, i64]),
        "q3tts_encode": (C.c_int, [vp, vp, i32, i64, vp]),
        "q3tts_encoder_launch_count": (i64, [vp]),
        "q3tts_encoder_set_taps": (C.c_int, [vp, i32]),
        "q3tts_encoder_tap": (C.c_int, [vp, cp, vp, i64, C.POINTER(i64 * 3)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._q3_symbols = tuple(sigs)
    _lib = L
    return L


class DecodeStream:
    """One chunked-decode stream (causal state lives on the model's GPU)."""

    def __init__(self, owner, handle):
        self._owner, self._h = owner, handle

    def push(self, codes: np.ndarray) -> np.ndarray:
        """codes [n,16] int32 -> PCM [n*1920] float32 for exactly these frames."""
        ac = np.ascontiguousarray(codes, dtype=np.int32).reshape(-1, self._owner.config.num_quantizers)
        out = np.empty(ac.shape[0] * self._owner.config.total_upsample, dtype=np.float32)
        _check(lib().q3tts_stream_push(self._h, ac.ctypes.data, ac.shape[0], out.ctypes.data))
        return out

    @property
    def frames(self) -> int:
        return int(lib().q3tts_stream_frames(self._h))

    def close(self) -> None:
        if self._h:
            lib().q3tts_stream_close(self._h)
            self._h = None


def _check(status: int) -> None:
    if status != 0:
        raise AudioDecodingFailed(status, lib().q3tts_last_error().decode("utf-8", "replace"))


def debug_conv_gemm(B: int, rows: int, Cin: int, N: int, taps: int, dil: int, mode: int, precision: int = PREC_FP16,
                    iters: int = 0) -> Tuple[float, float, float]:
    """Kernel-level check: (ms per launch, max|tc - simt| stream, max|tc - simt| operand) of one multi-tap GEMM."""
    ms, dy, da = C.c_float(0), C.c_float(0), C.c_float(0)
    _check(lib().q3tts_debug_conv_gemm(B, rows, Cin, N, taps, dil, mode, precision, iters, C.byref(ms), C.byref(dy), C.byref(da)))
    return float(ms.value), float(dy.value), float(da.value)


def debug_fused_unit(B: int, rows: int, Cch: int, dil: int, with_operand: int = 1, precision: int = PREC_FP16,
                     iters: int = 0) -> Tuple[float, float, float]:
    """Kernel-level check of the GEMM-fused residual unit: (ms, max|diff| stream, max|diff| operand)."""
    ms, dy, da = C.c_float(0), C.c_float(0), C.c_float(0)
    _check(lib().q3tts_debug_fused_unit(B, rows, Cch, dil, with_operand, precision, iters, C.byref(ms), C.byref(dy), C.byref(da)))
    return float(ms.value), float(dy.value), float(da.value)


def debug_resunit(B: int, rows: int, dil: int, out_snake: int = 0, precision: int = PREC_FP16, iters: int = 0) -> Tuple[float, float]:
    """Kernel-level check of the fused residual unit: (ms per launch, max |fused - composed|)."""
    ms, d = C.c_float(0), C.c_float(0)
    _check(lib().q3tts_debug_resunit(B, rows, dil, out_snake, precision, iters, C.byref(ms), C.byref(d)))
    return float(ms.value), float(d.value)


def debug_attention(qkv: np.ndarray, nh: int, nkv: int, hd: int, lens=None, row_begin=None, window: int = 0,
                    precision: int = PREC_FP16) -> np.ndarray:
    """The attention kernel alone (production dispatch): qkv [B,T,(nh+2nkv)*hd] float32 -> [B,T,nh*hd] float32."""
    a = np.ascontiguousarray(qkv, dtype=np.float32)
    B, T, ld = a.shape
    assert ld == (nh + 2 * nkv) * hd
    out = np.zeros((B, T, nh * hd), dtype=np.float32)
    ln = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
    rb = None if row_begin is None else np.ascontiguousarray(row_begin, dtype=np.int32)
    _check(lib().q3tts_debug_attention(a.ctypes.data, B, T, nh, nkv, hd, ln.ctypes.data if ln is not None else None,
                                       rb.ctypes.data if rb is not None else None, window, precision, out.ctypes.data))
    return out


class CodecEmbedder:
    """Codec-embedding sum for the Talker's next-step input (Qwen3.swift:720-728, 485-491): 16 table rows per frame,
    added left to right in the tables' dtype.  `model_dir` holds the main checkpoint's *.safetensors."""

    def __init__(self, model_dir: str, device: int = 0):
        h = C.c_void_p()
        _check(lib().q3tts_codec_embedder_load(model_dir.encode(), device, C.byref(h)))
        self._h = h
        hid, grp, prec = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        vocab = (C.c_int32 * 32)()
        _check(lib().q3tts_codec_embedder_info(self._h, C.byref(hid), C.byref(grp), C.byref(prec), vocab))
        self.hidden, self.groups, self.precision = int(hid.value), int(grp.value), int(prec.value)
        self.vocab = [int(vocab[i]) for i in range(self.groups)]

    def __call__(self, codes: np.ndarray) -> np.ndarray:
        """codes [n, groups] int32 -> [n, hidden]: float32, float16, or the raw bf16 bit patterns as uint16."""
        ac = np.ascontiguousarray(codes, dtype=np.int32)
        if ac.ndim != 2 or ac.shape[1] != self.groups:
            raise AudioDecodingFailed(1, f"codes must be [n,{self.groups}], got {ac.shape}")
        dt = {PREC_FP32: np.float32, PREC_FP16: np.float16, PREC_BF16: np.uint16}[self.precision]
        out = np.empty((ac.shape[0], self.hidden), dtype=dt)
        _check(lib().q3tts_codec_embed_sum(self._h, ac.ctypes.data, ac.shape[0], out.ctypes.data))
        return out

    def close(self) -> None:
        if self._h:
            lib().q3tts_codec_embedder_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Qwen3TTSSpeechTokenizerEncoder:
    """Speech-tokenizer encoder, audio -> codes (SpeechTokenizerEncoder.swift:955-1056): same name and `encode` shapes as the
    reference -- audio [B, 1, samples] float -> codes [B, 16, T] int32 at 12.5 frames per second."""

    def __init__(self, speech_tokenizer_dir: str, device: int = 0, precision: int = PREC_FP16):
        """precision: PREC_FP16 = tensor-core engine (every GEMM as three tcgen05 products of split fp16 operands, float32-equivalent),
        PREC_FP32 = CUDA-core float32 engine."""
        opts = Options()
        lib().q3tts_options_default(C.byref(opts))
        opts.device = device
        opts.precision = precision
        h = C.c_void_p()
        _check(lib().q3tts_encoder_load(speech_tokenizer_dir.encode(), C.byref(opts), C.byref(h)))
        self._h = h
        nq, cb, hop, sr, npar = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int64(0)
        _check(lib().q3tts_encoder_info(self._h, C.byref(nq), C.byref(cb), C.byref(hop), C.byref(sr), C.byref(npar)))
        self.valid_num_quantizers, self.codebook_size, self.hop = int(nq.value), int(cb.value), int(hop.value)
        self.sampling_rate, self.num_parameters = int(sr.value), int(npar.value)

    def frames(self, samples: int) -> int:
        return int(lib().q3tts_encode_frames(self._h, samples))

    def encode(self, audio: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(audio, dtype=np.float32)
        if a.ndim == 3 and a.shape[1] == 1:
            a = a[:, 0, :]
        if a.ndim != 2:
            raise AudioDecodingFailed(1, f"audio must be [B, 1, samples] or [B, samples], got {audio.shape}")
        a = np.ascontiguousarray(a)
        B, S = a.shape
        codes = np.empty((B, self.valid_num_quantizers, self.frames(S)), dtype=np.int32)
        _check(lib().q3tts_encode(self._h, a.ctypes.data, B, S, codes.ctypes.data))
        return codes

    def launch_count(self) -> int:
        return int(lib().q3tts_encoder_launch_count(self._h))

    def set_taps(self, on: bool) -> None:
        _check(lib().q3tts_encoder_set_taps(self._h, 1 if on else 0))

    def stage_tap(self, name: str) -> np.ndarray:
        """Stage output of the last encode as [B, C, rows] (the reference's NCL layout)."""
        dims = (C.c_int64 * 3)()
        _check(lib().q3tts_encoder_tap(self._h, name.encode(), None, 0, C.byref(dims)))
        out = np.empty((dims[0], dims[1], dims[2]), dtype=np.float32)
        _check(lib().q3tts_encoder_tap(self._h, name.encode(), out.ctypes.data, out.size, C.byref(dims)))
        return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))

    def close(self) -> None:
        if self._h:
            lib().q3tts_encoder_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count() -> int:
    return int(lib().q3tts_device_count())


def checkpoint_inspect(speech_tokenizer_dir: str) -> Config:
    cfg = Config()
    _check(lib().q3tts_checkpoint_inspect(speech_tokenizer_dir.encode(), C.byref(cfg)))
    return cfg


def partition_lpt(frames: Sequence[int], n_parts: int) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(frames, dtype=np.int64))
    out = np.zeros(len(f), dtype=np.int32)
    _check(lib().q3tts_partition_lpt(f.ctypes.data, len(f), int(n_parts), out.ctypes.data))
    return out


def trim_length(n_samples: int, valid_len: int) -> int:
    return int(lib().q3tts_trim_length(int(n_samples), int(valid_len)))


def voice_clone_cut(ref_frames: int, total_frames: int, n_samples: int) -> int:
    return int(lib().q3tts_voice_clone_cut(int(ref_frames), int(total_frames), int(n_samples)))


def pcm_to_int16(pcm: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
    out = np.empty(p.shape[0], dtype=np.int16)
    _check(lib().q3tts_pcm_to_int16(p.ctypes.data, p.shape[0], out.ctypes.data))
    return out


def write_wav(path: str, pcm: np.ndarray, sample_rate: int = 24000) -> None:
    p = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
    _check(lib().q3tts_write_wav(path.encode(), p.ctypes.data, p.shape[0], int(sample_rate)))


class Qwen3TTSSpeechTokenizerDecoder:
    """codes [B,16,T] int32 -> [B,1,samples] float32, clipped (SpeechTokenizer.swift:754-784)."""

    def __init__(self, owner: "Qwen3TTSSpeechTokenizer"):
        self._owner = owner

    def __call__(self, codes: np.ndarray) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        if codes.ndim != 3 or codes.shape[1] != self._owner.config.num_quantizers:
            raise AudioDecodingFailed(1, f"codes must be [B,{self._owner.config.num_quantizers},T], got {codes.shape}")
        B, _, T = codes.shape
        out = np.empty((B, 1, T * self._owner.config.total_upsample), dtype=np.float32)
        _check(lib().q3tts_decode(self._owner._h, codes.ctypes.data, B, T, CODES_BQT, out.ctypes.data, None))
        return out

    # stage internals the reference's test walks (Tests.swift:57-257), as NCT float32
    def stage(self, name: str) -> np.ndarray:
        return self._owner.stage_tap(name)


class Qwen3TTSSpeechTokenizer:
    """Mirror of ``Qwen3TTSSpeechTokenizer`` (SpeechTokenizer.swift:790-852), decode side only."""

    def __init__(self, speech_tokenizer_dir: str, precision: int = PREC_FP16, attn_mode: int = ATTN_REFERENCE,
                 device: int = -1, workspace_bytes: int = 0, max_frames_per_launch: int = 0):
        L = lib()
        opts = Options()
        L.q3tts_options_default(C.byref(opts))
        opts.device, opts.precision, opts.attn_mode = device, precision, attn_mode
        opts.workspace_bytes, opts.max_frames_per_launch = workspace_bytes, max_frames_per_launch
        h = C.c_void_p()
        _check(L.q3tts_model_load(speech_tokenizer_dir.encode(), C.byref(opts), C.byref(h)))
        self._h = h
        self.config = Config()
        _check(L.q3tts_model_config(self._h, C.byref(self.config)))
        self.decoder = Qwen3TTSSpeechTokenizerDecoder(self)
        self.decode_upsample_rate = int(self.config.decode_upsample_rate)

    @classmethod
    def from_pretrained(cls, model_dir: str, **kw) -> "Qwen3TTSSpeechTokenizer":
        """Like postLoadHook (Qwen3.swift:1461-1494): ``<model_dir>/speech_tokenizer``."""
        return cls(os.path.join(model_dir, "speech_tokenizer"), **kw)

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().q3tts_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def has_encoder(self) -> bool:      # SpeechTokenizer.swift:816 (the encoder itself is out of scope)
        return False

    def encode(self, audio):            # SpeechTokenizer.swift:841-846
        raise AudioDecodingFailed(6, "Speech tokenizer encoder is not available for the loaded model.")

    def decode(self, audio_codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """audio_codes [B,T,16] -> (audio [B, T*1920] float32, audio_lengths [B] int32)."""
        ac = np.ascontiguousarray(audio_codes, dtype=np.int32)
        if ac.ndim != 3 or ac.shape[2] != self.config.num_quantizers:
            raise AudioDecodingFailed(1, f"audio_codes must be [B,T,{self.config.num_quantizers}], got {ac.shape}")
        B, T, _ = ac.shape
        audio = np.empty((B, T * self.config.total_upsample), dtype=np.float32)
        lengths = np.zeros(B, dtype=np.int32)
        _check(lib().q3tts_decode(self._h, ac.ctypes.data, B, T, CODES_BTQ, audio.ctypes.data, lengths.ctypes.data))
        return audio, lengths

    def set_graphs(self, mode: int) -> None:
        """1: replay a CUDA graph for launch chains seen before; -1: only for chains of <= 2048 frames; 0: never (default)."""
        _check(lib().q3tts_set_graphs(self._h, mode))

    def decode_int16(self, audio_codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """audio_codes [B,T,16] -> (audio [B, T*1920] int16 = Int16(clamp(x) * 32767), audio_lengths [B]); main.swift:158-160."""
        ac = np.ascontiguousarray(audio_codes, dtype=np.int32)
        if ac.ndim != 3 or ac.shape[2] != self.config.num_quantizers:
            raise AudioDecodingFailed(1, f"audio_codes must be [B,T,{self.config.num_quantizers}], got {ac.shape}")
        B, T, _ = ac.shape
        audio = np.empty((B, T * self.config.total_upsample), dtype=np.int16)
        lengths = np.zeros(B, dtype=np.int32)
        _check(lib().q3tts_decode_int16(self._h, ac.ctypes.data, B, T, CODES_BTQ, audio.ctypes.data, lengths.ctypes.data))
        return audio, lengths

    def decode_varlen(self, utterances: Sequence[np.ndarray], int16: bool = False):
        """List of [T_i,16] code arrays -> (list of [T_i*1920] PCM arrays (float32, or int16 when `int16`), lengths [N])."""
        n = len(utterances)
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, u in enumerate(utterances):
            u = np.asarray(u)
            if u.ndim != 2 or u.shape[1] != self.config.num_quantizers:
                raise AudioDecodingFailed(1, f"utterance {i} must be [T,{self.config.num_quantizers}]")
            offs[i + 1] = offs[i] + u.shape[0]
        total = int(offs[-1])
        packed = (np.concatenate([np.asarray(u, dtype=np.int32) for u in utterances], axis=0)
                  if total else np.zeros((0, self.config.num_quantizers), np.int32))
        packed = np.ascontiguousarray(packed, dtype=np.int32)
        up = self.config.total_upsample
        pcm = np.empty(total * up, dtype=np.int16 if int16 else np.float32)
        lengths = np.zeros(n, dtype=np.int32)
        fn = lib().q3tts_decode_varlen_int16 if int16 else lib().q3tts_decode_varlen
        _check(fn(self._h, packed.ctypes.data, offs.ctypes.data, n, pcm.ctypes.data, lengths.ctypes.data))
        return [pcm[offs[i] * up: offs[i + 1] * up] for i in range(n)], lengths

    # ---- chunked streaming (Q3TTS_ATTN_CAUSAL_SW; the reference only streams token ids, Qwen3+Streaming.swift) ----
    def open_stream(self) -> "DecodeStream":
        h = C.c_void_p()
        _check(lib().q3tts_stream_open(self._h, C.byref(h)))
        return DecodeStream(self, h)

    def push_streams(self, streams: Sequence["DecodeStream"], chunks: Sequence[np.ndarray]):
        """One chunk [n_i,16] for each stream, decoded in ONE batched launch chain; returns the list of PCM chunks."""
        n = len(streams)
        if n != len(chunks):
            raise AudioDecodingFailed(1, "streams and chunks differ in length")
        up = self.config.total_upsample
        arrs = [np.ascontiguousarray(ch, dtype=np.int32).reshape(-1, self.config.num_quantizers) for ch in chunks]
        outs = [np.empty(a.shape[0] * up, dtype=np.float32) for a in arrs]
        hs = (C.c_void_p * n)(*[s._h for s in streams])
        cs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ps = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        ns = (C.c_int32 * n)(*[a.shape[0] for a in arrs])
        _check(lib().q3tts_stream_push_batch(hs, n, cs, ns, ps))
        return outs

    # ---- device-pointer path (used by bench.py: inputs already resident in HBM) ----
    def decode_device(self, d_codes_ptr: int, B: int, T: int, d_pcm_ptr: int, d_lengths_ptr: int = 0,
                      stream: int = 0, layout: int = CODES_BQT) -> None:
        _check(lib().q3tts_decode_device(self._h, d_codes_ptr, B, T, layout, d_pcm_ptr,
                                         d_lengths_ptr or None, stream or None))

    def decode_varlen_device(self, d_codes_ptr: int, frame_offsets: np.ndarray, d_pcm_ptr: int, d_lengths_ptr: int = 0,
                             stream: int = 0) -> None:
        offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        _check(lib().q3tts_decode_varlen_device(self._h, d_codes_ptr, offs.ctypes.data, len(offs) - 1, d_pcm_ptr,
                                                d_lengths_ptr or None, stream or None))

    def sync(self, stream: int = 0) -> None:
        _check(lib().q3tts_sync(self._h, stream or None))

    # ---- taps / probes ----
    def set_taps(self, enable: bool) -> None:
        _check(lib().q3tts_set_taps(self._h, int(enable)))

    def stage_tap(self, name: str) -> np.ndarray:
        B, Cc, Ln = C.c_int32(), C.c_int32(), C.c_int64()
        _check(lib().q3tts_stage_tap_shape(self._h, name.encode(), C.byref(B), C.byref(Cc), C.byref(Ln)))
        out = np.empty((B.value, Cc.value, Ln.value), dtype=np.float32)
        _check(lib().q3tts_stage_tap(self._h, name.encode(), out.ctypes.data, out.size))
        return out

    def weight_shape(self, swift_key: str) -> Tuple[int, ...]:
        nd = C.c_int32()
        dims = (C.c_int64 * 4)()
        _check(lib().q3tts_weight_shape(self._h, swift_key.encode(), C.byref(nd), C.byref(dims)))
        return tuple(int(dims[i]) for i in range(nd.value))

    # ---- measurement ----
    def profile_enable(self, enable: bool) -> None:
        _check(lib().q3tts_profile_enable(self._h, int(enable)))

    def profile_get(self):
        arr = (StageTime * 32)()
        n = lib().q3tts_profile_get(self._h, arr, 32)
        return [dict(name=arr[i].name.decode(), ms=float(arr[i].ms), launches=int(arr[i].launches),
                     flops=float(arr[i].flops), bytes=float(arr[i].bytes)) for i in range(min(n, 32))]

    def profile_kernels(self):
        """Per-kernel timing of the last profiled decode, aggregated by "stage.op"."""
        arr = (KernelTime * 128)()
        n = lib().q3tts_profile_kernels(self._h, arr, 128)
        return [dict(name=arr[i].name.decode(), ms=float(arr[i].ms), launches=int(arr[i].launches),
                     flops=float(arr[i].flops), bytes=float(arr[i].bytes)) for i in range(min(n, 128))]

    def launch_count(self) -> int:
        return int(lib().q3tts_launch_count(self._h))


def _pack_utterances(utterances: Sequence[np.ndarray], Q: int):
    n = len(utterances)
    offs = np.zeros(n + 1, dtype=np.int64)
    for i, u in enumerate(utterances):
        u = np.asarray(u)
        if u.ndim != 2 or u.shape[1] != Q:
            raise AudioDecodingFailed(1, f"utterance {i} must be [T,{Q}]")
        offs[i + 1] = offs[i] + u.shape[0]
    total = int(offs[-1])
    packed = (np.concatenate([np.asarray(u, dtype=np.int32) for u in utterances], axis=0)
              if total else np.zeros((0, Q), np.int32))
    return np.ascontiguousarray(packed, dtype=np.int32), offs


class Qwen3TTSDecoderPool:
    """One decoder replica + worker thread per GPU of this box; a batch of utterances is LPT-sharded by utterance
    (q3tts_pool_*).  The reference decodes one utterance at a time on one device (Qwen3.swift:744, 951, 1186)."""

    def __init__(self, speech_tokenizer_dir: str, devices: Optional[Sequence[int]] = None, n_devices: int = 0,
                 precision: int = PREC_FP16, attn_mode: int = ATTN_REFERENCE, workspace_bytes: int = 0,
                 max_frames_per_launch: int = 0):
        L = lib()
        opts = Options()
        L.q3tts_options_default(C.byref(opts))
        opts.precision, opts.attn_mode = precision, attn_mode
        opts.workspace_bytes, opts.max_frames_per_launch = workspace_bytes, max_frames_per_launch
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*devices)
            _check(L.q3tts_pool_open(speech_tokenizer_dir.encode(), C.byref(opts), arr, len(devices), C.byref(h)))
        else:
            _check(L.q3tts_pool_open(speech_tokenizer_dir.encode(), C.byref(opts), None, n_devices, C.byref(h)))
        self._h = h
        self.config = checkpoint_inspect(speech_tokenizer_dir)
        self.size = int(L.q3tts_pool_size(self._h))

    def decode_varlen(self, utterances: Sequence[np.ndarray], int16: bool = False, out: Optional[np.ndarray] = None):
        """List of [T_i,16] code arrays -> (list of PCM arrays, lengths [N]); `out` = optional preallocated flat PCM buffer."""
        packed, offs = _pack_utterances(utterances, self.config.num_quantizers)
        return self.decode_packed(packed, offs, int16, out)

    def decode_packed(self, packed: np.ndarray, offs: np.ndarray, int16: bool = False, out: Optional[np.ndarray] = None):
        n, up = len(offs) - 1, self.config.total_upsample
        total = int(offs[-1])
        pcm = out if out is not None else np.empty(total * up, dtype=np.int16 if int16 else np.float32)
        assert pcm.size >= total * up and pcm.dtype == (np.int16 if int16 else np.float32)
        lengths = np.zeros(n, dtype=np.int32)
        fn = lib().q3tts_pool_decode_varlen_int16 if int16 else lib().q3tts_pool_decode_varlen
        _check(fn(self._h, packed.ctypes.data, offs.ctypes.data, n, pcm.ctypes.data, lengths.ctypes.data))
        return [pcm[offs[i] * up: offs[i + 1] * up] for i in range(n)], lengths

    def decode(self, audio_codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """audio_codes [B,T,16] -> (audio [B, T*1920], audio_lengths [B]), rows sharded over the GPUs."""
        ac = np.ascontiguousarray(audio_codes, dtype=np.int32)
        B, T, _ = ac.shape
        audio = np.empty((B, T * self.config.total_upsample), dtype=np.float32)
        lengths = np.zeros(B, dtype=np.int32)
        _check(lib().q3tts_pool_decode(self._h, ac.ctypes.data, B, T, CODES_BTQ, audio.ctypes.data, lengths.ctypes.data))
        return audio, lengths

    def last_stats(self):
        ms = (C.c_float * 64)()
        fr = (C.c_int64 * 64)()
        n = lib().q3tts_pool_last_stats(self._h, ms, fr, 64)
        return [dict(ms=float(ms[i]), frames=int(fr[i])) for i in range(min(n, 64))]

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().q3tts_pool_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
