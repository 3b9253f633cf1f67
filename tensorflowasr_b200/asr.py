"""Host-side mirror of the reference's Python inference surface, running on the B200 engine.

Same names, arguments and return conventions as
  * utils/user_config.py:13-25            -> UserConfig(common_yml, model_yml)  (dict.update merge, missing key -> None)
  * utils/text_featurizers.py:7-98        -> TextFeaturizer(decoder_config)
  * utils/speech_featurizers.py:55-77     -> SpeechFeaturizer(speech_config).load_wav(path)
  * Inference/PythonInference/asr/src/asr.py:11-94 and test_asr.py:14-225
                                          -> ASR(config): compile / extract_feature / softmax / greedy_decode /
                                             remove_blank / decode / stt / offline_stt / stream_stt
The same am_data.yml / conformerS.yml / Streaming_ConformerS.yml files are read unchanged.  Differences, all forced
by scope: the translator (pinyin -> characters) is not part of this path, so `decode()` returns the
space-joined phone tokens and `stt()` returns (phones, "").  `inp_config.beam_width` (> 1) selects the device prefix
beam search -- the key exists in the reference's YAML but is never read there.
"""
from __future__ import annotations

import os
import wave
from collections import UserDict
from typing import List, Optional, Sequence

import numpy as np
import yaml

from . import engine as E


def load_yaml(path: str):
    with open(os.path.expanduser(path), "r", encoding="utf-8") as f:
        return yaml.load(f, Loader=yaml.FullLoader)


class UserConfig(UserDict):
    """utils/user_config.py:13-25."""

    def __init__(self, common: str, model: str):
        custom = load_yaml(common)
        custom.update(load_yaml(model))
        super().__init__(custom)

    def __missing__(self, key):
        return None


class TextFeaturizer:
    """utils/text_featurizers.py:7-98 without TensorFlow: vocabulary file -> token <-> index maps.
    blank_at_zero False (the shipped configs) puts the blank LAST: num_classes = len(vocab) + 1."""

    def __init__(self, decoder_config: dict, show: bool = False):
        self.decoder_config = decoder_config
        path = os.path.expanduser(decoder_config["vocabulary"])
        self.decoder_config["vocabulary"] = path
        self.token_to_index = {}
        self.index_to_token = {}
        self.vocab_array = []
        index = 0
        if decoder_config["blank_at_zero"]:
            self.blank = 0
            index = 1
        with open(path, "r", encoding="utf-8") as fin:
            for line in fin.readlines():
                line = line.strip()
                if line.startswith("#") or not line:
                    continue
                if line == "[SPACE]":
                    line = " "
                self.token_to_index[line] = index
                self.index_to_token[index] = line
                self.vocab_array.append(line)
                index += 1
        self.num_classes = index
        if not decoder_config["blank_at_zero"]:
            self.blank = index
            self.num_classes += 1
        self.pad = 0
        self.stop = -1

    def startid(self):
        return self.token_to_index["<S>"]

    def endid(self):
        return self.token_to_index["</S>"]

    def extract(self, tokens):
        return [self.token_to_index[t] for t in tokens]

    def iextract(self, feat):
        if isinstance(feat, list):
            return [self.index_to_token[i] for i in feat]
        return self.index_to_token[feat]


def read_raw_audio(audio, sample_rate: int = 16000) -> np.ndarray:
    """utils/speech_featurizers.py:10-22: path / bytes / ndarray -> float32 mono in [-1, 1] at `sample_rate`
    (PCM WAV through the stdlib; other rates are resampled with scipy's polyphase filter)."""
    if isinstance(audio, np.ndarray):
        return audio
    if isinstance(audio, (bytes, bytearray)):
        import io
        f = wave.open(io.BytesIO(audio))
    elif isinstance(audio, str):
        f = wave.open(os.path.expanduser(audio))
    else:
        raise ValueError("input audio must be either a path or bytes")
    n, sw, ch, sr = f.getnframes(), f.getsampwidth(), f.getnchannels(), f.getframerate()
    raw = f.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported sample width {sw}")
    if ch > 1:
        x = x.reshape(-1, ch).mean(axis=1)
    if sr != sample_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(sr, sample_rate)
        x = resample_poly(x, sample_rate // g, sr // g).astype(np.float32)
    return x


class SpeechFeaturizer:
    """utils/speech_featurizers.py:55-77 (the Mel features themselves are computed inside the model / engine)."""

    def __init__(self, speech_config: dict):
        self.sample_rate = speech_config["sample_rate"]
        try:
            self.frame_length = int(self.sample_rate * (speech_config["frame_ms"] / 1000))
            self.frame_step = int(self.sample_rate * (speech_config["stride_ms"] / 1000))
            self.num_feature_bins = speech_config["num_feature_bins"]
        except Exception:
            pass

    def load_wav(self, path):
        return read_raw_audio(path, self.sample_rate)

    def pad_signal(self, wavs: Sequence[np.ndarray], max_length: int) -> np.ndarray:
        out = np.zeros((len(wavs), int(max_length)), dtype=np.float32)    # 'post' padding / truncation (:75-77)
        for i, w in enumerate(wavs):
            n = min(len(w), int(max_length))
            out[i, :n] = w[:n]
        return out


class ASR:
    """Inference/PythonInference/asr/src/asr.py:11-94 + test_asr.py:14-225 over the B200 engine."""

    def __init__(self, config, device: int = 0, precision: int = E.PRECISION_TF32):
        self.running_config = config["running_config"]
        self.speech_config = config["speech_config"]
        self.model_config = config["model_config"]
        self.opt_config = config["optimizer_config"]
        self.phone_featurizer = TextFeaturizer(config["inp_config"])
        self.text_featurizer = TextFeaturizer(config["tar_config"]) if config["tar_config"] else None
        self.speech_featurizer = SpeechFeaturizer(self.speech_config)
        self.chunk = int(self.speech_config["sample_rate"] * self.speech_config["streaming_bucket"])
        self.beam_width = int((config["inp_config"] or {}).get("beam_width") or 1)
        self.device = device
        self.precision = precision
        self.engine: Optional[E.Engine] = None

    # ------------------------------------------------------------------ model loading
    def compile(self, path: str, chunked: Optional[bool] = None):
        """`path` holds encoder.onnx + ctc_model.onnx (+ translator.onnx) exactly as the reference's deployment directory does
        (asr.py:22-25).  `chunked` (default: speech_config.streaming) makes the engine split every utterance into
        `streaming_bucket` chunks itself (test_asr.py:116-165); the session layer passes False: it cuts the chunks and hands each one,
        whatever its length, to extract_feature() as the reference's sessions do."""
        if chunked is None:
            chunked = bool(self.speech_config["streaming"])
        chunk = self.chunk if chunked else 0
        self.engine = E.engine_from_onnx(path, device=self.device, precision=self.precision, chunk_samples=chunk)
        mc = self.model_config or {}
        geo = self.engine.enc_geo
        for key, have in (("dmodel", geo.dmodel), ("num_blocks", geo.num_blocks), ("num_heads", geo.num_heads),
                          ("head_size", geo.head_size), ("kernel_size", geo.kernel_size)):
            if mc.get(key) is not None and int(mc[key]) != have:
                raise ValueError(f"model_config.{key}={mc[key]} does not match the weights in {path} ({have})")
        if self.engine.ctc_geo.vocab != self.phone_featurizer.num_classes:
            raise ValueError("vocabulary size does not match the CTC head of the model")

    # ------------------------------------------------------------------ reference-shaped primitives
    def softmax(self, logits):
        m = np.max(logits, axis=1, keepdims=True)
        e = np.exp(logits - m)
        return e / np.sum(e, axis=1, keepdims=True)

    def extract_feature(self, wav: np.ndarray) -> np.ndarray:
        """wav float32 [L] -> encoder states np.float32 [1, T', D]   (asr.py:34-39)."""
        wav = np.asarray(wav, dtype=np.float32).reshape(1, -1)
        return self.engine.encode(wav).cpu().numpy()

    def remove_blank(self, labels, blank=0):
        out, prev = [], None
        for l in labels:
            if l != prev:
                out.append(l)
                prev = l
        return [l for l in out if l != blank]

    def greedy_decode(self, y, blank=1331):
        """y [T, V] -> token ids: frame argmax + merge + drop blank, on the device (asr.py:56-61)."""
        ids, lens = self.engine.ctc_greedy(np.asarray(y, dtype=np.float32)[None], blank=int(blank))
        return ids[0, :int(lens[0])].cpu().tolist()

    def decode_ids(self, enc_features: List[np.ndarray]) -> List[int]:
        enc = np.hstack(enc_features) if len(enc_features) > 1 else enc_features[0]
        logits = self.engine.ctc_logits(np.asarray(enc, dtype=np.float32))
        blank = self.phone_featurizer.num_classes - 1 if not self.phone_featurizer.decoder_config["blank_at_zero"] else 0
        if self.beam_width > 1:
            ids, lens, _ = self.engine.ctc_beam(logits, self.beam_width, blank=blank)
            return ids[0, 0, :int(lens[0, 0])].cpu().tolist()
        ids, lens = self.engine.ctc_greedy(logits, blank=blank)
        return ids[0, :int(lens[0])].cpu().tolist()

    def has_translator(self) -> bool:
        return self.engine is not None and self.engine.tr_geo is not None and self.text_featurizer is not None

    def translate_ids(self, phone_ids: List[int], enc: np.ndarray, pad: int = 10) -> List[int]:
        """Greedy phone ids + `pad` zeros and the encoder states through the translator -> per-position character argmax
        (Inference/PythonInference/asr/src/asr.py:77-84)."""
        seq = np.asarray([list(phone_ids) + [0] * pad], dtype=np.int32)
        logits = self.engine.translate(seq, np.asarray(enc, dtype=np.float32))
        return logits[0].argmax(-1).cpu().tolist()

    def decode(self, enc_features: List[np.ndarray]) -> str:
        """List of [1, T_i, D] encoder states (hstacked along time) -> text (asr.py:62-94): CTC greedy phone ids, ten zeros appended,
        translator, argmax; ids 0 and </S> are dropped and </S> ends the sentence; characters joined without separator.  Without
        translator weights / tar_config the phone string is returned instead (space separated)."""
        ids = self.decode_ids(enc_features)
        if not self.has_translator():
            return " ".join(self.phone_featurizer.iextract(ids))
        enc = np.hstack(enc_features) if len(enc_features) > 1 else enc_features[0]
        end = self.text_featurizer.endid()
        txt = []
        for n in self.translate_ids(ids, enc):
            if n != 0 and n != end:
                txt.append(n)
            if n == end:
                break
        return "".join(self.text_featurizer.iextract(txt))

    def _text_of(self, phone_ids: List[int], enc: np.ndarray) -> str:
        """test_asr.py:203-218: translator on the decoded ids as they are (no padding); 0 dropped, everything up to and including </S> kept."""
        if not self.has_translator() or not phone_ids:
            return ""
        end = self.text_featurizer.endid()
        txt = []
        for n in self.translate_ids(phone_ids, enc, pad=0):
            if n != 0:
                txt.append(n)
            if n == end:
                break
        return "".join(self.text_featurizer.iextract(txt))

    # ------------------------------------------------------------------ test_asr.py entry points
    def offline_stt(self, wav_path):
        data = self.speech_featurizer.load_wav(wav_path)
        if self.beam_width <= 1 and not self.has_translator():
            ids, lens = self.engine.recognize(np.asarray(data, dtype=np.float32)[None])      # one fused call: no logits, no encoder read-back
            raw_ids, enc = ids[0, :int(lens[0])].cpu().tolist(), None
        else:
            enc = self.extract_feature(data)
            raw_ids = self.decode_ids([enc])
        result = [i for i in raw_ids if i != 0]                                              # test_asr.py:206-209 drops id 0 too
        text = self._text_of(raw_ids, enc) if enc is not None else ""
        return " ".join(self.phone_featurizer.iextract(result)), text

    def stream_stt(self, wav_path):
        """test_asr.py:116-165: encode chunk by chunk (each chunk alone), re-decode everything seen so far."""
        data = self.speech_featurizer.load_wav(wav_path)
        enc_outputs, result, raw_ids = None, [], []
        for s in range(0, len(data), self.chunk):
            enc = self.extract_feature(data[s:s + self.chunk])
            enc_outputs = enc if enc_outputs is None else np.hstack((enc_outputs, enc))
            raw_ids = self.decode_ids([enc_outputs])
            result = [i for i in raw_ids if i != 0]
        text = self._text_of(raw_ids, enc_outputs) if enc_outputs is not None else ""
        return " ".join(self.phone_featurizer.iextract(result)), text

    def stt(self, wav_path):
        return self.stream_stt(wav_path) if self.speech_config["streaming"] else self.offline_stt(wav_path)

    # ------------------------------------------------------------------ batched form (asr/tester/am_tester.py:34-40)
    def recognize_batch(self, wavs: np.ndarray) -> List[List[int]]:
        """Equal-length batch [B, L] -> greedy ids per utterance (no padding mask in the reference)."""
        ids, lens = self.engine.recognize(np.asarray(wavs, dtype=np.float32))
        ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
        return [ids[b, :lens[b]].tolist() for b in range(len(lens))]
