"""Session layer of the reference's deployment: voice-activity gating in front of the recogniser.

Mirrors Inference/PythonInference/offline_asr_session.py (`ASRSession.send(wav_path)` -> one response per voiced segment, `OfflineVAD`)
and stream_asr_session.py (`ASRSession.send(pcm_bytes)` / `final_send()` -> 'sentence begin' / 'inter break' / 'sentence end' events,
`TaskContent`) with the same thresholds, time arithmetic and event dictionaries, so that a caller of the reference's sessions sees the
same events.  The models behind it are this package's GPU engines (asr.ASR over libb200asr.so, vad_model.VAD over b200asr_vad_*,
punc_model.Punc over b200asr_punc_*); nothing here imports onnxruntime.  `punc` is any object with the reference's
`punc_recover(text) -> list` method; without one (None) the text is returned unpunctuated.

Behaviours of the reference that are kept on purpose (they define what "the same events" means):
  * the offline segmenter never closes a segment on silence: its silence counter is fed only `if self.sound_pick` and nothing sets
    sound_pick (offline_asr_session.py:113-116), so a recording yields at most one segment, from the first voiced 100 ms block (minus
    200 ms) to the end (minus 100 ms); only lists of >= 2 segments go through the merge / split rule (`recover`, :88-89);
  * times are accumulated in float steps (+= 0.1 per block, += packet / rate per packet) in the reference's order: the millisecond
    values of the events come out identical, truncation included;
  * a 'sentence begin' event returns before the chunk-length check of that packet (stream_asr_session.py:111-115).
Deviations: a waveform whose length is not a whole number of 160-sample frames is trimmed (the reference's reshape raises);
`final_send()` does not need a 'task_id' left behind by an earlier event (the reference raises KeyError without one).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

BLOCK = 10          # VAD decisions per 100 ms block (one decision per 10 ms frame)


def _decisions(logits) -> List[int]:
    """logits >= 0 -> 1 (offline_asr_session.py:84-86, stream_asr_session.py:338-340)."""
    return np.where(np.asarray(logits).reshape(-1) >= 0.0, 1, 0).tolist()


def _vad_frames(wav16k: np.ndarray, keep_tail: bool = False) -> np.ndarray:
    """16 kHz samples -> the VAD input [1, N, 80]: every second sample, 80 per frame (a ragged remainder is dropped at the end, or at
    the front for the streaming window whose LAST decisions are the ones used)."""
    usable = len(wav16k) // 160 * 160
    part = wav16k[len(wav16k) - usable:] if keep_tail else wav16k[:usable]
    return np.asarray(part, dtype=np.float32)[::2].reshape(1, -1, 80)


def _apply_punc(punc, text, min_len: int):
    if punc is not None and len(text) >= min_len:
        return punc.punc_recover(text)
    return text


# ------------------------------------------------------------------------------------------------------------------ offline
class OfflineVAD:
    """offline_asr_session.py:53-217."""

    def __init__(self, min_duration=0.5, sr=8000, recover_thread=0.1, recover_max_duration=15.):
        self.init_params()
        self.min_duration = min_duration
        self.sample_rate = sr
        self.recover_thread = recover_thread
        self.recover_max_duration = recover_max_duration

    def compile(self, sd):
        self.sd = sd

    def init_params(self):
        self.data = None
        self.live_result = {"start_time": 0., "end_time": 0.}
        self.vad_result = []
        self.sil_record: List[int] = []
        self.sound_record: List[int] = []
        self.sound_pick = 0          # never set by the reference either: see the module docstring
        self.sound_start = 0
        self.sil_times = 0
        self.wav_length = 0

    def vad(self, wav):
        self.init_params()
        self.wav = wav
        preds = _decisions(self.sd.inference(_vad_frames(np.asarray(wav))))
        self.parse(preds)
        segments = [[round(r["start_time"], 3), round(r["end_time"], 3)] for r in self.vad_result]
        return self.recover(segments) if len(segments) >= 2 else segments

    def parse(self, vad_preds: Sequence[int]):
        self.wav_length = 0
        held = 0                                              # samples of the recording covered so far (the reference concatenates them)
        for blk in range(len(vad_preds) // BLOCK + 1):
            preds = list(vad_preds[blk * BLOCK:(blk + 1) * BLOCK])
            held = min(len(self.wav), (blk + 1) * BLOCK * 160)
            (self.sil_record if self.sound_pick else self.sound_record).extend(preds)
            if self.sound_start:
                if len(self.sil_record) >= 2 * BLOCK:
                    quiet = int(np.sum(self.sil_record[-BLOCK:]))
                    if (quiet <= 8 and self.sil_times == 0) or (quiet <= 5 and self.sil_times >= 1):
                        self.sil_times += 1
                    else:
                        self.sil_times = 0
                    self.sil_record = self.sil_record[-BLOCK:]
                if self.sil_times == 3:
                    self.live_result["end_time"] = self.wav_length - 3 * 0.1 + 0.1
                    self.sil_record = []
                    self.sound_start = 0
                    self.sil_times = 0
                    self.vad_result.append(self.live_result)
            elif len(self.sound_record) == 2 * BLOCK:
                if np.sum(self.sound_record[-BLOCK:]) >= 5.:
                    self.sound_start = 1
                    self.sound_record = []
                    self.live_result["start_time"] = self.wav_length - 0.2
                else:
                    self.sound_record = self.sound_record[-BLOCK:]
            self.wav_length += 0.1
        self.data = held
        self.final_parse()

    def final_parse(self):
        if self.data is None:
            return
        if self.data > int(8000 * 0.2) and self.sound_start:
            self.live_result["end_time"] = self.wav_length - 0.1
            self.vad_result.append(self.live_result)
        self.data = None

    def recover(self, results):
        """Merge neighbours closer than `recover_thread` while the merged span stays below `recover_max_duration`, then cut spans
        longer than that into equal integer-second pieces (offline_asr_session.py:184-217)."""
        merged = []
        s, e = results[0]
        last = len(results) - 1
        for i in range(1, len(results)):
            ns, ne = results[i]
            if ns - e < self.recover_thread and ne - s < self.recover_max_duration:
                e = ne
            else:
                merged.append([s, e])
                s, e = ns, ne
            if i == last:
                merged.append([s, e])
        out = []
        for s, e in merged:
            span = e - s
            if span <= self.recover_max_duration:
                out.append([s, e])
                continue
            pieces = span // self.recover_max_duration
            if span % self.recover_max_duration != 0:
                pieces += 1
            step = int(span / pieces)
            a = s
            for k in range(int(pieces)):
                b = a + step if k != pieces - 1 else e
                out.append([a, b])
                a = b
        return out


class OfflineASRSession:
    """offline_asr_session.py:15-50 (`ASRSession`): VAD segments -> one recognition per segment."""

    def __init__(self, asr, vad, punc=None, session="asr_1", sample_rate=16000):
        self.session = session
        self.sample_rate = sample_rate
        self.asr = asr
        self.punc = punc
        self.offline_vad = OfflineVAD(sr=sample_rate)
        self.offline_vad.compile(vad)

    def send(self, wav_path):
        wav = self.asr.speech_featurizer.load_wav(wav_path) if isinstance(wav_path, (str, os.PathLike)) else np.asarray(wav_path, dtype=np.float32)
        wav = wav[:len(wav) // 160 * 160]
        responses = []
        for idx, (s, e) in enumerate(self.offline_vad.vad(wav)):
            data = wav[int(s * self.sample_rate):int(e * self.sample_rate)]
            result = self.asr.decode([self.asr.extract_feature(data)])
            if len(result) > 5:
                result = _apply_punc(self.punc, result, 0)
            responses.append({"session": "asr_1", "sentence_index": idx, "sentence_begin_time": int(s * 1000), "best_text": result,
                              "sentence_end_time": int(e * 1000)})
        return responses


# ---------------------------------------------------------------------------------------------------------------- streaming
class TaskContent:
    """stream_asr_session.py:275-461: per-connection audio buffer + voice-activity state machine."""

    def __init__(self, session, chunk_max_duration, sr=8000, wait_sil=5, vad_time=1, start_thread=5, end_thread=2):
        self.session = session
        self.chunk_max_duration = chunk_max_duration * sr
        self.wait_sil = wait_sil
        self.sr = sr
        self.vad_time = vad_time
        self.start_thread = start_thread
        self.end_thread = end_thread
        self.init_params()

    def compile(self, sd):
        self.sd = sd

    @staticmethod
    def _fresh_result():
        return {"start_time": 0., "end_time": 0., "live_text": "", "decoded_result": []}

    def init_params(self):
        self.chunk = np.array([], "float32")
        self.wav_length = 0
        self.live_result = self._fresh_result()
        self.vad_point = 0
        self.voice_data = np.zeros(2400)
        self.inter_break = self.start_event = self.end_event = 0
        self.send_flag = 0
        self.sil_record: List[int] = []
        self.sil_times = 0
        self.sound_record: List[int] = []
        self.chunk_point = 0
        self.sound_start = self.sound_end = 0
        self.enc_outputs = []

    def vad(self, wav):
        return _decisions(self.sd.inference(_vad_frames(wav, keep_tail=True)))[-int(BLOCK * self.vad_time):]

    def parse(self, new_data: bytes):
        pcm = np.frombuffer(new_data, "int16").astype("float32")
        pcm /= 32768
        self.wav_length += len(pcm) / self.sr
        if self.sound_start:
            self.chunk = np.concatenate([self.chunk, pcm], 0)
        self.voice_data = np.hstack((self.voice_data, pcm))[-int((self.vad_time + 2) * self.sr):]
        if self.wav_length - self.vad_point >= 0.1 * self.vad_time:
            (self.sil_record if self.sound_start else self.sound_record).extend(self.vad(self.voice_data))
            self.vad_point = self.wav_length
        if not self.sound_start:
            self._look_for_start()
            return
        if len(self.sil_record) >= 2 * BLOCK:
            quiet = int(np.sum(self.sil_record[-BLOCK:]))
            if quiet <= 8 and self.sil_times == 0:
                self.sil_times = 1
                self.inter_break = 1
                self.live_result["end_time"] = self.wav_length
            elif quiet <= 5 and self.sil_times == 1:
                self.sil_times = 2
            elif quiet <= self.end_thread and self.sil_times >= 2:
                self.sil_times += 1
            else:
                self.sil_times = 0
            self.sil_record = self.sil_record[-BLOCK:]
        pending = len(self.chunk) - self.chunk_point
        if self.sil_times == self.wait_sil:
            self.sound_end = self.end_event = 1
            self.live_result["end_time"] = self.wav_length - self.wait_sil * 0.1 + 0.1
            self.sil_record = []
            self.sound_start = self.sil_times = self.inter_break = 0
            self.send_flag = 1
        elif pending >= self.chunk_max_duration:
            self.send_flag = 1
            self.chunk_point = len(self.chunk)
        elif pending == 0:
            self.send_flag = 0

    def _look_for_start(self):
        if len(self.sound_record) != 2 * BLOCK:
            return
        if np.sum(self.sound_record[-BLOCK:]) >= self.start_thread:
            self.sound_start = self.start_event = 1
            self.sound_record = []
            self.chunk = self.voice_data[-int(self.sr * 0.2):]
            self.live_result["start_time"] = self.wav_length - 0.2
        else:
            self.sound_record = self.sound_record[-BLOCK:]

    def reset_chunk(self):
        self.chunk = np.array([], "float32")
        self.chunk_point = 0

    def reset_chunk_end(self):
        self.reset_chunk()
        self.enc_outputs = []

    def chunk_length_check(self):
        if len(self.chunk) >= self.chunk_max_duration:
            self.reset_chunk()

    def final_parse(self):
        if len(self.chunk) > 800 and self.sound_start:
            self.send_flag = self.sound_end = 1
            self.live_result["end_time"] = self.wav_length

    def streaming_live_out(self):
        return self.live_result

    def reset_live_result(self):
        self.live_result = self._fresh_result()
        self.end_event = self.sound_end = self.sound_start = self.send_flag = 0
        self.reset_chunk_end()

    def send_asr(self):
        return self.send_flag


class StreamASRSession:
    """stream_asr_session.py:15-273 (`ASRSession`): packets in, sentence events out."""

    MIN_TAIL = 800          # samples: a shorter remainder is not worth an encoder call (stream_asr_session.py:131,167,233)

    def __init__(self, asr, vad, punc=None, session="asr_1", sample_rate=16000):
        self.session = session
        self.sample_rate = sample_rate
        self.asr = asr
        self.punc = punc
        self.task_content = TaskContent(session, 0.5, sample_rate, 5)
        self.task_content.compile(vad)
        self.sentence_id = 0

    # ---- event constructors (stream_asr_session.py:45-94)
    def on_sentence_begin(self, message):
        return dict(session=self.session, event_type="sentence begin", sentence_index=int(message["index"]),
                    sentence_begin_time=int(message["start_time"]))

    def on_inter_break(self, message):
        return dict(session=self.session, event_type="inter break", sentence_begin_time=int(message["begin_time"]),
                    sentence_end_time=int(message["end_time"]), best_text=str(message["text"]))

    def on_sentence_end(self, message):
        return dict(session=self.session, event_type="sentence end", sentence_index=int(message["index"]),
                    sentence_begin_time=int(message["begin_time"]), best_text=str(message["text"]), sentence_end_time=int(message["end_time"]))

    def _transcribe(self, keep_long_tail: bool = False):
        """Text of the sentence so far: the encoder states of the finished 0.5 s chunks + (if longer than MIN_TAIL) the open chunk."""
        tc = self.task_content
        audio = np.array(tc.chunk, "float32")
        encs = tc.enc_outputs
        if len(audio) > self.MIN_TAIL:
            tail = self.asr.extract_feature(audio)
            text = self.asr.decode(encs + [tail])
            if keep_long_tail and len(audio) >= tc.chunk_max_duration:
                encs.append(tail)
                tc.enc_outputs = encs
        else:
            text = self.asr.decode(encs)
        return "".join(_apply_punc(self.punc, text, 5))

    def _close_sentence(self):
        tc = self.task_content
        live = tc.streaming_live_out()
        live["live_text"] = self._transcribe()
        event = self.on_sentence_end({"index": self.sentence_id, "begin_time": live["start_time"] * 1000, "end_time": live["end_time"] * 1000,
                                      "text": live["live_text"]})
        self.sentence_id += 1
        return event

    def send(self, audio_data: bytes):
        tc = self.task_content
        tc.parse(audio_data)
        if tc.start_event:
            tc.start_event = 0
            return self.on_sentence_begin({"index": self.sentence_id, "start_time": tc.wav_length * 1000 - 200})
        event = None
        if tc.send_flag and tc.sound_end:
            event = self._close_sentence()
            tc.reset_live_result()
        elif tc.send_flag and tc.inter_break and tc.sil_times == 1:
            tc.inter_break = 0
            live = tc.streaming_live_out()
            live["live_text"] = self._transcribe(keep_long_tail=True)
            event = self.on_inter_break({"begin_time": live["start_time"] * 1000, "end_time": live["end_time"] * 1000, "text": live["live_text"]})
            tc.send_flag = 0
        elif tc.send_flag:                        # a 0.5 s chunk is full: encode it now, decode later
            tc.enc_outputs += [self.asr.extract_feature(np.array(tc.chunk, "float32"))]
            tc.send_flag = 0
        tc.chunk_length_check()
        return event

    def final_send(self):
        tc = self.task_content
        tc.final_parse()
        event = None
        if tc.send_asr():
            event = self._close_sentence()
            tc.reset_live_result()
        tc.init_params()
        return event


# ------------------------------------------------------------------------------------------------------- construction helpers
def _resolve(cfg: dict, root: str):
    for key in ("inp_config", "tar_config"):
        c = cfg[key]
        if c and c.get("vocabulary") and not os.path.isabs(c["vocabulary"]):
            c["vocabulary"] = os.path.normpath(os.path.join(root, c["vocabulary"]))
    return cfg


def sessions_from_reference_layout(root: str, kind: str = "offline", device: int = 0, model_root: Optional[str] = None,
                                   with_punctuation: bool = True):
    """Build a session from the reference's deployment tree `root` (= Inference/PythonInference: asr/src/configs/am_data.yml,
    asr/models/{offline,streaming}/*.onnx, vad/models/vad.onnx, punc_recover/models/punc.onnx + its configs), as
    offline_asr_session.py:21-36 / stream_asr_session.py:20-37 do."""
    from . import asr as A
    from . import vad_model as V
    cfg = A.UserConfig(os.path.join(root, "asr/src/configs/am_data.yml"), os.path.join(root, "asr/src/configs/am_data.yml"))
    _resolve(cfg, root)
    model_root = model_root or root
    recogniser = A.ASR(cfg, device=device)
    recogniser.compile(os.path.join(model_root, "asr/models", kind), chunked=False)
    vad = V.VAD(model_path=os.path.join(model_root, "vad/models/vad.onnx"), device=device)
    punc = None
    punc_path = os.path.join(model_root, "punc_recover/models/punc.onnx")
    if with_punctuation and os.path.isfile(punc_path):
        from . import punc_model as P
        pcfg = A.UserConfig(os.path.join(root, "punc_recover/src/configs/data.yml"), os.path.join(root, "punc_recover/src/configs/punc_settings.yml"))
        for key in ("punc_vocab", "punc_biaodian"):
            if not os.path.isabs(pcfg[key]["vocabulary"]):
                pcfg[key]["vocabulary"] = os.path.normpath(os.path.join(root, pcfg[key]["vocabulary"]))
        punc = P.Punc(pcfg, model_path=punc_path, device=device)
    cls = OfflineASRSession if kind == "offline" else StreamASRSession
    return cls(recogniser, vad, punc)
