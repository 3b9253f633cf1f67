"""Generates tests/golden/offline_golden.npz from the REFERENCE ITSELF (run in the build container):
the shipped ONNX graphs through the vendored onnxruntime 1.10.0 (oracle/ort_ref.py) and the reference's
externals/ctc_decoders C++ (oracle/ctcdec_ref.py).  Inputs: asr/BAC009S0764W0121.wav and seeded noise."""
import os, sys, wave
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ort_ref, ctcdec_ref, ctc_ref

w = wave.open(os.path.join(ROOT, "tests/golden/BAC009S0764W0121.wav"))
x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768
md = ort_ref.model_dir("offline")
enc_m = ort_ref.OrtModel(os.path.join(md, "encoder.onnx"), 1, taps=["melspectrogram/Reshape_2:0", "conv_subsampling/dense/BiasAdd:0"])
ctc_m = ort_ref.OrtModel(os.path.join(md, "ctc_model.onnx"), 1)
feed = {"inputs": x.reshape(1, -1, 1)}
mel = enc_m.run(feed, "melspectrogram/Reshape_2:0").reshape(-1, 80)
sub = enc_m.run(feed, "conv_subsampling/dense/BiasAdd:0")[0]
enc = enc_m.run(feed)[0]
logits = ctc_m.run({"inputs": enc[None]})[0]
frames = np.arange(0, logits.shape[0], 7)
probs = ctc_ref.softmax(logits)
beam = ctcdec_ref.beam_search(probs, 16)
rng = np.random.default_rng(7)
noise = np.clip(rng.standard_normal((2, 32000)).astype(np.float32) * 0.1, -1, 1)
enc_noise = enc_m.run({"inputs": noise[..., None]})
logits_noise = ctc_m.run({"inputs": enc_noise})
out = dict(
    wav_mel=mel.astype(np.float32), wav_sub=sub.astype(np.float32), wav_enc=enc.astype(np.float32),
    wav_logit_frames=frames.astype(np.int32), wav_logits=logits[frames].astype(np.float32),
    wav_argmax=logits.argmax(-1).astype(np.int32),
    wav_ids=np.asarray(ctc_ref.greedy_decode(logits, 1331), dtype=np.int32),
    wav_beam_scores=np.asarray([b[0] for b in beam[:4]], dtype=np.float64),
    wav_beam_ids=np.asarray([b[1] for b in beam[:4]], dtype=np.int32),
    noise_enc=enc_noise.astype(np.float32), noise_argmax=logits_noise.argmax(-1).astype(np.int32),
)
# translator: greedy phone ids + [0] * 10 (Inference/PythonInference/asr/src/asr.py:77) and the encoder states -> character logits
tpath = os.path.join(md, "translator.onnx")
if os.path.isfile(tpath):
    tr_m = ort_ref.OrtModel(tpath, 1)
    tr_in = np.asarray([out["wav_ids"].tolist() + [0] * 10], dtype=np.int32)
    tr_out = tr_m.run({"inputs": tr_in, "enc": enc[None].astype(np.float32)})[0]
    out.update(wav_tr_in=tr_in[0], wav_tr_argmax=tr_out.argmax(-1).astype(np.int32), wav_tr_logits_rows=tr_out[:4].astype(np.float32))
# streaming models: nine chunks (8 x 8000 + 3263 samples), encoded independently, one CTC decode over all frames
sd = ort_ref.model_dir("streaming")
if sd:
    se = ort_ref.OrtModel(os.path.join(sd, "encoder.onnx"), 1)
    sc = ort_ref.OrtModel(os.path.join(sd, "ctc_model.onnx"), 1)
    parts = [se.run({"inputs": x[s:s + 8000].reshape(1, -1, 1)}) for s in range(0, len(x), 8000)]
    senc = np.concatenate(parts, axis=1)
    slog = sc.run({"inputs": senc})[0]
    out.update(stream_enc=senc[0].astype(np.float32), stream_argmax=slog.argmax(-1).astype(np.int32),
               stream_ids=np.asarray(ctc_ref.greedy_decode(slog, 1331), dtype=np.int32))
np.savez_compressed(os.path.join(ROOT, "tests/golden/offline_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
print(out["wav_ids"], out.get("stream_ids"))
