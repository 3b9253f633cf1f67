"""TEST INFRASTRUCTURE ONLY (oracle) -- stage/build the reference's own binaries into oracle/_ref/ (git-ignored).

Run in the build container, where /root/reference exists; the GPU box only uses the staged results.
  * libonnxruntime.so.1.10.0  : the reference's vendored ONNX Runtime (binary, copied as is)
  * libortref.so              : oracle/ort_ref.cpp compiled against the vendored ORT headers
  * models/{offline,streaming}/{encoder,ctc_model,translator}.onnx : the reference's shipped weights (binary, copied as is)
  * libctcdec_ref.so          : the reference's externals/ctc_decoders C++ (beam/greedy), compiled from the zip where it
                                lies with the stub headers in oracle/ctcdec_stubs/ (openfst/kenlm are not vendored and
                                the ext_scorer == nullptr path never touches them) + oracle/ctcdec_wrap.cpp
No reference *source* is copied into the repository.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
ORT = os.path.join(REFERENCE, "Inference/CppInference/onnx/ext/onnxruntime")
MODELS = os.path.join(REFERENCE, "Inference/PythonInference/asr/models")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout + r.stderr)


def build(verbose: bool = False) -> bool:
    """Returns True when everything under oracle/_ref is in place (built now or earlier)."""
    if not os.path.isdir(REFERENCE):
        return os.path.isfile(os.path.join(REF, "libortref.so"))
    os.makedirs(REF, exist_ok=True)
    ort_so = os.path.join(REF, "libonnxruntime.so.1.10.0")
    if not os.path.isfile(ort_so):
        shutil.copyfile(os.path.join(ORT, "lib", "libonnxruntime.so.1.10.0"), ort_so)
    shim = os.path.join(REF, "libortref.so")
    src = os.path.join(HERE, "ort_ref.cpp")
    if not os.path.isfile(shim) or os.path.getmtime(shim) < os.path.getmtime(src):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ORT, "include"), src, "-o", shim, ort_so,
              "-Wl,-rpath,$ORIGIN"])
    for kind in ("offline", "streaming"):
        os.makedirs(os.path.join(REF, "models", kind), exist_ok=True)
        for m in ("encoder.onnx", "ctc_model.onnx", "translator.onnx"):
            dst = os.path.join(REF, "models", kind, m)
            if not os.path.isfile(dst) and os.path.isfile(os.path.join(MODELS, kind, m)):
                shutil.copyfile(os.path.join(MODELS, kind, m), dst)
    # session layer: the voice-activity model (binary, copied as is)
    os.makedirs(os.path.join(REF, "models", "vad"), exist_ok=True)
    vad_src = os.path.join(REFERENCE, "Inference/PythonInference/vad/models/vad.onnx")
    if os.path.isfile(vad_src) and not os.path.isfile(os.path.join(REF, "models", "vad", "vad.onnx")):
        shutil.copyfile(vad_src, os.path.join(REF, "models", "vad", "vad.onnx"))
    os.makedirs(os.path.join(REF, "models", "punc"), exist_ok=True)
    punc_src = os.path.join(REFERENCE, "Inference/PythonInference/punc_recover/models/punc.onnx")
    if os.path.isfile(punc_src) and not os.path.isfile(os.path.join(REF, "models", "punc", "punc.onnx")):
        shutil.copyfile(punc_src, os.path.join(REF, "models", "punc", "punc.onnx"))
    os.makedirs(os.path.join(REF, "dict"), exist_ok=True)
    for f in ("lm_tokens_ch.txt", "lm_tokens_bd.txt"):
        src_f = os.path.join(REFERENCE, "Inference/PythonInference/punc_recover/src/configs/dict", f)
        if os.path.isfile(src_f) and not os.path.isfile(os.path.join(REF, "dict", f)):
            shutil.copyfile(src_f, os.path.join(REF, "dict", f))
    # vocabulary files used by the reference's TextFeaturizer
    os.makedirs(os.path.join(REF, "dict"), exist_ok=True)
    for f in ("pinyin.txt", "lm_tokens.txt"):
        s = os.path.join(REFERENCE, "asr/configs/dict", f)
        if os.path.isfile(s) and not os.path.isfile(os.path.join(REF, "dict", f)):
            shutil.copyfile(s, os.path.join(REF, "dict", f))
    wrap = os.path.join(HERE, "ctcdec_wrap.cpp")
    if os.path.isfile(wrap):
        out = os.path.join(REF, "libctcdec_ref.so")
        if not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(wrap):
            with tempfile.TemporaryDirectory() as tmp:
                with zipfile.ZipFile(os.path.join(REFERENCE, "externals/ctc_decoders.zip")) as z:
                    for n in ("ctc_beam_search_decoder.cpp", "ctc_beam_search_decoder.h", "path_trie.cpp", "path_trie.h",
                              "decoder_utils.cpp", "decoder_utils.h", "ctc_greedy_decoder.cpp", "ctc_greedy_decoder.h",
                              "ThreadPool/ThreadPool.h"):
                        data = z.read("ctc_decoders/" + n)
                        p = os.path.join(tmp, n)
                        os.makedirs(os.path.dirname(p), exist_ok=True)
                        with open(p, "wb") as f:
                            f.write(data)
                stubs = os.path.join(HERE, "ctcdec_stubs")
                _run(["g++", "-O2", "-std=c++14", "-shared", "-fPIC", "-pthread", "-I" + stubs, "-I" + tmp, "-I" + os.path.join(tmp, "ThreadPool"),
                      os.path.join(tmp, "ctc_beam_search_decoder.cpp"), os.path.join(tmp, "path_trie.cpp"),
                      os.path.join(tmp, "decoder_utils.cpp"), os.path.join(tmp, "ctc_greedy_decoder.cpp"), wrap, "-o", out])
    if verbose:
        print("oracle/_ref:", sorted(os.listdir(REF)))
    return True


if __name__ == "__main__":
    build(verbose=True)
