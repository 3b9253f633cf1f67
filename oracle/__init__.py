"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is the checker for the B200 path: a CPU restatement of the reference's
algorithm (``conformer_ref.py``, ``ctc_ref.py``; ``chunk_conformer_ref.py`` for the not-yet-built ChunkConformer
streaming row, parity unpinned), and thin drivers for the reference's own binaries
(``ort_ref.py`` -> vendored ONNX Runtime 1.10.0 running the shipped ONNX graphs, ``ctcdec_ref.py`` -> the
reference's externals/ctc_decoders C++ compiled from its zip).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it.  The product package
``tensorflowasr_b200`` never does.
"""
