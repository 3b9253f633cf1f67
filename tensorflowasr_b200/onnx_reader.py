"""Minimal ONNX (protobuf wire-format) reader -- no `onnx` / `onnxruntime` dependency.

The reference ships its trained weights only as tf2onnx exports
(/root/reference/Inference/PythonInference/asr/models/{offline,streaming}/*.onnx, loaded by
Inference/PythonInference/asr/src/asr.py:22-25 through onnxruntime.InferenceSession).  This module
decodes just enough of the ModelProto to list nodes and pull initializers so the B200 path can build
its own flat weight blob from the very same files.

Wire-format fields used (onnx.proto3):
  ModelProto.graph = 7
  GraphProto: node = 1, name = 2, initializer = 5, input = 11, output = 12, value_info = 13
  NodeProto: input = 1, output = 2, name = 3, op_type = 4, attribute = 5
  AttributeProto: name = 1, f = 2, i = 3, s = 4, t = 5, floats = 7, ints = 8, type = 20
  TensorProto: dims = 1, data_type = 2, float_data = 4, int32_data = 5, int64_data = 7, name = 8, raw_data = 9
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Tuple

import numpy as np


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:  # pragma: no cover - groups are not used by ONNX
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64,
           9: np.bool_, 10: np.float16, 11: np.float64, 12: np.uint32, 13: np.uint64}


def _packed_varints(v) -> List[int]:
    out = []
    pos = 0
    mv = memoryview(v)
    while pos < len(mv):
        x, pos = _varint(mv, pos)
        out.append(_signed(x))
    return out


def _parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(_signed(val))
            else:
                dims.extend(_packed_varints(val))
        elif fno == 2:
            dtype = val
        elif fno == 8:
            name = bytes(val).decode()
        elif fno == 9:
            raw = bytes(val)
        elif fno == 4:
            if wt == 2:
                floats.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif fno in (5, 7):
            if wt == 0:
                ints.append(_signed(val))
            else:
                ints.extend(_packed_varints(val))
    np_dtype = _DTYPES.get(dtype)
    if np_dtype is None:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    else:
        arr = np.asarray(ints, dtype=np_dtype)
    return name, arr.reshape(dims) if dims else arr.reshape(())


@dataclass
class Node:
    name: str
    op_type: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)


def _parse_attr(buf: memoryview) -> Tuple[str, object]:
    name = ""
    val: object = None
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = _parse_tensor(v)[1]
        elif fno == 7:
            if wt == 2:
                floats.extend(np.frombuffer(bytes(v), dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            if wt == 0:
                ints.append(_signed(v))
            else:
                ints.extend(_packed_varints(v))
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


def _parse_node(buf: memoryview) -> Node:
    node = Node("", "", [], [])
    for fno, _wt, v in _fields(buf):
        if fno == 1:
            node.inputs.append(bytes(v).decode())
        elif fno == 2:
            node.outputs.append(bytes(v).decode())
        elif fno == 3:
            node.name = bytes(v).decode()
        elif fno == 4:
            node.op_type = bytes(v).decode()
        elif fno == 5:
            k, a = _parse_attr(v)
            node.attrs[k] = a
    return node


def _value_info_name(buf: memoryview) -> str:
    for fno, _wt, v in _fields(buf):
        if fno == 1:
            return bytes(v).decode()
    return ""


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[str]
    outputs: List[str]

    def producer(self) -> Dict[str, Node]:
        """tensor name -> node that produces it"""
        return {o: n for n in self.nodes for o in n.outputs}

    def consumers(self) -> Dict[str, List[Node]]:
        out: Dict[str, List[Node]] = {}
        for n in self.nodes:
            for i in n.inputs:
                out.setdefault(i, []).append(n)
        return out


def load_graph(path: str) -> Graph:
    with open(path, "rb") as f:
        data = memoryview(f.read())
    graph_buf = None
    for fno, _wt, v in _fields(data):
        if fno == 7:
            graph_buf = v
    if graph_buf is None:
        raise ValueError(f"{path}: no GraphProto in model")
    g = Graph([], {}, [], [])
    for fno, _wt, v in _fields(graph_buf):
        if fno == 1:
            g.nodes.append(_parse_node(v))
        elif fno == 5:
            name, arr = _parse_tensor(v)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_value_info_name(v))
        elif fno == 12:
            g.outputs.append(_value_info_name(v))
    # Constant nodes behave like initializers for our purposes
    for n in g.nodes:
        if n.op_type == "Constant" and "value" in n.attrs:
            g.initializers.setdefault(n.outputs[0], n.attrs["value"])
    return g
