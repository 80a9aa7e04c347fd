"""Reader/writer for the BSON.jl checkpoints the reference saves and loads.

The reference persists its model with ``@save "trained_model.bson" model opt``
(/root/reference/src/train_brain.jl:295-300) and restores it with
``@load "trained_model.bson" model`` (/root/reference/src/generate_images.jl:250).
The Julia host keeps doing exactly that; this module exists so the Python host
mirror and the tests can read (and re-write) the very same files.

Wire format: plain BSON (little endian).  BSON.jl lowers Julia values to tagged
documents:

* struct   -> {tag:"struct",   type:<datatype>, data:[fields...]}
* array    -> {tag:"array",    type:<datatype>, size:[d1,...], data:<binary col-major>}
* datatype -> {tag:"datatype", name:[module path...], params:[...]}
* tuple    -> {tag:"tuple",    data:[...]}
* backref  -> {tag:"backref",  ref:k}   (1-based index into top-level "_backrefs")

Only host I/O lives here -- no arithmetic of the DDPM path.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Any, List

import numpy as np

# ----------------------------------------------------------------------------- wire level


class BsonArray(list):
    """A BSON array element (type 0x04); distinguishes it from a Python list of fields."""


def _read_cstring(buf: bytes, pos: int):
    end = buf.index(b"\x00", pos)
    return buf[pos:end].decode("utf-8"), end + 1


def _parse_document(buf: bytes, pos: int, as_array: bool = False):
    (size,) = struct.unpack_from("<i", buf, pos)
    end = pos + size
    pos += 4
    out: Any = BsonArray() if as_array else OrderedDict()
    while pos < end - 1:
        etype = buf[pos]
        pos += 1
        key, pos = _read_cstring(buf, pos)
        if etype == 0x01:
            (val,) = struct.unpack_from("<d", buf, pos)
            pos += 8
        elif etype == 0x02:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = buf[pos + 4:pos + 4 + n - 1].decode("utf-8")
            pos += 4 + n
        elif etype == 0x03:
            val, pos = _parse_document(buf, pos, False)
        elif etype == 0x04:
            val, pos = _parse_document(buf, pos, True)
        elif etype == 0x05:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = bytes(buf[pos + 5:pos + 5 + n])
            pos += 5 + n
        elif etype == 0x08:
            val = bool(buf[pos])
            pos += 1
        elif etype == 0x0A:
            val = None
        elif etype == 0x10:
            (val,) = struct.unpack_from("<i", buf, pos)
            pos += 4
        elif etype == 0x12:
            (val,) = struct.unpack_from("<q", buf, pos)
            pos += 8
        else:
            raise ValueError(f"unsupported BSON element type 0x{etype:02x} at {pos}")
        if as_array:
            out.append(val)
        else:
            out[key] = val
    if buf[end - 1] != 0:
        raise ValueError("BSON document not NUL-terminated")
    return out, end


def parse_bson(data: bytes) -> OrderedDict:
    doc, end = _parse_document(data, 0)
    if end != len(data):
        raise ValueError("trailing bytes after BSON document")
    return doc


def _emit_cstring(s: str) -> bytes:
    return s.encode("utf-8") + b"\x00"


def _emit_element(key: str, val: Any) -> bytes:
    k = _emit_cstring(key)
    if isinstance(val, bool):
        return b"\x08" + k + (b"\x01" if val else b"\x00")
    if val is None:
        return b"\x0a" + k
    if isinstance(val, float):
        return b"\x01" + k + struct.pack("<d", val)
    if isinstance(val, int):
        return b"\x12" + k + struct.pack("<q", val)
    if isinstance(val, str):
        b = val.encode("utf-8") + b"\x00"
        return b"\x02" + k + struct.pack("<i", len(b)) + b
    if isinstance(val, (bytes, bytearray)):
        return b"\x05" + k + struct.pack("<i", len(val)) + b"\x00" + bytes(val)
    if isinstance(val, BsonArray) or isinstance(val, (list, tuple)):
        return b"\x04" + k + _emit_document(OrderedDict((str(i), v) for i, v in enumerate(val)))
    if isinstance(val, dict):
        return b"\x03" + k + _emit_document(val)
    raise TypeError(f"cannot BSON-encode {type(val)}")


def _emit_document(doc) -> bytes:
    body = b"".join(_emit_element(k, v) for k, v in doc.items())
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def emit_bson(doc) -> bytes:
    return _emit_document(doc)


# ----------------------------------------------------------------------------- BSON.jl level

_JULIA_DTYPES = {
    ("Core", "Float32"): np.float32,
    ("Core", "Float64"): np.float64,
    ("Core", "Int64"): np.int64,
    ("Core", "Int32"): np.int32,
    ("Core", "UInt8"): np.uint8,
}


class _Resolver:
    def __init__(self, doc):
        self.backrefs = doc.get("_backrefs", BsonArray())

    def deref(self, node):
        while isinstance(node, dict) and node.get("tag") == "backref":
            node = self.backrefs[node["ref"] - 1]
        return node

    def typename(self, tnode):
        tnode = self.deref(tnode)
        if not isinstance(tnode, dict) or tnode.get("tag") != "datatype":
            return None
        return tuple(tnode["name"])


def _collect_arrays(node, res: _Resolver, out: List[np.ndarray]):
    """Depth-first walk in field order, collecting every Float32 array."""
    node = res.deref(node)
    if isinstance(node, dict):
        tag = node.get("tag")
        if tag == "array":
            name = res.typename(node["type"])
            if name in _JULIA_DTYPES and isinstance(node["data"], (bytes, bytearray)):
                dt = _JULIA_DTYPES[name]
                dims = [int(d) for d in node["size"]]
                arr = np.frombuffer(node["data"], dtype=dt).copy()
                # column-major (Julia) -> keep flat + remember the Julia dims
                out.append(JuliaArray(arr, dims))
            else:
                for v in node["data"]:
                    _collect_arrays(v, res, out)
        elif tag in ("struct", "tuple", "svec"):
            for v in node["data"]:
                _collect_arrays(v, res, out)
        elif tag is None:
            for v in node.values():
                _collect_arrays(v, res, out)
    elif isinstance(node, (list, BsonArray)):
        for v in node:
            _collect_arrays(v, res, out)


class JuliaArray:
    """Flat column-major payload plus the Julia ``size``; ``.flat`` is what crosses the C ABI."""

    __slots__ = ("flat", "dims")

    def __init__(self, flat: np.ndarray, dims):
        self.flat = flat
        self.dims = tuple(dims)

    def rowmajor(self) -> np.ndarray:
        """View with reversed dims: Julia (d1,..,dk) column-major == NumPy (dk,..,d1) row-major."""
        return self.flat.reshape(tuple(reversed(self.dims)))

    def __repr__(self):
        return f"JuliaArray{self.dims}"


# Layer table of SimpleUNet in the order BSON stores the arrays
# (/root/reference/src/train_brain.jl:109-145).
# kind, Julia dims of the weight
UNET_LAYERS = [
    ("conv", (3, 3, 129, 64)), ("bn", 64), ("conv", (3, 3, 64, 64)), ("bn", 64),          # down1
    ("conv", (3, 3, 64, 128)), ("bn", 128), ("conv", (3, 3, 128, 128)), ("bn", 128),      # down2
    ("conv", (3, 3, 128, 128)), ("bn", 128), ("conv", (3, 3, 128, 128)), ("bn", 128),     # mid
    ("convT", (2, 2, 64, 128)), ("conv", (3, 3, 64, 64)), ("bn", 64),
    ("conv", (3, 3, 64, 64)), ("bn", 64),                                                  # up2
    ("conv", (3, 3, 128, 64)), ("bn", 64), ("conv", (3, 3, 64, 64)), ("bn", 64),          # up1
    ("conv", (1, 1, 64, 1)),                                                               # final
]


def expected_array_dims():
    """Julia dims of all 64 arrays in ABI order: conv W,b ; bn beta,gamma,mu,sigma2."""
    dims = []
    for kind, spec in UNET_LAYERS:
        if kind == "conv":
            dims += [spec, (spec[3],)]
        elif kind == "convT":
            dims += [spec, (spec[2],)]
        else:
            dims += [(spec,)] * 4
    return dims


def load_checkpoint(path: str):
    """Return (arrays, meta): the 64 Float32 arrays of ``model`` in ABI order and
    ``meta = {"eta":..., "beta":(b1,b2), "eps":..., "epoch": int|None}`` read from ``opt``."""
    with open(path, "rb") as fh:
        doc = parse_bson(fh.read())
    res = _Resolver(doc)
    arrays: List[JuliaArray] = []
    _collect_arrays(doc["model"], res, arrays)
    want = expected_array_dims()
    got = [a.dims for a in arrays]
    if got != want:
        raise ValueError(f"unexpected SimpleUNet array layout in {path}: {got[:6]}...")
    meta = {"epoch": doc.get("epoch")}
    opt = res.deref(doc.get("opt"))
    if isinstance(opt, dict) and opt.get("tag") == "struct":
        data = opt["data"]
        eta = res.deref(data[0])
        if isinstance(eta, dict):  # Float32 scalar lowered as struct{Core.Float32}(bytes)
            raw = eta["data"]
            if isinstance(raw, (list, BsonArray)):
                raw = raw[0]
            eta = float(np.frombuffer(raw, dtype=np.float32)[0]) if isinstance(raw, (bytes, bytearray)) else None
        meta["eta"] = eta
        try:
            meta["beta"] = tuple(res.deref(data[1])["data"])
            meta["eps"] = data[2]
        except Exception:  # pragma: no cover - informational only
            pass
    return arrays, meta


def save_checkpoint(path: str, template_path: str, arrays, epoch=None, eta=None):
    """Write a checkpoint with the exact document structure of ``template_path`` but the
    Float32 payloads replaced by ``arrays`` (ABI order).  Because structure, type tags and
    backrefs are copied verbatim from a file BSON.jl wrote, ``@load`` accepts the result.

    ``@save path model opt [epoch]`` (src/train_brain.jl:295-300) writes the LIVE rule and the current epoch:
    ``eta`` replaces the Float32 learning rate inside ``opt`` (``Optimisers.Adam`` field 1), ``epoch`` sets the
    ``epoch`` key and ``epoch=None`` removes one inherited from the template (``trained_model.bson`` has none)."""
    with open(template_path, "rb") as fh:
        doc = parse_bson(fh.read())
    res = _Resolver(doc)
    it = iter(arrays)

    def patch(node):
        node = res.deref(node)
        if isinstance(node, dict):
            tag = node.get("tag")
            if tag == "array":
                name = res.typename(node["type"])
                if name in _JULIA_DTYPES and isinstance(node["data"], (bytes, bytearray)):
                    new = next(it)
                    flat = new.flat if isinstance(new, JuliaArray) else np.asarray(new)
                    flat = np.ascontiguousarray(flat, dtype=_JULIA_DTYPES[name]).reshape(-1)
                    if flat.nbytes != len(node["data"]):
                        raise ValueError("array size mismatch while patching checkpoint")
                    node["data"] = flat.tobytes()
                else:
                    for v in node["data"]:
                        patch(v)
            elif tag in ("struct", "tuple", "svec"):
                for v in node["data"]:
                    patch(v)
            elif tag is None:
                for v in node.values():
                    patch(v)
        elif isinstance(node, (list, BsonArray)):
            for v in node:
                patch(v)

    patch(doc["model"])
    if epoch is not None:
        doc["epoch"] = int(epoch)
    elif "epoch" in doc:
        del doc["epoch"]
    if eta is not None:
        opt = res.deref(doc.get("opt"))
        if not (isinstance(opt, dict) and opt.get("tag") == "struct"):
            raise ValueError("template checkpoint has no Optimisers.Adam rule to patch")
        node = res.deref(opt["data"][0])
        payload = np.float32(eta).tobytes()
        if isinstance(node, dict):            # Float32 scalar lowered as struct{Core.Float32}(bytes)
            raw = node["data"]
            if isinstance(raw, (list, BsonArray)):
                raw[0] = payload
            else:
                node["data"] = payload
        else:
            opt["data"][0] = float(np.float32(eta))
    with open(path, "wb") as fh:
        fh.write(emit_bson(doc))


# ----------------------------------------------------------------------------- optimiser state (resume)
def _dt(name):
    return OrderedDict([("tag", "datatype"), ("name", BsonArray(name)), ("params", BsonArray())])


def _f32_vector(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    return OrderedDict([("tag", "array"), ("type", _dt(["Core", "Float32"])), ("size", BsonArray([int(a.size)])),
                        ("data", a.tobytes())])


def save_adam_state(path: str, m, v, beta_t, steps: int, eta: float):
    """Adam moments for a true resume, as a BSON.jl-shaped document (``BSON.load(path)`` yields a Dict with
    ``:m``/``:v`` = 64 ``Vector{Float32}`` in the weight order, ``:beta_t``, ``:steps``, ``:eta``).  The reference
    cannot resume: it saves the rule only (src/train_brain.jl:295-300)."""
    if len(m) != 64 or len(v) != 64:
        raise ValueError("expected 64 moment arrays")

    def vec_of_vec(arrs):
        any_t = OrderedDict([("tag", "datatype"), ("name", BsonArray(["Core", "Any"])), ("params", BsonArray())])
        return OrderedDict([("tag", "array"), ("type", any_t), ("size", BsonArray([len(arrs)])),
                            ("data", BsonArray([_f32_vector(a) for a in arrs]))])

    doc = OrderedDict([
        ("m", vec_of_vec(m)), ("v", vec_of_vec(v)),
        ("beta_t", OrderedDict([("tag", "tuple"), ("data", BsonArray([float(beta_t[0]), float(beta_t[1])]))])),
        ("steps", int(steps)), ("eta", float(eta)),
    ])
    with open(path, "wb") as fh:
        fh.write(emit_bson(doc))


def load_adam_state(path: str):
    """Inverse of :func:`save_adam_state`: (m, v, (bt1, bt2), steps, eta)."""
    with open(path, "rb") as fh:
        doc = parse_bson(fh.read())
    res = _Resolver(doc)
    out = {}
    for key in ("m", "v"):
        arrs: List[JuliaArray] = []
        _collect_arrays(doc[key], res, arrs)
        out[key] = [a.flat for a in arrs]
    bt = tuple(np.float32(x) for x in doc["beta_t"]["data"])
    return out["m"], out["v"], (float(bt[0]), float(bt[1])), int(doc["steps"]), float(doc["eta"])
