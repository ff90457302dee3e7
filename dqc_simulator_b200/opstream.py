"""Serialised form of a compiled circuit (``qpu_op_dict``).

The reference's compilers return ``{node_name: [[op tuple, ...] per time slice]}``
(``software/compilers.py:665-792``) whose tuples hold live instruction objects.  To move a
compiled circuit between processes (the GPU box has no copy of the reference), each tuple is
written as a JSON list whose elements are ints, strings, int lists, or an instruction record
``{"i": name}`` / ``{"i": name, "m": [re, im, re, im, ...]}`` for an
``(instruction, operator)`` pair.  Decoding rebuilds tuples that use this package's
instruction tokens, so the executor sees the same shapes either way.
"""
import gzip
import json

import numpy as np

from . import instructions as ins

FORMAT = "dqc-opstream-v1"


def _enc_elem(e):
    if isinstance(e, (int, np.integer)):
        return int(e)
    if isinstance(e, str):
        return e
    if isinstance(e, range):
        return list(e)
    if isinstance(e, (list, tuple)) and all(isinstance(x, (int, np.integer)) for x in e):
        return [int(x) for x in e]
    if isinstance(e, (list, tuple)) and len(e) == 2 and hasattr(e[0], "name"):
        m = e[1].arr if hasattr(e[1], "arr") else np.asarray(e[1])
        flat = np.asarray(m, dtype=complex).reshape(-1)
        return {"i": e[0].name, "m": np.stack([flat.real, flat.imag], 1).reshape(-1).tolist()}
    if hasattr(e, "name"):
        return {"i": e.name}
    raise TypeError(f"cannot encode {e!r}")


def _dec_elem(e):
    if isinstance(e, dict):
        tok = ins.BY_NAME[e["i"]]
        if "m" in e:
            v = np.asarray(e["m"], dtype=float).reshape(-1, 2)
            d = int(round(np.sqrt(v.shape[0])))
            return (tok, (v[:, 0] + 1j * v[:, 1]).reshape(d, d))
        return tok
    return e


def encode(qpu_op_dict, meta=None):
    return {
        "format": FORMAT,
        "meta": meta or {},
        "qpu_ops": {n: [[[_enc_elem(e) for e in op] for op in sl] for sl in slices]
                    for n, slices in qpu_op_dict.items()},
    }


def decode(doc):
    if doc.get("format") != FORMAT:
        raise ValueError("not a dqc-opstream-v1 document")
    ops = {n: [[tuple(_dec_elem(e) for e in op) for op in sl] for sl in slices]
           for n, slices in doc["qpu_ops"].items()}
    return ops, doc.get("meta", {})


def save(path, qpu_op_dict, meta=None):
    data = json.dumps(encode(qpu_op_dict, meta), separators=(",", ":"))
    if str(path).endswith(".gz"):
        # mtime = 0 and no file name in the header: regenerating a fixture is byte-identical
        with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", compresslevel=9, fileobj=raw, mtime=0) as f:
            f.write(data.encode())
    else:
        with open(path, "w") as f:
            f.write(data)


def load(path):
    if str(path).endswith(".gz"):
        with gzip.open(path, "rt") as f:
            doc = json.load(f)
    else:
        with open(path) as f:
            doc = json.load(f)
    return decode(doc)
