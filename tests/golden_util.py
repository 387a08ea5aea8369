"""Load the golden fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load():
    if "cases" not in _cache:
        with open(os.path.join(HERE, "golden", "cases.json")) as f:
            _cache["meta"] = json.load(f)
        _cache["data"] = numpy.load(os.path.join(HERE, "golden", "data.npz"))
        _cache["cases"] = dict((c["name"], c) for c in _cache["meta"]["cases"])
    return _cache["cases"]


def names():
    return [c["name"] for c in (load() and _cache["meta"]["cases"])]


def fills(case):
    """list of dict name -> ndarray | scalar"""
    data = _cache["data"]
    out = []
    for fm in case["fills"]:
        out.append(dict((n, data[v["npz"]] if "npz" in v else v["scalar"]) for n, v in fm.items()))
    return out


def outputs(case):
    """per hist: list of (keypath tuple, dense ndarray) in the reference's dict order"""
    data = _cache["data"]
    out = []
    for entry in case["outputs"]:
        flat = []
        for e in entry:
            if "npz" in e:
                arr = data[e["npz"]]
            else:
                arr = numpy.zeros(int(numpy.prod(e["shape"])), dtype=numpy.float64)
                arr[data[e["sparse"] + "_idx"]] = data[e["sparse"] + "_val"]
                arr = arr.reshape(e["shape"])
            flat.append((tuple(e["keys"]), arr))
        out.append(flat)
    return out


def flatten(content, path=()):
    """nested dict content -> list of (keypath, ndarray) in insertion order (same key normalisation as make_golden)"""
    if isinstance(content, numpy.ndarray):
        return [(tuple(path), content)]
    out = []
    for k, v in content.items():
        kk = k if isinstance(k, str) else (int(k) if isinstance(k, (int, numpy.integer)) else float(k))
        out.extend(flatten(v, tuple(path) + (kk,)))
    return out
