"""Neutral, JSON-able description of what a fill has to compute ("fill spec").

The spec is the seam between histbook's untouched front end (``expr.py``,
``axis.py``, the layout logic in ``Hist.__init__``) and this backend.  It holds,
per histogram, the already-canonical goal trees histbook derives
(``CallGraphGoal.goal`` = ``instr.totree(parsed)``, histbook/instr.py:117-235) and
the content layout (histbook/hist.py:167-258).  Both the CUDA lowering
(``lower.py``) and the CPU oracle (``oracle/fill_oracle.py``) consume it, and
golden fixtures under ``tests/golden`` store it, so neither needs histbook at
run time.

Tree grammar (plain lists so that it survives ``json``)::

    ["name", "x"]                    field            (expr.Name)
    ["pred", "p", positive]          boolean field    (expr.Predicate)
    ["bconst", name_or_None, value]  broadcast const  (expr.BroadcastConst)
    ["const", value]                 literal          (expr.Const)
    ["call", fcn, tree, ...]         library call     (expr.Call)

``value`` is a Python int/float/bool/str/None, or ``{"tuple": [...]}`` /
``{"set": [...]}`` for split edges / ``in`` sets.  ints and floats are kept
distinct because numpy's weak-scalar promotion (NEP 50) depends on it.
"""
import json
import math

import numpy


# ----------------------------------------------------------------------------- values

def _plain(value):
    """numpy scalar -> Python scalar (keeps int/float/bool distinction)."""
    if isinstance(value, (bool, numpy.bool_)):
        return bool(value)
    if isinstance(value, numpy.integer):
        return int(value)
    if isinstance(value, numpy.floating):
        return float(value)
    return value


def encode_value(value):
    value = _plain(value)
    if isinstance(value, tuple):
        return {"tuple": [encode_value(x) for x in value]}
    if isinstance(value, list):
        return {"tuple": [encode_value(x) for x in value]}
    if isinstance(value, (set, frozenset)):
        return {"set": sorted(encode_value(x) for x in value)}
    if value is None or isinstance(value, (bool, int, float, str)):
        return value
    raise TypeError("constant of type {0} cannot be used in a fill program".format(type(value).__name__))


def decode_value(value):
    if isinstance(value, dict):
        if "tuple" in value:
            return tuple(decode_value(x) for x in value["tuple"])
        if "set" in value:
            return set(decode_value(x) for x in value["set"])
        raise ValueError(value)
    return value


# ----------------------------------------------------------------------------- trees

def tree_from_expr(expr):
    """histbook Expr (only Const/Name/Predicate/BroadcastConst/Call, i.e. after
    ``instr.totree``) -> tree.  Class names are compared by name so that this
    module never needs to import histbook itself."""
    cls = expr.__class__.__name__
    if cls == "Const":
        return ["const", encode_value(expr.value)]
    if cls == "BroadcastConst":
        return ["bconst", expr.name, encode_value(expr.value)]
    if cls == "Name":
        return ["name", expr.value]
    if cls == "Predicate":
        return ["pred", expr.value, bool(expr.positive)]
    if cls in ("Call", "BinOp"):
        return ["call", expr.fcn] + [tree_from_expr(x) for x in expr.args]
    raise AssertionError("not a lowered expression: {0!r}".format(expr))


def tree_key(tree):
    """Hashable canonical key of a tree (used for CSE)."""
    if isinstance(tree, list):
        return tuple(tree_key(x) for x in tree)
    if isinstance(tree, dict):
        return tuple(sorted((k, tree_key(v)) for k, v in tree.items()))
    if isinstance(tree, float):
        if math.isnan(tree):
            return ("float", "nan")
        return ("float", tree.hex())
    if isinstance(tree, bool):
        return ("bool", tree)
    if isinstance(tree, int):
        return ("int", tree)
    return tree


def tree_str(tree):
    kind = tree[0]
    if kind == "const":
        return repr(decode_value(tree[1]))
    if kind == "bconst":
        return str(tree[1]) if tree[1] is not None else repr(decode_value(tree[2]))
    if kind == "name":
        return tree[1]
    if kind == "pred":
        return tree[1] if tree[2] else "not " + tree[1]
    return "{0}({1})".format(tree[1], ", ".join(tree_str(x) for x in tree[2:]))


def tree_fields(tree, out=None):
    """Names of the fields (Name / Predicate leaves) a tree reads."""
    if out is None:
        out = set()
    kind = tree[0]
    if kind in ("name", "pred"):
        out.add(tree[1])
    elif kind == "call":
        for x in tree[2:]:
            tree_fields(x, out)
    return out


# ----------------------------------------------------------------------------- hist -> spec

def _goal_trees(axis):
    return [tree_from_expr(g.goal) for g in axis._goals(axis._parsed)]


def hist_spec(hist):
    """Spec of one histbook ``Hist`` (layout per histbook/hist.py:167-258)."""
    group = []
    for axis in hist._group:
        kind = axis.__class__.__name__
        goal, = _goal_trees(axis)
        group.append({"kind": kind, "goal": goal,
                      "keeporder": bool(getattr(axis, "keeporder", False))})

    fixed = []
    for axis in hist._fixed:
        kind = axis.__class__.__name__
        if kind == "_nullaxis":
            fixed.append({"kind": kind, "goal": None, "totbins": int(axis.totbins)})
        else:
            goal, = _goal_trees(axis)
            fixed.append({"kind": kind, "goal": goal, "totbins": int(axis.totbins)})

    profile = []
    for axis in hist._profile:
        sumx, sumx2 = _goal_trees(axis)
        profile.append({"sumx": sumx, "sumx2": sumx2,
                        "sumwx": int(axis._sumwxindex), "sumwx2": int(axis._sumwx2index)})

    wp = hist._weightparsed
    if wp is None:
        weight = None
    elif wp.__class__.__name__ == "Const":
        # hist.py:416-418: weight = ones*value ; weight2 = ones*value**2
        value = _plain(wp.value)
        weight = {"const": value, "const2": value ** 2}
    else:
        # hist.py:225-249: goals are totree(w) and totree(Call(multiply, w, w))
        from histbook_b200 import _ref
        hb = _ref.load()
        w = hb.instr.CallGraphGoal(wp).goal
        w2 = hb.instr.CallGraphGoal(hb.expr.Call("numpy.multiply", wp, wp)).goal
        weight = {"w": tree_from_expr(w), "w2": tree_from_expr(w2)}

    return {"group": group, "fixed": fixed, "profile": profile, "weight": weight,
            "shape": [int(x) for x in hist._shape],
            "sumw": int(hist._sumwindex),
            "sumw2": int(hist._sumw2index) if weight is not None else None}


def book_spec(hists):
    """Spec of a fill over several histograms (one pass, histbook/book.py:540-586)."""
    return {"version": 1, "hists": [hist_spec(h) for h in hists]}


def spec_fields(spec):
    """Sorted field names a spec needs (``Fillable.fields``, histbook/fill.py:41-59)."""
    out = set()
    for h in spec["hists"]:
        for a in h["group"]:
            tree_fields(a["goal"], out)
        for a in h["fixed"]:
            if a["goal"] is not None:
                tree_fields(a["goal"], out)
        for a in h["profile"]:
            tree_fields(a["sumx"], out)
            tree_fields(a["sumx2"], out)
        if h["weight"] is not None and "w" in h["weight"]:
            tree_fields(h["weight"]["w"], out)
            tree_fields(h["weight"]["w2"], out)
    return sorted(out)


def dumps(spec):
    return json.dumps(spec, sort_keys=True, allow_nan=True)


def loads(text):
    return json.loads(text)
