"""Locate and import the (unmodified) reference ``histbook`` package.

histbook_b200 is a *backend* for histbook's fill hot path: expression parsing
(``histbook/expr.py``), axis descriptors (``histbook/axis.py``) and the whole
read side (``proj.py``, ``export.py``, ``vega.py``) stay the reference's own code
(BASELINE.json north_star: "Nothing else moves").  This module only finds that
package and applies the one-line Python >= 3.10 compatibility shim it needs
(``histbook/book.py:51`` uses ``collections.MutableMapping``); it never edits a
reference file.

Search order: an already imported / installed ``histbook``; ``$HISTBOOK_REF``;
``<repo>/baseline/_ref`` (pip ``--target`` install, git-ignored, shipped to the
GPU box); ``/root/reference`` (only exists in the build container).
"""
import collections
import collections.abc
import importlib
import os
import sys
import warnings

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_histbook = None


def _shim():
    # histbook/book.py:51 -> class GenericBook(collections.MutableMapping)
    for name in ("MutableMapping", "Mapping", "Iterable", "Sequence", "Callable"):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))


def candidates():
    out = []
    env = os.environ.get("HISTBOOK_REF")
    if env:
        out.append(env)
    out.append(os.path.join(_REPO, "baseline", "_ref"))
    out.append("/root/reference")
    return out


def available():
    """True if the reference package can be imported (without importing it)."""
    if "histbook" in sys.modules:
        return True
    try:
        if importlib.util.find_spec("histbook") is not None:
            return True
    except (ImportError, ValueError):
        pass
    return any(os.path.isfile(os.path.join(c, "histbook", "__init__.py")) for c in candidates())


def load():
    """Import and return the reference ``histbook`` module (cached)."""
    global _histbook
    if _histbook is not None:
        return _histbook
    _shim()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # ast.Num/ast.Str deprecations in expr.py:127-131
        try:
            _histbook = importlib.import_module("histbook")
        except ImportError:
            for c in candidates():
                if os.path.isfile(os.path.join(c, "histbook", "__init__.py")):
                    sys.path.insert(0, c)
                    try:
                        _histbook = importlib.import_module("histbook")
                        break
                    except ImportError:
                        sys.path.remove(c)
            else:
                raise ImportError(
                    "histbook_b200 is a fill backend for the histbook package, which is not "
                    "importable; install histbook or set HISTBOOK_REF (looked in: %s)"
                    % ", ".join(candidates()))
        # submodules used by the backend
        for sub in ("axis", "expr", "instr", "fill", "hist", "book", "util", "calc"):
            importlib.import_module("histbook." + sub)
    return _histbook
