"""Bind the user's ``fill`` arguments to kernel columns.

Mirrors the Param handling of ``Fillable._fill`` (histbook/fill.py:88-146): look every field up in the
dict-like ``arrays`` (``KeyError`` -> the ``maybeconstants`` pi/e/inf/nan, else the reference's
``ValueError``), turn non-arrays into numpy arrays, broadcast scalars, and insist on one common
length.  New here: values may live on the device (anything with ``__cuda_array_interface__`` or
``__dlpack__``, e.g. torch CUDA tensors); those are used in place, never copied to the host.
"""
import numpy

from histbook_b200.codegen import DTYPE_CODE, MAYBECONSTANTS, UnsupportedError


class StringCodes(object):
    """Dictionary encoding of one string-valued ``groupby`` field (histbook_groupby accepts strings:
    histbook/calc/__init__.py:149-154, tests/test_hist.py:394-407).  The device only ever sees int64 codes; the
    code of a string never changes, so codes from different fills (and contents uploaded from host dicts) agree."""

    def __init__(self):
        self.code_of = {}
        self.strings = []

    def code(self, key):
        key = str(key) if not isinstance(key, bytes) else key
        c = self.code_of.get(key)
        if c is None:
            c = self.code_of[key] = len(self.strings)
            self.strings.append(key)
        return c

    def encode(self, values):
        """array-like of str / bytes -> int64 codes (numpy.unique once, then a table lookup)."""
        arr = numpy.asarray(values)
        if arr.ndim != 1:
            raise UnsupportedError("only one-dimensional string arrays can be filled")
        if arr.dtype.kind == "O":
            arr = arr.astype(str)
        uniques, inverse = numpy.unique(arr, return_inverse=True)
        table = numpy.array([self.code(u.item() if hasattr(u, "item") else u) for u in uniques], dtype=numpy.int64)
        return table[inverse] if len(uniques) else numpy.zeros(0, dtype=numpy.int64)

    def decode(self, code):
        s = self.strings[int(code)]
        return numpy.bytes_(s) if isinstance(s, bytes) else numpy.str_(s)


def is_stringlike(value):
    """A fill argument that numpy would turn into a string array (list of str, '<U' / 'S' / object-of-str arrays)."""
    if isinstance(value, numpy.ndarray):
        if value.dtype.kind in "US":
            return True
        return value.dtype.kind == "O" and value.ndim == 1 and len(value) > 0 and isinstance(value[0], (str, bytes))
    if isinstance(value, (list, tuple)) and len(value) > 0:
        return isinstance(value[0], (str, bytes))
    return False


_CODE_OF = {}          # numpy dtype -> ABI code (dtype.name is slow; this sits on the per-fill path)


def _code(dtype):
    code = _CODE_OF.get(dtype)
    if code is None:
        code = DTYPE_CODE.get(dtype.name)
        if code is not None and dtype.byteorder != ">":
            _CODE_OF[dtype] = code
    return code


_TORCH_DTYPES = None   # torch dtype -> (numpy dtype, ABI code), built at the first torch column


def _torch_dtypes():
    global _TORCH_DTYPES
    import torch
    table = {}
    for name in DTYPE_CODE:
        t = getattr(torch, name, None)
        if isinstance(t, torch.dtype):
            table[t] = (numpy.dtype(name), DTYPE_CODE[name])
    _TORCH_DTYPES = table
    return table


class Column(object):
    __slots__ = ("name", "dtype", "code", "is_scalar", "ptr", "location", "device", "stream", "length", "scalar_bits", "keepalive")

    def __init__(self, name):
        self.name = name
        self.dtype = None
        self.code = None           # DTYPE_CODE of dtype
        self.is_scalar = False
        self.ptr = None
        self.location = 0          # 0 host, 1 device
        self.device = None         # CUDA device index of a device column, when the object tells (torch tensors)
        self.stream = None         # __cuda_array_interface__ v3 "stream": the producer's stream the consumer has to wait for
        self.length = None
        self.scalar_bits = 0
        self.keepalive = None      # whatever owns the memory `ptr` points into

    def abi(self):
        if self.is_scalar:
            return ("scalar", self.code, self.scalar_bits)
        return (self.ptr, self.code, self.location)


def _scalar(col, value):
    arr = numpy.array(value)
    if arr.dtype.name not in DTYPE_CODE:
        raise UnsupportedError("field {0!r}: scalars of dtype {1} are not supported".format(col.name, arr.dtype))
    col.dtype = arr.dtype
    col.code = DTYPE_CODE[arr.dtype.name]
    col.is_scalar = True
    raw = arr.tobytes() + b"\0" * 8
    col.scalar_bits = int.from_bytes(raw[:8], "little")
    return col


def _from_cuda_array_interface(col, obj, cai):
    shape = tuple(cai["shape"])
    dtype = numpy.dtype(cai["typestr"])
    if len(shape) == 0:
        # 0-d device value: fetch it (one element) and broadcast
        import torch
        return _scalar(col, torch.as_tensor(obj).item())
    if len(shape) != 1:
        raise UnsupportedError("field {0!r}: only one-dimensional device arrays can be filled".format(col.name))
    strides = cai.get("strides")
    if strides is not None and shape[0] > 1 and tuple(strides) != (dtype.itemsize,):
        raise UnsupportedError("field {0!r}: device arrays must be contiguous".format(col.name))
    code = _code(dtype)
    if code is None:
        raise UnsupportedError("field {0!r}: dtype {1} is not supported by the B200 fill backend".format(col.name, dtype))
    col.dtype = dtype
    col.code = code
    col.ptr = int(cai["data"][0]) if shape[0] > 0 else 0
    col.location = 1
    col.length = int(shape[0])
    col.keepalive = obj
    stream = cai.get("stream")
    if stream is not None and int(stream) != 1:                # None: nothing to wait for; 1: the legacy default stream (ours)
        col.stream = int(stream)
    return col


def bind(name, value):
    """One fill argument -> Column."""
    col = Column(name)
    mod = type(value).__module__ or ""
    if mod.startswith("torch"):
        import torch
        if isinstance(value, torch.Tensor):
            if value.is_cuda:
                # read the tensor directly: __cuda_array_interface__ builds a dict and refuses tensors that require grad
                ndim = value.dim()
                if ndim == 0:
                    return _scalar(col, value.item())
                if ndim != 1:
                    raise UnsupportedError("field {0!r}: only one-dimensional device arrays can be filled".format(name))
                known = (_TORCH_DTYPES or _torch_dtypes()).get(value.dtype)
                if known is None:
                    raise UnsupportedError("field {0!r}: dtype {1} is not supported by the B200 fill backend".format(name, value.dtype))
                if not value.is_contiguous():
                    value = value.contiguous()
                col.dtype, col.code = known
                col.length = value.shape[0]
                col.ptr = value.data_ptr() if col.length > 0 else 0
                col.location = 1
                col.device = value.device.index
                col.keepalive = value
                return col
            value = value.detach().numpy()
    cai = getattr(value, "__cuda_array_interface__", None)
    if cai is not None:
        return _from_cuda_array_interface(col, value, cai)
    if mod.startswith("pyarrow"):
        # Arrow arrays / chunked arrays: the values buffer is used in place when there are no nulls and one chunk
        # (zero copy); nulls become NaN for floating point (to_numpy), other types with nulls are refused
        if hasattr(value, "combine_chunks") and hasattr(value, "num_chunks"):
            value = value.chunk(0) if value.num_chunks == 1 else value.combine_chunks()
        if getattr(value, "null_count", 0) and not str(getattr(value, "type", "")).startswith(("float", "double", "halffloat")):
            raise UnsupportedError("field {0!r}: Arrow {1} array with nulls".format(name, value.type))
        value = value.to_numpy(zero_copy_only=False)
    if not isinstance(value, numpy.ndarray):
        if hasattr(value, "__dlpack__") and not isinstance(value, (list, tuple)):
            import torch
            return bind(name, torch.from_dlpack(value))
        value = numpy.asarray(value)                       # fill.py:101-102,138-139 (numpy.array there; no copy needed here)
    if value.shape == ():
        return _scalar(col, value)
    if value.ndim != 1:
        raise UnsupportedError("field {0!r}: only one-dimensional arrays can be filled".format(name))
    code = _code(value.dtype)                              # (cached per dtype: dtype.name costs 3 us, this sits on the per-fill path)
    if code is None and value.dtype.name not in DTYPE_CODE:
        raise UnsupportedError("field {0!r}: dtype {1} is not supported by the B200 fill backend "
                               "(strings/objects need dictionary encoding first)".format(name, value.dtype))
    if not value.flags.c_contiguous or value.dtype.byteorder == ">":
        value = numpy.ascontiguousarray(value, dtype=value.dtype.newbyteorder("="))
        code = _code(value.dtype)
    col.dtype = value.dtype
    col.code = code
    col.ptr = value.__array_interface__["data"][0] if value.shape[0] > 0 else 0      # (value.ctypes.data builds a ctypes object: 1.5 us)
    col.location = 0
    col.length = int(value.shape[0])
    col.keepalive = value
    return col


def bind_all(fields, arrays):
    """fields: sorted names the program needs.  Returns (columns by name, length)."""
    cols = {}
    for name in fields:
        try:
            value = arrays[name]
        except KeyError:
            if name in MAYBECONSTANTS:                     # fill.py:96,133-134
                cols[name] = _scalar(Column(name), numpy.float64(MAYBECONSTANTS[name]))
                continue
            raise ValueError("required field {0} not found in fill arguments".format(repr(name)))
        cols[name] = bind(name, value)

    length = None
    for name in fields:                                    # fill.py:88-110
        if not cols[name].is_scalar:
            length = cols[name].length
            break
    if length is None:
        length = 1
    for name in fields:                                    # fill.py:143-144
        c = cols[name]
        if not c.is_scalar and c.length != length:
            raise ValueError("array {0} has len {1} but other arrays have len {2}".format(repr(name), c.length, length))
    return cols, length
