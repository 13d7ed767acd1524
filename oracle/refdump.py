"""TEST INFRASTRUCTURE: reader for the dump container written by oracle/ref_driver.cpp."""
import struct
import numpy as np

_DTYPES = {0: np.uint32, 1: np.int32, 2: np.float64}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos); pos += 4
        name = data[pos:pos + nl].decode(); pos += nl
        dtype = data[pos]; pos += 1
        (nd,) = struct.unpack_from("<I", data, pos); pos += 4
        dims = struct.unpack_from("<%dQ" % nd, data, pos); pos += 8 * nd
        dt = np.dtype(_DTYPES[dtype])
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(data, dtype=dt, count=n, offset=pos).reshape(dims).copy()
        pos += n * dt.itemsize
        out[name] = arr
    return out


def hash_int(a):
    """Position-weighted multiplicative hash used by ref_driver.cpp (mod 2^64)."""
    a = np.ascontiguousarray(a).reshape(-1)
    v = a.astype(np.uint32).astype(np.uint64) + np.uint64(1)
    i = np.arange(v.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    with np.errstate(over="ignore"):
        h = (v * i * np.uint64(0x9E3779B97F4A7C15)).sum(dtype=np.uint64)
    return int(h)
