"""ORACLE / TEST INFRASTRUCTURE — reader for the "SBD1" dump container written by oracle/ref_driver.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import struct
import numpy as np

_DT = {b"i": np.int32, b"d": np.float64, b"b": np.uint8}


def read_sbd(path):
    """Return {record name: numpy array or str}."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"SBD1", "not an SBD1 file"
    pos = 4
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos); pos += 4
        name = data[pos:pos + nl].decode(); pos += nl
        dt = data[pos:pos + 1]; pos += 1
        (cnt,) = struct.unpack_from("<Q", data, pos); pos += 8
        if dt == b"s":
            out[name] = data[pos:pos + cnt].decode(); pos += cnt
        else:
            t = np.dtype(_DT[dt])
            out[name] = np.frombuffer(data, dtype=t, count=cnt, offset=pos).copy(); pos += cnt * t.itemsize
    return out
