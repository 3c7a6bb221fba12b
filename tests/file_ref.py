"""Independent numpy restatement of the container format of csrc/sfh_file.h (test infrastructure only).

Written from the format DESCRIPTION (the comment block at the top of sfh_file.h / include/sfhcuda.h), not from the C++:
the library's writer is checked by reading its files with `read_file`, its reader by opening files `write_file` made.
"""
import struct

import numpy as np

MAGIC = b"SFHFILE1"
ALIGN = 4096
DT = {0: np.dtype("<f4"), 1: np.dtype("<f8"), 2: np.dtype("<i8"), 3: np.dtype("u1")}
CODE = {v: k for k, v in DT.items()}
M64 = (1 << 64) - 1


def mix64(z):
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def checksum(raw: bytes) -> int:
    raw = raw + b"\0" * (-len(raw) % 8)
    w = np.frombuffer(raw, dtype="<u8")
    with np.errstate(over="ignore"):
        idx = (np.arange(1, w.shape[0] + 1, dtype=np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
        return int(mix64(w ^ idx).sum(dtype=np.uint64)) & M64


def _round_up(a, b):
    return (a + b - 1) // b * b


def write_file(path, arrays, kind=0, attrs=(0,) * 8):
    names = list(arrays)
    hb = _round_up(128 + 128 * len(names), ALIGN)
    off = hb
    table, blobs = b"", []
    for n in names:
        a = np.asarray(arrays[n])
        raw = a.tobytes(order="F")
        dims = list(a.shape) + [1] * (4 - a.ndim)
        table += struct.pack("<48sii4qQQQ2Q", n.encode(), CODE[a.dtype.newbyteorder("<") if a.dtype.itemsize > 1 else a.dtype],
                             a.ndim, *dims, off, len(raw), checksum(raw), 0, 0)
        blobs.append((off, raw))
        off = _round_up(off + len(raw), ALIGN)
    hdr = struct.pack("<8sIIQQii8qQ2Q", MAGIC, 1, 0x01020304, hb, off, len(names), kind, *attrs, checksum(table), 0, 0)
    buf = bytearray(off)
    buf[:128] = hdr
    buf[128:128 + len(table)] = table
    for o, raw in blobs:
        buf[o:o + len(raw)] = raw
    with open(path, "wb") as f:
        f.write(buf)


def read_file(path):
    """Returns (kind, attrs, {name: array}) after checking every structural rule and checksum of the format."""
    raw = open(path, "rb").read()
    magic, ver, endian, hb, fb, n, kind, *rest = struct.unpack("<8sIIQQii8qQ2Q", raw[:128])
    attrs, tsum = rest[:8], rest[8]
    assert magic == MAGIC and ver == 1 and endian == 0x01020304
    assert hb % ALIGN == 0 and hb >= 128 + 128 * n and fb == len(raw)
    assert checksum(raw[128:128 + 128 * n]) == tsum
    out = {}
    end = hb
    for i in range(n):
        name, dt, nd, d0, d1, d2, d3, off, nb, cs, _, _ = struct.unpack("<48sii4qQQQ2Q", raw[128 + 128 * i:256 + 128 * i])
        dims = (d0, d1, d2, d3)
        assert all(d == 1 for d in dims[nd:])
        assert off % ALIGN == 0 and off == end, "arrays are laid out in table order, each on the next 4096-byte boundary"
        assert nb == int(np.prod(dims[:nd], dtype=np.int64)) * DT[dt].itemsize
        blob = raw[off:off + nb]
        assert checksum(blob) == cs
        assert not any(raw[off + nb:_round_up(off + nb, ALIGN)]), "padding must be zero"
        end = _round_up(off + nb, ALIGN)
        out[name.rstrip(b"\0").decode()] = np.frombuffer(blob, dtype=DT[dt]).reshape(dims[:nd], order="F")
    assert end == len(raw)
    return kind, list(attrs), out
