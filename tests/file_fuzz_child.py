"""Child process of tests/test_file_format.py::test_structured_fuzz_never_crashes (test infrastructure).

Builds a valid container with the numpy restatement, then applies seeded hostile mutations to header and table FIELDS, re-sealing
the table checksum so that the mutation reaches the structural validation behind it; every mutant is opened through the library
(sfh_file_open / info / array / verify / close).  Each must either be refused with a status or behave as a consistent file:
when it opens, every array the library describes must lie inside the file and be readable end to end.  A crash kills this process."""
import ctypes as C
import faulthandler
import os
import struct
import sys

import numpy as np

faulthandler.enable()
ROOT, workdir, nmut = sys.argv[1], sys.argv[2], int(sys.argv[3])
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import file_ref as R   # noqa: E402
import sfh_b200        # noqa: E402

L = sfh_b200._lib
rng = np.random.default_rng(20261017)
base = os.path.join(workdir, "base.sfh")
R.write_file(base, {"models": np.asfortranarray(rng.random((300, 7))), "logAge": np.linspace(6.6, 10.1, 7),
                    "mask": np.array([1, 0, 1], dtype=np.uint8), "idx": np.arange(5, dtype=np.int64)}, kind=1, attrs=(300, 7, 0, 0, 0, 0, 0, 0))
good = bytearray(open(base, "rb").read())
NARR = 4
HOSTILE = [0, 1, -1, 2 ** 31 - 1, -2 ** 31, 2 ** 63 - 1, -2 ** 63, 2 ** 62, 4096, 4095, len(good), len(good) + 4096, 2 ** 40, 7, 300 * 7 * 8]
HDR_FMT, ENT_FMT = "<8sIIQQii8qQ2Q", "<48sii4qQQQ2Q"


def wrap(v, code):
    bits = {"I": 32, "i": 32, "Q": 64, "q": 64}[code]
    v &= (1 << bits) - 1
    return v - (1 << bits) if code in "iq" and v >> (bits - 1) else v


def seal(buf):
    n = struct.unpack_from("<i", buf, 32)[0]
    n = max(0, min(n, (len(buf) - 128) // 128))
    struct.pack_into("<Q", buf, 104, R.checksum(bytes(buf[128:128 + 128 * n])))


opened = refused = 0
for it in range(nmut):
    buf = bytearray(good)
    for _ in range(int(rng.integers(1, 4))):
        if rng.random() < 0.35:                                   # a header field
            f = list(struct.unpack_from(HDR_FMT, buf, 0))
            k = int(rng.choice([1, 2, 3, 4, 5, 6]))               # version, endian, header_bytes, file_bytes, narrays, kind
            f[k] = wrap(int(rng.choice(HOSTILE)), "IIQQii"[k - 1])
            struct.pack_into(HDR_FMT, buf, 0, *f)
        else:                                                     # a field of one table entry
            e = int(rng.integers(0, NARR))
            f = list(struct.unpack_from(ENT_FMT, buf, 128 + 128 * e))
            k = int(rng.integers(0, 10))
            if k == 0:
                f[0] = bytes(rng.integers(1, 256, 48, dtype=np.uint8)) if rng.random() < 0.5 else b"models"   # unterminated / duplicate name
            else:
                f[k] = wrap(int(rng.choice(HOSTILE)), "ii4qQQQ"[k - 1] if k - 1 < 2 else ("q" if k - 1 < 6 else "Q"))
            struct.pack_into(ENT_FMT, buf, 128 + 128 * e, *f)
    if rng.random() < 0.15:
        buf = buf[:int(rng.integers(0, len(buf)))]                # truncation on top
    if len(buf) >= 256 and rng.random() < 0.9:
        seal(buf)                                                 # most mutants carry a valid table checksum
    path = os.path.join(workdir, "m.sfh")
    with open(path, "wb") as fh:
        fh.write(buf)
    sh = C.c_void_p()                                              # the stack loader parses the same file plus its own attributes
    st = L.lib.sfh_stack_create_from_file(C.byref(sh), path.encode(), int(rng.integers(0, 2)), None)
    if st == 0:                                                   # (only on a machine with a device)
        assert L.lib.sfh_stack_destroy(sh) == 0
    else:
        assert not sh.value
    h = C.c_void_p()
    st = L.lib.sfh_file_open(path.encode(), C.byref(h))
    if st != 0:
        refused += 1
        assert not h.value, "no handle may be returned with an error status"
        continue
    opened += 1
    kind, n = C.c_int(), C.c_int()
    at = (C.c_int64 * 8)()
    assert L.lib.sfh_file_info(h, C.byref(kind), C.byref(n), at) == 0
    assert 0 <= n.value <= (len(buf) - 128) // 128
    for i in range(n.value):
        d = L.sfh_array_desc()
        p = C.c_void_p()
        assert L.lib.sfh_file_array(h, i, C.byref(d), C.byref(p)) == 0
        assert 0 <= d.nbytes <= len(buf)
        if d.nbytes:
            blob = C.string_at(p.value, d.nbytes)                 # touches every byte the library says belongs to the array
            assert len(blob) == d.nbytes
        L.lib.sfh_file_verify(h, i)                               # either answer is fine; it must not crash
    assert L.lib.sfh_file_array(h, n.value, C.byref(L.sfh_array_desc()), None) != 0    # out-of-range index is refused
    L.lib.sfh_file_verify(h, -1)
    assert L.lib.sfh_file_close(h) == 0
print(f"FUZZ DONE mutants={nmut} opened={opened} refused={refused}")
