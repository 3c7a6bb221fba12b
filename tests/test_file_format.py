"""On-disk container (include/sfhcuda.h sfh_file_*, csrc/sfh_file.h) -- CPU tests.

The library's writer is read back by the independent numpy restatement tests/file_ref.py and vice versa; corruption,
truncation and argument errors must be reported as SFH_ERR_IO / SFH_ERR_INVALID_ARG, never crash.  Moving a stack between
a file and the DEVICE needs a GPU (tests/test_zz_gpu_file.py); here sfh_stack_create_from_file must get as far as the
device and fail loudly with SFH_ERR_NO_DEVICE.
"""
import ctypes as C
import os

import numpy as np
import pytest

import sfh_b200
from sfh_b200 import io as sio

import file_ref

L = sfh_b200._lib


def _arrays(rng):
    return {
        "models": np.asfortranarray(rng.random((37, 5))),
        "f32": rng.random((3, 4, 2)).astype(np.float32),          # C-ordered input, odd byte count
        "counts": rng.integers(-5, 50, size=11),
        "note": np.frombuffer(b"free-form metadata", dtype=np.uint8),
        "empty": np.zeros((0, 7)),
        "four": rng.random((2, 3, 2, 2)),
    }


def test_checksum_known_answers():
    # the definition, evaluated by hand for tiny inputs
    def mix(z):
        z &= file_ref.M64
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & file_ref.M64
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & file_ref.M64
        return z ^ (z >> 31)
    g = 0x9E3779B97F4A7C15
    assert sio.checksum64(np.zeros(0)) == 0
    one = np.array([1], dtype=np.uint64)
    assert sio.checksum64(one) == mix(1 ^ g)
    two = np.array([7, 9], dtype=np.uint64)
    assert sio.checksum64(two) == (mix(7 ^ g) + mix(9 ^ (2 * g & file_ref.M64))) & file_ref.M64
    tail = np.frombuffer(b"\x01\x02\x03", dtype=np.uint8)           # zero-padded to one word
    assert sio.checksum64(tail) == mix(0x030201 ^ g)
    assert sio.checksum64(two) != sio.checksum64(two[::-1].copy())  # position-dependent


def test_checksum_matches_numpy_restatement_and_threads():
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 4097, (1 << 23) + 5):                     # the last one takes the multi-threaded path
        b = rng.integers(0, 256, size=n, dtype=np.uint8)
        assert sio.checksum64(b) == file_ref.checksum(b.tobytes()), n


def test_library_writer_read_by_numpy_restatement(tmp_path):
    rng = np.random.default_rng(1)
    arrs = _arrays(rng)
    p = tmp_path / "a.sfh"
    sio.write_arrays(p, arrs, kind=sio.KIND_GENERIC, attrs=[3, -4, 5])
    kind, attrs, got = file_ref.read_file(p)
    assert kind == 0 and attrs == [3, -4, 5, 0, 0, 0, 0, 0]
    assert list(got) == list(arrs)
    for k, v in arrs.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(got[k], v)
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]     # written under a temporary name, then renamed


def test_numpy_written_file_read_by_library(tmp_path):
    rng = np.random.default_rng(2)
    arrs = _arrays(rng)
    p = tmp_path / "b.sfh"
    file_ref.write_file(p, arrs, kind=2, attrs=(1, 2, 3, 4, 5, 6, 7, 8))
    with sio.SFHFile(p) as f:
        assert f.kind == 2 and f.attrs == [1, 2, 3, 4, 5, 6, 7, 8] and f.names == list(arrs)
        assert f.verify()
        for k, v in arrs.items():
            a = f[k]
            assert a.dtype == np.asarray(v).dtype and a.shape == np.asarray(v).shape and not a.flags.writeable
            np.testing.assert_array_equal(a, v)
            assert f.describe(k)["checksum"] == file_ref.checksum(np.asarray(v).tobytes(order="F"))
        assert "models" in f and "nope" not in f
        with pytest.raises(KeyError):
            f["nope"]
    got = sio.read_arrays(p)
    np.testing.assert_array_equal(got["four"], arrs["four"])


def test_corruption_and_truncation_are_detected(tmp_path):
    rng = np.random.default_rng(3)
    p = tmp_path / "c.sfh"
    sio.write_arrays(p, {"x": rng.random(1000), "y": rng.random(10)})
    raw = bytearray(open(p, "rb").read())

    def expect_io(data, at_open):
        q = tmp_path / "bad.sfh"
        open(q, "wb").write(data)
        if at_open:
            with pytest.raises(sfh_b200.SFHError) as ei:
                sio.SFHFile(q)
            assert ei.value.status == L.SFH_ERR_IO
        else:
            with sio.SFHFile(q) as f:
                with pytest.raises(sfh_b200.SFHError) as ei:
                    f.verify()
                assert ei.value.status == L.SFH_ERR_IO and "checksum" in str(ei.value)

    b = bytearray(raw); b[4096 + 80] ^= 0x10                        # one payload bit
    expect_io(b, at_open=False)
    b = bytearray(raw); b[128 + 60] ^= 1                            # the array table
    expect_io(b, at_open=True)
    b = bytearray(raw); b[0] = ord("X")                             # magic
    expect_io(b, at_open=True)
    expect_io(raw[:-4096], at_open=True)                            # truncated
    expect_io(raw[:64], at_open=True)                               # shorter than a header
    expect_io(b"", at_open=True)
    with pytest.raises(sfh_b200.SFHError) as ei:
        sio.SFHFile(tmp_path / "does-not-exist.sfh")
    assert ei.value.status == L.SFH_ERR_IO
    # a crafted table entry whose dims product WRAPS mod 2^64 onto the true byte count (table checksum recomputed, so only the
    # overflow check can catch it): dims (2^61 + 125, 8) x 8 bytes == 1000 x 8 bytes mod 2^64  
    import struct
    import file_ref
    b = bytearray(raw)
    ent = list(struct.unpack("<48sii4qQQQ2Q", bytes(b[128:256])))
    assert ent[0].rstrip(b"\0") == b"x" and ent[3] == 1000 and ent[8] == 8000
    ent[2], ent[3], ent[4] = 2, (1 << 61) + 125, 8
    assert ((ent[3] * ent[4] * 8) & ((1 << 64) - 1)) == 8000
    b[128:256] = struct.pack("<48sii4qQQQ2Q", *ent)
    table = bytes(b[128:128 + 2 * 128])
    hdr = list(struct.unpack("<8sIIQQii8qQ2Q", bytes(b[:128])))
    hdr[15] = file_ref.checksum(table)
    b[:128] = struct.pack("<8sIIQQii8qQ2Q", *hdr)
    expect_io(b, at_open=True)


def test_writer_argument_errors(tmp_path):
    with pytest.raises(ValueError):
        sio.write_arrays(tmp_path / "d.sfh", {"x" * 48: np.zeros(3)})
    with pytest.raises(ValueError):
        sio.write_arrays(tmp_path / "d.sfh", {"c": np.zeros(3, dtype=complex)})
    with pytest.raises(ValueError):
        sio.write_arrays(tmp_path / "d.sfh", {"five": np.zeros((1, 1, 1, 1, 1))})
    with pytest.raises(sfh_b200.SFHError) as ei:                     # unwritable directory -> I/O error, nothing left behind
        sio.write_arrays(tmp_path / "no" / "such" / "dir" / "d.sfh", {"x": np.zeros(3)})
    assert ei.value.status == L.SFH_ERR_IO
    # raw C-ABI: duplicate names, NULL data, bad ndim
    d = (L.sfh_array_desc * 2)()
    a = np.zeros(4)
    for k in range(2):
        d[k].name, d[k].dtype, d[k].ndim = b"same", L.SFH_F64, 1
        d[k].dims[0] = 4
    ptrs = (C.c_void_p * 2)(a.ctypes.data, a.ctypes.data)
    path = str(tmp_path / "e.sfh").encode()
    assert L.lib.sfh_file_write(path, 0, None, 2, d, ptrs) == L.SFH_ERR_INVALID_ARG and b"duplicate" in L.lib.sfh_last_error()
    d[1].name = b"other"
    ptrs[1] = None
    assert L.lib.sfh_file_write(path, 0, None, 2, d, ptrs) == L.SFH_ERR_INVALID_ARG
    ptrs[1] = a.ctypes.data
    d[1].ndim = 5
    assert L.lib.sfh_file_write(path, 0, None, 2, d, ptrs) == L.SFH_ERR_INVALID_ARG
    assert not os.listdir(tmp_path)
    assert L.lib.sfh_file_close(None) == L.SFH_OK
    assert L.lib.sfh_file_open(None, None) == L.SFH_ERR_INVALID_ARG


def _stack_file(path, rng, nb=24, nt=3, rows=(0, 24), dtype=np.float64, with_grid=True):
    M = np.asfortranarray(rng.random((rows[1] - rows[0], nt)).astype(dtype))
    arrays = {"models": M, "data": rng.poisson(5.0, rows[1] - rows[0]).astype(np.float64)}
    if with_grid:
        arrays["logAge"] = np.repeat([9.0, 8.0, 7.0], 1)[:nt].astype(np.float64)
        arrays["MH"] = np.full(nt, -1.0)
    file_ref.write_file(path, arrays, kind=1, attrs=(nb, rows[0], rows[1], 6, 4, 0 if dtype == np.float32 else 1, 0, 0))
    return arrays


def test_stack_file_reaches_the_device_or_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (tests/test_zz_gpu_file.py covers the device side)")
    rng = np.random.default_rng(4)
    p = tmp_path / "s.sfh"
    _stack_file(p, rng)
    h = C.c_void_p()
    assert L.lib.sfh_stack_create_from_file(C.byref(h), str(p).encode(), 1, None) == L.SFH_ERR_NO_DEVICE
    assert not h.value
    with pytest.raises(sfh_b200.SFHError) as ei:
        sfh_b200.DeviceStack.from_file(p, verify=True)
    assert ei.value.status == L.SFH_ERR_NO_DEVICE


def test_stack_file_validation_happens_before_the_device(tmp_path):
    rng = np.random.default_rng(6)
    h = C.c_void_p()
    p = tmp_path / "g.sfh"
    sio.write_arrays(p, {"models": rng.random((4, 2)), "data": np.ones(4)})          # generic kind: not a stack file
    assert L.lib.sfh_stack_create_from_file(C.byref(h), str(p).encode(), 0, None) == L.SFH_ERR_IO
    assert b"not a stack file" in L.lib.sfh_last_error()
    q = tmp_path / "s.sfh"
    _stack_file(q, rng, nb=48, rows=(8, 32))
    o = L.sfh_opts(); o.struct_size = C.sizeof(L.sfh_opts); o.row_begin, o.row_end = 0, 16   # outside the file's rows [8, 32)
    assert L.lib.sfh_stack_create_from_file(C.byref(h), str(q).encode(), 0, C.byref(o)) == L.SFH_ERR_SHAPE
    raw = bytearray(open(q, "rb").read()); raw[4096 + 9] ^= 4
    open(q, "wb").write(raw)
    assert L.lib.sfh_stack_create_from_file(C.byref(h), str(q).encode(), 1, None) == L.SFH_ERR_IO   # verify=1 catches it
    assert b"checksum" in L.lib.sfh_last_error()
    r = tmp_path / "r.sfh"                                                                 # attributes disagree with the arrays
    file_ref.write_file(r, {"models": np.asfortranarray(rng.random((5, 2))), "data": np.ones(5)}, kind=1, attrs=(5, 0, 4, 0, 0, 1, 0, 0))
    assert L.lib.sfh_stack_create_from_file(C.byref(h), str(r).encode(), 0, None) == L.SFH_ERR_IO
    assert L.lib.sfh_stack_save(None, b"x", 0, 0, None, None) == L.SFH_ERR_INVALID_ARG


class _R:
    def __init__(self, x):
        self.x = x


def test_result_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    mz = sfh_b200.PowerLawMZR(1.0, -2.0, 6.0, (True, False))
    dp = sfh_b200.GaussianDispersion(0.2)
    mk = lambda: sfh_b200.solvers.BFGSResult(rng.random(6), rng.random(6), rng.random((5, 5)), _R(rng.random(5)), mz, dp)
    res = {"map": mk(), "mle": mk()}
    p = tmp_path / "fit.sfh"
    sfh_b200.save_result(p, res)
    back = sfh_b200.load_result(p)
    for k in ("map", "mle"):
        np.testing.assert_array_equal(back[k]["mu"], res[k].mu)
        np.testing.assert_array_equal(back[k]["sigma"], res[k].sigma)
        np.testing.assert_array_equal(back[k]["invH"], res[k].invH)
        np.testing.assert_array_equal(back[k]["x"], res[k].result.x)
        assert back[k]["MH_class"] == "PowerLawMZR" and back[k]["disp_class"] == "GaussianDispersion"
        np.testing.assert_array_equal(back[k]["MH_params"], [1.0, -2.0])
        np.testing.assert_array_equal(back[k]["MH_free"], [1, 0])
        assert back[k]["MH_fixed"][0] == 6.0
    with sio.SFHFile(p) as f:
        assert f.kind == sio.KIND_RESULT
    chain = {"posterior_matrix": rng.random((6, 40)), "logp": rng.random(40), "step_size": 0.05}
    q = tmp_path / "chain.sfh"
    sfh_b200.save_result(q, chain)
    back = sfh_b200.load_result(q)
    np.testing.assert_array_equal(back["posterior_matrix"], chain["posterior_matrix"])
    assert back["step_size"][0] == 0.05


def test_structured_fuzz_never_crashes(tmp_path):
    """Hostile header / table fields with the table checksum re-sealed (so the mutation reaches the structural validation), plus
    truncations: every mutant is either refused with a status or opens as a consistent file whose arrays can be read end to end.
    Runs in a child process (tests/file_fuzz_child.py): a crash of the parser is a failure here, not the end of the test run."""
    import subprocess
    import sys
    child = os.path.join(os.path.dirname(os.path.abspath(__file__)), "file_fuzz_child.py")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, child, root, str(tmp_path), "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUZZ DONE" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
    opened, refused = (int(tok.split("=")[1]) for tok in r.stdout.split()[-2:])
    assert opened > 100 and refused > 1000          # the mutations reach both outcomes
