"""Property-based round trips of the container format (hypothesis): arbitrary sets of arrays -- any supported dtype, 1 to 4
dimensions, empty dimensions, odd byte counts, any memory order -- written by the library and read by the independent numpy
restatement (tests/file_ref.py), and the other way round; the checksum is invariant under how the bytes are split."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import sfh_b200
from sfh_b200 import io as sio

import file_ref

DTYPES = [np.float32, np.float64, np.int64, np.uint8]


@st.composite
def array_sets(draw):
    n = draw(st.integers(0, 5))
    out = {}
    for k in range(n):
        dt = draw(st.sampled_from(DTYPES))
        shape = tuple(draw(st.lists(st.integers(0, 7), min_size=1, max_size=4)))
        seed = draw(st.integers(0, 2**31 - 1))
        rng = np.random.default_rng(seed)
        a = (rng.integers(0, 255, size=shape).astype(dt) if dt in (np.int64, np.uint8) else rng.standard_normal(shape).astype(dt))
        if draw(st.booleans()) and a.ndim > 1:
            a = np.asfortranarray(a)
        out[f"a{k}_" + draw(st.text(alphabet="abcXYZ_09/", min_size=0, max_size=20))] = a
    return out, draw(st.integers(0, 2)), draw(st.lists(st.integers(-2**62, 2**62), min_size=8, max_size=8))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(array_sets())
def test_round_trips_both_ways(tmp_path, case):
    arrays, kind, attrs = case
    p, q = tmp_path / "lib.sfh", tmp_path / "np.sfh"
    sio.write_arrays(p, arrays, kind=kind, attrs=attrs)
    k, at, got = file_ref.read_file(p)
    assert k == kind and at == attrs and list(got) == list(arrays)
    for name, a in arrays.items():
        assert got[name].dtype == a.dtype and got[name].shape == a.shape
        np.testing.assert_array_equal(got[name], a)
    file_ref.write_file(q, arrays, kind=kind, attrs=tuple(attrs))
    assert open(p, "rb").read() == open(q, "rb").read()              # the two writers agree byte for byte
    with sio.SFHFile(q) as f:
        assert f.kind == kind and f.attrs == attrs and f.names == list(arrays) and f.verify()
        for name, a in arrays.items():
            np.testing.assert_array_equal(f[name], a)


@settings(max_examples=60, deadline=None)
@given(st.binary(min_size=0, max_size=300), st.integers(0, 300))
def test_checksum_is_additive_over_word_aligned_splits(blob, cut):
    whole = sio.checksum64(np.frombuffer(blob, dtype=np.uint8))
    assert whole == file_ref.checksum(blob)
    cut = min(cut, len(blob)) // 8 * 8                               # split on a word boundary: the second part needs its word offset
    w = np.frombuffer(blob[cut:] + b"\0" * (-len(blob[cut:]) % 8), dtype="<u8")
    with np.errstate(over="ignore"):
        idx = (np.arange(cut // 8 + 1, cut // 8 + 1 + w.shape[0], dtype=np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
        second = int(file_ref.mix64(w ^ idx).sum(dtype=np.uint64))
    first = file_ref.checksum(blob[:cut])
    assert (first + second) & file_ref.M64 == whole
