"""GPU tests of the stack container: sfh_stack_save / sfh_stack_create_from_file through
the C-ABI.  Bit-exactness is the bar (a file round trip moves bytes): the saved "models" / "data" arrays equal what was
uploaded, a stack loaded from the file evaluates fg! to the same bits as the stack it was saved from, and bin-row shards
loaded from ONE file add up to the whole (the loader of rank g touches only its rows).  Files written by the library
are re-read by the independent numpy restatement (tests/file_ref.py) and the other way round.
"""
import numpy as np
import pytest

import file_ref
from conftest import make_flat_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    assert sfh_b200.device_count() >= 1
    return sfh_b200


@pytest.mark.parametrize("dtype,nb,nt", [(np.float64, 1200, 37), (np.float32, 777, 130), (np.float64, 9, 1)])
def test_save_load_round_trip_is_bit_exact(S, tmp_path, dtype, nb, nt):
    M, x, data = make_flat_problem(nb, nt, dtype=dtype)
    ds = S.DeviceStack(M, data)
    la, mh = np.linspace(10.0, 7.0, nt), np.linspace(-2.0, 0.0, nt)
    p = tmp_path / "stack.sfh"
    ds.save(p, logAge=la, MH=mh)
    kind, attrs, arrs = file_ref.read_file(p)                      # the numpy reader checks layout and checksums
    assert kind == 1 and attrs[:3] == [nb, 0, nb] and attrs[5] == (0 if dtype == np.float32 else 1)
    assert arrs["models"].dtype == np.dtype(dtype)
    np.testing.assert_array_equal(arrs["models"], M)
    np.testing.assert_array_equal(arrs["data"], data.astype(np.float64))
    np.testing.assert_array_equal(arrs["logAge"], la)
    ds2 = S.DeviceStack.from_file(p, verify=True)
    assert ds2.shape == ds.shape and ds2.dtype == ds.dtype
    np.testing.assert_array_equal(ds2.logAge, la)
    M2, d2 = ds2.download()
    np.testing.assert_array_equal(M2, M)
    np.testing.assert_array_equal(d2, data.astype(np.float64))
    f1, G1, _ = ds.eval_fg(x)
    f2, G2, _ = ds2.eval_fg(x)
    assert f1 == f2 and np.array_equal(G1, G2)                     # same bytes, same kernel configuration, same bits


def test_numpy_written_stack_file_loads(S, tmp_path):
    M, x, data = make_flat_problem(640, 21)
    p = tmp_path / "np.sfh"
    file_ref.write_file(p, {"models": M, "data": data.astype(np.float64)}, kind=1, attrs=(640, 0, 640, 0, 0, 1, 0, 0))
    ds = S.DeviceStack.from_file(p, verify=True)
    ref = S.DeviceStack(M, data)
    assert ds.eval_fg(x)[0] == ref.eval_fg(x)[0]
    assert ds.logAge is None and ds.hess_shape is None


def test_row_shards_from_one_file_add_up(S, tmp_path):
    nb, nt = 3000, 64
    M, x, data = make_flat_problem(nb, nt)
    whole = S.DeviceStack(M, data)
    p = tmp_path / "whole.sfh"
    whole.save(p, hess_shape=(60, 50))
    f0, G0, _ = whole.eval_fg(x)
    cuts = [0, 700, 701, 2048, nb]
    logl, G = 0.0, np.zeros(nt)
    for a, b in zip(cuts[:-1], cuts[1:]):
        sh = S.DeviceStack.from_file(p, rows=(a, b))
        assert sh.rows == b - a and sh.hess_shape == (60, 50)
        Ms, ds_ = sh.download()
        np.testing.assert_array_equal(Ms, M[a:b])
        np.testing.assert_array_equal(ds_, data[a:b])
        f, g, _ = sh.eval_fg(x)
        logl += f
        G += g
        # a shard saved on its own can only be reloaded as (a subset of) itself
        q = tmp_path / f"shard_{a}.sfh"
        sh.save(q)
        _, attrs, arrs = file_ref.read_file(q)
        assert attrs[:3] == [nb, a, b]
        np.testing.assert_array_equal(arrs["models"], M[a:b])
        again = S.DeviceStack.from_file(q)
        assert again.rows == b - a and again.eval_fg(x)[0] == f
        if b - a > 2:
            sub = S.DeviceStack.from_file(q, rows=(a + 1, b - 1))
            np.testing.assert_array_equal(sub.download()[0], M[a + 1:b - 1])
        with pytest.raises(ValueError):
            S.DeviceStack.from_file(q, rows=(max(a - 1, 0), b + 1))
    assert abs(logl - f0) <= 1e-12 * abs(f0)
    assert np.all(np.abs(G - G0) <= 1e-10 * (np.abs(M).T @ np.abs(1 - data / np.maximum(M @ x, 1e-300))))


def test_corrupt_stack_file_is_refused(S, tmp_path):
    M, x, data = make_flat_problem(256, 8)
    p = tmp_path / "c.sfh"
    S.DeviceStack(M, data).save(p)
    raw = bytearray(open(p, "rb").read())
    raw[4096 + 33] ^= 0x40
    open(p, "wb").write(raw)
    with pytest.raises(S.SFHError) as ei:
        S.DeviceStack.from_file(p, verify=True)
    assert ei.value.status == S._lib.SFH_ERR_IO
    S.DeviceStack.from_file(p, verify=False)                       # header and table are intact: loads when not asked to verify
