"""N>1 host logic on CPU: world_size-2 `gloo` process group.  Each rank evaluates its bin-row
shard with the CPU ORACLE as the local evaluator (tests may use the oracle; the product's local evaluator is the
GPU), then the product's sharding code all-reduces [logL, G] and applies the guards; the result must equal the
whole-stack oracle.  Also covers the partition function and the unique-id broadcast plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_flat_problem


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nb, nt, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        import sfh_b200 as S
        M, x, data = make_flat_problem(nb, nt, seed=17)
        b, e = S.shard_rows(nb, world, rank)
        if e > b:
            Cm = O.composite(x, M[b:e])
            logl = float(np.sum(np.where(data[b:e] > 0, data[b:e] - Cm - data[b:e] * np.log(np.where(data[b:e] > 0, data[b:e], 1) / Cm), -Cm)))
            _, G, _ = O.fg(x, M[b:e], data[b:e])
        else:
            logl, G = 0.0, np.zeros(nt)
        out = np.concatenate([[logl], G])
        nl, Gall = S.allreduce_fg(out)
        uid = S.sharding.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
        q.put((rank, nl, Gall, uid, (b, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nb,nt", [(1000, 7), (37, 5)])
def test_two_rank_gloo_sharded_fg(nb, nt):
    import oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, nt, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    M, x, data = make_flat_problem(nb, nt, seed=17)
    nl_ref, G_ref, _ = O.fg(x, M, data)
    spans = sorted(r[4] for r in res)
    assert spans[0][0] == 0 and spans[-1][1] == nb and spans[0][1] == spans[1][0]
    for rank, nl, G, uid, _ in res:
        assert nl == pytest.approx(float(nl_ref), rel=1e-13)
        assert np.allclose(G, G_ref, rtol=1e-11, atol=1e-9)
        assert uid == bytes(range(128))


def test_shard_rows_partition_properties():
    import sfh_b200 as S
    for nb in (0, 1, 31, 32, 33, 1000, 60000, 10**6 + 7):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_rows(nb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nb
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b
                assert b % 32 == 0 or b == nb
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(s for s in sizes if s > 0 or True) <= max(32 * world, max(sizes))
    with pytest.raises(ValueError):
        S.shard_rows(10, 2, 2)


def test_guard_after_reduction():
    import sfh_b200 as S
    assert S.guard_neg_logl(0.0) == np.inf           # fitting_base.jl:95
    assert S.guard_neg_logl(-3.5) == 3.5
    # two shards whose partial sums cancel exactly must trigger the guard only AFTER the sum
    nl, G = S.allreduce_fg(np.array([0.0, 1.0, 2.0]))
    assert nl == np.inf and np.array_equal(G, [1.0, 2.0])


def _file_worker(rank, world, port, path, q):
    """Each rank maps the ONE stack file and evaluates only the bin rows shard_rows assigns to it (the host side of
    sfh_stack_create_from_file's per-shard loading; the oracle stands in for the device as the local evaluator)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        import sfh_b200 as S
        with S.SFHFile(path) as f:
            nb = f.attrs[0]
            b, e = S.shard_rows(nb, world, rank)
            M = np.array(f["models"][b:e])                       # touches only these rows of the mapping
            data = np.array(f["data"][b:e])
            x = f.read("x_eval")
            csum = S.io.checksum64(np.asfortranarray(M))
        Cm = O.composite(x, M)
        logl = float(np.sum(np.where(data > 0, data - Cm - data * np.log(np.where(data > 0, data, 1) / Cm), -Cm)))
        _, G, _ = O.fg(x, M, data)
        nl, Gall = S.allreduce_fg(np.concatenate([[logl], G]))
        q.put((rank, nl, Gall, (b, e), csum))
    finally:
        dist.destroy_process_group()


def test_two_ranks_load_their_shards_from_one_file(tmp_path):
    import oracle as O
    import sfh_b200 as S
    nb, nt = 1500, 9
    M, x, data = make_flat_problem(nb, nt, seed=23)
    path = str(tmp_path / "stack.sfh")
    S.write_arrays(path, {"models": M, "data": data.astype(np.float64), "x_eval": x}, kind=S.io.KIND_STACK, attrs=[nb, 0, nb, 0, 0, 1])
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_file_worker, args=(r, world, port, path, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    nl_ref, G_ref, _ = O.fg(x, M, data)
    for rank, nl, G, (b, e), csum in res:
        assert nl == pytest.approx(float(nl_ref), rel=1e-13) and np.allclose(G, G_ref, rtol=1e-11, atol=1e-9)
        assert csum == S.io.checksum64(np.asfortranarray(M[b:e]))   # the shard each rank read is the shard of the original


def _p2p_worker(rank, world, port, fail_rank, q):
    """init_p2p's agreement protocol on `gloo`: the library calls are replaced by a hook that fails on `fail_rank`."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sfh_b200 as S
        seen = []

        def open_handles(r, blob):
            seen.append((r, len(blob)))
            if r == fail_rank:
                raise ValueError("peer inbox cannot be opened on this rank")

        ok = S.sharding.init_p2p(None, open_handles=open_handles)
        q.put((rank, ok, seen))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 0, 1])
def test_p2p_switch_is_all_or_nothing(fail_rank):
    """if opening the peer inboxes fails on SOME ranks only, those that succeeded must not switch to the one-shot
    exchange (they would spin on flags that never come).  Every rank reports, the group takes the MIN: all switch or none."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, world, port, fail_rank, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = fail_rank < 0
    assert [r[1] for r in res] == [want, want]
    assert all(r[2] == [(r[0], 64 * world)] for r in res)        # every rank saw the whole table of handles


def test_shard_rows_c_abi_matches_host_partition():
    """sfh_shard_rows (what sfh_group_create uses) == sharding.shard_rows; shards tile [0, nbins) on `align` boundaries."""
    import ctypes as C
    import sfh_b200 as S
    b, e = C.c_int64(), C.c_int64()
    for nb in (0, 1, 127, 128, 129, 1000, 60000, 10**6, 10**6 + 7):
        for n in (1, 2, 3, 4, 8):
            for align in (32, 128):
                prev = 0
                for i in range(n):
                    S._lib.check(S._lib.lib.sfh_shard_rows(nb, n, i, align, C.byref(b), C.byref(e)))
                    assert (b.value, e.value) == S.shard_rows(nb, n, i, align=align)
                    assert b.value == prev and e.value >= b.value and (e.value % align == 0 or e.value == nb)
                    prev = e.value
                assert prev == nb
    assert S._lib.lib.sfh_shard_rows(10, 2, 2, 32, C.byref(b), C.byref(e)) != 0
