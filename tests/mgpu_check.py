"""Multi-GPU parity check, one process per GPU:  torchrun --nproc-per-node N tests/mgpu_check.py
Bin-row shards + in-library NCCL all-reduce must reproduce the single-GPU whole-stack answers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sfh_b200 as S
from conftest import make_flat_problem, make_hier_problem


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb, nt = 20011, 300
    M, x, data = make_flat_problem(nb, nt, seed=99)
    whole = S.DeviceStack(M, data, device=local)
    nl0, G0, _ = whole.eval_fg(x * 1.1)
    shard = S.DeviceStack(M, data, device=local, rows=S.shard_rows(nb, world, rank))
    S.init_library_comm(shard.ctx())
    for _ in range(3):
        nl, G, _ = shard.eval_fg(x * 1.1)
        assert abs(nl - nl0) <= 1e-12 * abs(nl0), (nl, nl0)
        assert np.allclose(G, G0, rtol=1e-10, atol=1e-10 * np.abs(G0).max())
    nl_f, _, _ = shard.eval_fg(x * 1.1, want_G=False)
    assert abs(nl_f - nl0) <= 1e-12 * abs(nl0)
    # the exchange the fused path uses must be the one-shot NVLink one (mode 2) unless it was switched off
    import ctypes as C
    mode = C.c_int()
    S._lib.check(S._lib.lib.sfh_ctx_comm_info(shard.ctx().handle, None, None, C.byref(mode)))
    assert mode.value == (1 if os.environ.get("SFH_NO_P2P") == "1" else 2), mode.value
    # many evaluations back to back: epochs / inbox parities / the replayed graph, alternating gradient and logL-only calls at
    # changing coefficients; every one must match the whole stack and be bit-identical across ranks
    rng_s = np.random.default_rng(123)
    for it in range(40):
        xi = x * (1.0 + 0.2 * rng_s.random(nt))
        wg = it % 3 != 2
        a = whole.eval_fg(xi, want_G=wg); b = shard.eval_fg(xi, want_G=wg)
        assert abs(a[0] - b[0]) <= 1e-12 * abs(a[0]), (it, a[0], b[0])
        if wg:
            assert np.allclose(a[1], b[1], rtol=1e-10, atol=1e-10 * np.abs(a[1]).max()), it
            t = torch.tensor([b[0]] + list(b[1]), dtype=torch.float64, device="cuda")
            ref = t.clone(); dist.broadcast(ref, 0)
            assert torch.equal(t, ref), it
    # every rank must hold the SAME reduced answer bit for bit (all-reduce is rank-symmetric)
    t = torch.tensor([nl] + list(G[:8]), dtype=torch.float64, device="cuda")
    ref = t.clone(); dist.broadcast(ref, 0)
    assert torch.equal(t, ref)
    # batched walkers over shards
    X = np.maximum(0.0, x[:, None] + np.random.default_rng(1).standard_normal((nt, 33)))
    X[5, 2] = -1.0
    a = whole.eval_logl_batched(X); b = shard.eval_logl_batched(X)
    assert b[2] == -np.inf and np.allclose(np.delete(a, 2), np.delete(b, 2), rtol=1e-12)
    # fg! for several coefficient vectors at once over shards (logL and the T x C gradient are all-reduced)
    nla, Ga_ = whole.eval_fg_batched(X[:, 3:12]); nlb, Gb_ = shard.eval_fg_batched(X[:, 3:12])
    assert np.allclose(nla, nlb, rtol=1e-12) and np.allclose(Ga_, Gb_, rtol=1e-9, atol=1e-10 * np.abs(Ga_).max())
    # the device-resident ensemble sampler over shards: same Philox streams on every rank, all-reduced logL ->
    # every rank walks the SAME chain (bit for bit), and it is the single-GPU chain
    X0 = np.maximum(0.0, x[:, None] + np.random.default_rng(4).standard_normal((nt, 24)))
    ca, la_, Xa, lfa, acca = whole.mcmc_run(X0, 6, 2, 2.0, seed=77)
    cb, lb_, Xb, lfb, accb = shard.mcmc_run(X0, 6, 2, 2.0, seed=77)
    assert acca == accb and np.allclose(ca, cb, rtol=1e-12, atol=0) and np.allclose(la_, lb_, rtol=1e-11)
    t = torch.tensor(np.ascontiguousarray(Xb).ravel(), dtype=torch.float64, device="cuda")
    ref = t.clone(); dist.broadcast(ref, 0)
    assert torch.equal(t, ref)
    # a stack built on the device from point lists, each rank building only its rows
    rng = np.random.default_rng(6)
    xe, ye = np.linspace(0.0, 1.0, 41), np.linspace(20.0, 25.0, 61)
    pls = [(rng.uniform(0, 1, 80), rng.uniform(20, 25, 80), rng.uniform(0.01, 0.08, 80), rng.uniform(0.03, 0.3, 80),
            rng.uniform(1, 5, 80), (0, 1, -1)[t % 3]) for t in range(9)]
    dpt = rng.poisson(30.0, 40 * 60).astype(np.float64)
    wp = S.DeviceStack.from_points((xe, ye), pls, data=dpt, device=local)
    sp = S.DeviceStack.from_points((xe, ye), pls, data=dpt, device=local, rows=S.shard_rows(40 * 60, world, rank))
    S.init_library_comm(sp.ctx())
    xc = rng.uniform(0.5, 2.0, 9)
    (fa, ga, _), (fb, gb, _) = wp.eval_fg(xc), sp.eval_fg(xc)
    assert abs(fa - fb) <= 1e-12 * abs(fa) and np.allclose(ga, gb, rtol=1e-10, atol=1e-10 * np.abs(ga).max())
    # hierarchical path over shards
    p = make_hier_problem(nj=12, nk=10, nb=5003)
    mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    xt = S.calculate_coeffs(mz, dp, p["R"], p["logAge"], p["MH"])
    d = np.random.default_rng(2).poisson(p["M"] @ xt).astype(np.float64)
    v = np.concatenate([p["R"], [1.0, -2.0, 0.2]]) * 1.07
    w2 = S.DeviceStack(p["M"], d, device=local)
    s2 = S.DeviceStack(p["M"], d, device=local, rows=S.shard_rows(5003, world, rank))
    S.init_library_comm(s2.ctx())
    Ga, Gb = np.empty(15), np.empty(15)
    na = S.fg_(True, Ga, mz, dp, v, w2, d, None, p["logAge"], p["MH"])
    nb_ = S.fg_(True, Gb, mz, dp, v, s2, d, None, p["logAge"], p["MH"])
    assert abs(na - nb_) <= 1e-12 * abs(na) and np.allclose(Ga, Gb, rtol=1e-8, atol=1e-10 * np.abs(Ga).max())
    # ... and for several variable vectors at once (sfh_eval_fg_hier_batched over shards)
    Vb = v[:, None] * (1 + 0.03 * np.random.default_rng(8).standard_normal((15, 7)))
    nla2, Gha = S.hierarchical.fg_batched_(mz, dp, Vb, w2, d, p["logAge"], p["MH"])
    nlb2, Ghb = S.hierarchical.fg_batched_(mz, dp, Vb, s2, d, p["logAge"], p["MH"])
    assert np.allclose(nla2, nlb2, rtol=1e-12) and np.allclose(Gha, Ghb, rtol=1e-8, atol=1e-10 * np.abs(Gha).max())
    g1 = np.empty(15); n1 = S.fg_(True, g1, mz, dp, Vb[:, 4], s2, d, None, p["logAge"], p["MH"])
    assert abs(n1 - nlb2[4]) <= 1e-12 * abs(n1) and np.allclose(g1, Ghb[:, 4], rtol=1e-8, atol=1e-10 * np.abs(g1).max())
    # the counter-based synthetic generator: shards of one big diagram == the whole
    xs = 10 * np.random.default_rng(3).random(500)
    ws = S.DeviceStack.synthetic(40000, 500, np.float32, 7, 1.0, xs, device=local)
    ss = S.DeviceStack.synthetic(40000, 500, np.float32, 7, 1.0, xs, device=local, rows=S.shard_rows(40000, world, rank))
    S.init_library_comm(ss.ctx())
    r0, r1 = ws.eval_fg(xs * 1.01), ss.eval_fg(xs * 1.01)
    assert abs(r0[0] - r1[0]) <= 1e-12 * abs(r0[0]) and np.allclose(r0[1], r1[1], rtol=1e-9, atol=1e-9 * np.abs(r0[1]).max())
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check OK on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
