"""Host restatement (test infrastructure) of the device-resident ensemble sampler sfh_mcmc_run: the same Philox4x32-10
streams (counter = (step, half, walker), stream = draw kind, key = seed) and the same stretch-move arithmetic, driven by
any per-walker log-likelihood (the tests pass the CPU oracle's MCMCModel, mcmc_sample.jl:12-23)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_u01(idx, seed, stream):
    idx = np.asarray(idx, dtype=np.uint64)
    r0, r1, _, _ = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(idx.shape, stream, np.uint64), np.zeros(idx.shape, np.uint64),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = ((r0 << np.uint64(32)) | r1) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


DRAW_STRETCH, DRAW_PARTNER, DRAW_ACCEPT = 16, 17, 18


def counters(step, h, half):
    return (np.uint64(step) << np.uint64(33)) | (np.uint64(h) << np.uint64(32)) | np.arange(half, dtype=np.uint64)


def stretch_move_reference(logl_batch, X0, nsteps, nthin=1, a_scale=2.0, seed=0):
    X = np.array(X0, dtype=np.float64, order="F")
    T, W = X.shape
    half = W // 2
    lp = np.asarray(logl_batch(X), dtype=np.float64).copy()
    chain, lps, acc = [], [], 0
    for step in range(nsteps):
        for h in (0, 1):
            ctr = counters(step, h, half)
            t = (a_scale - 1.0) * philox_u01(ctr, seed, DRAW_STRETCH) + 1.0
            z = t * t / a_scale
            j = np.minimum((philox_u01(ctr, seed, DRAW_PARTNER) * half).astype(np.int64), half - 1)
            act = np.arange(h * half, (h + 1) * half)
            pj = X[:, (1 - h) * half + j]
            P = pj + z[None, :] * (X[:, act] - pj)
            lpp = np.asarray(logl_batch(np.asfortranarray(P)), dtype=np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                lnr = (T - 1) * np.log(z) + lpp - lp[act]
                ok = (np.log(philox_u01(ctr, seed, DRAW_ACCEPT)) < lnr) & np.isfinite(lpp)
            X[:, act[ok]] = P[:, ok]
            lp[act[ok]] = lpp[ok]
            acc += int(ok.sum())
        if (step + 1) % nthin == 0:
            chain.append(X.copy()); lps.append(lp.copy())
    return np.array(chain).reshape(-1, T, W), np.array(lps).reshape(-1, W), X, lp, acc / max(nsteps * W, 1)
