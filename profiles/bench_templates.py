"""Template construction: the whole config-3 grid (60 ages x 40 metallicities = 2400 templates, 200 x 300 bins) built on
the device from isochrone point lists vs the CPU oracle's bin_cmd_smooth (1 thread; the reference reports ~1.25 ms per
template in bin_cmd_smooth for a 75 x 100 diagram, test/templates/template_test.jl:109-111)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
import oracle as O
T = S.templates

def isochrone(age_shift, mh_shift, n=400):
    m = np.linspace(0.1, 1.4, n)
    F1 = -0.6 - 8.2 * np.log10(m) + 0.05 * np.sin(6 * m) + 0.02 * age_shift + 0.10 * mh_shift
    F2 = -1.0 - 7.5 * np.log10(m) + 0.03 * age_shift
    return m, [F1, F2]

def main(nx=200, ny=300, nage=60, nmh=40, cov_y=1):
    dmod = 25.0
    edges = (np.linspace(-0.2, 1.2, nx + 1), np.linspace(dmod - 6.0, dmod + 5.0, ny + 1))
    comp = [lambda m, m50=m50: T.Martin2016_complete(m, 1.0, m50, 0.7) for m50 in (28.5, 27.5)]
    err = [lambda m, c=c: np.minimum(T.exp_photerr(m, 1.03, 15.0, c, 0.02), 0.4) for c in (36.0, 35.0)]
    imf = lambda m: np.asarray(m) ** -2.35 / 11.0
    isos = [isochrone(a / 10.0, k / 10.0) for a in range(nage) for k in range(nmh)]
    t0 = time.perf_counter()
    pts = [S.template_points(m, mg, err, cov_y, (0, 1), imf, comp, None, dmod, 1e7, 0.35, edges) for (m, mg) in isos]
    t_prep = time.perf_counter() - t0
    npts = sum(len(p[0]) for p in pts)
    import gc
    S.DeviceStack.from_points(edges, pts[:8])                                   # warm-up (context, module load)
    t_dev = float("inf")
    for _ in range(3):                                                          # best of 3; earlier stacks are freed OUTSIDE the timed region
        ds = None
        gc.collect()
        t0 = time.perf_counter()
        ds = S.DeviceStack.from_points(edges, pts)
        t_dev = min(t_dev, time.perf_counter() - t0)
    # CPU oracle on a sample of templates, 1 thread
    sample = list(range(0, len(pts), max(1, len(pts) // 24)))[:24]
    t0 = time.perf_counter()
    cols = {k: O.bin_cmd_smooth(*pts[k][:4], pts[k][5], pts[k][4], nx, edges[0][0], edges[0][1] - edges[0][0], ny, edges[1][0],
                                edges[1][1] - edges[1][0]) for k in sample}
    t_cpu = (time.perf_counter() - t0) / len(sample)
    M, _ = ds.download()
    err_max = max(np.abs(M[:, k] - cols[k].reshape(-1, order="F")).max() / cols[k].max() for k in sample)
    # what the upload path costs for the same stack (host-built templates -> sfh_stack_create)
    gc.collect()
    t0 = time.perf_counter()
    up = S.DeviceStack(M, np.zeros(M.shape[0]))
    t_up = time.perf_counter() - t0
    del up
    print(json.dumps({"what": f"build {len(pts)} templates of {nx}x{ny} bins from {npts} isochrone points (cov_mult={pts[0][5]})",
                      "host_point_prep_s": t_prep, "device_build_s": t_dev, "templates_per_s_device": len(pts) / t_dev,
                      "cpu_oracle_ms_per_template_1thread": t_cpu * 1e3, "cpu_oracle_s_all_templates_1thread": t_cpu * len(pts),
                      "upload_of_host_built_stack_s": t_up, "max_rel_err_vs_oracle": err_max}), flush=True)

if __name__ == "__main__":
    main(cov_y=1)     # y = second filter of the colour: covariant kernel
    main(nx=75, ny=100, nage=24, nmh=10, cov_y=1)
