"""One device build of 240 covariant templates on a 200 x 300 grid (run under ncu for the scatter kernel's profile)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
T = S.templates
dmod = 25.0
edges = (np.linspace(-0.2, 1.2, 201), np.linspace(dmod - 6.0, dmod + 5.0, 301))
comp = [lambda m, m50=m50: T.Martin2016_complete(m, 1.0, m50, 0.7) for m50 in (28.5, 27.5)]
err = [lambda m, c=c: np.minimum(T.exp_photerr(m, 1.03, 15.0, c, 0.02), 0.4) for c in (36.0, 35.0)]
imf = lambda m: np.asarray(m) ** -2.35 / 11.0
m = np.linspace(0.1, 1.4, 400)
isos = [(m, [-0.6 - 8.2 * np.log10(m) + 0.02 * a, -1.0 - 7.5 * np.log10(m) + 0.03 * a]) for a in range(240)]
pts = [S.template_points(mm, mg, err, 1, (0, 1), imf, comp, None, dmod, 1e7, 0.35, edges) for (mm, mg) in isos]
S.DeviceStack.from_points(edges, pts)
