"""Run one forced kernel configuration of the fused kernel on the config-3 shape (for ncu captures).
usage: one_config.py NW BT C VARIANT [reps [NB NT DTYPE]]   (0 = let the library choose)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
nw, bt, c, var = map(int, sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
nb, nt = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (60000, 2400)
dt = np.dtype(sys.argv[8] if len(sys.argv) > 8 else 'float64')
x = 100 * np.random.default_rng(0).random(nt)
ds = S.DeviceStack.synthetic(nb, nt, dt.type, seed=1, scale=1.0, x_true=x, tile_bins=bt, cluster=c, consumer_warps=nw, variant=var)
i = ds.info()
want_g = os.environ.get('SFH_WANT_G', '1') != '0'
ms, msk = ds.time_fg(x, reps=reps, want_G=want_g, flush_l2=False)
print(f"nb={nb} nt={nt} {dt} GB/s={nb*nt*dt.itemsize/msk/1e6:.0f} nw={i.consumer_warps} variant={i.variant} bt={i.tile_bins} c={i.cluster} kt={i.chunks_per_tile} ring={i.ring_slots} ncl={i.n_clusters} kernel={msk*1e3:.1f} us eval={ms*1e3:.1f} us want_G={int(want_g)}")
