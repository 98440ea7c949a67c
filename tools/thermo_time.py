"""Dev utility (GPU): throughput of the batched thermal history (cb200_thermo) and of the theta -> H0 bisection.
usage: thermo_time.py [n_points ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cosmomc_b200 import lib, params as P
ns = [int(a) for a in sys.argv[1:]] or [1024, 8192]
h = lib.Handle(lmax_computed_cl=0, max_points=8)
rng = np.random.default_rng(1)
for n in ns:
    ombh2 = 0.02237 * (1 + 0.02 * rng.standard_normal(n)); omch2 = 0.12 * (1 + 0.03 * rng.standard_normal(n))
    H0 = 67.3 * (1 + 0.03 * rng.standard_normal(n)); tau = np.clip(0.055 + 0.01 * rng.standard_normal(n), 0.02, 0.2)
    bg = P.background_batch(ombh2, omch2, H0)
    h.thermo(bg, 0.245, optical_depth=tau)   # first call: allocates the work tables
    t = time.time(); out, st = h.thermo(bg, 0.245, optical_depth=tau); dt = time.time() - t
    print("thermo: %6d points %8.1f ms  -> %7.1f us/point (%d failed); rdrag %.3f +- %.3f" % (n, 1e3 * dt, 1e6 * dt / n, (st != 0).sum(), out[:, 18].mean(), out[:, 18].std()), flush=True)
    omnuh2 = bg[0, 3] * (H0[0] / 100) ** 2
    t = time.time(); b2 = h.theta_to_background(ombh2, omch2, np.full(n, 1.0409), omnuh2, bg[0, 7:15]); dt = time.time() - t
    print("theta->H0: %6d points %8.1f ms -> %7.1f us/point; H0 %.3f +- %.3f" % (n, 1e3 * dt, 1e6 * dt / n, b2[:, 0].mean(), b2[:, 0].std()), flush=True)
