"""Dev utility (GPU): config-2 likelihood phase of every tuning build variants/lib_*.so against the in-tree library.
usage: k6variants.py [npts]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "1024"
CHILD = r'''
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cosmomc_b200 import lib, datasets as D, synthetic as syn, params as P
DATA = os.path.join(ROOT, "tests", "golden", "data")
npts = int(sys.argv[1]); name = sys.argv[2]
rng = np.random.default_rng(12345)
bg = P.background_batch(rng.normal(0.02237737, 0.0001, npts), rng.normal(0.1201035, 0.001, npts),
                        rng.normal(67.32, 0.6, npts), rng.normal(147.05, 0.3, npts))
nuis = np.stack([rng.normal(0.14, 0.01, npts), rng.normal(3.1, 0.1, npts)], axis=1)
zj = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
covs = syn.synthetic_sn_covs({"zcmb": zj})
for warps in (0, 8):
    h = lib.Handle(lmax_computed_cl=0, max_points=npts, chunk_points=min(npts, 256))
    jla = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs=covs)
    jla.register(h, 0, 1)
    h.set_option("sn_chol_warps", warps)
    h.set_background(bg)
    for _ in range(2):
        ll, tot, st = h.loglike_batch(npts, nuis)
    h.timing(reset=True)
    for _ in range(3):
        ll, tot, st = h.loglike_batch(npts, nuis)
    t = h.timing(reset=True)
    print("%%-10s warps %%d: like %%.3f ms per %%d points = %%.2f us/pt  sum lnL %%.9f" %% (name, warps, t["ms_like"] / 3, npts, 1e3 * t["ms_like"] / 3 / npts, tot.sum()), flush=True)
    del h
''' % ROOT
libs = [("base", os.path.join(ROOT, "cosmomc_b200", "libcosmob200.so"))]
libs += [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))]
for name, path in libs:
    env = dict(os.environ, CB200_LIB=path)
    r = subprocess.run([sys.executable, "-c", CHILD, n, name], env=env, capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout if r.returncode == 0 else "%-10s FAILED: %s\n" % (name, r.stderr[-400:]))
    sys.stdout.flush()
