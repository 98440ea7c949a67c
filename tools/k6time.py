"""Dev utility (GPU): time the config-2 likelihood phase (JLA Cholesky) for combinations of warps per CTA and chunk size.
usage: k6time.py [npts]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cosmomc_b200 import lib, datasets as D, synthetic as syn, params as P
DATA = os.path.join(ROOT, "tests", "golden", "data")
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(12345)
bg = P.background_batch(rng.normal(0.02237737, 0.0001, npts), rng.normal(0.1201035, 0.001, npts),
                        rng.normal(67.32, 0.6, npts), rng.normal(147.05, 0.3, npts))
nuis = np.stack([rng.normal(0.14, 0.01, npts), rng.normal(3.1, 0.1, npts)], axis=1)
zj = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
covs = syn.synthetic_sn_covs({"zcmb": zj})
ref = None
for warps, chunk in ((8, 256), (8, 296), (8, 1024), (4, 256), (4, 512), (4, 592), (4, 1024)):
    h = lib.Handle(lmax_computed_cl=0, max_points=npts, chunk_points=min(npts, chunk))
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
    if ref is None: ref = tot.copy()
    print("warps %d chunk %4d: like %.3f ms per %d points = %.2f us/pt   max |dlnL| vs first %.2e" %
          (warps, chunk, t["ms_like"] / 3, npts, 1e3 * t["ms_like"] / 3 / npts, np.abs(tot - ref).max()), flush=True)
    del h
