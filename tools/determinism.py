"""Dev utility (GPU): repeat the semi-slow step on the same resident batch and compare the C_l bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
h.upload_sources(th, n_k, k, syn.make_sources(th, tau, k, pert).numpy())
ref = None
for r in range(reps):
    cls, der, st = h.powers(ip, al)
    if ref is None:
        ref = cls.copy()
    else:
        bad = np.argwhere(cls != ref)
        nz = ref != 0
        print("rep %d: %d differing entries, max rel %.3e%s" % (r, len(bad), np.abs(cls[nz] / ref[nz] - 1).max(),
              (" first at (pt, spec, l) = %s" % (tuple(bad[0]),)) if len(bad) else ""), flush=True)
