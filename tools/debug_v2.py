"""Dev utility: compare the windowed projection kernel against the L2-gather kernel on one point."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = 1
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
src = syn.make_sources(th, tau, k, pert).numpy()
h.upload_sources(th, n_k, k, src)
h.keep_transfers(True)
D = {}
for pk in (1, 2):
    h.set_option("proj_kernel", pk); h.set_option("count_triples", 1)
    h.timing(reset=True)
    h.powers_resident(ip, al)
    t = h.timing(reset=True)
    nq = len(h.debug_fetch(4, 0))
    D[pk] = h.debug_fetch(3, 0).reshape(nq, 96, 3)
    print("v%d triples %d" % (pk, t["proj_triples"]))
d = np.abs(D[1] - D[2])
sc = np.abs(D[1]).max(axis=(0, 1))
print("max rel diff per source", d.max(axis=(0, 1)) / sc)
bad = np.argwhere(d > 1e-9 * sc[None, None, :])
print("n bad", len(bad), "of", d.size)
if len(bad):
    qs = np.unique(bad[:, 0]); ls = np.unique(bad[:, 1])
    print("bad q idx: n=%d min %d max %d first %s" % (len(qs), qs.min(), qs.max(), qs[:40]))
    print("bad l idx:", ls)
    for b in bad[:10]:
        print(b, D[1][tuple(b)], D[2][tuple(b)])
    print("zero pattern equal:", np.array_equal(D[1] == 0, D[2] == 0))
