"""Dev utility: clean timing of the projection kernel(s) on a small synthetic batch (GPU). usage: k1time.py [n] [kernels]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kernels = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "4")]
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
src = syn.make_sources(th, tau, k, pert).numpy()
h.upload_sources(th, n_k, k, src)
for pk in kernels:
    h.set_option("proj_kernel", pk)
    h.powers_resident(ip, al); h.timing(reset=True)
    for rep in range(2):
        h.powers_resident(ip, al)
        t = h.timing(reset=True)
        print("v%d: project %.2f ms for %d points -> %.1f us/point (spline %.2f us/pt)" % (pk, t["ms_project"], n, 1e3 * t["ms_project"] / n, 1e3 * t["ms_spline"] / n))
