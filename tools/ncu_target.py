"""Dev utility (GPU): one pass of the hot path over a 64-point batch - the command the ncu captures under profiles/ wrap.
usage: ncu_target.py [n_points]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates()
h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5)
ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
src = syn.make_sources(th, tau, k, pert).numpy()
h.upload_sources(th, n_k, k, src)
for rep in range(2):
    h.powers_resident(ip, al)
h.sync()
print("ok", h.timing()["n_launches"])
