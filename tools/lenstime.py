"""Dev utility (GPU): lensing-phase time of every tuning build variants/lib_*.so against the in-tree library (1 024 points)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = 1024; name = sys.argv[1]
h = lib.Handle(max_points=n, chunk_points=1024, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
import torch
src = syn.make_sources(th, tau, k, pert, device="cuda:0")
torch.cuda.synchronize()
h.upload_sources(th, n_k, k, None, src_device_ptr=src.data_ptr())
h.powers_resident(ip, al); h.timing(reset=True)
best = 1e9
for rep in range(3):
    h.powers_resident(ip, al)
    t = h.timing(reset=True)
    best = min(best, 1e3 * t["ms_lens"] / n)
cls, der, st = h.powers(ip, al)
print("%%-8s lens %%.3f us/point  checksum %%.12e" %% (name, best, float(np.abs(cls).sum())), flush=True)
''' % (ROOT, ROOT)
libs = [("base", os.path.join(ROOT, "cosmomc_b200", "libcosmob200.so"))]
libs += [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))]
for name, path in libs:
    env = dict(os.environ, CB200_LIB=path)
    r = subprocess.run([sys.executable, "-c", CHILD, name], env=env, capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout if r.returncode == 0 else "%-10s FAILED: %s\n" % (name, r.stderr[-400:]))
