"""Dev utility (GPU): time the projection kernel of every tuning build variants/lib_*.so against the in-tree library and
compare their C_l with it. usage: variants.py [n_points]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "256"
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = int(sys.argv[1]); name = sys.argv[2]
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
src = syn.make_sources(th, tau, k, pert).numpy()
h.upload_sources(th, n_k, k, src)
h.powers_resident(ip, al); h.timing(reset=True)
best = 1e9; bs = 1e9
for rep in range(3):
    h.powers_resident(ip, al)
    t = h.timing(reset=True)
    best = min(best, 1e3 * t["ms_project"] / n); bs = min(bs, 1e3 * t["ms_spline"] / n)
cls, der, st = h.powers(ip, al)
ref = "/tmp/variants_base.npy"
if name == "base":
    np.save(ref, cls); err = 0.0
else:
    b = np.load(ref); nz = b != 0
    err = float(np.abs(cls[nz] / b[nz] - 1).max())
print("%%-10s %%8.1f us/point (+ spline / source-q %%5.1f)   max rel C_l diff vs base %%.2e" %% (name, best, bs, err), flush=True)
''' % (ROOT, ROOT)
libs = [("base", os.path.join(ROOT, "cosmomc_b200", "libcosmob200.so"))]
libs += [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))]
for name, path in libs:
    env = dict(os.environ, CB200_LIB=path)
    r = subprocess.run([sys.executable, "-c", CHILD, n, name], env=env, capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout if r.returncode == 0 else "%-10s FAILED: %s\n" % (name, r.stderr[-400:]))
    sys.stdout.flush()
