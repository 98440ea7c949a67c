"""Dev utility (GPU): run the same batch in N fresh processes per library and compare the C_l bit for bit across processes."""
import os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, hashlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = 192
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
h.upload_sources(th, n_k, k, syn.make_sources(th, tau, k, pert).numpy())
out = []
for r in range(3):
    cls, der, st = h.powers(ip, al)
    out.append(hashlib.md5(cls.tobytes()).hexdigest()[:10])
print(" ".join(out))
''' % (ROOT, ROOT)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for path in sys.argv[2:]:
    seen = {}
    for i in range(N):
        r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CB200_LIB=path), capture_output=True, text=True)
        key = r.stdout.strip() or ("FAILED " + r.stderr[-200:])
        seen[key] = seen.get(key, 0) + 1
    print(os.path.basename(path), seen, flush=True)
