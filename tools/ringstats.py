"""Dev utility: windowed-projection statistics on a small synthetic batch (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from cosmomc_b200 import lib, synthetic as syn
n = 64
h = lib.Handle(max_points=n, lmax_out=H.LMAX_OUT, n_tau_max=576, n_k_max=224)
T = H.load_templates(); h.set_templates(T["highl_unlensed"], T["highl_lensed"])
th = syn.draw_thermo(n, 5); ip, al, cal, pert = syn.draw_params(n, 5)
tau, dtau, n_tau, k, n_k = syn.build_grids(h, th)
src = syn.make_sources(th, tau, k, pert).numpy()
h.upload_sources(th, n_k, k, src)
for pk in (4, 3):
    h.set_option("proj_kernel", pk)
    h.set_option("ring_stats", 1); h.set_option("count_triples", 1)
    h.powers_resident(ip, al); h.timing(reset=True)
    h.powers_resident(ip, al)
    t = h.timing(reset=True)
    print("kernel v%d:" % pk, {k_: (round(v, 3) if isinstance(v, float) else v) for k_, v in t.items()})
    pc = np.array(t["phase_cycles"], dtype=float)
    if pc.sum() > 0:
        print("  PHASES prologue/prefetch/barrier/ring/accumulate/metadata %%: %s" % np.round(100 * pc / pc.sum(), 1))
    if t["ring_slabs"]:
        print("  (v4: ring_pairs = slabs whose fill could not be issued ahead)")
        print("  direct frac %.4f rows/slab %.1f pairs/row %.2f triples/pair %.1f" % (t["ring_direct"] / t["ring_slabs"], t["ring_rows"] / t["ring_slabs"], t["ring_pairs"] / max(1, t["ring_rows"]), t["proj_triples"] / max(1, t["ring_pairs"])))
    h.set_option("ring_stats", 0); h.set_option("count_triples", 0)
    h.powers_resident(ip, al); h.timing(reset=True)
    h.powers_resident(ip, al)
    t = h.timing(reset=True)
    print("  CLEAN timing: project %.2f ms for %d points -> %.1f us/point" % (t["ms_project"], n, 1e3 * t["ms_project"] / n))
