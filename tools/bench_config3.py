#!/usr/bin/env python
"""Evidence tool (not the driver's bench.py): BASELINE configs[2] - BK15 B-mode likelihood with CAMB tensors, fixed
cosmology (block_semi_fast: transfer functions shared by the batch), r / n_t and the foreground parameters vary,
lensed BB to l = 600, 4096-point batch.  One JSON line: evaluations/s of cb200_powers_shared + cb200_loglike_batch.
The 702 x 702 covariance is the documented synthetic stand-in (the blob is absent from the reference checkout)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers as H  # noqa: E402
from cosmomc_b200 import lib  # noqa: E402
from cosmomc_b200.datasets import BK15Plan  # noqa: E402

P0 = np.array([3.0, 1.0, -0.42, 1.59, 19.6, -0.6, -3.1, 0.2, 2.0, 2.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0])


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    T = H.load_templates()
    h = lib.Handle(max_points=B, chunk_points=min(B, 1024), lmax_out=H.LMAX_OUT, compute_tensors=1, lmax_tensor=H.MAX_L_T)
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    plan = BK15Plan.from_pack(os.path.join(ROOT, "tests", "golden", "bk15_pack.npz"))
    plan.register(h, nuis_offset=0)
    batch = H.small_batch(1, seed=21, NT=h.info.n_tau_max, NK=h.info.n_k_max)
    tb = H.small_batch_tensor(batch["thermo"], seed=21, NT=h.cfg.n_tau_max_tensor, NK=h.cfg.n_k_max_tensor)
    h.upload_sources(batch["thermo"], batch["n_k"], batch["k"], batch["src"])
    h.upload_sources(tb["thermo"], tb["n_k"], tb["k"], tb["src"], kind=1)
    rng = np.random.default_rng(3)
    ip = np.tile(batch["initpower"][0], (B, 1))
    ip[:, 4] = rng.uniform(0.0, 0.5, B)      # r
    ip[:, 5] = rng.uniform(-1.0, 1.0, B)     # n_t
    ip[:, 9] = 0.0
    nuis = np.tile(P0, (B, 1))
    nuis[:, 0] = rng.uniform(2.0, 6.0, B)
    nuis[:, 1] = rng.uniform(0.0, 3.0, B)
    nuis[:, 2] = rng.uniform(-0.8, -0.2, B)
    nuis[:, 3] = rng.normal(1.59, 0.11, B)
    nuis[:, 5] = rng.uniform(-1.0, -0.2, B)
    nuis[:, 6] = rng.normal(-3.1, 0.3, B)
    nuis[:, 7] = rng.uniform(-0.5, 0.5, B)
    al = np.ones(B)

    def step():
        h.powers_shared(ip, al, src_point=0, first=0, want_cls=False)
        return h.loglike_batch(B, nuis)

    step()
    h.timing(reset=True)
    h.timer_start()
    t0 = time.time()
    for _ in range(steps):
        ll, tot, st = step()
    ms = h.timer_stop()
    wall = time.time() - t0
    tm = h.timing(reset=True)
    out = {"config": "BASELINE configs[2]: BK15 (HL, 12 B-mode maps, foreground model) + tensors, shared transfers, %d points" % B,
           "evaluations_per_s_device": B * steps / (ms * 1e-3), "evaluations_per_s_wall": B * steps / wall,
           "ms_per_step": ms / steps, "phase_ms_per_step": {k: tm[k] / steps for k in
                                                            ["ms_spline", "ms_project", "ms_contract", "ms_interp", "ms_lens", "ms_like"]},
           "status_nonzero": int((st != 0).sum()), "mean_loglike": float(np.mean(tot))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
