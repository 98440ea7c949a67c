#!/usr/bin/env python
"""Evidence tool (not the driver's bench.py): BASELINE configs[1] - background-only chains, JLA + DR12 BAO + HST,
1024 parameter points per batch - timed on one GPU.  Prints one JSON line with evaluations/s, the share of the blocked
DMMA Cholesky (K6) and its FP64 rate against the DMMA peak measured by cb200_measure_fp64_peaks.
JLA covariance blocks are the documented synthetic stand-ins (the blobs are absent from the reference checkout)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cosmomc_b200 import lib, datasets as D, synthetic as syn, params as P  # noqa: E402

DATA = os.path.join(ROOT, "tests", "golden", "data")


def main():
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rng = np.random.default_rng(12345)
    bg = P.background_batch(rng.normal(0.02237737, 0.0001, npts), rng.normal(0.1201035, 0.001, npts),
                            rng.normal(67.32, 0.6, npts), rng.normal(147.05, 0.3, npts))
    nuis = np.stack([rng.normal(0.14, 0.01, npts), rng.normal(3.1, 0.1, npts)], axis=1)
    h = lib.Handle(lmax_computed_cl=0, max_points=npts, chunk_points=min(npts, 256))
    zj = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
    jla = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs=syn.synthetic_sn_covs({"zcmb": zj}))
    dr12 = D.BAOPlan(os.path.join(DATA, "DR12", "sdss_DR12Consensus_bao.dataset"))
    hst = D.HSTPlan(os.path.join(DATA, "HST_Riess2018.ini"))
    jla.register(h, 0, 1); dr12.register(h); hst.register(h)

    def step():
        h.set_background(bg)
        return h.loglike_batch(npts, nuis)

    for _ in range(2):
        step()
    h.timing(reset=True)
    h.timer_start()
    t0 = time.time()
    for _ in range(steps):
        ll, tot, st = step()
    ms = h.timer_stop()
    wall = time.time() - t0
    tm = h.timing(reset=True)
    dfma, dmma = h.measure_fp64_peaks()
    n = 740
    flop = npts * steps * (n ** 3 / 3.0 + 3 * 2.0 * n * n)   # potrf + three triangular solves (d, A1, A2)
    out = {"config": "BASELINE configs[1]: JLA (740 SNe, alpha/beta per point) + DR12 BAO + HST, %d points per batch" % npts,
           "evaluations_per_s_device": npts * steps / (ms * 1e-3), "evaluations_per_s_wall": npts * steps / wall,
           "ms_per_step": ms / steps, "ms_background_phase": tm["ms_background"] / steps, "ms_like_phase": tm["ms_like"] / steps,
           "cholesky_gflop_per_point": (n ** 3 / 3.0) / 1e9, "fp64_tflops_if_all_time_were_cholesky": flop / (ms * 1e-3) / 1e12,
           "dmma_peak_tflops": dmma, "dfma_peak_tflops": dfma, "status_nonzero": int((st != 0).sum()),
           "mean_loglike": float(tot.mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
