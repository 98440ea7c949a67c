"""BK15 (BASELINE configs[2]): Hamimeche-Lewis likelihood over 12 B-mode maps + the BICEP/Keck foreground model.

CPU: the oracle's HL path on the real BK15 band powers / noise / fiducial / windows (fixture bk15_pack.npz, made by
tests/golden/make_golden_bk15.py) reproduces the chi^2 of the reference's own python port python/CMBlikes.py at the
golden Planck best-fit C_l (foreground-free theory; the covariance blob is missing from the reference checkout, so both
sides use the documented synthetic stand-in).  The foreground model has no golden vector in the reference (no port,
test.ini's 663.808 needs the missing blobs): literal restatement, "parity unpinned" for row a17.
GPU: the CUDA path through the C ABI against the oracle, foregrounds on, band-centre errors, decorrelation."""
import os

import numpy as np
import pytest

import helpers as H

PACK = os.path.join(H.ROOT, "tests", "golden", "bk15_pack.npz")
# batch3/BK15.ini centres: BBdust BBsync alphadust betadust Tdust alphasync betasync corr EEtoBB_d EEtoBB_s Dd Ds g_corr g95 g150 g220
P0 = np.array([3.0, 1.0, -0.42, 1.59, 19.6, -0.6, -3.1, 0.2, 2.0, 2.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0])


@pytest.fixture(scope="module")
def plan():
    from cosmomc_b200.datasets import BK15Plan
    return BK15Plan.from_pack(PACK)


def theory_cls(templates, r_scale=1.0):
    cls = np.zeros((5, H.LMAX_OUT + 1))
    cls[:, :] = templates["theory_cl"].T[:, :H.LMAX_OUT + 1]
    cls[3] *= r_scale
    return cls


def test_oracle_hl_vs_reference_port(plan, templates):
    import pyoracle as o
    want = float(np.load(PACK)["port_chi2_nofg"])
    binned = plan.binned_theory(theory_cls(templates))
    got = o.cmblikes_chisq(plan.nmaps, plan.nbins_used, plan.cl_use_index, plan.like_approx, plan.noise, plan.chat,
                           plan.sqrt_fid, plan.invcov, binned)
    assert plan.nmaps == 12 and plan.ncl_used == 78 and plan.nbins_used == 9 and plan.like_approx == 1
    assert abs(got / want - 1) < 1e-9, (got, want)


def test_foreground_model_properties(plan):
    """size-independent properties of the restated foreground model"""
    import pyoracle as o
    fg = o.bk_foregrounds(plan, P0)
    # auto spectra are positive, pivot-frequency scaling: dust at 353 GHz ~ A_dust (l/80)^alpha within the bandpass colour correction
    i353 = plan.used_map_order.index("P353_B")
    c = i353 * (i353 + 1) // 2 + i353
    assert np.all(fg[c, plan.pcl_lmin:] > 0)
    assert abs(fg[c, 80] / P0[0] - 1) < 0.35
    # linear in the amplitudes when the correlation is off
    Pa = P0.copy(); Pa[7] = 0.0
    Pb = Pa.copy(); Pb[0] *= 2; Pb[1] *= 2
    assert np.allclose(o.bk_foregrounds(plan, Pb), 2 * o.bk_foregrounds(plan, Pa), rtol=1e-13)
    # decorrelation leaves auto spectra untouched and suppresses cross spectra
    Pd = P0.copy(); Pd[10] = 0.9
    fd = o.bk_foregrounds(plan, Pd)
    assert np.array_equal(fd[c], fg[c])
    i217 = plan.used_map_order.index("P217_B")
    cx = i353 * (i353 + 1) // 2 + i217
    assert np.all(fd[cx, plan.pcl_lmin:] < fg[cx, plan.pcl_lmin:])
    # band-centre error only touches the BK bands
    Pg = P0.copy(); Pg[13] = 0.02
    fgm = o.bk_foregrounds(plan, Pg)
    i95 = plan.used_map_order.index("BK15_95_B")
    assert not np.allclose(fgm[i95 * (i95 + 1) // 2 + i95], fg[i95 * (i95 + 1) // 2 + i95], rtol=1e-6)
    assert np.array_equal(fgm[c], fg[c])


@pytest.mark.gpu
def test_gpu_bk15_loglike(plan, templates):
    import pyoracle as o
    from cosmomc_b200 import lib
    h = lib.Handle(max_points=16, chunk_points=8, lmax_out=H.LMAX_OUT)
    lid = plan.register(h, nuis_offset=1)      # nuisance vector: [dummy, 16 BK parameters]
    rng = np.random.default_rng(3)
    B = 6
    P = np.tile(P0, (B, 1))
    P[:, 0] = rng.uniform(2.0, 6.0, B)
    P[:, 1] = rng.uniform(0.0, 3.0, B)
    P[:, 2] = rng.uniform(-0.8, -0.2, B)
    P[:, 3] = rng.normal(1.59, 0.11, B)
    P[:, 5] = rng.uniform(-1.0, -0.2, B)
    P[:, 6] = rng.normal(-3.1, 0.3, B)
    P[:, 7] = rng.uniform(-0.5, 0.5, B)
    P[1, 10], P[1, 11] = 0.92, 0.97           # decorrelation on (lin form in BK15_dust.dataset)
    P[2, 12:16] = [0.01, 0.02, -0.015, 0.01]  # band-centre errors
    P[3, 10] = 1.04                           # non-physical branch Delta > 1
    rs = rng.uniform(0.5, 3.0, B)
    cls = np.stack([theory_cls(templates, r) for r in rs])
    nuis = np.concatenate([np.zeros((B, 1)), P], axis=1)
    ll, tot, st = h.loglike_cls(cls, nuis)
    for i in range(B):
        want = o.bk_loglike(plan, cls[i], P[i])
        assert abs(ll[i, lid] / want - 1) < 1e-9, (i, ll[i, lid], want)   # north_star: |Delta lnL| < 0.01


@pytest.mark.gpu
def test_gpu_bk15_foregrounds_gemm_form_matches_scalar_kernel(plan, templates):
    """Without frequency decorrelation the foreground band powers go through one GEMM for the batch (three l-shapes per
    point x the transposed windows); with it, or with option bk_scalar_foregrounds, through the per-point kernel.  Both
    against the oracle, and against each other, incl. band-centre errors and E/B ratios."""
    import pyoracle as o
    from cosmomc_b200 import lib
    h = lib.Handle(max_points=16, chunk_points=8, lmax_out=H.LMAX_OUT)
    lid = plan.register(h, nuis_offset=0)
    rng = np.random.default_rng(11)
    B = 5
    P = np.tile(P0, (B, 1))
    P[:, 0] = rng.uniform(2.0, 6.0, B); P[:, 1] = rng.uniform(0.0, 3.0, B); P[:, 2] = rng.uniform(-0.8, -0.2, B)
    P[:, 3] = rng.normal(1.59, 0.11, B); P[:, 5] = rng.uniform(-1.0, -0.2, B); P[:, 6] = rng.normal(-3.1, 0.3, B)
    P[:, 7] = rng.uniform(-0.5, 0.5, B); P[:, 8] = rng.uniform(1.5, 2.5, B); P[:, 9] = rng.uniform(1.5, 2.5, B)
    P[2, 12:16] = [0.01, 0.02, -0.015, 0.01]
    assert np.all(P[:, 10:12] == 1.0)                      # no decorrelation: the GEMM form is taken
    cls = np.stack([theory_cls(templates, r) for r in rng.uniform(0.5, 3.0, B)])
    ll_gemm, _, _ = h.loglike_cls(cls, P)
    h.set_option("bk_scalar_foregrounds", 1)
    ll_scalar, _, _ = h.loglike_cls(cls, P)
    assert np.abs(ll_gemm[:, lid] / ll_scalar[:, lid] - 1).max() < 1e-11
    for i in range(B):
        want = o.bk_loglike(plan, cls[i], P[i])
        assert abs(ll_gemm[i, lid] / want - 1) < 1e-9
