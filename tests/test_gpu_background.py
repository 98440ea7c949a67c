"""GPU parity tests for the background path (BASELINE configs[1]): K5 distances, BAO / MGS /
HST, and the JLA / Pantheon supernova likelihood (K6, blocked DMMA Cholesky), through the C ABI, against the oracle
(itself pinned to the golden .minimum derived parameters and to the reference's python port SN.py).

Tolerances: distances 1e-10 relative (same Romberg sequence; device libm vs glibc in the neutrino table lookup);
-lnL 1e-8 relative (north_star: |Delta lnL| < 0.01)."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
DATA = os.path.join(H.ROOT, "tests", "golden", "data")
NPTS = 5


def draw_bg(npts, seed=12345):
    """BASELINE configs[1] synthetic points: 6 LCDM parameters ~ N(best fit, propose widths)."""
    from cosmomc_b200 import params as P
    rng = np.random.default_rng(seed)
    ombh2 = rng.normal(0.02237737, 0.0001, npts)
    omch2 = rng.normal(0.1201035, 0.001, npts)
    H0 = rng.normal(67.32, 0.6, npts)
    rdrag = rng.normal(147.05, 0.3, npts)
    bg = P.background_batch(ombh2, omch2, H0, rdrag)
    alpha = rng.normal(0.14, 0.01, npts)
    beta = rng.normal(3.1, 0.1, npts)
    return bg, alpha, beta


@pytest.fixture(scope="module")
def handle():
    from cosmomc_b200 import lib
    return lib.Handle(lmax_computed_cl=0, max_points=64, chunk_points=3)  # background-only handle; chunk 3 < NPTS


def test_distances_and_scalars(handle):
    import pyoracle as o
    bg, _, _ = draw_bg(NPTS)
    # an open, a closed and a w != -1 model as well
    bg[1, 4] += 0.02
    bg[2, 4] -= 0.02
    bg[3, 5] = -0.9
    z = np.array([0.01, 0.106, 0.15, 0.38, 0.51, 0.61, 1.3, 2.33, 1089.0])
    DA, Hz, sc = handle.background(bg, z, want_scalars=True)
    for i in range(NPTS):
        da, hz, ex = o.background(bg[i], z)
        assert np.abs(DA[i] / da - 1).max() < 1e-10
        assert np.abs(Hz[i] / hz - 1).max() < 1e-12
        assert np.abs(sc[i] / ex - 1).max() < 1e-9


def test_golden_minimum_direct(handle):
    """the GPU path against the reference's golden derived parameters themselves (7 significant figures)"""
    from cosmomc_b200 import params as P
    bg = P.cmb_to_background(0.2237737E-01, 0.1201035E+00, 0.6732178E+02, rdrag=147.0552)
    z = np.array([0.15, 0.38, 0.51, 0.61, 2.33])
    DA, Hz, sc = handle.background(bg, z, want_scalars=True)
    gH = np.array([0.7265234E+02, 0.8284762E+02, 0.8961363E+02, 0.9527038E+02, 0.2366354E+03])
    gDM = np.array([0.6436578E+03, 0.1534040E+04, 0.1986535E+04, 0.2311050E+04, 0.5763646E+04])
    assert np.abs(Hz[0] * 2.99792458e5 / gH - 1).max() < 5e-7
    assert np.abs(DA[0] * (1 + z) / gDM - 1).max() < 5e-7
    assert abs(sc[0, 1] / 13.79731 - 1) < 5e-7 and abs(100 * sc[0, 2] / 1.040920 - 1) < 1e-5


def test_config2_likelihoods():
    """JLA (synthetic covariance blocks, real light curves) + DR12 BAO + 6DF + MGS + HST + Pantheon"""
    import pyoracle as o
    from cosmomc_b200 import lib, datasets as D, synthetic as syn
    h = lib.Handle(lmax_computed_cl=0, max_points=64, chunk_points=3)
    bg, alpha, beta = draw_bg(NPTS, seed=4)
    zj = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
    jla = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs=syn.synthetic_sn_covs({"zcmb": zj}))
    zp = np.loadtxt(os.path.join(DATA, "Pantheon", "lcparam_full_long_zhel.txt"), usecols=1)
    pan = D.SNPlan(os.path.join(DATA, "Pantheon", "full_long.dataset"),
                   covs=syn.synthetic_sn_covs({"zcmb": zp}, names=("mag",), seed=2025))
    dr12 = D.BAOPlan(os.path.join(DATA, "DR12", "sdss_DR12Consensus_bao.dataset"))
    sdf = D.BAOPlan(os.path.join(DATA, "sdss_6DF_bao.dataset"))
    mgs = D.BAOPlan(os.path.join(DATA, "sdss_MGS_bao.dataset"), tag="MGS")
    hst = D.HSTPlan(os.path.join(DATA, "HST_Riess2018.ini"))
    ids = [jla.register(h, 0, 1), dr12.register(h), sdf.register(h), mgs.register(h), hst.register(h), pan.register(h)]
    assert ids == list(range(6))
    h.set_background(bg)
    nuis = np.stack([alpha, beta], axis=1)
    ll, tot, st = h.loglike_batch(NPTS, nuis)
    sj = o.SN(jla.lc, jla.covs, pecz=jla.pecz, twoscriptmfit=True, scriptmcut=jla.scriptmcut)
    sp = o.SN(pan.lc, pan.covs, pecz=pan.pecz)
    for i in range(NPTS):
        DAj, _, _ = o.background(bg[i], jla.lc["zcmb"])
        DAp, _, _ = o.background(bg[i], pan.lc["zcmb"])
        want = [sj.loglike(DAj, alpha[i], beta[i]),
                o.bao_loglike(bg[i], bg[i, 15], dr12.rs_rescale, dr12.types, dr12.z, dr12.obs, dr12.invcov),
                o.bao_loglike(bg[i], bg[i, 15], sdf.rs_rescale, sdf.types, sdf.z, sdf.obs, sdf.invcov),
                o.mgs_loglike(bg[i], bg[i, 15], mgs.z[0], mgs.alpha_prob),
                o.hst_loglike(bg[i], hst.H0, hst.H0_err),
                sp.loglike(DAp)]
        for k, w in enumerate(want):
            assert abs(ll[i, k] - w) < 1e-8 * max(1.0, abs(w)), (i, k, ll[i, k], w)
        assert abs(tot[i] - sum(want)) < 1e-7 * abs(sum(want))
    t = h.timing()
    assert t["n_launches"] > 0 and t["ms_background"] > 0


@pytest.mark.gpu
def test_sn_cholesky_variants_agree():
    """K6 launch shapes (supernovae_JLA.f90:1074-1168 per point): the 4-warp and the 8-warp instance of the blocked Cholesky, one
    launch for the whole batch or several of 3 points (option "sn_chunk") give the same -lnL (1e-10) and match the oracle."""
    import pyoracle as o
    from cosmomc_b200 import lib, datasets as D, synthetic as syn
    bg, alpha, beta = draw_bg(7, seed=9)
    zj = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
    covs = syn.synthetic_sn_covs({"zcmb": zj})
    nuis = np.stack([alpha, beta], axis=1)
    res = []
    for warps, chunk in ((0, 1024), (8, 1024), (4, 1024), (4, 3), (8, 3)):
        h = lib.Handle(lmax_computed_cl=0, max_points=16, chunk_points=2)
        jla = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs=covs)
        jla.register(h, 0, 1)
        h.set_option("sn_chol_warps", warps)
        h.set_option("sn_chunk", chunk)
        h.set_background(bg)
        ll, tot, st = h.loglike_batch(7, nuis)
        assert (st == 0).all()
        res.append(ll[:, 0].copy())
    for r in res[1:]:
        assert np.abs(r - res[0]).max() < 1e-10 * np.abs(res[0]).max(), np.abs(r - res[0]).max()
    sj = o.SN(jla.lc, jla.covs, pecz=jla.pecz, twoscriptmfit=True, scriptmcut=jla.scriptmcut)
    for i in (0, 6):
        DAj, _, _ = o.background(bg[i], jla.lc["zcmb"])
        w = sj.loglike(DAj, alpha[i], beta[i])
        assert abs(res[2][i] - w) < 1e-8 * max(1.0, abs(w)), (res[2][i], w)
