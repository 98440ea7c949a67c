"""CPU tests (no GPU): pin the ORACLE's background functions and background-only likelihoods
against the reference's golden vectors, so that GPU-vs-oracle parity means GPU-vs-reference parity.

Pinned here
  * H(z), D_M(z) at z = 0.15, 0.38, 0.51, 0.61, 2.33, the age and Omega_Lambda / omeganuh2 of the Planck 2018 best fit:
    the derived-parameter block of data/base_plikHM_TTTEEE_lowl_lowE.minimum (7 significant figures), produced by the
    reference (CAMB background functions, camb/modules.f90:519-751, massive neutrino, normal hierarchy)
  * 100 theta_MC of the same file (CosmomcTheta; the file's H0 is rounded to 7 figures, so 1e-5)
  * JLA / Pantheon -lnL against the reference's Python port python/planck/SN.py run on the real light-curve tables
    with the documented synthetic covariance stand-ins (fixture tests/golden/sn_py.npz, make_golden_sn.py)
  * host-side .dataset readers for DR12 BAO / 6DF / MGS / HST / JLA / Pantheon against the shipped files
"""
import os

import numpy as np
import pytest

import helpers as H

DATA = os.path.join(H.ROOT, "tests", "golden", "data")
# data/base_plikHM_TTTEEE_lowl_lowE.minimum: parameters 1,2,72 and derived 73,76,97,110-119, 3
BESTFIT = dict(ombh2=0.2237737E-01, omch2=0.1201035E+00, H0=0.6732178E+02, rdrag=0.1470552E+03)
GOLD_Z = np.array([0.15, 0.38, 0.51, 0.61, 2.33])
GOLD_H = np.array([0.7265234E+02, 0.8284762E+02, 0.8961363E+02, 0.9527038E+02, 0.2366354E+03])
GOLD_DM = np.array([0.6436578E+03, 0.1534040E+04, 0.1986535E+04, 0.2311050E+04, 0.5763646E+04])


@pytest.fixture(scope="module")
def o():
    import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def bg():
    from cosmomc_b200 import params as P
    return P.cmb_to_background(BESTFIT["ombh2"], BESTFIT["omch2"], BESTFIT["H0"], rdrag=BESTFIT["rdrag"])


def test_parameter_mapping_matches_golden_derived(bg):
    h2 = (bg[0] / 100) ** 2
    assert abs(bg[4] / 0.6842033 - 1) < 5e-7           # omegal (inputs carry 7 figures)
    assert abs(bg[3] * h2 / 0.6451439E-03 - 1) < 2e-7  # omeganuh2
    assert bg[8] == 1 and abs(bg[7] - (3.046 - 3.046 / 3)) < 1e-12  # one massive eigenstate (mnu = 0.06, normal)


def test_background_golden_minimum(o, bg):
    DA, Hz, ex = o.background(bg, GOLD_Z)
    assert np.abs(Hz * o.CONST_C / 1e3 / GOLD_H - 1).max() < 5e-7   # 7 significant figures in the file
    assert np.abs(DA * (1 + GOLD_Z) / GOLD_DM - 1).max() < 5e-7
    assert abs(ex[1] / 0.1379731E+02 - 1) < 5e-7                   # age / Gyr
    assert abs(100 * ex[2] / 0.1040920E+01 - 1) < 1e-5             # 100 theta_MC (H0 rounded in the file)


def test_hierarchy_two_eigenstates():
    from cosmomc_b200 import params as P
    massless, n_eig, deg, frac = P.neutrino_hierarchy(0.3 / 94.07 * (3.046 / 3) ** 0.75, 0.0, 3.046, "normal")
    assert n_eig == 2 and abs(sum(frac) - 1) < 1e-12 and abs(deg[0] - 2 * deg[1]) < 1e-12
    massless, n_eig, deg, frac = P.neutrino_hierarchy(0.06 / 94.07 * (3.046 / 3) ** 0.75, 0.0, 3.046, "degenerate", 3)
    assert n_eig == 1 and abs(deg[0] - 3.046) < 1e-12 and massless == 0


def test_dataset_readers():
    from cosmomc_b200 import datasets as D
    b = D.BAOPlan(os.path.join(DATA, "DR12", "sdss_DR12Consensus_bao.dataset"))
    assert list(b.types) == [10, 7, 10, 7, 10, 7] and b.rs_rescale == .6766815537e-2 and b.invcov.shape == (6, 6)
    assert np.allclose(b.invcov @ np.loadtxt(os.path.join(DATA, "DR12", "BAO_consensus_covtot_dM_Hz.txt")), np.eye(6), atol=1e-9)
    m = D.BAOPlan(os.path.join(DATA, "sdss_MGS_bao.dataset"), tag="MGS")
    assert len(m.alpha_prob) == 399 and m.z[0] == 0.15
    s = D.BAOPlan(os.path.join(DATA, "sdss_6DF_bao.dataset"))
    assert list(s.types) == [3] and s.rs_rescale == 1.027369826 and abs(s.invcov[0, 0] - 1 / 0.015 ** 2) < 1e-9
    hst = D.HSTPlan(os.path.join(DATA, "HST_Riess2018.ini"))
    assert (hst.H0, hst.H0_err, hst.zeff) == (73.45, 1.66, 0.0)
    j = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs={k: np.eye(740) for k in D.SN_COV_NAMES})
    assert j.nsn == 740 and j.twoscriptmfit and j.alphabeta_covmat and j.A1.sum() + j.A2.sum() == 740
    p = D.SNPlan(os.path.join(DATA, "Pantheon", "full_long.dataset"), covs={"mag": np.eye(1048)})
    assert p.nsn == 1048 and not p.twoscriptmfit and not p.alphabeta_covmat


def _fit(z):  # python/planck/SN.py:303-304
    return -338.65487197 * z ** 4 + 1972.59141641 * z ** 3 - 4310.60442428 * z ** 2 + 4357.72542145 * z


def test_sn_oracle_vs_reference_port(o):
    from cosmomc_b200 import datasets as D, synthetic as syn
    G = np.load(os.path.join(H.ROOT, "tests", "golden", "sn_py.npz"))
    z = np.loadtxt(os.path.join(DATA, "jla_lcparams.txt"), usecols=1)
    j = D.SNPlan(os.path.join(DATA, "jla.dataset"), covs=syn.synthetic_sn_covs({"zcmb": z}))
    sn = o.SN(j.lc, j.covs, pecz=j.pecz, twoscriptmfit=True, scriptmcut=j.scriptmcut)
    assert np.array_equal(sn.A1, j.A1) and np.allclose(sn.pre_vars, j.pre_vars, rtol=1e-15)
    for (a, b), want in zip(G["jla_ab"], G["jla_lnl"]):
        got = sn.loglike(_fit(j.lc["zcmb"]), a, b)
        assert abs(got - want) < 1e-7 * abs(want), (a, b, got, want)
    zp = np.loadtxt(os.path.join(DATA, "Pantheon", "lcparam_full_long_zhel.txt"), usecols=1)
    p = D.SNPlan(os.path.join(DATA, "Pantheon", "full_long.dataset"),
                 covs=syn.synthetic_sn_covs({"zcmb": zp}, names=("mag",), seed=2025))
    snp = o.SN(p.lc, p.covs, pecz=p.pecz, twoscriptmfit=False)
    assert abs(snp.loglike(_fit(zp)) - G["pantheon_lnl"][0]) < 1e-7 * G["pantheon_lnl"][0]
    assert abs(snp.loglike(1.02 * _fit(zp) + 3.0) - G["pantheon_lnl_scaled"][0]) < 1e-7 * G["pantheon_lnl_scaled"][0]


def test_bao_hst_oracle_known_values(o, bg):
    """-lnL of the shipped BAO / H0 data at the Planck 2018 best fit: regression anchors for the GPU parity tests
    (the reference ships no number for these; the theory vector is pinned by test_background_golden_minimum)."""
    from cosmomc_b200 import datasets as D
    b = D.BAOPlan(os.path.join(DATA, "DR12", "sdss_DR12Consensus_bao.dataset"))
    ll = o.bao_loglike(bg, BESTFIT["rdrag"], b.rs_rescale, b.types, b.z, b.obs, b.invcov)
    # independent evaluation straight from the golden derived parameters
    th = np.zeros(6)
    th[0::2] = GOLD_DM[1:4] / (BESTFIT["rdrag"] * b.rs_rescale)
    th[1::2] = GOLD_H[1:4] * (BESTFIT["rdrag"] * b.rs_rescale)
    d = th - b.obs
    assert abs(ll - 0.5 * d @ b.invcov @ d) < 1e-4   # golden values carry 7 figures
    hst = D.HSTPlan(os.path.join(DATA, "HST_Riess2018.ini"))
    assert abs(o.hst_loglike(bg, hst.H0, hst.H0_err) - (BESTFIT["H0"] - 73.45) ** 2 / (2 * 1.66 ** 2)) < 1e-12
