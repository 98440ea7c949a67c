"""cosmomc_b200.LikeCalculator: a top-level CosmoMC .ini + .paramnames -> likelihood list, parameter layout and one
batched -lnL, driven by the reference's own batch3 configuration files (fixtures under tests/golden/data).
CPU: what source/DataLikelihoods.f90:9-41 / BaseParameters.f90:90-200 build at start-up.  GPU: loglike(P) against the
separate C-ABI calls + the oracle's restatement of GetLogLike (source/calclike.f90:97-151)."""
import os

import numpy as np
import pytest

import helpers as H

DATA = os.path.join(H.ROOT, "tests", "golden", "data")


def write_ini(tmp_path, body):
    p = tmp_path / "run.ini"
    p.write_text(body)
    return str(p)


LENSING_RUN = """
DEFAULT(%s/batch3/lensing.ini)
DEFAULT(%s/batch3/params_CMB_defaults.ini)
temperature = 1.5
prior[ns] = 0.96 0.02
linear_combination[nsrun] = ns nrun
linear_combination_weights[nsrun] = 1 1
prior[nsrun] = 0.95 0.05
param[nrun] = 0 -1 1 0.01 0.01
""" % (DATA, DATA)


def make_calc(tmp_path, body, **kw):
    from cosmomc_b200.likecalc import LikeCalculator
    return LikeCalculator(write_ini(tmp_path, body), data_dir=DATA, local_dir=DATA, **kw)


def test_ini_to_layout_lensing(tmp_path):
    c = make_calc(tmp_path, LENSING_RUN, create_handle=False)
    assert c.like_names() == ["lensing"]
    # base block = the non-derived names of params_CMB.paramnames in file order, then the likelihood's nuisance block
    assert c.names[:4] == ["omegabh2", "omegach2", "theta", "tau"] and c.names[c.n_base:] == ["calPlanck"]
    i = c.names.index
    assert c.n_base == 24 and c.num_params == 25 and c.nuis_first == 24 and c.n_nuis == 1
    assert (c.center[i("logA")], c.pmin[i("logA")], c.pmax[i("logA")]) == (3.05, 1.61, 3.91)
    assert c.propose_width[i("ns")] == 0.002 and c.varying[i("ns")] and not c.varying[i("mnu")]
    assert c.center[i("mnu")] == c.pmin[i("mnu")] == c.pmax[i("mnu")] == 0.06
    # later definitions do not override earlier ones for DEFAULT() files: the run file's own nrun line wins
    assert c.varying[i("nrun")] and (c.pmin[i("nrun")], c.pmax[i("nrun")]) == (-1.0, 1.0)
    assert (c.prior_mean[i("calPlanck")], c.prior_std[i("calPlanck")]) == (1.0, 0.0025)   # planck_calibration.ini
    assert (c.prior_mean[i("ns")], c.prior_std[i("ns")]) == (0.96, 0.02)
    assert len(c.lincomb) == 1 and c.lincomb[0][i("ns")] == 1 and c.lincomb[0][i("nrun")] == 1
    assert c.lincomb_mean == [0.95] and c.lincomb_std == [0.05] and c.temperature == 1.5
    assert c.columns["logA"] == i("logA") and c.columns["Aphiphi"] == i("Aphiphi")
    P = c.full_params(np.zeros((2, int(c.varying.sum()))))
    assert P.shape == (2, 25) and P[0, i("mnu")] == 0.06 and P[0, i("ns")] == 0.0


def test_ini_likelihood_order_and_keys(tmp_path):
    """reference order CMB -> Hubble -> supernovae -> BAO whatever the order in the file; use_* switches respected"""
    body = """
use_BAO = T
bao_dataset[DR12BAO] = %s/DR12/sdss_DR12Consensus_bao.dataset
bao_dataset[MGS] = %s/sdss_MGS_bao.dataset
DEFAULT(%s/batch3/HST_Riess2018.ini)
DEFAULT(%s/batch3/lensing.ini)
DEFAULT(%s/batch3/params_CMB_defaults.ini)
""" % (DATA, DATA, DATA, DATA, DATA)
    c = make_calc(tmp_path, body, create_handle=False)
    assert c.like_names() == ["lensing", "H073p45", "DR12BAO", "MGS"]
    kinds = [k for k, _, _, _ in c.likes]
    assert kinds == ["cmb", "hst", "bao", "bao"]
    c2 = make_calc(tmp_path, body.replace("use_BAO = T", "use_BAO = F"), create_handle=False)
    assert c2.like_names() == ["lensing", "H073p45"]
    with pytest.raises(KeyError):   # a nuisance parameter without a param[...] line stops the run, as in the reference
        make_calc(tmp_path, "cmb_dataset[lensing] = %s/planck_lensing_2018/smicadx12_Dec5_ftl_mv2_ndclpp_p_teb_consext8.dataset\n"
                  "DEFAULT(%s/batch3/params_CMB_defaults.ini)\n" % (DATA, DATA), create_handle=False)


@pytest.mark.gpu
def test_loglike_matches_get_loglike(tmp_path):
    import pyoracle as o
    T = H.load_templates()
    c = make_calc(tmp_path, LENSING_RUN, handle_kw=dict(max_points=4, chunk_points=2, lmax_out=H.LMAX_OUT))
    c.set_templates(T["highl_unlensed"], T["highl_lensed"])
    n = 3
    b = H.small_batch(n, seed=11, NT=c.handle.info.n_tau_max, NK=c.handle.info.n_k_max)
    c.upload_sources(b["thermo"], b["n_k"], b["k"], b["src"])
    i = c.names.index
    P = c.full_params(np.zeros((n, int(c.varying.sum()))))
    P[:] = c.center
    P[:, i("logA")] = np.log(1e10 * b["initpower"][:, 0])
    P[:, i("ns")] = b["initpower"][:, 1]
    P[:, i("nrun")] = [0.0, 0.01, -0.02]
    P[:, i("calPlanck")] = b["cal"]
    P[1, i("ns")] = 1.5                       # leaves the prior box: logZero
    ll, likes, prior, st = c.loglike(P, full_output=True)
    # the same through the separate calls + the oracle's GetLogLike control flow
    ip = b["initpower"].copy()
    ip[:, 1] = P[:, i("ns")]
    ip[:, 2] = P[:, i("nrun")]
    ip[:, 7] = ip[:, 8] = c.pivot_k
    c.handle.powers_resident(ip, np.ones(n))
    want_likes, tot, _ = c.handle.loglike_batch(n, P[:, c.nuis_first:])
    want, wprior, wst = o.get_loglike(P, want_likes, pmin=c.pmin, pmax=c.pmax, prior_mean=c.prior_mean,
                                      prior_std=c.prior_std, use_prior=c.use_prior, lincomb=np.array(c.lincomb),
                                      lincomb_mean=c.lincomb_mean, lincomb_std=c.lincomb_std, temperature=c.temperature)
    assert st.tolist() == wst.tolist() == [0, 1, 0]
    assert ll[1] == 1e30
    assert np.allclose(likes[[0, 2]], want_likes[[0, 2]], rtol=1e-12)
    assert np.abs(ll[[0, 2]] - want[[0, 2]]).max() < 1e-9 and np.allclose(prior[[0, 2]], wprior[[0, 2]], rtol=1e-13)


@pytest.mark.gpu
def test_background_from_params_feeds_bao_with_device_rdrag(tmp_path):
    """theta -> H0 -> thermal history -> r_drag -> BAO / H0 likelihoods, all on the device, against the oracle chain
    (h0_from_theta, orc_thermo, bao_loglike, hst_loglike) - the data flow of a background-only CosmoMC run."""
    import pyoracle as o
    from cosmomc_b200 import params as prm
    body = """
use_BAO = T
bao_dataset[DR12BAO] = %s/DR12/sdss_DR12Consensus_bao.dataset
DEFAULT(%s/batch3/HST_Riess2018.ini)
DEFAULT(%s/batch3/params_CMB_defaults.ini)
""" % (DATA, DATA, DATA)
    c = make_calc(tmp_path, body, handle_kw=dict(max_points=4, lmax_computed_cl=0))
    assert c.like_names() == ["H073p45", "DR12BAO"]
    i = c.names.index
    P = np.tile(c.center, (3, 1))
    P[:, i("omegabh2")] = [0.02237737, 0.0221, 0.0226]
    P[:, i("omegach2")] = [0.1201035, 0.1180, 0.1230]
    P[:, i("theta")] = [1.040920, 1.0405, 3.0]           # the last one has no H0 in [20, 100]
    P[:, i("tau")] = [0.05430138, 0.06, 0.05]
    bg, th, ok = c.set_background_from_params(P, yhe=0.2453985)
    assert ok.tolist() == [True, True, False] and bg[2, 0] == 0
    ll, tot, st = c.handle.loglike_batch(3, np.zeros((3, max(1, c.n_nuis))))
    for k in range(2):
        mk = lambda h: prm.cmb_to_background(P[k, i("omegabh2")], P[k, i("omegach2")], h)
        H0 = o.h0_from_theta(P[k, i("theta")], mk)
        assert abs(bg[k, 0] / H0 - 1) < 1e-9
        r = o.thermo(mk(H0), 0.2453985, optical_depth=P[k, i("tau")])
        assert abs(bg[k, 15] / r["derived"]["rdrag"] - 1) < 2e-6
        kinds = {kind: plan for kind, tag, plan, _ in c.likes}
        bao, hst = kinds["bao"], kinds["hst"]
        want = (o.hst_loglike(bg[k], hst.H0, hst.H0_err) +
                o.bao_loglike(bg[k], bg[k, 15], bao.rs_rescale, bao.types, bao.z, bao.obs, bao.invcov))
        assert abs(tot[k] - want) < 1e-6, (tot[k], want)
