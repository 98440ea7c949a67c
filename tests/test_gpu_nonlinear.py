"""GPU parity tests for the non-linear lensing rescale and sigma_8 (cb200_nonlinear_lensing, csrc/nonlin.cuh) through the
C ABI, against the oracle (oracle/orc_nonlin.hpp; its known-answer tests are in tests/test_nonlinear_oracle.py)."""
import numpy as np
import pytest

import helpers as H
from test_nonlinear_oracle import lcdm_transfer

pytestmark = pytest.mark.gpu
NPTS = 3
Z = np.array([9.0, 8.0, 7.0, 6.0, 5.0, 4.0, 3.0, 2.0, 1.0, 0.0])     # Transfer_SetForNonlinearLensing: NLL_redshifts


@pytest.fixture(scope="module")
def setup():
    import pyoracle as o
    from cosmomc_b200 import lib
    T = H.load_templates()
    h = lib.Handle(max_points=4, chunk_points=2, lmax_out=H.LMAX_OUT)
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    b = H.small_batch(NPTS, seed=17, NT=h.info.n_tau_max, NK=h.info.n_k_max)
    hub = np.array([0.6732, 0.70, 0.65])
    n_kt = int(b["n_k"].max()) + 40
    kh = np.zeros((NPTS, n_kt))
    tr = np.zeros((NPTS, len(Z), n_kt))
    tautf = np.zeros((NPTS, len(Z)))
    cosmo = np.zeros((NPTS, 6))
    for i in range(NPTS):
        nk = b["n_k"][i]
        ksrc = b["k"][i, :nk] / hub[i]
        extra = np.exp(np.linspace(np.log(ksrc[-1] * 1.05), np.log(8.0), n_kt - nk))
        kh[i] = np.concatenate([ksrc, extra])
        tr[i] = lcdm_transfer(kh[i], hub[i], Z) * (1 + 0.03 * i)
        tau0 = b["thermo"][i, 0]
        tautf[i] = tau0 * (1 - 0.7 * (Z / (1 + Z)) ** 0.8)      # ascending, ends at tau0 (z = 0)
        cosmo[i] = [hub[i], 0.3158 - 0.01 * i, 0.6842 + 0.01 * i, 0.0045, -1.0, 0.0]
    return dict(o=o, h=h, b=b, kh=kh, tr=tr, tautf=tautf, cosmo=cosmo, hub=hub, n_kt=n_kt)


def test_sigma8_ratios_and_rescaled_sources_match_oracle(setup):
    o, h, b = setup["o"], setup["h"], setup["b"]
    h.upload_sources(b["thermo"], b["n_k"], b["k"], b["src"])
    tau_dev = [h.debug_fetch(6, i) for i in range(NPTS)]
    before = [h.debug_fetch(10, i, max_n=h.info.n_tau_max * h.info.n_k_max) for i in range(NPTS)]
    r = h.nonlinear_lensing(b["initpower"], setup["cosmo"], setup["kh"], Z, setup["tr"], tautf=setup["tautf"], rescale_sources=True)
    assert np.all(r["status"] == 0)
    NK = h.info.n_k_max
    for i in range(NPTS):
        nk, nt = b["n_k"][i], b["n_tau"][i]
        c = setup["cosmo"][i]
        src = np.ascontiguousarray(b["src"][i, :nt, :, :nk])
        want = o.nonlinear(b["initpower"][i], c[0], c[1], c[2], c[3], setup["kh"][i], Z, setup["tr"][i], k=b["k"][i, :nk],
                           tau=tau_dev[i], tautf=setup["tautf"][i], src=src)
        assert np.abs(r["sigma8"][i] / want["sigma8"] - 1).max() < 1e-12
        assert np.allclose(r["spec"][i], want["spec"], rtol=1e-9, atol=0)   # same bisection path (zeros where still linear)
        assert np.abs(r["ratio"][i] / want["ratio"] - 1).max() < 1e-9
        got = h.debug_fetch(10, i, max_n=h.info.n_tau_max * NK).reshape(nt, NK)[:, :nk]
        assert np.abs(got - src[:, 2]).max() <= 1e-11 * np.abs(src[:, 2]).max()
        assert not np.array_equal(got, before[i].reshape(nt, NK)[:, :nk])       # something was rescaled
        assert 0.3 < r["sigma8"][i, -1] < 1.5 and np.all(np.diff(r["sigma8"][i]) > 0)


def test_rescale_raises_the_small_scale_lensing_power(setup):
    h, b = setup["h"], setup["b"]
    h.upload_sources(b["thermo"], b["n_k"], b["k"], b["src"])
    lin, _, st0 = h.powers(b["initpower"], b["alens"])
    h.nonlinear_lensing(b["initpower"], setup["cosmo"], setup["kh"], Z, setup["tr"], tautf=setup["tautf"], rescale_sources=True)
    nl, _, st1 = h.powers(b["initpower"], b["alens"])
    assert np.all(st0 == 0) and np.all(st1 == 0)
    pp_lin, pp_nl = lin[:, 4], nl[:, 4]
    assert np.all(pp_nl[:, 1500:2000] > pp_lin[:, 1500:2000])                   # non-linear growth: more small-scale power
    assert np.abs(pp_nl[:, 2:20] / pp_lin[:, 2:20] - 1).max() < 2e-2            # large scales stay (nearly) linear
    assert np.array_equal(lin[:, 0] != 0, nl[:, 0] != 0)


def test_error_paths(setup):
    from cosmomc_b200 import lib
    h, b = setup["h"], setup["b"]
    with pytest.raises(lib.CB200Error):     # rescaling without the transfer times
        h.nonlinear_lensing(b["initpower"], setup["cosmo"], setup["kh"], Z, setup["tr"], tautf=None, rescale_sources=True)
    # a wildly non-linear spectrum takes halofit's error exit (global_error_flag = 349 in the reference)
    r = h.nonlinear_lensing(b["initpower"], setup["cosmo"], setup["kh"], Z, setup["tr"] * 1e7)
    assert np.all(r["status"] == 349)
    # a linear one leaves every ratio at one
    r = h.nonlinear_lensing(b["initpower"], setup["cosmo"], setup["kh"], Z, setup["tr"] * 1e-4)
    assert np.all(r["status"] == 0) and np.all(r["ratio"] == 1.0) and np.all(r["spec"] == 0)
