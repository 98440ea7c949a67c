"""CPU tests (no GPU): the ORACLE's non-linear lensing rescale and sigma_8 (oracle/orc_nonlin.hpp) against
known answers that do not depend on the restatement.  No reference-produced vector exists for this stage (matter
transfer functions come from the ODE stage of the Fortran build): parity is unpinned by reference output, and these
tests bound the restatement from the outside:
  * a pure power law P(k) = A k^n has sigma_G^2(R) = A Gamma((n+3)/2) / (4 pi^2 R^(n+3)), so halofit's search must return
    n_eff = n, zero curvature and k_NL = the analytic root (wint + the bisection of NonLinear_GetNonLinRatios);
  * sigma_R converges to the quadrature of its integrand (Transfer_Get_SigmaR);
  * the Takahashi fitting formula re-evaluated in numpy from the returned (k_NL, n_eff, curvature) (published formula);
  * MakeNonlinearSources leaves k/h <= 0.005 and tau < tau(z_max) alone and passes through the ratios at the knots."""
import numpy as np
import pytest
from math import gamma, pi

import helpers as H  # noqa: F401

IP = np.array([2.1e-9, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.05, 0.05, 1.0])   # ns = 1: flat primordial spectrum


@pytest.fixture(scope="module")
def o():
    import pyoracle
    return pyoracle


def power_law_transfer(kh, h, n, A):
    """T(k) such that Transfer_GetMatterPowerData gives P(k/h) = A (k/h)^n for the flat spectrum IP."""
    k = kh * h
    return np.sqrt(A * kh ** n / (k * pi * 2 * pi * h ** 3 * IP[0]))


@pytest.mark.parametrize("n", [-2.0, -1.5, -1.0])
def test_power_law_gives_analytic_nonlinear_scale(o, n):
    h = 0.67
    kh = np.exp(np.linspace(np.log(1e-4), np.log(20.0), 300))
    A = 2 * pi ** 2 * 0.5                                   # Delta^2(k) = 0.5 (k h/Mpc)^(n+3)
    T = power_law_transfer(kh, h, n, A)
    assert np.allclose(o.matter_power_at(IP, h, kh, T, [1e-6, 0.3, 500.0]), A * np.array([1e-6, 0.3, 500.0]) ** n, rtol=1e-10)
    r = o.nonlinear(IP, h, 1.0, 0.0, 0.0, kh, [0.0], T[None, :])
    rknl, rneff, rncur = r["spec"][0]
    # sigma_G^2(R) = (A / 2 pi^2) Gamma((n+3)/2) / (2 R^(n+3)) = 1 ; the search stops at |sigma - 1| <= 1e-3
    R = (A / (2 * pi ** 2) * gamma((n + 3) / 2) / 2) ** (1 / (n + 3))
    assert abs(rknl * R - 1) < 2 * 1.2e-3 / (n + 3)
    assert abs(rneff - n) < 2e-3 and abs(rncur) < 5e-3       # 3000-midpoint rule of wint
    assert r["err"] == 0


def bbks(kh, h, omm=0.315):
    q = kh / (omm * h)
    return np.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25


def lcdm_transfer(kh, h, z):
    """CAMB-like matter transfer (Delta / k^2 convention) with an EdS-like growth, amplitude set for sigma_8 ~ 0.8."""
    return 1.365e7 * bbks(kh, h)[None, :] / (1 + np.asarray(z))[:, None] ** 0.9


def test_sigma_r_converges_to_the_quadrature(o):
    from scipy.integrate import quad
    h = 0.6732
    ip = IP.copy(); ip[1] = 0.9649
    z = np.array([1.0, 0.0])
    def s8(nk):
        kh = np.exp(np.linspace(np.log(1e-5), np.log(30.0), nk))
        return o.nonlinear(ip, h, 0.3158, 0.6842, 0.0, kh, z, lcdm_transfer(kh, h, z))["sigma8"]
    def integrand(lnk, zz):
        k = np.exp(lnk); kh = k / h; x = kh * 8.0
        win = 3 * (np.sin(x) - x * np.cos(x)) / x ** 3
        T = lcdm_transfer(np.array([kh]), h, [zz])[0, 0]
        return (win * k * k) ** 2 * ip[0] * (k / 0.05) ** (ip[1] - 1) * T * T
    want = np.sqrt([quad(integrand, np.log(1e-5 * h), np.log(30.0 * h), args=(zz,), limit=400)[0] for zz in z])
    coarse, fine = s8(300), s8(4000)
    assert np.abs(fine / want - 1).max() < 2e-5 and np.abs(coarse / want - 1).max() < 3e-3
    assert 0.5 < fine[1] < 1.2 and fine[0] < fine[1]


def takahashi(rk, rn, rncur, rknl, plin, om_m, om_v, fnu, w=-1.0):
    """Takahashi et al. 2012 (ApJ 761, 152) fitting formula with the massive-neutrino terms of Bird et al. as CAMB applies them."""
    gam = 0.1971 - 0.0843 * rn + 0.8460 * rncur
    a = 10 ** (1.5222 + 2.8553 * rn + 2.3706 * rn ** 2 + 0.9903 * rn ** 3 + 0.2250 * rn ** 4 - 0.6038 * rncur + 0.1749 * om_v * (1 + w))
    b = 10 ** (-0.5642 + 0.5864 * rn + 0.5716 * rn ** 2 - 1.5474 * rncur + 0.2279 * om_v * (1 + w))
    c = 10 ** (0.3698 + 2.0404 * rn + 0.8161 * rn ** 2 + 0.5869 * rncur)
    xnu = 10 ** (5.2105 + 3.6902 * rn)
    alpha = abs(6.0835 + 1.3373 * rn - 0.1959 * rn ** 2 - 5.5274 * rncur)
    beta = 2.0379 - 0.7354 * rn + 0.3157 * rn ** 2 + 1.2490 * rn ** 3 + 0.3980 * rn ** 4 - 0.1682 * rncur + fnu * (1.081 + 0.395 * rn ** 2)
    frac = om_v / (1 - om_m)
    f1 = frac * om_m ** -0.0307 + (1 - frac) * om_m ** -0.0732
    f2 = frac * om_m ** -0.0585 + (1 - frac) * om_m ** -0.1423
    f3 = frac * om_m ** 0.0743 + (1 - frac) * om_m ** 0.0725
    y = rk / rknl
    ph = a * y ** (3 * f1) / (1 + b * y ** f2 + (f3 * c * y) ** (3 - gam)) / (1 + xnu / y ** 2) * (1 + fnu * 0.977)
    plinaa = plin * (1 + fnu * 47.48 * rk ** 2 / (1 + 1.5 * rk ** 2))
    pq = plin * (1 + plinaa) ** beta / (1 + plinaa * alpha) * np.exp(-y / 4 - y ** 2 / 8)
    return pq + ph


def test_ratios_follow_the_published_fitting_formula(o):
    h, omm0, omv, fnu = 0.6732, 0.3158, 0.6842, 0.0045
    ip = IP.copy(); ip[1] = 0.9649
    kh = np.exp(np.linspace(np.log(1e-5), np.log(8.0), 260))
    z = np.array([2.0, 0.5, 0.0])
    T = lcdm_transfer(kh, h, z)
    r = o.nonlinear(ip, h, omm0, omv, fnu, kh, z, T)
    for i, zz in enumerate(z):
        a = 1 / (1 + zz)
        om_m = omm0 / a ** 3 / (omm0 / a ** 3 + omv)      # flat LCDM: omega_m(a) of halofit_ppf.f90:336-341
        rknl, rn, rc = r["spec"][i]
        assert rknl > 0 and -3 < rn < 0
        plin = o.matter_power_at(ip, h, kh, T[i], kh) * kh ** 3 / (2 * pi ** 2)
        want = np.sqrt(takahashi(kh, rn, rc, rknl, plin, om_m, 1 - om_m, fnu) / plin)
        want[kh <= np.float32(0.005)] = 1.0
        assert np.abs(r["ratio"][i] / want - 1).max() < 1e-10
    assert np.all(r["ratio"][:, -1] > 1.5) and np.all(np.abs(r["ratio"][:, kh < 0.004] - 1) < 1e-12)
    assert r["spec"][0, 0] > r["spec"][2, 0]              # the non-linear scale moves to smaller k with time


def test_make_nonlinear_sources_properties(o):
    h = 0.6732
    ip = IP.copy(); ip[1] = 0.9649
    kh_t = np.exp(np.linspace(np.log(1e-5), np.log(8.0), 200))
    z = np.array([9.0, 6.0, 3.0, 1.0, 0.0])
    T = lcdm_transfer(kh_t, h, z)
    tautf = np.array([4500.0, 6000.0, 8500.0, 11000.0, 14100.0])
    n_k = 120
    k = kh_t[:n_k] * h
    tau = np.sort(np.concatenate([np.linspace(300.0, 14099.0, 57), tautf[:4]]))
    rng = np.random.default_rng(2)
    src = rng.standard_normal((len(tau), 3, n_k))
    before = src.copy()
    r = o.nonlinear(ip, h, 0.3158, 0.6842, 0.0, kh_t, z, T, k=k, tau=tau, tautf=tautf, src=src)
    assert np.array_equal(src[:, :2], before[:, :2])                       # temperature and polarisation sources untouched
    assert np.array_equal(src[tau < tautf[0]], before[tau < tautf[0]])     # before the first transfer redshift
    assert np.array_equal(src[-1], before[-1])                             # the loop stops at npoints - 1
    lin = (k / h <= np.float32(0.005)) | np.all(np.abs(r["ratio"][:, :n_k] - 1) < 5e-4, axis=0)
    assert np.array_equal(src[:, 2][:, lin], before[:, 2][:, lin])
    for j in range(4):                                                     # the spline passes through the ratios
        i = int(np.where(tau == tautf[j])[0][0])
        got = src[i, 2] / before[i, 2]
        assert np.allclose(got[~lin], r["ratio"][j, :n_k][~lin], rtol=1e-12)
    scale = (src[:-1, 2] / before[:-1, 2])[tau[:-1] >= tautf[0]]
    assert scale[:, ~lin].min() > 0.99 and scale[:, ~lin].max() < r["ratio"].max() * 1.05
