"""Known-answer test of the line-of-sight projection (camb/cmbmain.f90:1440-1562), independent of the C++ restatement.

The reference ships no golden vector for Delta_l(q).  A source that is 1 at ONE time sample (all wavenumbers) and 0
elsewhere collapses the time integral to a single term, so for every (q, l) whose integration window contains that
sample
        Delta_l(q) = dtau_n * J_l(q (tau0 - tau_n)),
with J_l the reference's cubic-spline interpolation of its Bessel table (cmbmain.f90:1515-1516).  The expectation
is evaluated here in plain numpy from the table arrays (textbook spline formula, numpy searchsorted for the
interval) and, as a sanity bound, against scipy's exact spherical Bessel function: the reference evaluator bjl
(bessels.f90:132-275) is itself only good to ~4e-4 absolute near its matching points (tests/test_oracle_golden.py),
so that second bound is loose by construction.  Where the window excludes the sample the result must be exactly 0.
"""
import numpy as np
import pytest

import helpers as H


def _single_sample_batch(seed=11):
    b = H.small_batch(1, seed=seed)
    nt, nk = int(b["n_tau"][0]), int(b["n_k"][0])
    ns = nt // 2  # 0-based sample (Fortran n = ns + 1 >= 2: Source_q(1,:) is forced to zero, cmbmain.f90:1380)
    b["src"][:] = 0.0
    b["src"][0, ns, 0, :nk] = 1.0   # temperature source; constant in k, so the k-spline reproduces it exactly
    b["src"][0, ns, 1, :nk] = -2.0  # polarisation source: same J_l, different weight
    return b, ns


def _expected(bessel, ls, q, tau0, tau_n, dtau_n):
    xs, ajl, ajlpr = bessel.arrays()
    x = q * (tau0 - tau_n)
    i = np.clip(np.searchsorted(xs, x, side="right") - 1, 0, len(xs) - 2)
    h = xs[i + 1] - xs[i]
    a = (xs[i + 1] - x) / h
    b = 1.0 - a
    J = (a[:, None] * ajl[:, i].T + b[:, None] * ajl[:, i + 1].T +
         ((a ** 3 - a)[:, None] * ajlpr[:, i].T + (b ** 3 - b)[:, None] * ajlpr[:, i + 1].T) * (h * h / 6.0)[:, None])
    return x, dtau_n * J  # [q][l]


def _check(D, q, ls, bessel, th, tau_n, dtau_n):
    from scipy.special import spherical_jn
    x, want = _expected(bessel, ls, q, th[0], tau_n, dtau_n)
    nz = D[..., 0] != 0
    assert 0.3 < nz.mean() < 0.6                      # the windows keep about 40 % of the (q, l) plane
    assert np.array_equal(nz, D[..., 1] != 0)
    scale = np.abs(want[nz]).max()
    # tolerance 1e-11 of the largest term: same table, same weights, only the association of the sum differs
    assert np.abs(D[..., 0] - want)[nz].max() < 1e-11 * scale
    assert np.abs(D[..., 1] + 2.0 * want)[nz].max() < 2e-11 * scale
    assert np.all(D[..., 2][:, np.asarray(ls) <= 400] == 0)  # lensing source is zero; l > 400 takes the Limber value
    exact = dtau_n * np.array([spherical_jn(int(l), x) for l in ls]).T
    assert np.abs(D[..., 0] - exact)[nz].max() < 5e-4 * dtau_n   # bjl's own accuracy (see module docstring)
    # a (q, l) outside its window contributes nothing even where j_l is far from negligible (x > 80 l cut)
    assert np.abs(exact[~nz]).max() > 1e-3 * dtau_n


def test_oracle_projection_single_sample():
    import pyoracle as o
    b, ns = _single_sample_batch()
    th = b["thermo"][0]
    nt, nk = int(b["n_tau"][0]), int(b["n_k"][0])
    ls = o.initlval(H.MAX_L)
    bessel = o.Bessel(ls, H.MAX_ETA_K)
    src = np.ascontiguousarray(b["src"][0, :nt, :, :nk])
    q, dq, D, triples = o.project(bessel, th[0], th[1], th[2], th[3], th[4], H.MAX_ETA_K, H.MAX_L, False,
                                  b["k"][0, :nk], src)
    _, dtau = o.time_steps(th[1], th[2], th[0], H.MAX_ETA_K, False, th[3], th[4])
    _check(D, q, ls, bessel, th, b["tau"][0, ns], dtau[ns])


@pytest.mark.gpu
@pytest.mark.parametrize("pk", [4, 3])
def test_kernel_projection_single_sample(pk):
    """the same known answer through the C ABI (projection kernel 4 = default, 3 = its fallback pass)"""
    import pyoracle as o
    from cosmomc_b200 import lib
    T = H.load_templates()
    h = lib.Handle(max_points=2, lmax_out=H.LMAX_OUT)
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    h.set_option("proj_kernel", pk)
    b, ns = _single_sample_batch()
    th = b["thermo"][0]
    ls = o.initlval(H.MAX_L)
    bessel = o.Bessel(ls, H.MAX_ETA_K)   # the checker's table; the library builds its own on the device
    h.keep_transfers(True)
    h.upload_sources(b["thermo"], b["n_k"], b["k"], b["src"])
    h.powers(b["initpower"], b["alens"])
    nt, nk = int(b["n_tau"][0]), int(b["n_k"][0])
    q = o.project(bessel, th[0], th[1], th[2], th[3], th[4], H.MAX_ETA_K, H.MAX_L, False, b["k"][0, :nk],
                  np.ascontiguousarray(b["src"][0, :nt, :, :nk]))[0]   # only the wavenumber grid is taken from here
    D = h.debug_fetch(3, 0).reshape(len(q), 96, 3)[:, :len(ls), :]
    _, dtau = o.time_steps(th[1], th[2], th[0], H.MAX_ETA_K, False, th[3], th[4])
    _check(D, q, ls, bessel, th, b["tau"][0, ns], dtau[ns])
