"""GPU parity tests for the tensor pass (CalcTensCls, camb/cmbmain.f90:2344-2397; tensors added in
SetPowersFromCAMB, source/Calculator_CAMB.f90:403-406) and for the shared-transfer semi-slow step
(cb200_powers_shared: k-contraction of a batch of initial-power points as one DMMA GEMM), BASELINE configs[2].
Synthetic sources on the real (bit-exact) tensor grids; tolerances next to each check."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
NPTS = 2


@pytest.fixture(scope="module")
def setup():
    import pyoracle as o
    from cosmomc_b200 import lib
    T = H.load_templates()
    h = lib.Handle(max_points=16, chunk_points=4, lmax_out=H.LMAX_OUT, compute_tensors=1, lmax_tensor=H.MAX_L_T)
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    ls = o.initlval(H.MAX_L)
    ls_t = o.initlval(H.MAX_L_T)
    bessel = o.Bessel(ls, H.MAX_ETA_K)
    bessel_t = o.Bessel(ls_t, H.MAX_ETA_K_T)
    batch = H.small_batch(NPTS, seed=21, NT=h.info.n_tau_max, NK=h.info.n_k_max)
    tb = H.small_batch_tensor(batch["thermo"], seed=21, NT=h.cfg.n_tau_max_tensor, NK=h.cfg.n_k_max_tensor)
    ip = batch["initpower"].copy()
    ip[:, 4] = [0.07, 0.2]       # r
    ip[1, 9] = 0.0               # point 1: free n_t (no inflation consistency), different tensor pivot
    ip[1, 5] = -0.3
    ip[1, 8] = 0.01
    batch["initpower"] = ip
    h.upload_sources(batch["thermo"], batch["n_k"], batch["k"], batch["src"])
    h.upload_sources(tb["thermo"], tb["n_k"], tb["k"], tb["src"], kind=1)
    return dict(h=h, T=T, ls=ls, ls_t=ls_t, bessel=bessel, bessel_t=bessel_t, batch=batch, tb=tb, o=o)


def test_tensor_tables_and_grids(setup):
    h, o = setup["h"], setup["o"]
    assert h.info.n_lsamp_tensor == 47 and h.info.max_l_tensor == 600 and h.info.max_eta_k_tensor == 1500
    assert np.array_equal(h.lsamples(1), setup["ls_t"])
    x, ajl, ajlpr = h.bessel_table(1)
    xo, ao, apo = setup["bessel_t"].arrays()
    assert np.array_equal(x, xo) and np.abs(ajl - ao).max() < 2e-12
    th = setup["batch"]["thermo"][0]
    t, dt = h.time_steps(th[0], th[1], th[2], th[3], th[4], kind=1)
    to, dto = o.time_steps(th[1], th[2], th[0], H.MAX_ETA_K_T, True, th[3], th[4])
    assert np.array_equal(t, to) and np.array_equal(dt, dto)       # bit-exact tensor time grid (n ~ 2100)
    q, dq = h.q_grid(th[0], kind=1)
    qo, dqo = o.q_grid(th[0], H.MAX_ETA_K_T, H.MAX_L_T)
    assert np.array_equal(q, qo) and np.array_equal(dq, dqo)
    k = h.source_k(th[0], th[1], kind=1)
    assert np.array_equal(k, o.source_k(th[0], th[1], H.MAX_ETA_K_T, True, H.MAX_L_T))


def test_per_point_tensor_pass(setup):
    h, b, tb = setup["h"], setup["batch"], setup["tb"]
    cls, derived, status = h.powers(b["initpower"], b["alens"])
    assert np.all(status == 0)
    for i in range(NPTS):
        ot = H.oracle_tensor_point(tb, i, setup["bessel_t"], setup["ls_t"], b["initpower"][i])
        icl = h.debug_fetch(8, i).reshape(4, -1)
        scale = np.abs(ot["iCl"]).max(axis=1, keepdims=True)
        assert (np.abs(icl - ot["iCl"]) / scale).max() < 1e-9          # sampled tensor C_l
        clt = h.debug_fetch(9, i).reshape(4, -1)
        assert (np.abs(clt - ot["cl"]) / np.abs(ot["cl"]).max(axis=1, keepdims=True)).max() < 1e-9
        ref = H.oracle_point_with_tensors(b, i, setup["bessel"], setup["ls"], setup["T"]["highl_unlensed"],
                                          setup["T"]["highl_lensed"], ot["cl"])
        c = ref["cls_out"]
        for X in (0, 1, 2, 3, 4):
            sc = np.abs(c[X]).max()
            # TT,TE,EE,PP 1e-9; lensed BB is a cancelling difference of two correlation sums: 1e-7 (north_star 1e-4)
            assert (np.abs(cls[i, X] - c[X]) / sc).max() < (1e-7 if X == 3 else 1e-9), X
        # tensor derived ratios (Calculator_CAMB.f90:451-455)
        o = setup["o"]
        ip = b["initpower"][i]
        assert abs(derived[i, 1] / (o.tensor_power(ip, [0.002])[0] / o.scalar_power(ip, [0.002])[0]) - 1) < 1e-12
        assert abs(derived[i, 2] / (o.tensor_power(ip, [0.01])[0] / o.scalar_power(ip, [0.01])[0]) - 1) < 1e-12
        assert abs(derived[i, 3] / o.tensor_power(ip, [ip[8]])[0] - 1) < 1e-12


def test_shared_transfers_gemm_path(setup):
    """cb200_powers_shared == per-point evaluation with the same sources and the batch's initial-power points"""
    h, b, tb, o = setup["h"], setup["batch"], setup["tb"], setup["o"]
    rng = np.random.default_rng(5)
    B = 6
    ip = np.tile(b["initpower"][0], (B, 1))
    ip[:, 0] *= np.exp(rng.normal(0, 0.02, B))
    ip[:, 1] += rng.normal(0, 0.01, B)
    ip[:, 4] = rng.uniform(0.0, 0.5, B)
    ip[:, 9] = 1.0
    al = np.ones(B)
    cls, derived, status = h.powers_shared(ip, al, src_point=0, first=4)
    assert np.all(status == 0)
    # oracle: project once, contract per initial-power point
    th = b["thermo"][0]
    nt, nk = b["n_tau"][0], b["n_k"][0]
    q, dq, Delta, _ = o.project(setup["bessel"], th[0], th[1], th[2], th[3], th[4], H.MAX_ETA_K, H.MAX_L, False,
                                b["k"][0, :nk], np.ascontiguousarray(b["src"][0, :nt, :, :nk]))
    ntt, nkt = tb["n_tau"][0], tb["n_k"][0]
    qt, dqt, Dt, _ = o.project(setup["bessel_t"], th[0], th[1], th[2], th[3], th[4], H.MAX_ETA_K_T, H.MAX_L_T, True,
                               tb["k"][0, :nkt], np.ascontiguousarray(tb["src"][0, :ntt, :, :nkt]))
    for j in (0, 3, 5):
        iclt = o.calc_cls(qt, dqt, setup["ls_t"], Dt, ip[j], 1.0, tensors=True)
        clt = np.stack([o.interp_cl(setup["ls_t"], iclt[X])[:H.MAX_L_T + 1] for X in range(4)])
        ref = H.oracle_from_delta(q, dq, Delta, setup["ls"], setup["T"]["highl_unlensed"], setup["T"]["highl_lensed"],
                                  ip[j], 1.0, clt)
        for X in range(5):
            sc = np.abs(ref["cls_out"][X]).max()
            assert (np.abs(cls[j, X] - ref["cls_out"][X]) / sc).max() < (1e-7 if X == 3 else 1e-9), (j, X)
