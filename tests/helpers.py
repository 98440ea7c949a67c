"""Shared test helpers: build a small synthetic batch and push ONE point through the oracle stage by stage.

The oracle (oracle/pyoracle.py) is the checker; the product path is only ever reached through the C ABI
(cosmomc_b200/lib.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

MAX_L = 2650
MAX_ETA_K = 14000
LMAX_COMPUTED = 2500
LMAX_OUT = 2508


def load_templates():
    return np.load(os.path.join(ROOT, "tests", "golden", "templates.npz"))


def oracle_grids(thermo_row, maximum_l=MAX_L, max_eta_k=MAX_ETA_K):
    import pyoracle as o
    tau, dtau = o.time_steps(thermo_row[1], thermo_row[2], thermo_row[0], max_eta_k, False, thermo_row[3], thermo_row[4])
    k = o.source_k(thermo_row[0], thermo_row[1], max_eta_k, False, maximum_l)
    return tau, dtau, k


def small_batch(npts, seed=7, NT=768, NK=256):
    """Synthetic batch built from ORACLE grids (no GPU needed): returns dict of padded arrays."""
    from cosmomc_b200 import synthetic as syn
    thermo = syn.draw_thermo(npts, seed)
    ip, alens, cal, pert = syn.draw_params(npts, seed)
    tau = np.zeros((npts, NT))
    k = np.zeros((npts, NK))
    n_tau = np.zeros(npts, dtype=np.int32)
    n_k = np.zeros(npts, dtype=np.int32)
    for i in range(npts):
        t, dt, kk = oracle_grids(thermo[i])
        n_tau[i], n_k[i] = len(t), len(kk)
        tau[i, :len(t)] = t
        tau[i, len(t):] = t[-1]
        k[i, :len(kk)] = kk
        k[i, len(kk):] = kk[-1]
    src = syn.make_sources(thermo, tau, k, pert).numpy()
    return dict(thermo=thermo, initpower=ip, alens=alens, cal=cal, tau=tau, k=k, n_tau=n_tau, n_k=n_k, src=src)


def oracle_point(batch, i, bessel, ls, tmpl_unl, highl_lensed, keep=False):
    """Run point i through every oracle stage.  Returns dict of stage outputs."""
    import pyoracle as o
    th = batch["thermo"][i]
    nt, nk = batch["n_tau"][i], batch["n_k"][i]
    src = np.ascontiguousarray(batch["src"][i, :nt, :, :nk])
    q, dq, Delta, triples = o.project(bessel, th[0], th[1], th[2], th[3], th[4], MAX_ETA_K, MAX_L, False,
                                      batch["k"][i, :nk], src)
    ip = batch["initpower"][i]
    iCl = o.calc_cls(q, dq, ls, Delta, ip, batch["alens"][i])
    cl = np.zeros((6, MAX_L + 1))
    for X in range(6):
        cl[X, :] = o.interp_cl(ls, iCl[X], template_index=X + 1 if X < 3 else 0, tmpl=tmpl_unl)[:MAX_L + 1]
    cl4 = np.stack([cl[0], cl[1], cl[2], cl[3]])
    lensed = o.lens_cls(ls, MAX_L, cl4, tmpl_unl)
    lens_pad = np.zeros((4, MAX_L + 1))
    lens_pad[:, :lensed.shape[1]] = lensed
    cl_lmax = [LMAX_OUT] * 5
    out, hn, rms = o.set_powers(lens_pad, cl[3], LMAX_COMPUTED, cl_lmax, highl_lensed, lmax_out=LMAX_OUT)
    # oracle order TT,TE,EE,BB,PP matches the library's cls_out
    res = dict(q=q, dq=dq, iCl=iCl, cl=cl, lensed=lens_pad, cls_out=out, rms=rms, triples=triples)
    if keep:
        res["Delta"] = Delta
    return res


# ---- tensor pass (CAMB's second pass: Max_l_tensor = 600, Max_eta_k_tensor = 1500) -------------------------------
MAX_L_T = 600
MAX_ETA_K_T = 1500


def small_batch_tensor(thermo, seed=7, NT=2304, NK=128):
    """Synthetic tensor sources (T, E, B shaped stand-ins) on the ORACLE's tensor grids for the given thermo rows."""
    import pyoracle as o
    from cosmomc_b200 import synthetic as syn
    npts = len(thermo)
    rng = np.random.default_rng(seed + 100)
    pert = rng.normal(0.0, 1.0, (npts, 6))
    tau = np.zeros((npts, NT))
    k = np.zeros((npts, NK))
    n_tau = np.zeros(npts, dtype=np.int32)
    n_k = np.zeros(npts, dtype=np.int32)
    for i in range(npts):
        th = thermo[i]
        t, dt = o.time_steps(th[1], th[2], th[0], MAX_ETA_K_T, True, th[3], th[4])
        kk = o.source_k(th[0], th[1], MAX_ETA_K_T, True, MAX_L_T)
        n_tau[i], n_k[i] = len(t), len(kk)
        tau[i, :len(t)] = t
        tau[i, len(t):] = t[-1]
        k[i, :len(kk)] = kk
        k[i, len(kk):] = kk[-1]
    src = 1e-3 * syn.make_sources(thermo, tau, k, pert).numpy()
    return dict(thermo=thermo, tau=tau, k=k, n_tau=n_tau, n_k=n_k, src=src)


def oracle_tensor_point(tb, i, bessel_t, ls_t, ip):
    """tensor pass of point i: Delta, iCl_tensor [4][47], Cl_tensor [4][601] (TT, EE, BB, TE; dimensionless)"""
    import pyoracle as o
    th = tb["thermo"][i]
    nt, nk = tb["n_tau"][i], tb["n_k"][i]
    src = np.ascontiguousarray(tb["src"][i, :nt, :, :nk])
    q, dq, Delta, triples = o.project(bessel_t, th[0], th[1], th[2], th[3], th[4], MAX_ETA_K_T, MAX_L_T, True,
                                      tb["k"][i, :nk], src)
    iCl = o.calc_cls(q, dq, ls_t, Delta, ip, 1.0, tensors=True)
    cl = np.zeros((4, MAX_L_T + 1))
    for X in range(4):
        cl[X, :] = o.interp_cl(ls_t, iCl[X])[:MAX_L_T + 1]
    return dict(q=q, iCl=iCl, cl=cl, triples=triples)


def oracle_point_with_tensors(batch, i, bessel, ls, tmpl_unl, highl_lensed, cl_tensor, ip=None, alens=None):
    """scalar chain of point i with initial power `ip`, tensors added in SetPowersFromCAMB"""
    import pyoracle as o
    th = batch["thermo"][i]
    nt, nk = batch["n_tau"][i], batch["n_k"][i]
    src = np.ascontiguousarray(batch["src"][i, :nt, :, :nk])
    q, dq, Delta, triples = o.project(bessel, th[0], th[1], th[2], th[3], th[4], MAX_ETA_K, MAX_L, False,
                                      batch["k"][i, :nk], src)
    return oracle_from_delta(q, dq, Delta, ls, tmpl_unl, highl_lensed, batch["initpower"][i] if ip is None else ip,
                             batch["alens"][i] if alens is None else alens, cl_tensor)


def oracle_from_delta(q, dq, Delta, ls, tmpl_unl, highl_lensed, ip, alens, cl_tensor=None):
    import pyoracle as o
    iCl = o.calc_cls(q, dq, ls, Delta, ip, alens)
    cl = np.zeros((6, MAX_L + 1))
    for X in range(6):
        cl[X, :] = o.interp_cl(ls, iCl[X], template_index=X + 1 if X < 3 else 0, tmpl=tmpl_unl)[:MAX_L + 1]
    lensed = o.lens_cls(ls, MAX_L, np.stack([cl[0], cl[1], cl[2], cl[3]]), tmpl_unl)
    lens_pad = np.zeros((4, MAX_L + 1))
    lens_pad[:, :lensed.shape[1]] = lensed
    out, hn, rms = o.set_powers(lens_pad, cl[3], LMAX_COMPUTED, [LMAX_OUT] * 5, highl_lensed, lmax_out=LMAX_OUT,
                                cl_tensor=cl_tensor, lmax_tensor=MAX_L_T if cl_tensor is not None else 0)
    return dict(iCl=iCl, cl=cl, cls_out=out, rms=rms)
