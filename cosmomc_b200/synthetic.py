"""Synthetic inputs of the named shapes (BASELINE.json configs / SURVEY 8d): there is no Fortran compiler in
this environment, so the Boltzmann source functions Src(k, source, tau) that CAMB's ODE stage would hand to the
hot path are replaced by analytic, CMB-shaped stand-ins (visibility-weighted acoustic oscillations with Silk
damping, a reionisation bump, an ISW tail and a lensing-potential kernel), one smooth random perturbation per
parameter point.  Grids (time steps, source wavenumbers) are the real ones, built by the library's own
bit-exact grid builders from per-point thermal-history scalars.

Nothing here is on the measured path: bench.py generates the batch before the timed region.
"""
import numpy as np
import torch


def draw_thermo(npts, seed=7):
    """thermo[npts][5] = tau0, taurst, taurend, reion_tau_start, reion_tau_complete (Mpc)."""
    rng = np.random.default_rng(seed)
    th = np.zeros((npts, 5))
    th[:, 0] = rng.normal(14160.0, 60.0, npts)
    th[:, 1] = rng.normal(231.0, 1.5, npts)
    th[:, 2] = rng.normal(465.0, 4.0, npts)
    th[:, 3] = rng.normal(4300.0, 50.0, npts)
    th[:, 4] = th[:, 3] + rng.normal(1100.0, 30.0, npts)
    return th


def draw_params(npts, seed=7):
    """initpower[npts][10], alens, calPlanck and the per-point source perturbation parameters."""
    rng = np.random.default_rng(seed + 1)
    ip = np.zeros((npts, 10))
    logA = rng.normal(3.044, 0.014, npts)
    ip[:, 0] = 1e-10 * np.exp(logA)          # cl_norm * As (CosmologyTypes.f90:18)
    ip[:, 1] = rng.normal(0.9649, 0.004, npts)
    ip[:, 7] = 0.05
    ip[:, 8] = 0.05
    ip[:, 9] = 1.0
    alens = np.ones(npts)
    cal = rng.normal(1.0, 0.0025, npts)
    pert = rng.normal(0.0, 1.0, (npts, 6))
    return ip, alens, cal, pert


def build_grids(handle, thermo, kind=0):
    """Per-point padded tau/dtau/k arrays from the library's grid builders."""
    npts = len(thermo)
    NT, NK = handle.info.n_tau_max, handle.info.n_k_max
    tau = np.zeros((npts, NT))
    dtau = np.zeros((npts, NT))
    k = np.zeros((npts, NK))
    n_tau = np.zeros(npts, dtype=np.int32)
    n_k = np.zeros(npts, dtype=np.int32)
    for i in range(npts):
        t, dt = handle.time_steps(thermo[i, 0], thermo[i, 1], thermo[i, 2], thermo[i, 3], thermo[i, 4], kind)
        kk = handle.source_k(thermo[i, 0], thermo[i, 1], kind)
        if len(t) > NT or len(kk) > NK:
            raise ValueError("grid exceeds capacity: n_tau=%d n_k=%d" % (len(t), len(kk)))
        n_tau[i] = len(t)
        n_k[i] = len(kk)
        tau[i, :len(t)] = t
        dtau[i, :len(t)] = dt
        k[i, :len(kk)] = kk
        k[i, len(kk):] = kk[-1]
        tau[i, len(t):] = t[-1]
    return tau, dtau, n_tau, k, n_k


def make_sources(thermo, tau, k, pert, device="cpu", out=None):
    """Src[p][n][s][i] for s = T, E, lensing-potential; float64 torch tensor on `device`."""
    dev = torch.device(device)
    th = torch.as_tensor(thermo, dtype=torch.float64, device=dev)
    t = torch.as_tensor(tau, dtype=torch.float64, device=dev)[:, :, None]      # [P][NT][1]
    kk = torch.as_tensor(k, dtype=torch.float64, device=dev)[:, None, :]       # [P][1][NK]
    pe = torch.as_tensor(pert, dtype=torch.float64, device=dev)
    P = th.shape[0]
    tau0 = th[:, 0].view(P, 1, 1)
    eps = 0.02 * pe
    tau_rec = (281.0 + 2.0 * pe[:, 0]).view(P, 1, 1)
    sig_rec = (17.0 * (1 + eps[:, 1])).view(P, 1, 1)
    rs = (144.4 * (1 + 0.3 * eps[:, 2])).view(P, 1, 1)
    kD = (0.14 * (1 + eps[:, 3])).view(P, 1, 1)
    a_isw = (1.0 + 5 * eps[:, 4]).view(P, 1, 1)
    a_phi = (1.0 + eps[:, 5]).view(P, 1, 1)
    chi = torch.clamp(tau0 - t, min=1e-3)
    # visibility: recombination peak + reionisation bump
    g_rec = torch.exp(-0.5 * ((t - tau_rec) / sig_rec) ** 2) / (sig_rec * np.sqrt(2 * np.pi))
    t_re = (0.5 * (th[:, 3] + th[:, 4])).view(P, 1, 1)
    s_re = (0.25 * (th[:, 4] - th[:, 3])).view(P, 1, 1)
    g_re = 0.055 * torch.exp(-0.5 * ((t - t_re) / s_re) ** 2) / (s_re * np.sqrt(2 * np.pi))
    damp = torch.exp(-(kk / kD) ** 2)
    q = kk / 0.143
    Tk = torch.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    # temperature: Sachs-Wolfe plateau + acoustic oscillation + late ISW
    grow = torch.clamp((t - 3000.0) / (tau0 - 3000.0), min=0.0)
    S_T = (g_rec + g_re) * (0.2 * Tk + 0.42 * torch.cos(kk * rs) * damp) + a_isw * 2.0e-5 * Tk * grow ** 2
    # E polarisation: quadrupole ~ k * velocity at last scattering, with the 1/x^2 of the E source
    S_E = (g_rec * 0.06 * (kk * sig_rec) * torch.sin(kk * rs) * damp + g_re * 0.05 * Tk) * (15.0 / 8.0) / (kk * chi) ** 2
    # lensing potential: 2 phi (chi* - chi) / (chi* chi) after last scattering
    chis = tau0 - tau_rec
    W = torch.where(t > tau_rec, (chis - chi) / (chis * chi), torch.zeros_like(chi))
    D = 1.0 - 0.25 * grow ** 2
    S_P = -2.0 * a_phi * 0.6 * Tk * D * W
    if out is None:
        out = torch.empty((P, t.shape[1], 3, kk.shape[2]), dtype=torch.float64, device=dev)
    out[:, :, 0, :] = S_T
    out[:, :, 1, :] = S_E
    out[:, :, 2, :] = S_P
    return out


def synthetic_pliklite(lmax=2508, seed=99, fiducial_cls=None):
    """plik-lite-shaped data set: 613 bins = 215 TT + 199 TE + 199 EE over l = 30..2508 with
    top-hat bins of width 5/10/17/33, flat weights, synthetic SPD covariance, data = fiducial + noise."""
    plmin = 30
    edges = []
    l = plmin
    while l <= lmax:
        if l < 100:
            w = 5
        elif l < 1504:
            w = 9
        elif l < 2014:
            w = 17
        else:
            w = 33
        hi = min(l + w - 1, lmax)
        edges.append((l, hi))
        l = hi + 1
    nb_tt = 215
    edges = edges[:nb_tt] if len(edges) >= nb_tt else edges
    # pad/trim to exactly 215 bins by splitting the widest bins from the top
    while len(edges) < nb_tt:
        j = max(range(len(edges)), key=lambda i: edges[i][1] - edges[i][0])
        lo, hi = edges[j]
        mid = (lo + hi) // 2
        edges[j:j + 1] = [(lo, mid), (mid + 1, hi)]
    blmin = np.array([e[0] for e in edges], dtype=np.int32)
    blmax = np.array([e[1] for e in edges], dtype=np.int32)
    nb = np.array([215, 199, 199], dtype=np.int32)
    ls = np.arange(lmax + 1, dtype=np.float64)
    weights = np.zeros(lmax + 1)
    for lo, hi in edges:
        weights[lo:hi + 1] = 1.0 / (hi - lo + 1)
    # the reference multiplies the file weights by 2pi/(l(l+1)) at load (CMB.f90:225-232)
    with np.errstate(divide="ignore", invalid="ignore"):
        weights[plmin:] = weights[plmin:] * 2 * np.pi / (ls[plmin:] * (ls[plmin:] + 1))
    weights[:plmin] = 0
    nused = int(nb.sum())
    rng = np.random.default_rng(seed)
    if fiducial_cls is None:
        fid = np.zeros(nused)
    else:
        fid = np.zeros(nused)
        ix = 0
        for s, spec in enumerate([0, 1, 2]):  # TT, TE, EE rows of fiducial_cls [5][lmax+1]
            for j in range(nb[s]):
                fid[ix] = np.dot(fiducial_cls[spec, blmin[j]:blmax[j] + 1], weights[blmin[j]:blmax[j] + 1])
                ix += 1
    sig = 0.02 * np.abs(fid) + 1e-3 * (np.abs(fid).max() if np.abs(fid).max() > 0 else 1.0) * 1e-3
    A = rng.normal(0, 1, (nused, 40)) / np.sqrt(40)
    corr = 0.15 * A @ A.T + np.eye(nused)
    d = np.sqrt(np.diag(corr))
    corr = corr / d[:, None] / d[None, :]
    cov = corr * sig[:, None] * sig[None, :]
    invcov = np.linalg.inv(cov)
    invcov = 0.5 * (invcov + invcov.T)
    x_data = fid + np.linalg.cholesky(cov) @ rng.normal(0, 1, nused)
    return dict(nb=nb, blmin=blmin, blmax=blmax, weights=weights, invcov=invcov, x_data=x_data, cov=cov)


def synthetic_sn_covs(lc, names=("mag", "stretch", "colour", "mag_stretch", "mag_colour", "stretch_colour"), seed=2024,
                      rank=24):
    """Documented stand-ins for the supernova covariance blocks that are MISSING from the reference checkout
    (all six JLA matrices and Pantheon's sys_full_long.txt).  Each diagonal block is a low-rank
    systematics matrix  B B^T  (rank `rank`, amplitude ~ a few 1e-2 mag) and the cross blocks are  B_x B_y^T
    symmetrised, so that V(alpha, beta) stays positive definite over the prior range, like the real matrices."""
    rng = np.random.default_rng(seed)
    n = len(lc["zcmb"])
    z = np.asarray(lc["zcmb"])
    basis = {}
    for nm, amp in (("mag", 0.02), ("stretch", 0.05), ("colour", 0.006)):
        modes = rng.normal(0.0, 1.0, (n, rank)) * amp / np.sqrt(rank)
        modes += amp * 0.5 * np.outer(np.sin(3 * z + rng.uniform(0, 6)), rng.normal(0, 1, rank)) / np.sqrt(rank)
        basis[nm] = modes
    out = {}
    for nm in names:
        if "_" in nm:
            a, b = nm.split("_")
            m = 0.3 * basis[a] @ basis[b].T
            out[nm] = 0.5 * (m + m.T)
        else:
            out[nm] = basis[nm] @ basis[nm].T
    return out


class LowRankCov:
    """cov = diag(d) + U U^T, addressed by index blocks (stand-in for a covariance file that is too big / missing)."""

    def __init__(self, d, U):
        self.d, self.U = np.asarray(d), np.asarray(U)

    def block(self, rows, cols):
        out = self.U[rows] @ self.U[cols].T
        same = np.asarray(rows)[:, None] == np.asarray(cols)[None, :]
        out[same] += self.d[np.asarray(rows)][np.where(same)[0]]
        return out

    def full(self):
        return np.diag(self.d) + self.U @ self.U.T


def synthetic_bk15_cov(scale_per_entry, seed=15, rank=40, frac=0.12):
    """Documented stand-in for data/BK15/BK15_covmat_dust.dat (MISSING from the reference checkout): band-power
    covariance over all (bin, spectrum) entries of the data set, sigma = frac * scale (|fiducial| + |noise| of that
    entry) on the diagonal plus a rank-`rank` correlated part of comparable size."""
    rng = np.random.default_rng(seed)
    sc = np.asarray(scale_per_entry, dtype=np.float64).reshape(-1)
    sig = frac * sc + 1e-6 * sc.max()
    U = rng.normal(0.0, 1.0, (len(sc), rank)) * (0.5 * sig[:, None] / np.sqrt(rank))
    return LowRankCov(sig ** 2, U)
