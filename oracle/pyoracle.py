"""ORACLE (test infrastructure, NOT product code): ctypes front-end of oracle/liborc.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
It wraps the C++ restatement of the reference routines (see oracle/orc_*.hpp for file:line cites).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LMAX_EXTRAP_HIGHL = 8000

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _declare(L):
    L.orc_bessel_create.restype = C.c_void_p
    L.orc_bessel_destroy.argtypes = [C.c_void_p]
    L.orc_bessel_numxx.argtypes = [C.c_void_p]
    L.orc_bessel_get.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
    L.orc_project.restype = C.c_longlong
    L.orc_quadform.restype = C.c_double
    L.orc_pliklite.restype = C.c_double
    L.orc_cmblikes_chisq.restype = C.c_double
    return L


_FAST = None


def fast_lib():
    """Timing build of the same sources (-O3 -march=native -fopenmp) for bench.py's CPU baseline.
    Rebuilt whenever the host CPU differs from the one it was built on (-march=native is host-specific)."""
    global _FAST
    if _FAST is None:
        so = os.path.join(_HERE, "liborc_fast.so")
        tag = os.path.join(_HERE, "liborc_fast.cpu")
        try:
            with open("/proc/cpuinfo") as f:
                cpu = "".join(l for l in f if l.startswith(("model name", "flags")))[:8192]
            import hashlib
            cpu = hashlib.sha1(cpu.encode()).hexdigest()
        except OSError:
            cpu = "unknown"
        have = open(tag).read().strip() if os.path.exists(tag) else ""
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".inc"))]
        if (have != cpu or not os.path.exists(so)
                or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs)):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "fast"])
            with open(tag, "w") as f:
                f.write(cpu)
        _FAST = _declare(C.CDLL(so))
    return _FAST


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _declare(C.CDLL(build()))
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(c_dp if a.dtype == np.float64 else c_ip)


def ranges_build(ops, max_points=200000):
    """ops: list of (kind, start, end, delta_or_nstep, islog); kind 0 = Add_delta, 1 = Add."""
    o = _d(ops).reshape(-1, 5)
    pts = np.zeros(max_points)
    dpts = np.zeros(max_points)
    reg = np.zeros((100, 6))
    n = C.c_int(0)
    cnt = lib().orc_ranges_build(len(o), _p(o), max_points, C.byref(n), _p(pts), _p(dpts), _p(reg))
    if cnt < 0:
        raise RuntimeError("orc_ranges_build failed")
    return pts[: n.value].copy(), dpts[: n.value].copy(), reg[:cnt].copy()


def ranges_indexof(ops, x):
    o = _d(ops).reshape(-1, 5)
    x = _d(x)
    idx = np.zeros(len(x), dtype=np.int32)
    if lib().orc_ranges_indexof(len(o), _p(o), len(x), _p(x), _p(idx)) != 0:
        raise RuntimeError("orc_ranges_indexof failed")
    return idx


def initlval(max_l, lSampleBoost=1.0, AccurateReionization=True):
    out = np.zeros(4100, dtype=np.int32)
    n = lib().orc_initlval(int(max_l), C.c_double(lSampleBoost), int(AccurateReionization), _p(out), len(out))
    if n < 0:
        raise RuntimeError("orc_initlval failed")
    return out[:n].copy()


def spline(x, y, d11=1e40, d1n=1e40):
    x = _d(x)
    y = _d(y)
    d2 = np.zeros_like(x)
    lib().orc_spline(_p(x), _p(y), len(x), C.c_double(d11), C.c_double(d1n), _p(d2))
    return d2


def bjl(L, x):
    L = _i(np.broadcast_to(L, np.shape(x)))
    x = _d(x)
    out = np.zeros_like(x)
    lib().orc_bjl(x.size, _p(L), _p(x), _p(out))
    return out


class Bessel:
    def __init__(self, ls, max_eta_k):
        self.ls = _i(ls)
        self.h = lib().orc_bessel_create(len(self.ls), _p(self.ls), C.c_double(max_eta_k))
        if not self.h:
            raise RuntimeError("orc_bessel_create failed")
        self.num_xx = lib().orc_bessel_numxx(self.h)

    def arrays(self):
        x = np.zeros(self.num_xx)
        ajl = np.zeros((len(self.ls), self.num_xx))
        ajlpr = np.zeros((len(self.ls), self.num_xx))
        lib().orc_bessel_get(self.h, _p(x), _p(ajl), _p(ajlpr))
        return x, ajl, ajlpr

    def __del__(self):
        try:
            if self.h:
                lib().orc_bessel_destroy(self.h)
                self.h = None
        except Exception:
            pass


def time_steps(taurst, taurend, tau0, maximum_qeta, want_tensors, reion_start, reion_complete, max_n=4000):
    tau = np.zeros(max_n)
    dtau = np.zeros(max_n)
    n = lib().orc_time_steps(C.c_double(taurst), C.c_double(taurend), C.c_double(tau0), C.c_double(maximum_qeta),
                             int(want_tensors), C.c_double(reion_start), C.c_double(reion_complete), max_n,
                             _p(tau), _p(dtau))
    if n < 0:
        raise RuntimeError("orc_time_steps failed")
    return tau[:n].copy(), dtau[:n].copy()


def source_k(tau0, taurst, maximum_qeta, want_tensors, maximum_l, max_n=4000):
    k = np.zeros(max_n)
    n = lib().orc_source_k(C.c_double(tau0), C.c_double(taurst), C.c_double(maximum_qeta), int(want_tensors),
                           int(maximum_l), max_n, _p(k))
    if n < 0:
        raise RuntimeError("orc_source_k failed")
    return k[:n].copy()


def q_grid(tau0, maximum_qeta, maximum_l, max_n=20000):
    q = np.zeros(max_n)
    dq = np.zeros(max_n)
    n = lib().orc_q_grid(C.c_double(tau0), C.c_double(maximum_qeta), int(maximum_l), max_n, _p(q), _p(dq))
    if n < 0:
        raise RuntimeError("orc_q_grid failed")
    return q[:n].copy(), dq[:n].copy()


def project(bessel, tau0, taurst, taurend, reion_start, reion_complete, maximum_qeta, maximum_l, want_tensors,
            k_src, src, max_q=20000):
    """src: [n_tau][n_src][n_k].  Returns q, dq, Delta[n_q][nl][n_src], triples."""
    k_src = _d(k_src)
    src = _d(src)
    n_src = src.shape[1]
    nl = len(bessel.ls)
    q = np.zeros(max_q)
    dq = np.zeros(max_q)
    Delta = np.zeros((max_q, nl, n_src))
    trip = C.c_longlong(0)
    n = lib().orc_project(C.c_void_p(bessel.h), nl, _p(bessel.ls), C.c_double(tau0), C.c_double(taurst),
                          C.c_double(taurend), C.c_double(reion_start), C.c_double(reion_complete),
                          C.c_double(maximum_qeta), int(maximum_l), int(want_tensors), len(k_src), _p(k_src), n_src,
                          _p(src), max_q, _p(q), _p(dq), _p(Delta), C.byref(trip))
    if n < 0:
        raise RuntimeError("orc_project failed")
    return q[:n].copy(), dq[:n].copy(), Delta[:n].copy(), trip.value


class BatchChain:
    """Whole per-point chain (projection ... plik-lite-shaped chi^2) for a batch of points in ONE C call
    (orc_eval_pliklite_batch); the CPU baseline of bench.py.  `fast` selects the -O3 -march=native timing build.
    mode 0: points one after the other, OpenMP inside the point; mode 1: one point per thread."""

    def __init__(self, ls, max_eta_k, Max_l, lmax_computed_cl, lmax_out, tmpl_unl, highl, data, fast=True,
                 threads=None):
        self.L = fast_lib() if fast else lib()
        if threads:
            self.L.orc_set_num_threads(int(threads))
        self.threads = int(self.L.orc_num_threads())
        self.ls = _i(ls)
        self.args = (float(max_eta_k), int(Max_l), int(lmax_computed_cl), int(lmax_out))
        self.tm, self.hl = _d(tmpl_unl), _d(highl)
        self.data = {k: (_d(v) if k in ("weights", "invcov", "x_data") else _i(v)) for k, v in data.items()
                     if k in ("nb", "blmin", "blmax", "weights", "invcov", "x_data")}
        self.h = self.L.orc_bessel_create(len(self.ls), _p(self.ls), C.c_double(max_eta_k))
        if not self.h:
            raise RuntimeError("orc_bessel_create failed")

    def run(self, batch, mode):
        L, d = self.L, self.data
        src = _d(batch["src"])
        npts, NT, _, NK = src.shape
        out = np.zeros(npts)
        mek, Max_l, lcc, lmo = self.args
        rc = L.orc_eval_pliklite_batch(
            C.c_void_p(self.h), len(self.ls), _p(self.ls), Max_l, C.c_double(mek), lcc, lmo,
            int(npts), int(NT), int(NK), _p(_d(batch["thermo"])), _p(_i(batch["n_tau"])), _p(_i(batch["n_k"])),
            _p(_d(batch["k"])), _p(src), _p(_d(batch["initpower"])), _p(_d(batch["alens"])), _p(_d(batch["cal"])),
            _p(self.tm), _p(self.hl), self.hl.shape[1], _p(d["nb"]), _p(d["blmin"]), _p(d["blmax"]),
            _p(d["weights"]), _p(d["invcov"]), _p(d["x_data"]), int(mode), _p(out))
        if rc != 0:
            raise RuntimeError("orc_eval_pliklite_batch failed")
        return out

    def close(self):
        if self.h:
            self.L.orc_bessel_destroy(C.c_void_p(self.h))
            self.h = None


def initpower_vec(As=2.1e-9, ns=0.96, nrun=0.0, nrunrun=0.0, r=0.0, nt=0.0, ntrun=0.0, pivot_k=0.05,
                  tensor_pivot_k=0.05, inflation_consistency=True):
    return _d([As, ns, nrun, nrunrun, r, nt, ntrun, pivot_k, tensor_pivot_k, float(inflation_consistency)])


def scalar_power(ip, k):
    k = _d(k)
    out = np.zeros_like(k)
    lib().orc_scalar_power(_p(_d(ip)), k.size, _p(k), _p(out))
    return out


def tensor_power(ip, k):
    k = _d(k)
    out = np.zeros_like(k)
    lib().orc_tensor_power(_p(_d(ip)), k.size, _p(k), _p(out))
    return out


def calc_cls(q, dq, ls, Delta, ip, ALens=1.0, tensors=False):
    q = _d(q)
    dq = _d(dq)
    ls = _i(ls)
    Delta = _d(Delta)
    nX = 4 if tensors else 6
    iCl = np.zeros((nX, len(ls)))
    if lib().orc_calc_cls(int(tensors), len(q), _p(q), _p(dq), len(ls), _p(ls), Delta.shape[2], _p(Delta),
                          _p(_d(ip)), C.c_double(ALens), _p(iCl)) != 0:
        raise RuntimeError("orc_calc_cls failed")
    return iCl


def load_highl_template(path):
    """camb/modules.f90:1162-1185: columns L, TT, EE, BB, TE, PP,... -> [4][8001] = TT, EE, TE, PP."""
    a = np.loadtxt(path)
    t = np.zeros((4, LMAX_EXTRAP_HIGHL + 1))
    L = a[:, 0].astype(int)
    m = L <= LMAX_EXTRAP_HIGHL
    t[0, L[m]] = a[m, 1]
    t[1, L[m]] = a[m, 2]
    t[2, L[m]] = a[m, 4]
    t[3, L[m]] = a[m, 5]
    return t


def interp_cl(ls, iCl, max_ind=None, template_index=0, tmpl=None):
    ls = _i(ls)
    iCl = _d(iCl)
    if max_ind is None:
        max_ind = len(ls)
    out = np.zeros(int(ls[max_ind - 1]) + 1)
    t = _d(tmpl) if tmpl is not None else None
    if lib().orc_interp_cl(len(ls), _p(ls), _p(iCl), int(max_ind), int(template_index), _p(t), _p(out)) != 0:
        raise RuntimeError("orc_interp_cl failed")
    return out


def lens_cls(ls, Max_l, cl_scalar, tmpl, accuracy_boost=None, accurate_bb=False):
    """cl_scalar: [4][Max_l+1] TT,EE,TE,PP(l^4 C_phi) dimensionless; returns ([4][lmax_lensed+1] TT,EE,BB,TE).
    accuracy_boost / accurate_bb: the reference's AccuracyBoost and accurate_BB knobs (default: the production settings)."""
    ls = _i(ls)
    cl_scalar = _d(cl_scalar)
    assert cl_scalar.shape == (4, Max_l + 1)
    out = np.zeros((4, Max_l + 1))
    if accuracy_boost is not None or accurate_bb:
        lml = lib().orc_lens_cls_opts(len(ls), _p(ls), int(Max_l), _p(cl_scalar), _p(_d(tmpl)), _p(out), Max_l + 1,
                                      C.c_double(accuracy_boost or 1.0), int(accurate_bb))
    else:
        lml = lib().orc_lens_cls(len(ls), _p(ls), int(Max_l), _p(cl_scalar), _p(_d(tmpl)), _p(out), Max_l + 1)
    if lml < 0:
        raise RuntimeError("orc_lens_cls failed")
    return out[:, : lml + 1].copy()


def set_powers(cl_lensed, cl_phi, lmax_computed_cl, cl_lmax, highl, Aphiphi=1.0, cl_tensor=None, lmax_tensor=0,
               highL_norm=0.0, lmax_out=None):
    cl_lensed = _d(cl_lensed)
    cl_phi = _d(cl_phi)
    highl = _d(highl)
    cl_lmax = _i(cl_lmax)
    if lmax_out is None:
        lmax_out = int(cl_lmax.max())
    out = np.zeros((5, lmax_out + 1))
    hn = C.c_double(highL_norm)
    rms = C.c_double(0)
    ct = _d(cl_tensor) if cl_tensor is not None else None
    if lib().orc_set_powers(_p(cl_lensed), cl_lensed.shape[1], _p(cl_phi), _p(ct), ct.shape[1] if ct is not None else 0,
                            int(lmax_tensor), int(lmax_computed_cl), _p(cl_lmax), _p(highl), highl.shape[1],
                            C.c_double(Aphiphi), C.byref(hn), _p(out), lmax_out, C.byref(rms)) != 0:
        raise RuntimeError("orc_set_powers failed")
    return out, hn.value, rms.value


def get_loglike(P, likes, pmin=None, pmax=None, prior_mean=None, prior_std=None, use_prior=None, lincomb=None,
                lincomb_mean=None, lincomb_std=None, temperature=1.0, soft_error=None):
    """TLikeCalculator%GetLogLike for rows of P given the per-likelihood -lnL (source/calclike.f90:97-151): hard bounds
    -> logZero, sum of likelihoods / Temperature (AddLikeTemp :80-94, logZero propagates), Gaussian and linear-
    combination priors / Temperature (GetLogPriors :111-134).  Returns (loglike, prior, status)."""
    P = np.atleast_2d(np.asarray(P, dtype=np.float64))
    likes = np.atleast_2d(np.asarray(likes, dtype=np.float64))
    npts, n = P.shape
    logZero = 1e30
    out = np.zeros(npts)
    prior = np.zeros(npts)
    st = np.zeros(npts, dtype=np.int32)
    for i in range(npts):
        lp = 0.0
        if prior_std is not None:
            for j in range(n):
                if (use_prior is None or use_prior[j]) and prior_std[j] != 0:
                    lp += ((P[i, j] - (prior_mean[j] if prior_mean is not None else 0.0)) / prior_std[j]) ** 2
        if lincomb is not None:
            for c, comb in enumerate(np.atleast_2d(lincomb)):
                if lincomb_std[c] != 0:
                    lp += ((np.dot(comb, P[i]) - lincomb_mean[c]) / lincomb_std[c]) ** 2
        prior[i] = lp / 2
        oob = (pmax is not None and np.any(P[i] > np.asarray(pmax))) or (pmin is not None and np.any(P[i] < np.asarray(pmin)))
        if oob:
            st[i] = 1
            out[i] = logZero
            continue
        if soft_error is not None and soft_error[i]:
            st[i] = 1 + int(soft_error[i])
        if st[i] != 0 or np.any(~(likes[i] < logZero)):
            out[i] = logZero
        else:
            out[i] = likes[i].sum() / temperature + prior[i] / temperature
    return out, prior, st


def quadform(M, v):
    M = _d(M)
    v = _d(v)
    return lib().orc_quadform(_p(M), _p(v), len(v))


def pliklite(cls, nb, blmin, blmax, weights, invcov, X_data, cal):
    cls = _d(cls)
    return lib().orc_pliklite(_p(cls), cls.shape[1], _p(_i(nb)), _p(_i(blmin)), _p(_i(blmax)), _p(_d(weights)),
                              _p(_d(invcov)), _p(_d(X_data)), C.c_double(cal))


def cmblikes_chisq(nmaps, nbins, cl_use_index, like_approx, NoiseM, ChatM, sqrt_fid, inv_cov, binnedC):
    cui = _i(cl_use_index)
    return lib().orc_cmblikes_chisq(int(nmaps), int(nbins), len(cui), _p(cui), int(like_approx),
                                    _p(_d(NoiseM)) if NoiseM is not None else None, _p(_d(ChatM)),
                                    _p(_d(sqrt_fid)) if sqrt_fid is not None else None, _p(_d(inv_cov)),
                                    _p(_d(binnedC)))


# ---------------------------------------------------------------- background (orc_bg.hpp) and background likelihoods
def background(bg, z):
    """bg [16] (layout: cosmomc_b200/params.py) -> D_A(z), H(z) [Mpc^-1], (tau0, age/Gyr, CosmomcTheta)."""
    lib().orc_background.restype = C.c_int
    bg = _d(bg)
    z = _d(np.atleast_1d(z))
    DA = np.zeros_like(z)
    H = np.zeros_like(z)
    ex = np.zeros(3)
    if lib().orc_background(_p(bg), len(z), _p(z), _p(DA), _p(H), _p(ex)) != 0:
        raise RuntimeError("orc_background failed")
    return DA, H, ex


def nu_table():
    r1 = np.zeros(2000)
    dr1 = np.zeros(2000)
    dl = C.c_double(0)
    lib().orc_nu_table(_p(r1), _p(dr1), C.byref(dl))
    return r1, dr1, dl.value


CONST_C = 2.99792458e8  # source/settings.f90 const_c

# measurement types of source/bao.f90:29-35 (1-based codes as in the reference)
BAO_TYPES = ['Az', 'DV_over_rs', 'rs_over_DV', 'DA_over_rs', 'F_AP', 'f_sigma8', 'bao_Hz_rs', 'bao_Hz_rs_103',
             'dilation', 'DM_over_rs']


def bao_loglike(bg, rs_drag, rs_rescale, types, zs, obs, invcov):
    """BAO_LnLike (source/bao.f90:265-308).  types: 1-based codes into BAO_TYPES."""
    rs = rs_drag * rs_rescale
    DA, H, _ = background(bg, zs)
    th = np.zeros(len(zs))
    for j, (t, z) in enumerate(zip(types, zs)):
        Dv = ((DA[j] * (1 + z)) ** 2 * z / H[j]) ** (1. / 3.)
        if t == 2:
            th[j] = Dv / rs
        elif t == 7:
            th[j] = CONST_C * H[j] / 1e3 * rs
        elif t == 8:
            th[j] = CONST_C * H[j] / 1e3 * rs * 1.0e-3
        elif t == 3:
            th[j] = rs / Dv
        elif t == 1:
            omegam = 1.0 - bg[4] - (1 - (bg[1] + bg[2] + bg[3] + bg[4]))
            omh2 = omegam * (bg[0] / 100) ** 2
            th[j] = 100 * Dv * np.sqrt(omh2) / (CONST_C / 1e3 * z)
        elif t == 4:
            th[j] = DA[j] / rs
        elif t == 10:
            th[j] = (1 + z) * DA[j] / rs
        elif t == 5:
            th[j] = (1 + z) * DA[j] * H[j]
        else:
            raise ValueError("unsupported BAO type")
    d = th - np.asarray(obs)
    return quadform(np.asarray(invcov), d) / 2


def mgs_loglike(bg, rs_drag, z, alpha_prob):
    """BAO_MGS_loglike (source/bao.f90:390-410)."""
    DA, H, _ = background(bg, [z])
    Dv = ((DA[0] * (1 + z)) ** 2 * z / H[0]) ** (1. / 3.)
    alphamgs = Dv / rs_drag / (638.9518 / 148.69)
    if alphamgs > 1.1985 or alphamgs < 0.8005:
        return 1e30
    ii = 1 + int(np.floor((alphamgs - 0.8005) / np.float64(np.float32(0.001))))
    return (alpha_prob[ii - 1] + alpha_prob[ii]) / 2.0 / 2.0


def hst_loglike(bg, H0_obs, H0_err, zeff=0.0, angconversion=0.0):
    """HST_LnLike (source/HST.f90:47-59)."""
    if zeff > 0:
        DA, _, _ = background(bg, [zeff])
        th = angconversion / DA[0]
    else:
        th = bg[0]
    return (th - H0_obs) ** 2 / (2 * H0_err ** 2)


class SN:
    """JLA / Pantheon likelihood (source/supernovae_JLA.f90:874-991 jla_prep, :773-866 invert_covariance_matrix,
    :1028-1168 JLA_alpha_beta_like, :1170-1228 jla_LnLike), LAPACK through numpy/scipy."""

    def __init__(self, lc, covs, pecz=0.0, twoscriptmfit=False, scriptmcut=10.0, intrinsicdisp=0.0):
        # lc: dict of columns zcmb zhel dz mb dmb x1 dx1 color dcolor 3rdvar cov_m_s cov_m_c cov_s_c ; covs: dict
        self.lc = {k: np.asarray(v, dtype=np.float64) for k, v in lc.items()}
        self.covs = covs
        L = self.lc
        self.nsn = len(L["zcmb"])
        zfacsq = 25.0 / np.float64(np.float32(np.log(np.float32(10.0)))) ** 2
        self.pre_vars = L["dmb"] ** 2 + intrinsicdisp ** 2 + zfacsq * pecz ** 2 * (
            (1.0 + L["zcmb"]) / (L["zcmb"] * (1 + 0.5 * L["zcmb"]))) ** 2
        self.twoscriptmfit = twoscriptmfit
        if twoscriptmfit:
            self.A1 = (L["3rdvar"] <= scriptmcut).astype(np.float64)
            self.A2 = 1.0 - self.A1
            if not self.A1.any():
                self.A1, self.A2 = self.A2, np.zeros(self.nsn)
                self.twoscriptmfit = False
            if not self.A2.any():
                self.twoscriptmfit = False

    def diag(self, alpha, beta):
        L = self.lc
        return (self.pre_vars + alpha * alpha * L["dx1"] ** 2 + beta * beta * L["dcolor"] ** 2
                + 2.0 * alpha * L["cov_m_s"] - 2.0 * beta * L["cov_m_c"] - 2.0 * alpha * beta * L["cov_s_c"])

    def covariance(self, alpha, beta):
        c = self.covs
        V = np.zeros((self.nsn, self.nsn))
        if "mag" in c: V = V + c["mag"]
        if "stretch" in c: V = V + alpha * alpha * c["stretch"]
        if "colour" in c: V = V + beta * beta * c["colour"]
        if "mag_stretch" in c: V = V + 2.0 * alpha * c["mag_stretch"]
        if "mag_colour" in c: V = V - 2.0 * beta * c["mag_colour"]
        if "stretch_colour" in c: V = V - 2.0 * alpha * beta * c["stretch_colour"]
        V[np.diag_indices(self.nsn)] += self.diag(alpha, beta)
        return V

    def alpha_beta_like(self, lumdists, alpha, beta):
        import scipy.linalg as sl
        L = self.lc
        invvars = 1.0 / self.diag(alpha, beta)
        wtval = invvars.sum()
        est = ((L["mb"] - lumdists) * invvars).sum() / wtval
        diffmag = L["mb"] - lumdists + alpha * L["x1"] - beta * L["color"] - est
        cf = sl.cho_factor(self.covariance(alpha, beta), lower=False)   # DPOTRF('U')
        inv = sl.cho_solve(cf, np.eye(self.nsn))                       # DPOTRI
        iv = inv @ diffmag                                             # DSYMV
        A = diffmag @ iv
        if self.twoscriptmfit:
            B = iv @ self.A1
            Cc = iv @ self.A2
            iv = inv @ self.A1
            D = iv @ self.A2
            E = iv @ self.A1
            iv = inv @ self.A2
            F = iv @ self.A2
            G = F - D * D / E
            chisq = (A + np.log(E / (2 * np.pi)) + np.log(G / (2 * np.pi)) - Cc * Cc / G - B * B * F / (E * G)
                     + 2.0 * B * Cc * D / (E * G))
        else:
            B = iv.sum()
            E = np.trace(inv) + 2.0 * np.triu(inv, 1).sum()
            chisq = A + np.log(E / (2 * np.pi)) - B ** 2 / E
        return chisq / 2

    def loglike(self, DA, alpha=0.0, beta=0.0):
        L = self.lc
        lumdists = 5.0 * np.log10((1.0 + L["zhel"]) * (1.0 + L["zcmb"]) * DA)
        return self.alpha_beta_like(lumdists, alpha, beta)


# ---------------------------------------------------------------- BICEP/Keck foregrounds (source/CMB_BK_Planck.f90)
_BK_T_CMB = 2.72548
_BK_GK = 6.62606957e-34 / 1.3806488e-23 * 1e9


def bk_dust_scaling(beta, Tdust, bp, nu0, bcerr):
    """DustScaling (:109-147). bp: dict nu, R, dnu, th_dust, nu_bar."""
    nu = bp["nu"]
    gb_int = np.sum(bp["dnu"] * bp["R"] * nu ** (3 + beta) / (np.exp(_BK_GK * nu / Tdust) - 1))
    gb0 = nu0 ** (3 + beta) / (np.exp(_BK_GK * nu0 / Tdust) - 1)
    th_err = gb_err = 1.0
    if bcerr != 1.:
        nb = bp["nu_bar"]
        th_err = bcerr ** 4 * np.exp(_BK_GK * nb * (bcerr - 1) / _BK_T_CMB) * (np.exp(_BK_GK * nb / _BK_T_CMB) - 1) ** 2 / \
            (np.exp(_BK_GK * nb * bcerr / _BK_T_CMB) - 1) ** 2
        gb_err = bcerr ** (3 + beta) * (np.exp(_BK_GK * nb / Tdust) - 1) / (np.exp(_BK_GK * nb * bcerr / Tdust) - 1)
    return (gb_int / gb0) / bp["th_dust"] * (gb_err / th_err)


def bk_sync_scaling(beta, bp, nu0, bcerr):
    """SyncScaling (:151-183)"""
    nu = bp["nu"]
    pl_int = np.sum(bp["dnu"] * bp["R"] * nu ** (2 + beta))
    pl0 = nu0 ** (2 + beta)
    th_err = pl_err = 1.0
    if bcerr != 1.:
        nb = bp["nu_bar"]
        th_err = bcerr ** 4 * np.exp(_BK_GK * nb * (bcerr - 1) / _BK_T_CMB) * (np.exp(_BK_GK * nb / _BK_T_CMB) - 1) ** 2 / \
            (np.exp(_BK_GK * nb * bcerr / _BK_T_CMB) - 1) ** 2
        pl_err = bcerr ** (2 + beta)
    return (pl_int / pl0) / bp["th_sync"] * (pl_err / th_err)


def bk_decorrelation(Delta, nu0, nu1, nupivot, l, lform):
    """Decorrelation (:187-227); lform 0 flat / 1 lin / 2 quad; l array"""
    scl_nu = np.log(nu0 / nu1) ** 2 / np.log(nupivot[0] / nupivot[1]) ** 2
    scl_ell = np.ones_like(l, dtype=float) if lform == 0 else (l / 80.0) if lform == 1 else (l / 80.0) ** 2
    if Delta > 1:
        return 2.0 - np.exp(np.log(2.0 - Delta) * scl_nu * scl_ell)
    return np.exp(np.log(Delta) * scl_nu * scl_ell)


def bk_foregrounds(plan, P):
    """TBK_planck_AddForegrounds (:229-340): foreground D_l added to every used map pair -> [ncl][lmax+1].
    plan: cosmomc_b200.datasets.BK15Plan (set-up data only); P: the 16 BK15.paramnames values."""
    Adust, Async, alphadust, betadust, Tdust, alphasync, betasync, corr, EEd, EEs, Dd, Ds = P[:12]
    n = plan.nmaps
    bcerr = [1.0 if c == 0 else P[12] + P[12 + c] + 1. for c in plan.bc_class]
    fd = [bk_dust_scaling(betadust, Tdust, plan.bandpasses[i], plan.fpivot_dust, bcerr[i]) for i in range(n)]
    fs = [bk_sync_scaling(betasync, plan.bandpasses[i], plan.fpivot_sync, bcerr[i]) for i in range(n)]
    l = np.arange(plan.pcl_lmin, plan.pcl_lmax + 1).astype(float)
    dustpow = Adust * (l / 80.0) ** alphadust
    syncpow = Async * (l / 80.0) ** alphasync
    dspow = corr * np.sqrt(Adust * Async) * (l / 80.0) ** ((alphadust + alphasync) / 2)
    need_d, need_s = abs(Dd - 1) > 1e-5, abs(Ds - 1) > 1e-5
    out = np.zeros((plan.ncl, plan.pcl_lmax + 1))
    ix = 0
    for i in range(n):
        for j in range(i + 1):
            fi, fj = plan.used_fields[i], plan.used_fields[j]
            if fi == fj and fi in (1, 2):
                dust, sync, ds = fd[i] * fd[j], fs[i] * fs[j], fd[i] * fs[j] + fs[i] * fd[j]
                if fi == 1:
                    dust, sync, ds = dust * EEd, sync * EEs, ds * np.sqrt(EEd * EEs)
                dd = dsy = 1.0
                if need_d and i != j:
                    dd = bk_decorrelation(Dd, plan.bandpasses[i]["nu_bar"] * bcerr[i], plan.bandpasses[j]["nu_bar"] * bcerr[j],
                                          plan.fpivot_dust_decorr, l, plan.lform_dust)
                if need_s and i != j:
                    dsy = bk_decorrelation(Ds, plan.bandpasses[i]["nu_bar"] * bcerr[i], plan.bandpasses[j]["nu_bar"] * bcerr[j],
                                           plan.fpivot_sync_decorr, l, plan.lform_sync)
                out[ix, plan.pcl_lmin:] = dust * dustpow * dd + sync * syncpow * dsy + ds * dspow
            ix += 1
    return out


def bk_loglike(plan, cls, P):
    """-lnL of a BK15-style data set: theory Cls [5][>=lmax+1] (CosmoMC units) + foregrounds -> binned -> HL chi^2/2"""
    binned = plan.binned_theory(cls) + np.einsum("bcl,cl->bc", plan.fgW, bk_foregrounds(plan, P))
    chi2 = cmblikes_chisq(plan.nmaps, plan.nbins_used, plan.cl_use_index, plan.like_approx, plan.noise, plan.chat,
                          plan.sqrt_fid, plan.invcov, binned)
    return chi2 / 2


# ---------------------------------------------------------------- thermal history (orc_thermo.hpp)
THERMO_DERIVED = ["age", "zstar", "rstar", "thetastar", "DAstar", "zdrag", "rdrag", "kd", "thetad", "zeq", "keq",
                  "thetaeq", "thetarseq"]


def thermo(bg, yhe, zre=0.0, optical_depth=0.0, max_eta_k=14000.0, want_tensors=False, transfer_kmax_h=5.0,
           accuracy_boost=1.0, tables=False):
    """CAMBParams_Set tail + cmbmain set-up + inithermo (camb/modules.f90:2682-2992) for one bg[16] row.
    optical_depth > 0: the reionisation redshift comes from the bisection of Reionization_zreFromOptDepth.
    Returns a dict (tau0, taurst, taurend, tau_start, tau_complete, dtaurec, tau_maxvis, zre, z_star, z_drag,
    actual_opt_depth, status, derived{...}, n_fcn[, tables [4][20000] = xe, dotmu, emmu, cs2])."""
    L = lib()
    L.orc_thermo.restype = C.c_int
    inp = _d([yhe, zre, optical_depth, max_eta_k, 1.0 if want_tensors else 0.0, transfer_kmax_h, accuracy_boost, 0.0])
    out = np.zeros(32)
    tab = np.zeros((4, 20000)) if tables else None
    if L.orc_thermo(_p(_d(bg)), _p(inp), _p(out), _p(tab) if tables else None) != 0:
        raise RuntimeError("orc_thermo failed")
    keys = ["tau0", "taurst", "taurend", "tau_start", "tau_complete", "dtaurec", "tau_maxvis", "zre", "z_star", "z_drag",
            "actual_opt_depth", "status"]
    r = {k: out[i] for i, k in enumerate(keys)}
    r["status"] = int(r["status"])
    r["derived"] = {k: out[12 + i] for i, k in enumerate(THERMO_DERIVED)}
    r["n_fcn"] = int(out[25])
    if tables:
        r["tables"] = tab
    return r


def recfast_xe(bg, yhe, a):
    """Recombination_xe (camb/recfast.f90:434-456) after Recombination_init, at scale factors a."""
    L = lib()
    L.orc_recfast_xe.restype = C.c_int
    L.orc_recfast_xe.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    a = _d(np.atleast_1d(a))
    xe = np.zeros_like(a)
    if L.orc_recfast_xe(_p(_d(bg)), float(yhe), len(a), _p(a), _p(xe)) != 0:
        raise RuntimeError("orc_recfast_xe failed")
    return xe


def h0_from_theta(theta100, make_bg, H0_min=20.0, H0_max=100.0):
    """ThetaParameterization%ParamArrayToTheoryParams (source/CosmologyParameterizations.f90:134-176): bisection on H0
    until successive CosmomcTheta values differ by < 1e-7.  make_bg(H0) -> bg[16] row.  Returns H0 (0: out of range)."""
    DA = theta100 / 100
    theta = lambda H0: background(make_bg(H0), [0.0])[2][2]
    try_b, try_t = H0_min, H0_max
    D_b, D_t = theta(try_b), theta(try_t)
    if DA < D_b or DA > D_t:
        return 0.0
    lasttry = -1.0
    while True:
        H0 = (try_b + try_t) / 2
        D_try = theta(H0)
        if D_try < DA:
            try_b = (try_b + try_t) / 2
        else:
            try_t = (try_b + try_t) / 2
        if abs(D_try - lasttry) < 1e-7:
            return H0
        lasttry = D_try


# ---------------------------------------------------------------- non-linear lensing rescale + sigma_8 (orc_nonlin.hpp)
def nonlinear(initpower, h, omm0, omegav, fnu, kh, z, transfer, w=-1.0, wa=0.0, k=None, tau=None, tautf=None, src=None):
    """Transfer_Get_SigmaR(8) + NonLinear_GetNonLinRatios (halofit, Takahashi) [+ MakeNonlinearSources on src in place].
    transfer [nz][nkt]; returns dict(sigma8 [nz], ratio [nz][nkt], spec [nz][3] = (rknl, rneff, rncur), err)."""
    L = lib()
    L.orc_nonlinear.restype = C.c_int
    kh, z, tr = _d(kh), _d(z), _d(transfer)
    nz, nkt = tr.shape
    par = _d([h, omm0, omegav, fnu, w, wa])
    s8, ratio, spec = np.zeros(nz), np.zeros((nz, nkt)), np.zeros((nz, 3))
    if src is not None:
        k, tau, tautf = _d(k), _d(tau), _d(tautf)
        assert src.flags["C_CONTIGUOUS"] and src.shape == (len(tau), 3, len(k))
        err = L.orc_nonlinear(_p(_d(initpower)), _p(par), nkt, nz, _p(kh), _p(z), _p(tr), _p(s8), _p(ratio), _p(spec),
                              len(k), _p(k), len(tau), _p(tau), _p(tautf), _p(src))
    else:
        err = L.orc_nonlinear(_p(_d(initpower)), _p(par), nkt, nz, _p(kh), _p(z), _p(tr), _p(s8), _p(ratio), _p(spec),
                              0, None, 0, None, None, None)
    if err < 0:
        raise RuntimeError("orc_nonlinear failed")
    return dict(sigma8=s8, ratio=ratio, spec=spec, err=err)


def matter_power_at(initpower, h, kh, transfer, kq):
    """MatterPowerData_k (log-log spline of P(k/h) built by Transfer_GetMatterPowerData) at k/h = kq, one redshift."""
    L = lib()
    L.orc_matter_power_at.restype = C.c_int
    L.orc_matter_power_at.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    kh, tr, kq = _d(kh), _d(transfer), _d(np.atleast_1d(kq))
    out = np.zeros_like(kq)
    if L.orc_matter_power_at(_p(_d(initpower)), float(h), len(kh), _p(kh), _p(tr), len(kq), _p(kq), _p(out)) != 0:
        raise RuntimeError("orc_matter_power_at failed")
    return out
