"""ctypes binding of libcosmob200.so (the C ABI declared in include/cosmob200.h).

This is the shim GetDist-side Python and the tests drive; it is the Python analogue of the Fortran
ISO_C_BINDING glue in cosmomc_b200/fortran/Calculator_B200.f90 (precedent in the reference: pycamb binds
camblib.so's `__handles_MOD_*` symbols through ctypes, camb/camb_python.f90 + camb/pycamb/camb/baseconfig.py).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but creating a
handle without a CUDA device raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB200_LIB", os.path.join(_HERE, "libcosmob200.so"))  # CB200_LIB: dev override for kernel-tuning builds
# the same sources plus the superseded projection kernels 1 and 2 (cross-checks of the parity tests; never the product)
TEST_LIB_PATH = os.path.join(_HERE, "libcosmob200_test.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

# every symbol include/cosmob200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "cb200_default_config", "cb200_create", "cb200_destroy", "cb200_last_error", "cb200_get_info",
    "cb200_get_lsamples", "cb200_set_templates", "cb200_make_q_grid", "cb200_make_time_steps",
    "cb200_make_source_k", "cb200_grid_build", "cb200_get_bessel_table", "cb200_upload_sources", "cb200_powers",
    "cb200_debug_fetch", "cb200_keep_transfers", "cb200_like_add_pliklite", "cb200_like_add_cmblikes",
    "cb200_loglike_batch", "cb200_loglike_cls", "cb200_get_timing", "cb200_sync", "cb200_set_option",
    "cb200_timer_start", "cb200_timer_stop", "cb200_measure_fp64_peaks", "cb200_config_size", "cb200_abi_version",
    "cb200_upload_sources_packed",
    "cb200_powers_shared", "cb200_like_set_bk_foregrounds", "cb200_background", "cb200_set_background", "cb200_like_add_bao", "cb200_like_add_hst", "cb200_like_add_sn",
    "cb200_eval_batch", "cb200_test_like_batch", "cb200_thermo", "cb200_theta_to_background", "cb200_nonlinear_lensing",
]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int), ("device", C.c_int), ("lmax_computed_cl", C.c_int), ("cmb_lensing", C.c_int),
                ("use_lensing_potential", C.c_int), ("use_nonlinear_lensing", C.c_int), ("compute_tensors", C.c_int),
                ("lmax_tensor", C.c_int), ("accurate_bb", C.c_int), ("k_eta_max_scalar", C.c_double),
                ("accuracy_level", C.c_double), ("lmax_out", C.c_int), ("highl_norm_first_call", C.c_int),
                ("max_points", C.c_int), ("chunk_points", C.c_int), ("n_tau_max", C.c_int), ("n_k_max", C.c_int),
                ("n_q_max", C.c_int), ("n_tau_max_tensor", C.c_int), ("n_k_max_tensor", C.c_int),
                ("n_q_max_tensor", C.c_int)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ["max_l", "max_eta_k", "max_l_tensor", "max_eta_k_tensor", "n_lsamp", "n_lsamp_tensor", "num_xx",
                 "lmax_lensed", "lens_lmax", "lens_npoints", "lens_jmax", "n_tau_max", "n_k_max", "n_q_max",
                 "max_points", "chunk_points", "num_xx_tensor"]]


class ParamLayout(C.Structure):
    """cb200_param_layout (include/cosmob200.h): bounds, priors and the CosmoMC columns that feed the initial power."""
    _fields_ = ([("num_params", C.c_int), ("pmin", C.POINTER(C.c_double)), ("pmax", C.POINTER(C.c_double)),
                 ("prior_mean", C.POINTER(C.c_double)), ("prior_std", C.POINTER(C.c_double)),
                 ("use_prior", C.POINTER(C.c_ubyte)), ("n_lincomb", C.c_int), ("lincomb", C.POINTER(C.c_double)),
                 ("lincomb_mean", C.POINTER(C.c_double)), ("lincomb_std", C.POINTER(C.c_double)),
                 ("temperature", C.c_double)] +
                [(n, C.c_int) for n in ["i_logA", "i_ns", "i_nrun", "i_nrunrun", "i_r", "i_nt", "i_ntrun", "i_Alens",
                                        "i_Aphiphi"]] +
                [(n, C.c_double) for n in ["def_logA", "def_ns", "def_nrun", "def_nrunrun", "def_r", "def_nt",
                                           "def_ntrun", "def_Alens", "def_Aphiphi", "pivot_scalar", "pivot_tensor"]] +
                [("inflation_consistency", C.c_int), ("i_nuis_first", C.c_int), ("n_nuis", C.c_int)])


class Timing(C.Structure):
    _fields_ = [("ms_spline", C.c_float), ("ms_project", C.c_float), ("ms_contract", C.c_float),
                ("ms_interp", C.c_float), ("ms_lens", C.c_float), ("ms_like", C.c_float), ("ms_total", C.c_float),
                ("n_launches", C.c_longlong), ("proj_triples", C.c_longlong), ("ring_slabs", C.c_longlong),
                ("ring_direct", C.c_longlong), ("ring_rows", C.c_longlong), ("ring_pairs", C.c_longlong),
                ("phase_cycles", C.c_longlong * 6), ("ms_background", C.c_float),
                ("proj_mask_mismatch", C.c_longlong), ("eval_points_powers", C.c_longlong),
                ("eval_points_reused", C.c_longlong)]


_lib = None


_libs = {}


def load(path=None):
    """Load the shared library; raises with a clear message if it has not been built."""
    global _lib
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError("cosmomc_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback." % path)
    L = C.CDLL(path)
    L.cb200_last_error.restype = C.c_char_p
    L.cb200_nonlinear_lensing.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp,
                                          C.c_int, c_dp, c_dp, c_dp, c_ip]
    L.cb200_thermo.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_ip]
    L.cb200_theta_to_background.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, C.c_double, c_dp]
    L.cb200_last_error.argtypes = [C.c_void_p]
    L.cb200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.cb200_destroy.argtypes = [C.c_void_p]
    L.cb200_destroy.restype = None
    L.cb200_default_config.argtypes = [C.POINTER(Config)]
    L.cb200_default_config.restype = None
    L.cb200_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.cb200_get_lsamples.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip]
    L.cb200_set_templates.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int]
    L.cb200_make_q_grid.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, c_dp, c_dp, c_ip]
    L.cb200_make_time_steps.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_int, c_dp, c_dp, c_ip]
    L.cb200_make_source_k.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, c_dp, c_ip]
    L.cb200_grid_build.argtypes = [C.c_int, c_dp, C.c_int, c_dp, c_dp, c_ip, C.c_int, c_dp, c_ip]
    L.cb200_get_bessel_table.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_dp]
    L.cb200_upload_sources.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_dp, C.c_void_p, C.c_int]
    L.cb200_upload_sources_packed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, C.c_void_p]
    L.cb200_powers.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip]
    L.cb200_powers_shared.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip]
    L.cb200_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_ip]
    L.cb200_keep_transfers.argtypes = [C.c_void_p, C.c_int]
    L.cb200_like_add_pliklite.argtypes = [C.c_void_p, c_ip, C.c_int, c_ip, c_ip, c_dp, C.c_int, c_dp, c_dp, C.c_int,
                                          c_ip]
    L.cb200_like_add_cmblikes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, C.c_int, C.c_int, c_dp, c_dp,
                                          c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_int, c_ip]
    L.cb200_loglike_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_ip]
    L.cb200_loglike_cls.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_ip]
    L.cb200_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing), C.c_int]
    L.cb200_sync.argtypes = [C.c_void_p]
    L.cb200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.cb200_timer_start.argtypes = [C.c_void_p]
    L.cb200_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cb200_measure_fp64_peaks.argtypes = [C.c_void_p, c_dp, c_dp]
    L.cb200_like_set_bk_foregrounds.argtypes = [C.c_void_p, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp,
                                                c_dp, c_dp, C.c_double, C.c_double, c_dp, c_dp, C.c_int, C.c_int,
                                                C.c_int, C.c_int, c_dp, C.c_int]
    L.cb200_background.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp]
    L.cb200_set_background.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
    L.cb200_like_add_bao.argtypes = [C.c_void_p, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp, C.c_double, C.c_double,
                                     c_dp, C.c_int, c_ip]
    L.cb200_like_add_hst.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, c_ip]
    L.cb200_like_add_sn.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, C.c_int, C.POINTER(c_dp), C.c_int, C.c_int,
                                    c_ip]
    L.cb200_eval_batch.argtypes = [C.c_void_p, C.POINTER(ParamLayout), C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip]
    L.cb200_test_like_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp]
    _libs[path] = L
    if path == LIB_PATH:
        _lib = L
    return L


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(c_ip)


class CB200Error(RuntimeError):
    pass


def grid_build(ops, query=None, max_n=200000):
    """Product-side grid builder (csrc/grids.hpp); no GPU needed."""
    L = load()
    o = _d(ops).reshape(-1, 5)
    x = np.zeros(max_n)
    dx = np.zeros(max_n)
    n = C.c_int(0)
    qv = _d(query if query is not None else [])
    idx = np.zeros(max(1, len(qv)), dtype=np.int32)
    rc = L.cb200_grid_build(len(o), _pd(o), max_n, _pd(x), _pd(dx), C.byref(n), len(qv), _pd(qv), _pi(idx))
    if rc != 0:
        raise CB200Error("cb200_grid_build failed (%d)" % rc)
    return x[:n.value].copy(), dx[:n.value].copy(), idx[:len(qv)].copy()


class Handle:
    """Owns one cb200_handle (device tables, resident sources, likelihood data)."""

    def __init__(self, lib_path=None, **kw):
        L = load(lib_path)
        self.L = L
        if L.cb200_config_size() != C.sizeof(Config):
            raise CB200Error("ctypes mirror of cb200_config (%d bytes) is out of step with the library (%d bytes)"
                             % (C.sizeof(Config), L.cb200_config_size()))
        cfg = Config()
        L.cb200_default_config(C.byref(cfg))
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown config field " + k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = L.cb200_create(C.byref(cfg), C.byref(self.h))
        if rc != 0 or not self.h:
            raise CB200Error("cb200_create failed (no CUDA device or bad config); there is no CPU fallback")
        self.info = Info()
        L.cb200_get_info(self.h, C.byref(self.info))
        self.n_like = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.cb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise CB200Error("%s failed (%d): %s" % (what, rc, self.L.cb200_last_error(self.h).decode()))

    # ---- tables / grids
    def lsamples(self, kind=0):
        l = np.zeros(128, dtype=np.int32)
        n = C.c_int(0)
        self._check(self.L.cb200_get_lsamples(self.h, kind, _pi(l), C.byref(n)), "get_lsamples")
        return l[:n.value].copy()

    def set_templates(self, highl_unlensed, highl_lensed):
        u = _d(highl_unlensed)
        assert u.shape == (4, 8001)
        le = _d(highl_lensed)
        self._check(self.L.cb200_set_templates(self.h, _pd(u), _pd(le), le.shape[1]), "set_templates")

    def q_grid(self, tau0, kind=0):
        q = np.zeros(self.info.n_q_max * 4)
        dq = np.zeros_like(q)
        n = C.c_int(0)
        self._check(self.L.cb200_make_q_grid(self.h, kind, tau0, len(q), _pd(q), _pd(dq), C.byref(n)), "make_q_grid")
        return q[:n.value].copy(), dq[:n.value].copy()

    def time_steps(self, tau0, taurst, taurend, reion_start, reion_complete, kind=0):
        t = np.zeros(8192)
        dt = np.zeros_like(t)
        n = C.c_int(0)
        self._check(self.L.cb200_make_time_steps(self.h, kind, tau0, taurst, taurend, reion_start, reion_complete,
                                                 len(t), _pd(t), _pd(dt), C.byref(n)), "make_time_steps")
        return t[:n.value].copy(), dt[:n.value].copy()

    def source_k(self, tau0, taurst, kind=0):
        k = np.zeros(8192)
        n = C.c_int(0)
        self._check(self.L.cb200_make_source_k(self.h, kind, tau0, taurst, len(k), _pd(k), C.byref(n)), "make_source_k")
        return k[:n.value].copy()

    def bessel_table(self, kind=0):
        nl = self.info.n_lsamp if kind == 0 else self.info.n_lsamp_tensor
        nx = self.info.num_xx if kind == 0 else self.info.num_xx_tensor
        x = np.zeros(nx)
        ajl = np.zeros((nl, nx))
        ajlpr = np.zeros_like(ajl)
        self._check(self.L.cb200_get_bessel_table(self.h, kind, _pd(x), _pd(ajl), _pd(ajlpr)), "get_bessel_table")
        return x, ajl, ajlpr

    # ---- calculator
    def upload_sources(self, thermo, n_k, k, src, first=0, kind=0, src_device_ptr=None, src_host_ptr=None):
        thermo = _d(thermo).reshape(-1, 5)
        npts = len(thermo)
        n_k = _i(n_k)
        k = _d(k)
        nkm = self.cfg.n_k_max_tensor if kind == 1 else self.info.n_k_max
        ntm = self.cfg.n_tau_max_tensor if kind == 1 else self.info.n_tau_max
        assert k.shape == (npts, nkm), (k.shape, nkm)
        if src_device_ptr is not None:
            ptr, isdev = C.c_void_p(src_device_ptr), 1
        elif src_host_ptr is not None:  # caller-owned (e.g. pinned) host buffer with the padded layout
            ptr, isdev = C.c_void_p(src_host_ptr), 0
        elif src is None:
            ptr, isdev = None, 0
        else:
            src = _d(src)
            assert src.shape == (npts, ntm, 3, nkm), src.shape
            self._keep = src
            ptr, isdev = C.c_void_p(src.ctypes.data), 0
        self._check(self.L.cb200_upload_sources(self.h, kind, first, npts, _pd(thermo), _pi(n_k), _pd(k), ptr, isdev),
                    "upload_sources")

    def upload_sources_packed(self, thermo, n_tau, n_k, k, src_packed=None, first=0, kind=0, src_host_ptr=None):
        """cb200_upload_sources_packed: point i's sources are [n_tau[i]][3][n_k[i]], back to back, no padding.
        `src_packed`: 1-D float64 array, or `src_host_ptr`: address of a caller-owned (pinned) buffer."""
        thermo = _d(thermo).reshape(-1, 5)
        npts = len(thermo)
        n_tau, n_k, k = _i(n_tau), _i(n_k), _d(k)
        if src_host_ptr is not None:
            ptr = C.c_void_p(src_host_ptr)
        else:
            src_packed = _d(src_packed).reshape(-1)
            assert src_packed.size == int((n_tau.astype(np.int64) * 3 * n_k).sum()), "packed size"
            self._keep = src_packed
            ptr = C.c_void_p(src_packed.ctypes.data)
        self._check(self.L.cb200_upload_sources_packed(self.h, kind, first, npts, _pd(thermo), _pi(n_tau), _pi(n_k),
                                                       _pd(k), ptr), "upload_sources_packed")

    def powers(self, initpower, alens=None, aphiphi=None, first=0, want_cls=True, want_derived=True):
        ip = _d(initpower).reshape(-1, 10)
        npts = len(ip)
        al = _d(alens) if alens is not None else None
        ap = _d(aphiphi) if aphiphi is not None else None
        lo = self.cfg.lmax_out + 1
        cls = np.zeros((npts, 5, lo)) if want_cls else None
        der = np.zeros((npts, 4)) if want_derived else None
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_powers(self.h, first, npts, _pd(ip), _pd(al), _pd(ap), _pd(cls), _pd(der), _pi(st)),
                    "powers")
        return cls, der, st

    def powers_shared(self, initpower, alens=None, aphiphi=None, src_point=0, first=0, want_cls=True):
        """Semi-slow step with one shared source point and many initial-power points (BK15-style chains)."""
        ip = _d(initpower).reshape(-1, 10)
        npts = len(ip)
        al = _d(alens) if alens is not None else None
        ap = _d(aphiphi) if aphiphi is not None else None
        cls = np.zeros((npts, 5, self.cfg.lmax_out + 1)) if want_cls else None
        der = np.zeros((npts, 4)) if want_cls else None
        st = np.zeros(npts, dtype=np.int32) if want_cls else None
        self._check(self.L.cb200_powers_shared(self.h, src_point, first, npts, _pd(ip), _pd(al), _pd(ap), _pd(cls),
                                               _pd(der), _pi(st)), "powers_shared")
        return cls, der, st

    def powers_resident(self, initpower, alens=None, aphiphi=None, first=0):
        """Run the semi-slow step leaving every output on the device (no D2H)."""
        ip = _d(initpower).reshape(-1, 10)
        al = _d(alens) if alens is not None else None
        ap = _d(aphiphi) if aphiphi is not None else None
        self._check(self.L.cb200_powers(self.h, first, len(ip), _pd(ip), _pd(al), _pd(ap), None, None, None), "powers")

    def powers_into(self, initpower, alens=None, aphiphi=None, first=0, cls_ptr=None, derived_ptr=None, status_ptr=None):
        """cb200_powers with caller-owned (e.g. pinned) host output buffers given as raw addresses."""
        ip = _d(initpower).reshape(-1, 10)
        al = _d(alens) if alens is not None else None
        ap = _d(aphiphi) if aphiphi is not None else None
        self._check(self.L.cb200_powers(self.h, first, len(ip), _pd(ip), _pd(al), _pd(ap),
                                        C.cast(C.c_void_p(cls_ptr), c_dp) if cls_ptr else None,
                                        C.cast(C.c_void_p(derived_ptr), c_dp) if derived_ptr else None,
                                        C.cast(C.c_void_p(status_ptr), c_ip) if status_ptr else None), "powers")

    def keep_transfers(self, on=True):
        self.L.cb200_keep_transfers(self.h, int(on))

    def debug_fetch(self, what, point=0, max_n=None):
        if max_n is None:
            max_n = 3 * 96 * self.info.n_q_max + 6 * (self.info.max_l + 8)
        out = np.zeros(max_n)
        n = C.c_int(0)
        self._check(self.L.cb200_debug_fetch(self.h, what, point, max_n, _pd(out), C.byref(n)), "debug_fetch")
        return out[:n.value].copy()

    # ---- likelihoods
    def add_pliklite(self, nb, blmin, blmax, weights, invcov, x_data, cal_index=0):
        nb = _i(nb)
        blmin = _i(blmin)
        blmax = _i(blmax)
        w = _d(weights)
        ic = _d(invcov)
        xd = _d(x_data)
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_pliklite(self.h, _pi(nb), len(blmin), _pi(blmin), _pi(blmax), _pd(w),
                                                   len(w) - 1, _pd(ic), _pd(xd), cal_index, C.byref(lid)),
                    "like_add_pliklite")
        self.n_like += 1
        return lid.value

    def add_cmblikes(self, nmaps, nbins, cl_use_index, like_approx, W, offset, chat, invcov, noise=None,
                     sqrt_fid=None, log_cal_prior=-1.0, cal_index=-1):
        cui = _i(cl_use_index)
        W = _d(W)
        ncl = nmaps * (nmaps + 1) // 2
        assert W.shape[0] == nbins and W.shape[1] == ncl and W.shape[2] == 5
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_cmblikes(
            self.h, nmaps, nbins, len(cui), _pi(cui), like_approx, W.shape[3] - 1, _pd(W), _pd(_d(offset)),
            _pd(_d(noise)) if noise is not None else None, _pd(_d(chat)),
            _pd(_d(sqrt_fid)) if sqrt_fid is not None else None, _pd(_d(invcov)), log_cal_prior, cal_index,
            C.byref(lid)), "like_add_cmblikes")
        self.n_like += 1
        return lid.value

    def set_bk_foregrounds(self, like_id, map_field, bc_class, bandpasses, th_dust, th_sync, nu_bar, fpivot_dust,
                           fpivot_sync, fpivot_dust_decorr, fpivot_sync_decorr, lform_dust, lform_sync, lmin, lmax, fgW,
                           nuis_offset):
        """bandpasses: list of (nu, R, dnu) arrays per used map."""
        off = np.zeros(len(bandpasses) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(b[0]) for b in bandpasses])
        nu = _d(np.concatenate([b[0] for b in bandpasses]))
        R = _d(np.concatenate([b[1] for b in bandpasses]))
        dnu = _d(np.concatenate([b[2] for b in bandpasses]))
        fgW = _d(fgW)
        self._check(self.L.cb200_like_set_bk_foregrounds(
            self.h, like_id, len(bandpasses), _pi(_i(map_field)), _pi(_i(bc_class)), _pi(off), _pd(nu), _pd(R), _pd(dnu),
            _pd(_d(th_dust)), _pd(_d(th_sync)), _pd(_d(nu_bar)), fpivot_dust, fpivot_sync, _pd(_d(fpivot_dust_decorr)),
            _pd(_d(fpivot_sync_decorr)), lform_dust, lform_sync, lmin, lmax, _pd(fgW), nuis_offset),
            "like_set_bk_foregrounds")

    # ---- background functions and likelihoods
    def background(self, bg, z, want_scalars=False):
        """D_A(z) [Mpc], H(z) [Mpc^-1] for every bg row; optionally (tau0, age/Gyr, CosmomcTheta)."""
        bg = _d(bg).reshape(-1, 16)
        z = _d(np.atleast_1d(z))
        npts = len(bg)
        DA = np.zeros((npts, len(z)))
        H = np.zeros((npts, len(z)))
        sc = np.zeros((npts, 3)) if want_scalars else None
        self._check(self.L.cb200_background(self.h, npts, _pd(bg), len(z), _pd(z), _pd(DA), _pd(H), _pd(sc)),
                    "background")
        return (DA, H, sc) if want_scalars else (DA, H)

    def nonlinear_lensing(self, initpower, cosmo, kh, z, transfer, tautf=None, rescale_sources=False, first=0):
        """sigma_8(z), halofit ratios sqrt(P_NL / P_L)(z, k), (k_NL, n_eff, curvature)(z) of every point, and - with
        rescale_sources - MakeNonlinearSources on the resident lensing sources of points [first, first + npts).
        cosmo [npts][6] = h, omm0, omegav, fnu, w, wa; kh [npts][n_kt]; z [n_z]; transfer [npts][n_z][n_kt]; tautf [npts][n_z]."""
        ip, cs, kh, z, tr = _d(initpower).reshape(-1, 10), _d(cosmo).reshape(-1, 6), _d(kh), _d(z), _d(transfer)
        npts, n_z, n_kt = tr.shape
        kh = np.ascontiguousarray(np.broadcast_to(kh.reshape(-1, n_kt), (npts, n_kt)))
        tf = _d(tautf).reshape(npts, n_z) if tautf is not None else None
        s8, ratio, spec = np.zeros((npts, n_z)), np.zeros((npts, n_z, n_kt)), np.zeros((npts, n_z, 3))
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_nonlinear_lensing(self.h, first, npts, _pd(ip), _pd(cs), n_kt, n_z, _pd(kh), _pd(z), _pd(tr),
                                                   _pd(tf), int(bool(rescale_sources)), _pd(s8), _pd(ratio), _pd(spec), _pi(st)),
                    "nonlinear_lensing")
        return dict(sigma8=s8, ratio=ratio, spec=spec, status=st)

    THERMO_DERIVED = ["age", "zstar", "rstar", "thetastar", "DAstar", "zdrag", "rdrag", "kd", "thetad", "zeq", "keq",
                      "thetaeq", "thetarseq"]

    def thermo(self, bg, yhe, zre=0.0, optical_depth=0.0, max_eta_k=14000.0, want_tensors=False, transfer_kmax_h=5.0,
               accuracy_boost=1.0):
        """Thermal history of every bg row (RECFAST, reionisation, inithermo): returns (out [npts][32], status [npts]);
        out[:, :12] = tau0, taurst, taurend, tau_start, tau_complete, dtaurec, tau_maxvis, zre, z_star, z_drag,
        actual_opt_depth, status; out[:, 12:25] = ThermoDerivedParams in the order of THERMO_DERIVED."""
        bg = _d(bg).reshape(-1, 16)
        npts = len(bg)
        tin = np.zeros((npts, 8))
        tin[:, 0], tin[:, 1], tin[:, 2] = yhe, zre, optical_depth
        tin[:, 3], tin[:, 4], tin[:, 5], tin[:, 6] = max_eta_k, 1.0 if want_tensors else 0.0, transfer_kmax_h, accuracy_boost
        out = np.zeros((npts, 32))
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_thermo(self.h, npts, _pd(bg), _pd(tin), _pd(out), _pi(st)), "thermo")
        return out, st

    def theta_to_background(self, ombh2, omch2, theta100, omnuh2, nu_split, omk=0.0, w=-1.0, H0_min=20.0, H0_max=100.0,
                            tcmb=2.7255, rdrag=None):
        """theta_MC -> H0 by the reference's bisection; returns bg [npts][16] (H0 = 0 rows: theta out of range).
        nu_split [npts][8] or [8] = bg[7:15] (massless degeneracy, eigenstates, degeneracies, mass fractions)."""
        ombh2 = _d(np.atleast_1d(ombh2))
        npts = len(ombh2)
        cs = np.zeros((npts, 8))
        cs[:, 0], cs[:, 1], cs[:, 2], cs[:, 3], cs[:, 4], cs[:, 5] = ombh2, omch2, omnuh2, omk, w, theta100
        cs[:, 6], cs[:, 7] = H0_min, H0_max
        nu = np.ascontiguousarray(np.broadcast_to(_d(nu_split).reshape(-1, 8), (npts, 8)))
        bg = np.zeros((npts, 16))
        if rdrag is not None:
            bg[:, 15] = rdrag
        self._check(self.L.cb200_theta_to_background(self.h, npts, _pd(cs), _pd(nu), float(tcmb), _pd(bg)), "theta_to_background")
        return bg

    def set_background(self, bg, first=0):
        bg = _d(bg).reshape(-1, 16)
        self._check(self.L.cb200_set_background(self.h, first, len(bg), _pd(bg)), "set_background")

    def add_bao(self, types, z, obs, invcov, rs_rescale=1.0, fixed_rs=-1.0):
        t, z, obs, ic = _i(types), _d(z), _d(obs), _d(invcov)
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_bao(self.h, 0, len(z), _pi(t), _pd(z), _pd(obs), _pd(ic), rs_rescale,
                                              fixed_rs, None, 0, C.byref(lid)), "like_add_bao")
        self.n_like += 1
        return lid.value

    def add_mgs(self, z, alpha_prob, fixed_rs=-1.0):
        z, ap = _d([z]), _d(alpha_prob)
        t = _i([2])
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_bao(self.h, 1, 1, _pi(t), _pd(z), None, None, 1.0, fixed_rs, _pd(ap), len(ap),
                                              C.byref(lid)), "like_add_bao(MGS)")
        self.n_like += 1
        return lid.value

    def add_hst(self, H0, H0_err, zeff=0.0, angconversion=0.0):
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_hst(self.h, H0, H0_err, zeff, angconversion, C.byref(lid)), "like_add_hst")
        self.n_like += 1
        return lid.value

    def add_sn(self, cols, covs, A1=None, A2=None, twoscriptm=False, alpha_index=-1, beta_index=-1):
        """cols [11][nsn] (include/cosmob200.h), covs: list of 6 arrays or None."""
        cols = _d(cols)
        nsn = cols.shape[1]
        keep = [None if c is None else _d(c) for c in covs]
        arr = (c_dp * 6)(*[None if c is None else c.ctypes.data_as(c_dp) for c in keep])
        a1 = _d(A1) if A1 is not None else None
        a2 = _d(A2) if A2 is not None else None
        lid = C.c_int(-1)
        self._check(self.L.cb200_like_add_sn(self.h, nsn, _pd(cols), _pd(a1), _pd(a2), int(twoscriptm), arr, alpha_index,
                                             beta_index, C.byref(lid)), "like_add_sn")
        self.n_like += 1
        return lid.value

    def loglike_batch(self, npts, nuisance, first=0):
        nu = _d(nuisance).reshape(npts, -1)
        ll = np.zeros((npts, self.n_like))
        tot = np.zeros(npts)
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_loglike_batch(self.h, first, npts, _pd(nu), nu.shape[1], _pd(ll), _pd(tot), _pi(st)),
                    "loglike_batch")
        return ll, tot, st

    def loglike_cls(self, cls, nuisance):
        cls = _d(cls)
        npts = cls.shape[0]
        assert cls.shape[1:] == (5, self.cfg.lmax_out + 1)
        nu = _d(nuisance).reshape(npts, -1)
        ll = np.zeros((npts, self.n_like))
        tot = np.zeros(npts)
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_loglike_cls(self.h, npts, _pd(cls), _pd(nu), nu.shape[1], _pd(ll), _pd(tot), _pi(st)),
                    "loglike_cls")
        return ll, tot, st

    def timing(self, reset=True):
        t = Timing()
        self._check(self.L.cb200_get_timing(self.h, C.byref(t), int(reset)), "get_timing")
        d = {n: getattr(t, n) for n, _ in Timing._fields_}
        d["phase_cycles"] = list(t.phase_cycles)
        return d

    def eval_batch(self, params, pmin, pmax, columns, prior_mean=None, prior_std=None, use_prior=None, lincomb=None,
                   lincomb_mean=None, lincomb_std=None, temperature=1.0, defaults=None, pivot_scalar=0.05,
                   pivot_tensor=0.05, inflation_consistency=True, nuis_first=0, n_nuis=0, first=0):
        """TLikeCalculator%GetLogLike for a batch (source/calclike.f90:97-151): params [npts][num_params];
        columns: dict name -> column index for logA, ns, nrun, nrunrun, r, nt, ntrun, Alens, Aphiphi (missing: default).
        Returns (loglike[npts], likelihoods[npts][n_like], prior[npts], status[npts])."""
        P = _d(params)
        npts, npar = P.shape
        keep = []

        def ptr(a):
            if a is None:
                return None
            a = _d(a)
            keep.append(a)
            return _pd(a)
        L = ParamLayout()
        L.num_params = npar
        L.pmin, L.pmax = ptr(pmin), ptr(pmax)
        L.prior_mean, L.prior_std = ptr(prior_mean), ptr(prior_std)
        if use_prior is not None:
            up = np.ascontiguousarray(use_prior, dtype=np.uint8)
            keep.append(up)
            L.use_prior = up.ctypes.data_as(C.POINTER(C.c_ubyte))
        if lincomb is not None:
            lc = _d(lincomb).reshape(-1, npar)
            L.n_lincomb = len(lc)
            L.lincomb, L.lincomb_mean, L.lincomb_std = ptr(lc), ptr(lincomb_mean), ptr(lincomb_std)
        L.temperature = temperature
        dflt = dict(logA=3.044, ns=0.965, nrun=0.0, nrunrun=0.0, r=0.0, nt=0.0, ntrun=0.0, Alens=1.0, Aphiphi=1.0)
        dflt.update(defaults or {})
        for name in dflt:
            setattr(L, "i_" + name, int(columns.get(name, -1)))
            setattr(L, "def_" + name, float(dflt[name]))
        L.pivot_scalar, L.pivot_tensor = pivot_scalar, pivot_tensor
        L.inflation_consistency = int(inflation_consistency)
        L.i_nuis_first, L.n_nuis = nuis_first, n_nuis
        ll = np.zeros(npts)
        likes = np.zeros((npts, max(self.n_like, 1)))
        pr = np.zeros(npts)
        st = np.zeros(npts, dtype=np.int32)
        self._check(self.L.cb200_eval_batch(self.h, C.byref(L), first, npts, _pd(P), _pd(ll), _pd(likes), _pd(pr),
                                            _pi(st)), "eval_batch")
        return ll, likes[:, :self.n_like], pr, st

    def test_like_batch(self, x, center, covinv):
        """test_likelihood = T analogue (source/calclike.f90:180-199): (x-c)^T covinv (x-c) / 2 per row of x."""
        x = _d(x)
        npts, n = x.shape
        c = _d(center)
        ci = _d(covinv)
        out = np.zeros(npts)
        self._check(self.L.cb200_test_like_batch(self.h, npts, n, _pd(x), _pd(c), _pd(ci), _pd(out)), "test_like_batch")
        return out

    def sync(self):
        self._check(self.L.cb200_sync(self.h), "sync")

    def set_option(self, name, value):
        self._check(self.L.cb200_set_option(self.h, name.encode(), float(value)), "set_option")

    def timer_start(self):
        self._check(self.L.cb200_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.L.cb200_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def measure_fp64_peaks(self):
        a = C.c_double(0)
        b = C.c_double(0)
        self._check(self.L.cb200_measure_fp64_peaks(self.h, C.byref(a), C.byref(b)), "measure_fp64_peaks")
        return a.value, b.value
