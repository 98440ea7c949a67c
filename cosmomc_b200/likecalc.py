"""LikeCalculator: a top-level CosmoMC `.ini` + `.paramnames` -> one batched -lnL function on the B200 library.

What the reference does at start-up for this path, restated on the host (nothing numerical happens here):
  * source/DataLikelihoods.f90:9-41   SetDataLikelihoods: the likelihood list, in this order -
        CMB data sets  `cmb_dataset[tag] = file`      (source/CMB.f90:60-108; tag BKPLANCK -> TBK_planck, else TCMBLikes)
        Hubble         `use_HST = T` + Hubble_* keys  (source/HST.f90:24-45)
        supernovae     `use_SN`, `use_JLA`, `jla_dataset` (source/supernovae.f90:16, supernovae_JLA.f90:228-260)
        BAO            `use_BAO`, `bao_dataset[tag] = file` (source/bao.f90:71-111)
  * source/CosmologyConfig.f90:40-52 / CosmologyParameterizations.f90:34-70: the base parameter names come from
    `paramnames/params_CMB.paramnames`; every likelihood appends its nuisance parameters (source/GeneralTypes.f90:620-700)
  * source/BaseParameters.f90:90-200: `param[name] = centre [min max start_width propose_width]`, `prior[name] = mean std`,
    `linear_combination[tag]` + `linear_combination_weights[tag]` + `prior[tag]`
  * source/calclike.f90:97-151: GetLogLike = bounds -> logZero, sum of -lnL / Temperature, priors  (cb200_eval_batch)

Placeholders `%DATASETDIR%` and `%LOCALDIR%` are resolved as the reference does (source/settings.f90 DataDir / LocalDir).
The Boltzmann source functions stay with CAMB (north_star): a caller uploads them with `upload_sources` /
`upload_sources_packed`; background-only likelihoods (BAO, H0, supernovae) take their distances from the background vectors
given to `set_background`.
"""
import os

import numpy as np

from . import datasets as ds

# columns of CosmoMC's InitPower block that feed CAMBCalc_SetCAMBInitPower (source/Calculator_CAMB.f90:839-877)
POWER_NAMES = ["logA", "ns", "nrun", "nrunrun", "r", "nt", "ntrun", "Alens", "Aphiphi"]
LOGZERO = 1e30


class _Ini(ds.IniFile):
    """IniFile with the reference's %DATASETDIR% / %LOCALDIR% placeholders."""

    def __init__(self, path, subst):
        self.subst = subst
        super().__init__(path)

    def string(self, key, default=None):
        v = super().string(key, default)
        for k, r in self.subst.items():
            v = v.replace(k, r)
        return v

    def tagged(self, name):
        """Ini%TagValuesForName: [(tag, value)] of the keys `name[tag]`, in file order."""
        out = []
        for k in self.order:
            if k.startswith(name + "[") and k.endswith("]") and self.has(k):
                out.append((k[len(name) + 1:-1], self.string(k)))
        return out

    def path(self, value):
        return value if os.path.isabs(value) else os.path.join(self.dir, value)


class LikeCalculator:
    def __init__(self, ini_path, data_dir=None, local_dir=None, paramnames=None, handle_kw=None, cov_providers=None,
                 sn_covs=None, create_handle=True):
        local_dir = local_dir or os.path.dirname(os.path.abspath(ini_path))
        data_dir = data_dir or os.path.join(local_dir, "data")
        self.subst = {"%DATASETDIR%": data_dir.rstrip(os.sep) + os.sep, "%LOCALDIR%": local_dir.rstrip(os.sep) + os.sep}
        ini = self.ini = _Ini(ini_path, self.subst)
        cov_providers = cov_providers or {}
        # ---- likelihood list, reference order (DataLikelihoods.f90:22-38)
        self.likes = []   # (kind, tag, plan, nuisance names)
        for tag, fn in ini.tagged("cmb_dataset"):
            path = ini.path(fn)
            if tag == "BKPLANCK":
                plan = ds.BK15Plan(path, cov_provider=cov_providers.get(tag))
            elif tag in ("WMAP", "SMICA", "PLIK_LITE", "SPTPOL_TEEE", "SPTPOL_BB"):
                raise NotImplementedError("cmb_dataset[%s]: this likelihood class is outside the B200 path" % tag)
            else:
                plan = ds.CMBLikesPlan(path, cov_provider=cov_providers.get(tag))
            self.likes.append(("cmb", tag, plan, list(plan.nuisance_names)))
        if ini.bool("use_HST", False):
            plan = ds.HSTPlan.__new__(ds.HSTPlan)
            plan.name = ini.string("Hubble_name")
            plan.H0, plan.H0_err = ini.float("Hubble_H0"), ini.float("Hubble_H0_err")
            plan.zeff = ini.float("Hubble_zeff", 0.0)
            plan.angconversion = ini.float("Hubble_angconversion", 0.0) if plan.zeff > 0 else 0.0
            self.likes.append(("hst", plan.name, plan, []))
        # supernovae: SNLikelihood_Add (source/supernovae.f90:16: use_SN) -> JLALikelihood_Add (supernovae_JLA.f90:228-260):
        # use_JLA, jla_dataset (default <data>/jla.dataset; batch3/Pantheon.ini points it at the Pantheon set), nuisance
        # names from <data>/JLA.paramnames unless JLA_marginalize
        if ini.bool("use_SN", False) and ini.bool("use_JLA", False):
            if ini.bool("JLA_marginalize", False):
                raise NotImplementedError("JLA_marginalize: grid marginalisation is not on the B200 path")
            path = ini.path(ini.string("jla_dataset", self.subst["%DATASETDIR%"] + "jla.dataset"))
            version = ini.string("jla_version", "JLA")
            plan = ds.SNPlan(path, covs=(sn_covs or {}).get(version))
            pn_file = os.path.join(self.subst["%DATASETDIR%"], "JLA.paramnames")
            names = [n for n, _, _ in ds.read_paramnames(pn_file)] if os.path.exists(pn_file) else ["alpha_JLA", "beta_JLA"]
            self.likes.append(("sn", version, plan, names))
        if ini.bool("use_BAO", False):
            tags = ini.tagged("bao_dataset")
            if not tags:
                raise ValueError("Use_BAO but no bao_dataset[NAMETAG] defined")
            for tag, fn in tags:
                if tag in ("DR11CMASS", "DR12CMASS", "DR12LOWZ"):
                    raise NotImplementedError("bao_dataset[%s]: the DR1x probability-grid class is outside the B200 path" % tag)
                self.likes.append(("bao", tag, ds.BAOPlan(ini.path(fn), tag=tag), []))
        self.fixed_rs = ini.float("BAO_fixed_rs", -1.0)   # late_time_only runs (source/bao.f90:85-87)

        # ---- parameter names: base block, then each likelihood's nuisance block (no duplicates, first wins)
        pn = paramnames or (ini.path(ini.string("paramnames")) if ini.has("paramnames")
                            else os.path.join(local_dir, "paramnames", "params_CMB.paramnames"))
        base = [(n, lab) for n, lab, derived in ds.read_paramnames(pn) if not derived]
        self.names = [n for n, _ in base]
        self.labels = [lab for _, lab in base]
        self.n_base = len(self.names)
        self.like_nuis = []   # per likelihood: indices of its nuisance parameters in the full vector
        for kind, tag, plan, nuis in self.likes:
            idx = []
            for n in nuis:
                if n not in self.names:
                    self.names.append(n)
                    self.labels.append(n)
                idx.append(self.names.index(n))
            self.like_nuis.append(idx)
        n = self.num_params = len(self.names)

        # ---- param[name], prior[name], linear combinations (BaseParameters.f90:90-200)
        self.center = np.zeros(n); self.pmin = np.zeros(n); self.pmax = np.zeros(n)
        self.start_width = np.zeros(n); self.propose_width = np.zeros(n)
        for i, name in enumerate(self.names):
            key = "param[%s]" % name
            if not ini.has(key):
                raise KeyError("parameter ranges not found: " + key)
            v = [float(x) for x in ini.string(key).split()]
            if len(v) == 1:   # fixed
                self.center[i] = self.pmin[i] = self.pmax[i] = v[0]
            elif len(v) == 5:
                self.center[i], self.pmin[i], self.pmax[i], self.start_width[i], self.propose_width[i] = v
                if self.pmax[i] < self.pmin[i]:
                    raise ValueError("You have param Max < Min: " + name)
            else:
                raise ValueError("Must have min max start_width propose_width for " + key)
        self.varying = self.pmax > self.pmin
        self.prior_mean = np.zeros(n); self.prior_std = np.zeros(n)
        for i, name in enumerate(self.names):
            key = "prior[%s]" % name
            if ini.has(key):
                self.prior_mean[i], self.prior_std[i] = [float(x) for x in ini.string(key).split()[:2]]
        self.include_fixed_priors = ini.bool("include_fixed_parameter_priors", False)
        self.use_prior = (self.varying | self.include_fixed_priors).astype(np.uint8)
        self.lincomb, self.lincomb_mean, self.lincomb_std = [], [], []
        for tag, plist in ini.tagged("linear_combination"):
            w = [float(x) for x in ini.string("linear_combination_weights[%s]" % tag).split()]
            row = np.zeros(n)
            for pname, wi in zip(plist.split(), w):
                row[self.names.index(pname)] = wi
            self.lincomb.append(row)
            pr = ini.string("prior[%s]" % tag, "0 0").split()
            self.lincomb_mean.append(float(pr[0])); self.lincomb_std.append(float(pr[1]))
        self.temperature = ini.float("temperature", 1.0)
        self.pivot_k = ini.float("pivot_k", 0.05)
        self.tensor_pivot_k = ini.float("tensor_pivot_k", self.pivot_k)
        self.inflation_consistency = ini.bool("inflation_consistency", True)
        self.compute_tensors = ini.bool("compute_tensors", False)
        self.columns = {p: self.names.index(p) for p in POWER_NAMES if p in self.names}

        # ---- nuisance block handed to the likelihoods: the registered order must match the columns
        self.nuis_first = self.n_base
        self.n_nuis = n - self.n_base
        self.handle = None
        if create_handle:
            self._create(handle_kw or {})

    # -------------------------------------------------------------------------------------------------------------
    def _create(self, kw):
        from . import lib
        ini = self.ini
        cfg = dict(lmax_computed_cl=ini.int("lmax_computed_cl", 2500) if ini.has("lmax_computed_cl")
                   else max(ini.int("lmin_store_all_cmb", 2500), 2500),
                   compute_tensors=int(self.compute_tensors),
                   use_nonlinear_lensing=int(ini.bool("use_nonlinear_lensing", True)))
        cfg.update(kw)
        h = self.handle = lib.Handle(**cfg)
        for (kind, tag, plan, nuis), idx in zip(self.likes, self.like_nuis):
            rel = [i - self.nuis_first for i in idx]   # position inside the nuisance block
            if kind == "cmb" and isinstance(plan, ds.BK15Plan):
                plan.register(h, nuis_offset=rel[0] if rel else 0)
            elif kind == "cmb":
                cal = self.names.index(plan.calibration_param) - self.nuis_first if plan.calibration_param else -1
                plan.register(h, cal_index=cal)
            elif kind == "hst":
                plan.register(h)
            elif kind == "sn":
                plan.register(h, alpha_index=rel[0] if len(rel) > 0 else -1, beta_index=rel[1] if len(rel) > 1 else -1)
            elif kind == "bao":
                plan.register(h, fixed_rs=self.fixed_rs)
        return h

    # ---- thin pass-throughs: what the CAMB side hands over per point
    def set_templates(self, highl_unlensed, highl_lensed):
        self.handle.set_templates(highl_unlensed, highl_lensed)

    def upload_sources(self, *a, **k):
        self.handle.upload_sources(*a, **k)

    def upload_sources_packed(self, *a, **k):
        self.handle.upload_sources_packed(*a, **k)

    def set_background(self, bg, first=0):
        self.handle.set_background(bg, first=first)

    def set_background_from_params(self, P, yhe=0.2453985, H0_min=20.0, H0_max=100.0, first=0, **nu_kw):
        """What ThetaParameterization%ParamArrayToTheoryParams + CAMB's thermal history do for every row of P before a
        background likelihood is called (source/CosmologyParameterizations.f90:114-187, camb/modules.f90:2682-2992), on
        the device: theta_MC -> H0 (cb200_theta_to_background), then RECFAST / inithermo (cb200_thermo) for r_drag, then
        the rows are made resident (cb200_set_background).  Columns used: omegabh2, omegach2, theta, tau[, omegak, mnu,
        nnu, w].  Y_He is an input (the BBN-consistency table of the golden run, PArthENoPE 880.2, does not ship).
        Returns (bg [B][16], thermo_out [B][32], ok [B]); rows with ok = False (theta out of range, thermal-history
        error) must be rejected by the caller as the reference rejects them (H0 = 0 / global_error_flag)."""
        from . import params as prm
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        col = lambda n, d=None: P[:, self.names.index(n)] if n in self.names else np.full(len(P), d)
        mnu, nnu = col("mnu", 0.06), col("nnu", 3.046)
        if np.ptp(mnu) or np.ptp(nnu):
            raise NotImplementedError("the neutrino split is built once per batch: mnu and nnu must be common to the rows")
        ref = prm.cmb_to_background(0.022, 0.12, 70.0, mnu=float(mnu[0]), nnu=float(nnu[0]), **nu_kw)
        omnuh2 = ref[3] * 0.7 ** 2
        bg = self.handle.theta_to_background(col("omegabh2"), col("omegach2"), col("theta"), omnuh2, ref[7:15],
                                             omk=col("omegak", 0.0), w=col("w", -1.0), H0_min=H0_min, H0_max=H0_max)
        ok = bg[:, 0] > 0
        safe = np.where(ok[:, None], bg, ref[None, :])
        th, st = self.handle.thermo(safe, yhe, optical_depth=col("tau", 0.0))
        ok &= st == 0
        bg[:, 15] = np.where(ok, th[:, 18], 0.0)
        safe[:, 15] = np.where(ok, th[:, 18], 147.0)
        self.handle.set_background(safe, first=first)
        return bg, th, ok

    # ---- the batched GetLogLike
    def full_params(self, varied):
        """[B][n_varying] -> [B][num_params] with the fixed parameters at their `param[...]` values."""
        varied = np.atleast_2d(np.asarray(varied, dtype=np.float64))
        P = np.tile(self.center, (len(varied), 1))
        P[:, self.varying] = varied
        return P

    def loglike(self, P, first=0, full_output=False):
        """-lnL (incl. priors, / Temperature; 1e30 = out of bounds or rejected) of every row of P [B][num_params]."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        if P.shape[1] != self.num_params:
            raise ValueError("P must have %d columns (%s ...)" % (self.num_params, ", ".join(self.names[:4])))
        defaults = {p: self.center[self.names.index(p)] for p in POWER_NAMES if p in self.names}
        out = self.handle.eval_batch(
            P, self.pmin, self.pmax, self.columns, prior_mean=self.prior_mean, prior_std=self.prior_std,
            use_prior=self.use_prior, lincomb=np.array(self.lincomb) if self.lincomb else None,
            lincomb_mean=self.lincomb_mean or None, lincomb_std=self.lincomb_std or None, temperature=self.temperature,
            defaults=defaults, pivot_scalar=self.pivot_k, pivot_tensor=self.tensor_pivot_k,
            inflation_consistency=self.inflation_consistency, nuis_first=self.nuis_first, n_nuis=self.n_nuis, first=first)
        return out if full_output else out[0]

    def like_names(self):
        return [tag for _, tag, _, _ in self.likes]
