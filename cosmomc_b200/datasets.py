"""Host-side data preparation: CosmoMC .ini / .dataset / .paramnames readers and the resolution of a binned
`CMBlikes` data set into the dense arrays the C ABI takes (cb200_like_add_cmblikes).

Mirrors the reference's file handling for this path so that the same files configure the B200 likelihoods:
  * ini dialect: source/IniObjects.f90:466-473 (DEFAULT(file) lower priority, INCLUDE(file)), key = value, '#' comments,
    relative file names resolved against the ini file's directory (ReadRelativeFileName)
  * .paramnames: source/ObjectParamNames.f90:82-86  ("name  latex  #comment", trailing '*' = derived)
  * CMBlikes data set: source/CMBlikes.f90:371-859 (CMBLikes_ReadIni, ReadClArr, ReadBinWindows, ReadCovmat)
The numerical evaluation is NOT here (it runs on the GPU); this module only builds arrays at set-up time.
"""
import os
import re

import numpy as np

FIELDS = ["T", "E", "B", "P"]
# theory spectrum slot of the device Cls array [TT, TE, EE, BB, PP] for a pair of theory fields (sorted)
SPEC_SLOT = {(0, 0): 0, (0, 1): 1, (1, 1): 2, (2, 2): 3, (3, 3): 4}


class IniFile:
    def __init__(self, path=None):
        self.params = {}
        self.order = []
        self.read_values = {}   # key -> value as handed out, in first-read order (Ini%ReadValues of the reference)
        self.dir = "."
        if path is not None:
            self.dir = os.path.dirname(os.path.abspath(path))
            self._read(path, override=True)

    def _read(self, path, override):
        defaults = []
        with open(path) as f:
            for raw in f:
                line = raw.strip()
                if not line or line.startswith("#") or line.startswith(";"):
                    continue
                m = re.match(r"^(DEFAULT|INCLUDE)\((.*)\)\s*$", line)
                if m:
                    fn = m.group(2).strip()
                    if not os.path.isabs(fn):
                        fn = os.path.join(os.path.dirname(os.path.abspath(path)), fn)
                    if m.group(1) == "INCLUDE":
                        self._read(fn, override=True)
                    else:
                        defaults.append(fn)
                    continue
                if "=" not in line:
                    continue
                k, v = line.split("=", 1)
                k = k.strip()
                v = v.split("#")[0].strip() if not v.strip().startswith("#") else ""
                if override or k not in self.params:
                    if k not in self.params:
                        self.order.append(k)
                    self.params[k] = v
        for fn in defaults:  # DEFAULT files never override what is already set
            self._read(fn, override=False)

    def has(self, key):
        return key in self.params and self.params[key] != ""

    def string(self, key, default=None):
        if self.has(key):
            self.read_values.setdefault(key, self.params[key])
            return self.params[key]
        if default is None:
            raise KeyError("ini key not found: " + key)
        self.read_values.setdefault(key, str(default))
        return default

    def save_read_values(self, path):
        """`<root>.inputparams` (Ini%SaveReadValues, source/IniObjects.f90:870-884; called from source/driver.F90:198):
        one `name = value` line per key the run actually read, defaults included, in the order of the first read."""
        with open(path, "w") as f:
            for k, v in self.read_values.items():
                f.write("%s = %s\n" % (k, v))

    def int(self, key, default=None):
        return int(self.string(key, None if default is None else str(default)))

    def float(self, key, default=None):
        return float(self.string(key, None if default is None else str(default)))

    def bool(self, key, default=False):
        v = self.string(key, "T" if default else "F").strip().upper()
        return v in ("T", "TRUE", ".TRUE.", "1", "YES")

    def split(self, key, default=None):
        if self.has(key):
            return self.params[key].split()
        if default is None:
            raise KeyError("ini key not found: " + key)
        return list(default)

    def relative_file(self, key):
        fn = self.string(key)
        fn = fn.replace("%DATASETDIR%", os.environ.get("DATASETDIR", self.dir + os.sep))
        return fn if os.path.isabs(fn) else os.path.join(self.dir, fn)


def read_paramnames(path):
    """-> list of (name, latex, is_derived)"""
    out = []
    with open(path) as f:
        for raw in f:
            line = raw.split("#")[0].strip()
            if not line:
                continue
            parts = line.split(None, 1)
            name = parts[0]
            derived = name.endswith("*")
            out.append((name.rstrip("*"), parts[1].strip() if len(parts) > 1 else "", derived))
    return out


def _top_comment(path):
    last = None
    with open(path) as f:
        for line in f:
            s = line.strip()
            if s.startswith("#"):
                last = s[1:].strip()
            elif s:
                break
    return last


class CMBLikesPlan:
    """Dense form of a binned CMBlikes data set (gaussian or HL), ready for cb200_like_add_cmblikes."""

    def __init__(self, dataset_path, overrides=None, cov_provider=None):
        """cov_provider: object with .block(rows, cols) standing in for the covmat_fiducial file when that blob is
        missing from the reference checkout (BK15_covmat_dust.dat; cosmomc_b200/synthetic.py documents the stand-in)."""
        ini = IniFile(dataset_path)
        if overrides:
            ini.params.update(overrides)
        self.name = os.path.basename(dataset_path)
        self.cov_provider = cov_provider
        self._read(ini)
        self._read_extra(ini)

    def _read_extra(self, ini):
        pass

    # -- naming helpers (CMBlikes.f90:196-330)
    def _pair_to_maps(self, s):
        if len(s) == 2 and not self.has_map_names:
            return self.map_names.index(s[0]), self.map_names.index(s[1])
        if "x" not in s:
            raise ValueError("CMBlikes: invalid spectrum name " + s)
        a, b = s.split("x", 1)
        return self.map_names.index(a), self.map_names.index(b)

    def _pair_to_used(self, index, s):
        i1, i2 = self._pair_to_maps(s)
        i1, i2 = index[i1], index[i2]
        return (i2, i1) if i2 > i1 else (i1, i2)

    def _element_index(self, i1, i2):
        if i1 < 0 or i2 < 0:
            return -1
        return i1 * (i1 + 1) // 2 + i2  # lower-triangle row-major, i1 >= i2 (MatrixToElements order)

    def _used_name(self, i, j):
        a, b = self.used_map_order[i], self.used_map_order[j]
        return a + "x" + b if self.has_map_names else a + b

    def _cols_from_order(self, names):
        cols = -np.ones(self.ncl, dtype=int)
        ix = 0
        for i in range(self.nmaps):
            for j in range(i + 1):
                nm = self._used_name(i, j)
                if nm not in names and i != j:
                    nm = self._used_name(j, i)
                if nm in names:
                    cols[ix] = names.index(nm)
                ix += 1
        return cols

    def _read_cl_arr(self, ini, stem):
        fn = ini.relative_file(stem + "_file")
        order = ini.string(stem + "_order", "")
        names = ("L " + order).split() if order else (_top_comment(fn) or "").split()
        if not names:
            raise ValueError("no column order for " + fn)
        cols = self._cols_from_order(names)
        data = np.loadtxt(fn)
        cl = np.zeros((self.ncl, self.nbins_used))
        Ls = data[:, 0].astype(int) - (1 if self.binned else 0)
        for r, L in enumerate(Ls):
            if self.bin_min <= L <= self.bin_max:
                for ix in range(self.ncl):
                    if cols[ix] != -1:
                        cl[ix, L - self.bin_min] = data[r, cols[ix]]
        if Ls[-1] < self.bin_max:
            raise ValueError("C_l file does not reach the last used bin: " + fn)
        return cl

    def _read_windows(self, ini, stem):
        in_cl = ini.split(stem + "_in_order")
        out_cl = ini.split(stem + "_out_order", in_cl)
        cols_in = [self._pair_to_used(self.map_required_index, s) for s in in_cl]
        cols_out = [self._element_index(*self._pair_to_used(self.map_used_index, s)) for s in out_cl]
        if len(cols_in) != len(cols_out):
            raise ValueError("_in_order and _out_order differ in length")
        nL = self.pcl_lmax - self.pcl_lmin + 1
        W = np.zeros((len(cols_in), self.nbins_used, nL))
        pattern = ini.relative_file(stem + "_files")
        for b in range(self.nbins_used):
            win = np.loadtxt(pattern.replace("%u", str(b + 1 + self.bin_min)))
            for row in win:
                L = int(row[0])
                if self.pcl_lmin <= L <= self.pcl_lmax:
                    W[:, b, L - self.pcl_lmin] = row[1:1 + len(cols_in)]
        return cols_in, cols_out, W

    def _read(self, ini):
        self.map_names = ini.split("map_names", [])
        self.has_map_names = len(self.map_names) > 0
        if self.has_map_names:
            mf = ini.split("map_fields")
            self.map_fields = [FIELDS.index(f) for f in mf]
        else:
            self.map_names = list(FIELDS)
            self.map_fields = list(range(4))
        fields_use = ini.split("fields_use", [])
        if fields_use:
            use_field = [FIELDS[i] in fields_use for i in range(4)]
        else:
            if not self.has_map_names:
                raise ValueError("CMBlikes: need fields_use or map_names")
            use_field = [True] * 4
        maps_use = ini.split("maps_use", [])
        if maps_use:
            self.use_map = [m in maps_use for m in self.map_names]
        else:
            self.use_map = [use_field[self.map_fields[i]] for i in range(len(self.map_names))]
        self.require_map = list(self.use_map)
        req = ini.split("maps_required" if self.has_map_names else "fields_required", [])
        for m in req:
            self.require_map[self.map_names.index(m)] = True
        self.like_approx = {"gaussian": 2, "HL": 1}[ini.string("like_approx", "gaussian")]
        self.nmaps = int(np.count_nonzero(self.use_map))
        self.nmaps_required = int(np.count_nonzero(self.require_map))
        self.required_order = [i for i, r in enumerate(self.require_map) if r]
        self.map_required_index = -np.ones(len(self.map_names), dtype=int)
        for k, i in enumerate(self.required_order):
            self.map_required_index[i] = k
        self.map_used_index = -np.ones(len(self.map_names), dtype=int)
        self.used_map_order = []
        for i, nm in enumerate(self.map_names):
            if self.use_map[i]:
                self.map_used_index[i] = len(self.used_map_order)
                self.used_map_order.append(nm)
        self.ncl = self.nmaps * (self.nmaps + 1) // 2
        self.pcl_lmax = ini.int("cl_lmax")
        self.pcl_lmin = ini.int("cl_lmin")
        self.binned = ini.bool("binned", True)
        if not self.binned:
            raise NotImplementedError("only binned CMBlikes data sets are supported")
        self.nbins = ini.int("nbins")
        self.bin_min = ini.int("use_min", 1) - 1
        self.bin_max = ini.int("use_max", self.nbins) - 1
        self.nbins_used = self.bin_max - self.bin_min + 1
        win_in, win_out, winW = self._read_windows(ini, "bin_window")
        self.bandpowers = self._read_cl_arr(ini, "cl_hat")
        cl_fid = self._read_cl_arr(ini, "cl_fiducial") if self.like_approx == 1 else None
        includes_noise = ini.bool("cl_hat_includes_noise", False)
        cl_noise = None
        if self.like_approx != 2 or includes_noise:
            cl_noise = self._read_cl_arr(ini, "cl_noise")
            if not includes_noise:
                self.bandpowers = self.bandpowers + cl_noise
            elif self.like_approx == 2:
                self.bandpowers = self.bandpowers - cl_noise
        if cl_fid is not None and not ini.bool("cl_fiducial_includes_noise", False):
            cl_fid = cl_fid + cl_noise

        def to_matrix(x):
            M = np.zeros((self.nmaps, self.nmaps))
            ix = 0
            for i in range(self.nmaps):
                for j in range(i + 1):
                    M[i, j] = M[j, i] = x[ix]
                    ix += 1
            return M

        def sqrtm(M):
            w, V = np.linalg.eigh(M)
            return (V * np.sqrt(w)) @ V.T

        nb = self.nbins_used
        self.chat = np.array([to_matrix(self.bandpowers[:, b]) for b in range(nb)])
        self.noise = np.array([to_matrix(cl_noise[:, b]) for b in range(nb)]) if cl_noise is not None and self.like_approx != 2 else None
        self.sqrt_fid = np.array([sqrtm(to_matrix(cl_fid[:, b])) for b in range(nb)]) if cl_fid is not None else None

        # covariance (ReadCovmat, CMBlikes.f90:752-859)
        covmat_cl = ini.split("covmat_cl")
        cl_in_index = [self._element_index(*self._pair_to_used(self.map_used_index, s)) for s in covmat_cl]
        used = [(k, ix) for k, ix in enumerate(cl_in_index) if ix >= 0]
        self.ncl_used = len(used)
        self.cl_use_index = np.array([ix for _, ix in used], dtype=np.int32)
        cov_cl_used = np.array([k for k, _ in used], dtype=int)
        scale = ini.float("covmat_scale", 1.0)
        num_in = len(cl_in_index)
        n = nb * self.ncl_used
        cov = np.zeros((n, n))
        if self.cov_provider is not None:
            rows = np.concatenate([(bx + self.bin_min) * num_in + cov_cl_used for bx in range(nb)])
            cov = scale * self.cov_provider.block(rows, rows)
        else:
            full_cov = np.loadtxt(ini.relative_file("covmat_fiducial"))
            for bx in range(nb):
                for by in range(nb):
                    cov[bx * self.ncl_used:(bx + 1) * self.ncl_used, by * self.ncl_used:(by + 1) * self.ncl_used] = \
                        scale * full_cov[np.ix_((bx + self.bin_min) * num_in + cov_cl_used, (by + self.bin_min) * num_in + cov_cl_used)]
        self.cov = cov
        inv = np.linalg.inv(cov)
        self.invcov = 0.5 * (inv + inv.T)

        # windows -> dense W[bin][cl][spec][l] on the device Cls layout [TT,TE,EE,BB,PP], l = 0..pcl_lmax
        self.lmax_w = self.pcl_lmax
        W = np.zeros((nb, self.ncl, 5, self.lmax_w + 1))
        offset = np.zeros((nb, self.ncl))

        def add(cols_in, cols_out, Wt):
            for k, ((i, j), out) in enumerate(zip(cols_in, cols_out)):
                if out < 0:
                    continue
                f = tuple(sorted((self.map_fields[self.required_order[i]], self.map_fields[self.required_order[j]])))
                if f not in SPEC_SLOT:
                    continue  # spectrum the theory never allocates (e.g. TB): contributes zero, CMBlikes.f90:1318
                W[:, out, SPEC_SLOT[f], self.pcl_lmin:self.pcl_lmax + 1] += Wt[k]

        add(win_in, win_out, winW)
        if ini.has("linear_correction_fiducial_file"):
            fid_corr = self._read_cl_arr(ini, "linear_correction_fiducial")  # [ncl][nbins]
            cin, cout, cW = self._read_windows(ini, "linear_correction_bin_window")
            add(cin, cout, cW)
            offset += fid_corr.T
        self.W, self.offset = W, offset

        # calibration / nuisance parameters (CMBlikes.f90:560-590)
        self.nuisance_names = []
        self.calibration_param = None
        if ini.has("nuisance_params"):
            self.nuisance_names = [n for n, _, _ in read_paramnames(ini.relative_file("nuisance_params"))]
            if ini.has("calibration_paramname"):
                self.calibration_param = ini.string("calibration_paramname")
        elif ini.has("calibration_param"):
            self.nuisance_names = [n for n, _, _ in read_paramnames(ini.relative_file("calibration_param"))]
            self.calibration_param = self.nuisance_names[0]
        self.log_cal_prior = ini.float("log_calibration_prior", -1.0) if self.calibration_param else -1.0
        if ini.float("aberration_coeff", 0.0) != 0.0:
            raise NotImplementedError("aberration_coeff != 0 is not folded into the dense windows yet")

    def register(self, handle, cal_index):
        return handle.add_cmblikes(self.nmaps, self.nbins_used, self.cl_use_index, self.like_approx, self.W, self.offset,
                                   self.chat, self.invcov, noise=self.noise, sqrt_fid=self.sqrt_fid,
                                   log_cal_prior=self.log_cal_prior, cal_index=cal_index)

    def binned_theory(self, cls, cal=1.0):
        """numpy evaluation of the dense form (for host-side checks): cls [5][>=lmax_w+1]."""
        c = np.array(cls[:, :self.lmax_w + 1], dtype=float)
        c[:4] = c[:4] / cal ** 2
        return np.einsum("bcxl,xl->bc", self.W, c) - self.offset


# ---------------------------------------------------------------------------------------------------------------
# background-only likelihoods: BAO / MGS (source/bao.f90:113-234), HST (source/HST.f90:20-45),
# JLA / Pantheon supernovae (source/supernovae_JLA.f90:602-765 read_jla_dataset, :874-991 jla_prep)
BAO_TYPES = ['Az', 'DV_over_rs', 'rs_over_DV', 'DA_over_rs', 'F_AP', 'f_sigma8', 'bao_Hz_rs', 'bao_Hz_rs_103',
             'dilation', 'DM_over_rs']   # 1-based codes, bao.f90:29-35


class BAOPlan:
    """BAO_ReadIni + BAO_InitProbDist for TBAOLikelihood and MGSLikelihood (tag 'MGS')."""

    def __init__(self, dataset_path, tag=None, overrides=None):
        ini = IniFile(dataset_path)
        if overrides:
            ini.params.update(overrides)
        self.name = ini.string("name", os.path.basename(dataset_path))
        self.tag = tag or self.name
        self.num_bao = ini.int("num_bao", 1)
        self.rs_rescale = ini.float("rs_rescale", 1.0)
        has_type = ini.has("measurement_type")
        self.types = np.zeros(self.num_bao, dtype=np.int32)
        if has_type:
            names = ini.split("measurement_type")
            if len(names) == 1:
                names = names * self.num_bao
            self.types[:] = [BAO_TYPES.index(n) + 1 for n in names]
        self.z = np.zeros(self.num_bao)
        self.obs = np.zeros(self.num_bao)
        self.err = np.zeros(self.num_bao)
        if ini.has("zeff"):
            zs = [float(x) for x in ini.split("zeff")]
            self.z[:] = zs if len(zs) == self.num_bao else zs[0]
            vals = [float(x) for x in ini.split("bao_measurement")]
            if self.num_bao > 1:
                self.obs[:] = vals[:self.num_bao]
            else:
                self.obs[0], self.err[0] = vals[0], vals[1]
        else:
            has_err = ini.bool("bao_measurements_file_has_error", True)
            rows = []
            with open(ini.relative_file("bao_measurements_file")) as f:
                for raw in f:
                    line = raw.split("#")[0].strip()
                    if line:
                        rows.append(line.split())
            for i in range(self.num_bao):
                r = rows[i]
                self.z[i], self.obs[i] = float(r[0]), float(r[1])
                k = 2
                if has_err:
                    self.err[i] = float(r[2])
                    k = 3
                if not has_type:
                    self.types[i] = BAO_TYPES.index(r[k]) + 1
        if np.any(self.z < 0.0001):
            raise ValueError("Error reading BAO measurements")
        self.invcov = np.zeros((self.num_bao, self.num_bao))
        self.alpha_prob = None
        if self.tag == "MGS":
            self.alpha_prob = np.loadtxt(ini.relative_file("prob_dist")).reshape(-1)
        elif ini.has("bao_invcov_file"):
            self.invcov = np.loadtxt(ini.relative_file("bao_invcov_file"))
        elif ini.has("bao_cov_file"):
            self.invcov = np.linalg.inv(np.loadtxt(ini.relative_file("bao_cov_file")))   # Matrix_Inverse at load
        else:
            self.invcov[np.diag_indices(self.num_bao)] = 1 / self.err ** 2

    def register(self, handle, fixed_rs=-1.0):
        if self.alpha_prob is not None:
            return handle.add_mgs(self.z[0], self.alpha_prob, fixed_rs)
        return handle.add_bao(self.types, self.z, self.obs, self.invcov, self.rs_rescale, fixed_rs)


class HSTPlan:
    def __init__(self, ini_path):
        ini = IniFile(ini_path)
        self.name = ini.string("Hubble_name")
        self.H0 = ini.float("Hubble_H0")
        self.H0_err = ini.float("Hubble_H0_err")
        self.zeff = ini.float("Hubble_zeff", 0.0)
        self.angconversion = ini.float("Hubble_angconversion", 0.0) if self.zeff > 0 else 0.0

    def register(self, handle):
        return handle.add_hst(self.H0, self.H0_err, self.zeff, self.angconversion)


SN_COV_NAMES = ["mag", "stretch", "colour", "mag_stretch", "mag_colour", "stretch_colour"]


class SNPlan:
    """read_jla_dataset + jla_prep.  `covs` (dict name -> [nsn][nsn]) overrides the *_covmat_file entries: the six JLA
    blocks and the Pantheon systematics matrix are missing from the reference checkout, so tests and the bench pass
    documented synthetic SPD stand-ins (cosmomc_b200/synthetic.py)."""

    def __init__(self, dataset_path, covs=None, data_file=None):
        ini = IniFile(dataset_path)
        self.name = ini.string("name")
        if data_file is None:
            df = ini.string("data_file")
            cand = [df, os.path.join(ini.dir, df), os.path.join(ini.dir, os.path.basename(df)),
                    os.path.join(ini.dir, "..", df)]
            data_file = next((c for c in cand if os.path.exists(c)), None)
            if data_file is None:
                raise FileNotFoundError(df)
        if ini.bool("absdist_file", False):
            raise ValueError("absdist_file not supported")
        self.pecz = ini.float("pecz", 0.001)
        self.twoscriptmfit = ini.bool("twoscriptmfit", False)
        self.scriptmcut = ini.float("scriptmcut", 10.0) if self.twoscriptmfit else 10.0
        idisp0 = ini.float("intrinsicdisp", 0.13)
        idisp = [ini.float("intrinsicdisp%d" % i, idisp0) for i in range(10)]
        cols, rows = None, []
        with open(data_file) as f:
            for raw in f:
                if raw.startswith("#"):
                    cols = raw[1:].split()
                elif raw.strip():
                    rows.append(raw.split())
        if cols is None:
            raise ValueError("SN data file must have a comment header")
        tab = {c: np.array([float(r[i]) if i > 0 else 0.0 for r in rows]) for i, c in enumerate(cols) if i < len(rows[0])}
        self.names = [r[0] for r in rows]
        self.nsn = len(rows)
        g = lambda k: tab.get(k, np.zeros(self.nsn))
        self.lc = {k: g(k) for k in ["zcmb", "zhel", "dz", "mb", "dmb", "x1", "dx1", "color", "dcolor", "3rdvar",
                                     "d3rdvar", "cov_m_s", "cov_m_c", "cov_s_c", "set"]}
        L = self.lc
        zfacsq = 25.0 / np.float64(np.float32(np.log(np.float32(10.0)))) ** 2   # single-precision literals, :882
        ds = L["set"].astype(int)
        intrinsicsq = np.array(idisp) ** 2
        self.pre_vars = L["dmb"] ** 2 + intrinsicsq[np.clip(ds, 0, 9)] + zfacsq * self.pecz ** 2 * (
            (1.0 + L["zcmb"]) / (L["zcmb"] * (1 + 0.5 * L["zcmb"]))) ** 2
        self.A1 = np.ones(self.nsn)
        self.A2 = np.zeros(self.nsn)
        if self.twoscriptmfit:
            self.A1 = (L["3rdvar"] <= self.scriptmcut).astype(np.float64)
            self.A2 = 1.0 - self.A1
            if not self.A1.any():
                self.A1, self.A2 = self.A2, np.zeros(self.nsn)
                self.twoscriptmfit = False
            if not self.A2.any():
                self.twoscriptmfit = False
        self.covs = {}
        for nm in SN_COV_NAMES:
            if ini.bool("has_%s_covmat" % nm, False):
                if covs is not None and nm in covs:
                    self.covs[nm] = np.ascontiguousarray(covs[nm], dtype=np.float64)
                else:
                    fn = ini.string("%s_covmat_file" % nm)
                    cand = [fn, os.path.join(ini.dir, fn), os.path.join(ini.dir, os.path.basename(fn))]
                    path = next((c for c in cand if os.path.exists(c)), None)
                    if path is None:
                        raise FileNotFoundError("%s (missing from the reference checkout; pass covs=...)" % fn)
                    c = np.loadtxt(path).reshape(-1)
                    if len(c) == self.nsn ** 2 + 1:
                        c = c[1:]
                    self.covs[nm] = c.reshape(self.nsn, self.nsn)
        self.alphabeta_covmat = len(self.covs) > 1 or "mag" not in self.covs

    def columns(self):
        L = self.lc
        return np.stack([L["zcmb"], L["zhel"], L["mb"], L["x1"], L["color"], self.pre_vars, L["dx1"] ** 2,
                         L["dcolor"] ** 2, L["cov_m_s"], L["cov_m_c"], L["cov_s_c"]])

    def register(self, handle, alpha_index=-1, beta_index=-1):
        covs = [self.covs.get(nm) for nm in SN_COV_NAMES]
        if not self.alphabeta_covmat:
            alpha_index = beta_index = -1
        return handle.add_sn(self.columns(), covs, self.A1, self.A2, self.twoscriptmfit, alpha_index, beta_index)


# ---------------------------------------------------------------------------------------------------------------
# BICEP/Keck (source/CMB_BK_Planck.f90): TBK_planck_ReadIni (:36-70), TBK_planck_Read_Bandpass (:72-105)
BK_T_CMB = 2.72548
BK_GHZ_KELVIN = 6.62606957e-34 / 1.3806488e-23 * 1e9
LFORM = {"flat": 0, "lin": 1, "quad": 2}


def read_bandpass(fname, fpivot_dust, fpivot_sync):
    R = np.loadtxt(fname)
    nu = R[:, 0]
    n = len(nu)
    dnu = np.zeros(n)
    dnu[0] = nu[1] - nu[0]
    dnu[1:n - 1] = (nu[2:] - nu[:n - 2]) / 2
    dnu[n - 1] = nu[n - 1] - nu[n - 2]
    ex = np.exp(BK_GHZ_KELVIN * nu / BK_T_CMB)
    th_int = np.sum(dnu * R[:, 1] * nu ** 4 * ex / (ex - 1) ** 2)

    def th0(nu0):
        e = np.exp(BK_GHZ_KELVIN * nu0 / BK_T_CMB)
        return nu0 ** 4 * e / (e - 1) ** 2
    nu_bar = np.sum(dnu * nu * R[:, 1]) / np.sum(dnu * R[:, 1])
    return dict(nu=nu, R=R[:, 1].copy(), dnu=dnu, th_dust=th_int / th0(fpivot_dust), th_sync=th_int / th0(fpivot_sync),
                nu_bar=nu_bar)


class BK15Plan(CMBLikesPlan):
    """BK15-style data set: generic binned HL CMBlikes + the foreground model of TBK_planck."""

    def _read_extra(self, ini):
        self.fpivot_dust = ini.float("fpivot_dust", 353.0)
        self.fpivot_sync = ini.float("fpivot_sync", 23.0)
        self.fpivot_dust_decorr = [ini.float("fpivot_dust_decorr(1)", 217.0), ini.float("fpivot_dust_decorr(2)", 353.0)]
        self.fpivot_sync_decorr = [ini.float("fpivot_sync_decorr(1)", 23.0), ini.float("fpivot_sync_decorr(2)", 33.0)]
        self.lform_dust = LFORM.get(ini.string("lform_dust_decorr", "flat"), 0)
        self.lform_sync = LFORM.get(ini.string("lform_sync_decorr", "flat"), 0)
        if self.nmaps != self.nmaps_required:
            raise NotImplementedError("BK foregrounds with maps_required beyond maps_use")
        self.bandpasses = [read_bandpass(ini.relative_file("bandpass[%s]" % m), self.fpivot_dust, self.fpivot_sync)
                           for m in self.used_map_order]
        self.bc_class = [1 if "95" in m else 2 if "150" in m else 3 if "220" in m else 0 for m in self.used_map_order]
        self.used_fields = [self.map_fields[self.map_names.index(m)] for m in self.used_map_order]
        # band-power window of every used map pair on its own EE / BB spectrum (what the foreground model is binned with)
        fgW = np.zeros((self.nbins_used, self.ncl, self.lmax_w + 1))
        ix = 0
        for i in range(self.nmaps):
            for j in range(i + 1):
                fi, fj = self.used_fields[i], self.used_fields[j]
                if fi == fj and fi in (1, 2):
                    fgW[:, ix, :] = self.W[:, ix, SPEC_SLOT[(fi, fi)], :]
                ix += 1
        self.fgW = fgW

    def register(self, handle, nuis_offset=0, cal_index=-1):
        lid = CMBLikesPlan.register(self, handle, cal_index)
        handle.set_bk_foregrounds(lid, self.used_fields, self.bc_class,
                                  [(b["nu"], b["R"], b["dnu"]) for b in self.bandpasses],
                                  [b["th_dust"] for b in self.bandpasses], [b["th_sync"] for b in self.bandpasses],
                                  [b["nu_bar"] for b in self.bandpasses], self.fpivot_dust, self.fpivot_sync,
                                  self.fpivot_dust_decorr, self.fpivot_sync_decorr, self.lform_dust, self.lform_sync,
                                  self.pcl_lmin, self.pcl_lmax, self.fgW, nuis_offset)
        return lid

    # ---- compact fixture (tests / bench on the GPU box have no reference tree)
    PACK_KEYS = ["nmaps", "nbins_used", "ncl", "ncl_used", "like_approx", "lmax_w", "pcl_lmin", "pcl_lmax", "cl_use_index",
                 "offset", "chat", "noise", "sqrt_fid", "fgW", "used_fields", "bc_class", "fpivot_dust", "fpivot_sync",
                 "fpivot_dust_decorr", "fpivot_sync_decorr", "lform_dust", "lform_sync", "cov"]

    def save_pack(self, path):
        d = {k: np.asarray(getattr(self, k)) for k in self.PACK_KEYS}
        d["used_map_order"] = np.array(self.used_map_order)
        for i, b in enumerate(self.bandpasses):
            for k, v in b.items():
                d["bp%d_%s" % (i, k)] = np.asarray(v)
        np.savez_compressed(path, **d)

    @classmethod
    def from_pack(cls, path):
        z = np.load(path)
        self = cls.__new__(cls)
        for k in cls.PACK_KEYS:
            v = z[k]
            setattr(self, k, v.item() if v.ndim == 0 else v)
        for k in ("nmaps", "nbins_used", "ncl", "ncl_used", "like_approx", "lmax_w", "pcl_lmin", "pcl_lmax", "lform_dust", "lform_sync"):
            setattr(self, k, int(getattr(self, k)))
        self.used_map_order = [str(x) for x in z["used_map_order"]]
        self.used_fields = [int(x) for x in self.used_fields]
        self.bc_class = [int(x) for x in self.bc_class]
        self.bandpasses = [{k: (z["bp%d_%s" % (i, k)] if z["bp%d_%s" % (i, k)].ndim else float(z["bp%d_%s" % (i, k)]))
                            for k in ("nu", "R", "dnu", "th_dust", "th_sync", "nu_bar")} for i in range(self.nmaps)]
        self.nmaps_required = self.nmaps
        inv = np.linalg.inv(self.cov)
        self.invcov = 0.5 * (inv + inv.T)
        self.log_cal_prior = -1.0
        W = np.zeros((self.nbins_used, self.ncl, 5, self.lmax_w + 1))
        ix = 0
        for i in range(self.nmaps):
            for j in range(i + 1):
                fi, fj = self.used_fields[i], self.used_fields[j]
                if fi == fj and fi in (1, 2):
                    W[:, ix, SPEC_SLOT[(fi, fi)], :] = self.fgW[:, ix, :]
                ix += 1
        self.W = W
        return self
