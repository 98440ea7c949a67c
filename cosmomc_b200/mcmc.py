"""Batched adaptive Metropolis driver: many chains stepped in lockstep, ONE batched likelihood call per step.

CosmoMC runs one chain per MPI rank and evaluates one parameter point per call (`source/driver.F90:54-57`,
`source/calclike.f90:136`).  The B200 library wants batches, so here K chains advance together and the K proposals of
a step are one call of `loglike_fn(P[K, n]) -> -lnL[K]` (`Handle.eval_batch`, `Handle.test_like_batch`, ...).  With
torch.distributed initialised, every rank owns K local chains; the only communication is the all-gather of the
per-chain (count, mean, covariance) records behind the proposal-covariance update and the R-1 convergence test
(`cosmomc_b200/chains.py`) - NCCL on the GPU box, gloo in the CPU tests.

Reference behaviour mirrored (host logic, not a kernel):
  * RandDirectionProposer%ProposeVec / Propose_r  (`source/propose.f90:97-134`): cycle through the columns of a random
    rotation, step length r = exponential (prob. 0.33) or sqrt(chi2_min(n,2) / min(n,2)), times propose_scale;
  * BlockProposer%UpdateParams                     (`source/propose.f90:137-144`): P += mapping_matrix . vec with
    mapping_matrix = sigma_i chol(corr) (`chains.proposal_mapping`, propose.f90:210-244);
  * TChainSampler_MetropolisAccept                 (`source/MCMC.f90:119-131`): accept if the new -lnL is lower, else
    with probability exp(-(Like - CurLike)); logZero (1e30) is never accepted;
  * chain rows: a point is written with its multiplicity when the chain leaves it (`source/MCMC.f90:133-180`,
    `(*(E16.7))` rows [weight, -lnL, params], `source/GeneralTypes.f90:254-274`);
  * TMpiChainCollector_UpdateCovAndCheckConverge    (`source/SampleCollector.f90:212-322`): every
    `update_every` stored samples, pooled covariance -> new proposal (MPI_LearnPropose), R-1 < converge_test stops.
The random streams are numpy Generators (one per chain), not CosmoMC's ranmar: chains are statistically, not
bit-wise, the reference's.
"""
import os

import numpy as np

from . import chains as _chains

LOG_ZERO = 1e30


class RandDirectionProposer:
    """source/propose.f90:97-134 for one block of n parameters."""

    def __init__(self, n, rng):
        self.n, self.rng = n, rng
        self.loopix = 0
        self.R = None

    def _rotation(self):
        n = self.n
        if n == 1:
            return np.array([[1.0 if self.rng.random() >= 0.5 else -1.0]])
        q, r = np.linalg.qr(self.rng.normal(size=(n, n)))
        return q * np.sign(np.diag(r))[None, :]          # Haar-distributed rotation (RandRotation)

    def propose_r(self):
        if self.rng.random() < 0.33:
            return self.rng.exponential()
        m = min(self.n, 2)
        return np.sqrt((self.rng.normal(size=m) ** 2).sum() / m)

    def propose_vec(self, wid):
        if self.loopix % self.n == 0:
            self.R = self._rotation()
            self.loopix = 0
        self.loopix += 1
        return self.R[:, self.loopix - 1] * (self.propose_r() * wid)


class BatchedMetropolis:
    def __init__(self, loglike_fn, start, propose_cov, pmin=None, pmax=None, propose_scale=2.4, seed=0, rank=0,
                 update_every=None, converge_test=0.01, learn_propose=True, chain_root=None, names=None,
                 chain_offset=None, max_R_propose_update=2.0, R_stop_propose_update=0.0, max_stored=500000):
        """start [K][n]: starting points of this rank's K chains (all n parameters vary).
        chain_offset: global index of this rank's first chain (default rank * K).  The random stream of a chain depends
        on (seed, global chain index) only, so K x world chains evolve identically however they are spread over ranks.
        max_R_propose_update / R_stop_propose_update: MPI_Max_R_ProposeUpdate / MPI_R_StopProposeUpdate
        (source/SampleCollector.f90:306-313): the proposal is re-learned only while R_stop < R-1 < max_R (or while the
        current proposal matrix is diagonal, or with a single chain)."""
        self.f = loglike_fn
        self.P = np.array(start, dtype=np.float64)
        self.K, self.n = self.P.shape
        self.pmin = None if pmin is None else np.asarray(pmin, dtype=np.float64)
        self.pmax = None if pmax is None else np.asarray(pmax, dtype=np.float64)
        self.scale = propose_scale
        self.set_covariance(propose_cov)
        self.chain_offset = rank * self.K if chain_offset is None else int(chain_offset)
        self.rngs = [np.random.default_rng([seed, self.chain_offset + k]) for k in range(self.K)]
        self.prop = [RandDirectionProposer(self.n, r) for r in self.rngs]
        self.like = self._eval(self.P)
        if np.any(self.like >= LOG_ZERO):
            raise ValueError("a starting point is rejected (logZero)")
        self.mult = np.ones(self.K)
        self.samples = [[self.P[k].copy()] for k in range(self.K)]   # every step's current point (thinning 1)
        self.update_every = update_every or 50 * self.n               # MPI_Sample_update_freq x n
        self.converge_test = converge_test
        self.learn = learn_propose
        self.R_history = []
        self.max_R_propose_update = max_R_propose_update
        self.R_stop_propose_update = R_stop_propose_update
        self.cov_is_diagonal = bool(np.count_nonzero(self.cov - np.diag(np.diag(self.cov))) == 0)
        self.flukecheck = False          # TMpiChainCollector%flukecheck: R-1 has to pass twice in a row
        self.converged = False
        self.thin_fac = 1
        self.max_stored = max_stored
        self.n_steps = 0
        self.n_accept = np.zeros(self.K, dtype=np.int64)
        self.rank = rank
        self.files = None
        if chain_root is not None:
            os.makedirs(os.path.dirname(os.path.abspath(chain_root)), exist_ok=True)
            self.files = [open("%s_%d.txt" % (chain_root, self.chain_offset + k + 1), "w") for k in range(self.K)]
            if rank == 0 and names is not None:
                write_paramnames(chain_root + ".paramnames", names)
                write_ranges(chain_root + ".ranges", names, self.pmin, self.pmax)

    def set_covariance(self, cov):
        self.cov = np.array(cov, dtype=np.float64)
        self.mapping = _chains.proposal_mapping(self.cov)

    def _eval(self, P):
        out = np.asarray(self.f(P), dtype=np.float64).copy()
        if self.pmin is not None:
            out[np.any(P < self.pmin[None, :], axis=1)] = LOG_ZERO   # GetLogLikeBounds
        if self.pmax is not None:
            out[np.any(P > self.pmax[None, :], axis=1)] = LOG_ZERO
        return out

    def step(self):
        trial = np.empty_like(self.P)
        for k in range(self.K):
            trial[k] = self.P[k] + self.mapping @ self.prop[k].propose_vec(self.scale)
        like = self._eval(trial)
        for k in range(self.K):
            ok = like[k] < LOG_ZERO and (self.like[k] > like[k] or self.rngs[k].exponential() > like[k] - self.like[k])
            if ok:
                if self.files is not None:
                    self.files[k].write(_chains.format_chain_row(self.mult[k], self.like[k], self.P[k]) + "\n")
                self.P[k], self.like[k], self.mult[k] = trial[k], like[k], 1.0
                self.n_accept[k] += 1
            else:
                self.mult[k] += 1.0
            if self.n_steps % self.thin_fac == 0:
                self.samples[k].append(self.P[k].copy())
        self.n_steps += 1

    def update(self):
        """TMpiChainCollector_UpdateCovAndCheckConverge (source/SampleCollector.f90:212-322) on this rank's chains + the
        all-gather.  Returns the pooled statistics dict; sets self.converged after two consecutive passes of R-1."""
        st = _chains.update_cov_and_check_converge([np.asarray(s) for s in self.samples], self.n)
        if not st.get("ready"):
            return st
        R = st.get("R")
        single = R is None                      # MPIChains == 1
        if not single:
            self.R_history.append(R)
            if R < self.converge_test and self.flukecheck:
                self.converged = True
            self.flukecheck = R < self.converge_test
            if max(len(x) for x in self.samples) > self.max_stored:   # Samples%Thin(2), MPI_thin_fac *= 2
                self.samples = [x[::2] for x in self.samples]
                self.thin_fac *= 2
        if self.learn and (single or ((self.cov_is_diagonal or R < self.max_R_propose_update)
                                      and R > self.R_stop_propose_update)):
            self.set_covariance(st["cov"])
            self.cov_is_diagonal = False
        return st

    def run(self, max_steps, min_steps=0):
        """Step until R-1 < converge_test at two consecutive checks (every update_every steps) or max_steps."""
        while self.n_steps < max_steps:
            self.step()
            if self.n_steps % self.update_every == 0:
                self.update()
                if self.converged and self.n_steps >= min_steps:
                    self.close()
                    return True
        self.close()
        return False

    def close(self):
        if self.files is not None:
            for k, f in enumerate(self.files):
                f.write(_chains.format_chain_row(self.mult[k], self.like[k], self.P[k]) + "\n")
                f.close()
            self.files = None


# ---- chain-side files GetDist reads (source/GeneralTypes.f90:618-736, source/ParamNames.f90, IO.f90) --------------
def write_paramnames(path, names, labels=None, derived=()):
    """`<root>.paramnames`: one `name<TAB>label` line per column after [weight, -lnL]; derived names end with `*`."""
    with open(path, "w") as f:
        for i, n in enumerate(names):
            lab = labels[i] if labels is not None else n
            f.write("%s%s\t%s\n" % (n, "*" if n in derived else "", lab))


def write_ranges(path, names, pmin, pmax):
    """`<root>.ranges`: `name  min  max` with N for an open end (source/ParamNames / getdist ParamBounds)."""
    with open(path, "w") as f:
        for i, n in enumerate(names):
            lo = "N" if pmin is None or not np.isfinite(pmin[i]) else "%.7E" % pmin[i]
            hi = "N" if pmax is None or not np.isfinite(pmax[i]) else "%.7E" % pmax[i]
            f.write("%-22s %16s %16s\n" % (n, lo, hi))


def read_chain(path):
    """Inverse of the row writer: returns (weights, loglikes, params[rows][n])."""
    a = np.atleast_2d(np.loadtxt(path))
    return a[:, 0], a[:, 1], a[:, 2:]


def write_likelihoods(path, likelihoods):
    """`<root>.likelihoods` (TLikelihoodList%OutputDescription, source/GeneralTypes.f90:796-817): one TAB-joined line
    `1 <LikelihoodType> <tag> <name> <version>` per likelihood, in the original (registration) order.
    likelihoods: iterable of dicts / tuples (type, tag, name, version)."""
    with open(path, "w") as f:
        for L in likelihoods:
            if isinstance(L, dict):
                L = (L.get("type", ""), L.get("tag", ""), L.get("name", ""), L.get("version", ""))
            f.write("\t".join(["1"] + [str(x).strip() for x in L]) + "\n")


def likelihood_derived_params(likelihoods, loglike, type_indices=(), derived=None):
    """addLikelihoodDerivedParams (source/GeneralTypes.f90:751-795) for a batch: appends to `derived` [npts][nd]
    chi2 of every likelihood (2 x -lnL, original order), chi2_prior = 2 (logLike - sum of the likelihoods), and the
    total chi2 of every likelihood type (type_indices: list of index lists).  logLike is the un-tempered total -lnL
    including priors (what cb200_eval_batch returns at temperature 1)."""
    likes = np.atleast_2d(np.asarray(likelihoods, dtype=np.float64))
    ll = np.asarray(loglike, dtype=np.float64).reshape(-1)
    cols = [likes * 2, (2 * (ll - likes.sum(axis=1)))[:, None]]
    for idx in type_indices:
        cols.append((2 * likes[:, list(idx)].sum(axis=1))[:, None])
    out = np.concatenate(cols, axis=1)
    if derived is not None:
        out = np.concatenate([np.atleast_2d(np.asarray(derived, dtype=np.float64)), out], axis=1)
    return out


def _fortran_e(v, width, digits):
    """Fortran `Ew.d` edit descriptor (0.dddE+ee form, no scale factor), as gfortran / ifort print it."""
    v = float(v)
    if v == 0.0 or not np.isfinite(v):
        s = ("0." + "0" * digits + "E+00") if v == 0.0 else str(v)
        return s.rjust(width)
    mant, exp = ("%.*e" % (digits - 1, abs(v))).split("e")   # d.ddd e+xx  ->  0.dddd E+(xx+1)
    e = int(exp) + 1
    body = "0." + mant.replace(".", "") + "E%s%02d" % ("+" if e >= 0 else "-", abs(e))
    return (("-" if v < 0 else "") + body).rjust(width)


def write_theory_cl(path, cls, lmax=None, fields=("TT", "TE", "EE", "BB", "PP"), digits=5):
    """`<root>.theory_cl` (TCosmoTheoryPredictions%WriteTextCls, source/CosmoTheory.f90:197-232): header
    `#    L    ` + one 15-wide column title per spectrum, then `(1I6,*(E15.5))` rows for L = 2..lmax.
    cls: [n_spectra][>= lmax+1] in the library's output order (l(l+1)C_l/2pi in muK^2, [L(L+1)]^2 C^pp/2pi for PP).
    The golden data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl was written with six digits (an older format
    statement): pass digits=6 to reproduce that file byte for byte."""
    cls = np.asarray(cls, dtype=np.float64)
    lmax = cls.shape[1] - 1 if lmax is None else int(lmax)
    with open(path, "w") as f:
        f.write(("#    L    " + "".join(t + " " * 13 for t in fields)).rstrip() + "\n")
        for L in range(2, lmax + 1):
            f.write("%6d" % L + "".join(_fortran_e(cls[i, L], 15, digits) for i in range(len(fields))) + "\n")


def read_theory_cl(path):
    """Inverse of write_theory_cl: returns (L[int], cls[n_spectra][len(L)])."""
    a = np.atleast_2d(np.loadtxt(path))
    return a[:, 0].astype(int), a[:, 1:].T.copy()


# ---- derived-parameter block and the .minimum file -------------------------------------------------
# ThermoDerivedParams order (camb/modules.f90:244-246; cb200_thermo out[12:25])
THERMO_DERIVED = ["age", "zstar", "rstar", "thetastar", "DAstar", "zdrag", "rdrag", "kd", "thetad", "zeq", "keq",
                  "thetaeq", "thetarseq"]
DERIVED_CL = (40, 220, 810, 1420, 2000)


def calc_derived_params(cmb, thermo_derived, rms_deflect, cl_TT=None, sigma8=0.0, pivot_k=0.05, bbn_dh=None,
                        background_outputs=(), lss_outputs=(), tensor=None):
    """TP_CalcDerivedParams (source/CosmologyParameterizations.f90:189-272) for ONE point, in the reference's column
    order: H0, omegal, omegam, omegamh2, omeganuh2, omegamh3, sigma8, S8, s8omegamp5, s8omegamp25, s8h5, rdragh,
    rmsdeflect, zrei, A, clamp, DL40..DL2000, ns02, yheused, YpBBN[, DHBBN], ThermoDerivedParams(1:13), then
    CAMB's background outputs (H(z), D_M(z) pairs), then (f sigma8, sigma8)(z) pairs, then the six tensor entries.
    cmb: dict with H0, omv, omdm, omb, omdmh2, ombh2, omnuh2, h, zre, tau, logA, ns, nrun, nrunrun, yhe.
    sigma8 comes from CAMB's transfer functions (not built: pass the CPU path's value)."""
    omm = cmb["omdm"] + cmb["omb"]
    d = [cmb["H0"], cmb["omv"], omm, cmb["omdmh2"] + cmb["ombh2"], cmb["omnuh2"], (cmb["omdmh2"] + cmb["ombh2"]) * cmb["h"],
         sigma8, sigma8 * (omm / 0.3) ** 0.5, sigma8 * omm ** 0.5, sigma8 * omm ** 0.25, sigma8 / cmb["h"] ** 0.5,
         thermo_derived[6] * cmb["H0"] / 100, rms_deflect, cmb["zre"]]
    As9 = 1e-10 * np.exp(cmb["logA"]) * 1e9
    d += [As9, As9 * np.exp(-2 * cmb["tau"])]
    d += [float(cl_TT[L]) if cl_TT is not None else 0.0 for L in DERIVED_CL]
    lograt = np.log(0.002 / pivot_k)
    d.append(cmb["ns"] + cmb.get("nrun", 0.0) * lograt + cmb.get("nrunrun", 0.0) * lograt ** 2 / 2)
    m_H, not4 = 1.673575e-27, 3.9715          # GetYpBBN (source/bbn.f90:26-36)
    m_He = m_H * not4
    d += [cmb["yhe"], 4 * m_H * cmb["yhe"] / (m_He - cmb["yhe"] * (m_He - 4 * m_H))]
    if bbn_dh is not None:
        d.append(bbn_dh)
    d += [float(x) for x in thermo_derived[:13]]
    d += [float(x) for x in background_outputs]
    d += [float(x) for x in lss_outputs]
    if tensor is not None:
        d += [float(x) for x in tensor]
    return np.array(d)


def _list_directed_real(v):
    """List-directed output of a double as the build that wrote the reference's golden files prints it (Intel
    Fortran: G24.15E3): F editing with 15 significant digits and 5 trailing blanks for 0.1 <= |v| < 1e15, E24.15E3 else."""
    v = float(v)
    a = abs(v)
    if v == 0.0:
        return "  0.000000000000000E+000"
    if 0.1 <= a < 1e15:
        k = int(np.floor(np.log10(a))) + 1 if a >= 1 else 0
        s = "%.*f" % (15 - k, a)
        if len(s.split(".")[0]) > max(k, 1):       # rounding carried into a new leading digit
            k += 1
            s = "%.*f" % (15 - k, a)
        return (("-" if v < 0 else "") + s).rjust(19) + " " * 5
    mant, exp = ("%.14e" % a).split("e")
    e = int(exp) + 1
    return (("-" if v < 0 else "") + "0." + mant.replace(".", "") + "E%s%03d" % ("+" if e >= 0 else "-", abs(e))).rjust(24)


def write_minimum(path, loglike, values, varying, names, labels, derived=(), derived_names=(), derived_labels=(),
                  like_contribs=(), weight=None):
    """`<root>.minimum` (TTheoryLike_WriteParamsHumanText, source/calclike.f90:208-236,436-462; called from
    source/driver.F90:230): -log(Like) and chi-sq, the varied parameters, the fixed ones, the derived block - each line
    `(1I5,1E15.7,"   ",1A22)` + label - and one `(2f11.3)   <type>: <tag> = <name> <version>` line per likelihood
    (WriteLikelihoodContribs, source/GeneralTypes.f90:562-583).  values / varying / names / labels: every base parameter in
    index order (1-based numbering in the file).  like_contribs: (-lnL, type, tag, name, version) tuples.
    tests/test_writers.py regenerates the reference's golden .minimum from its own contents byte for byte."""
    with open(path, "w") as f:
        if weight is not None:
            f.write("  weight    = " + _list_directed_real(weight) + "\n")
        f.write(" -log(Like) = " + _list_directed_real(loglike) + "\n")
        f.write("  chi-sq    = " + _list_directed_real(loglike * 2) + "\n")
        f.write(" \n")
        n = len(values)
        for used in (True, False):
            for i in range(n):
                if bool(varying[i]) == used:
                    f.write("%5d%s   %-22s%s\n" % (i + 1, _fortran_e(values[i], 15, 7), names[i][:22], labels[i].rstrip()))
            f.write(" \n")
        for i in range(len(derived)):
            f.write("%5d%s   %-22s%s\n" % (n + i + 1, _fortran_e(derived[i], 15, 7), derived_names[i][:22], derived_labels[i].rstrip()))
        f.write(" \n")
        f.write(" -log(Like)     chi-sq   data\n")
        for L in like_contribs:
            v, typ, tag, name, version = (list(L) + [""] * 5)[:5]
            tagname = name.strip()
            if tag and tag != name:
                tagname = tag + " = " + tagname
            line = "%11.3f%11.3f" % (v, v * 2) + "   " + typ.strip() + ": " + tagname
            if version:
                line += " " + version.strip()
            f.write(line + "\n")


def read_minimum(path):
    """Parse a .minimum file back into (loglike, rows, like_contribs); rows = [(index, value, name, label, block)] with
    block 0 = varied, 1 = fixed, 2 = derived."""
    lines = open(path).read().split("\n")
    loglike = float(lines[0].split("=")[1])
    rows, contribs, block, i = [], [], 0, 3
    while i < len(lines) and not lines[i].startswith(" -log(Like)     chi-sq"):
        ln = lines[i]
        if ln.strip() == "":
            block += 1
        else:
            rows.append((int(ln[:5]), float(ln[5:20]), ln[23:45].rstrip(), ln[45:], block))
        i += 1
    for ln in lines[i + 1:]:
        if ln.strip():
            typ, rest = ln[25:].split(": ", 1)
            contribs.append((float(ln[:11]), typ, rest))
    return loglike, rows, contribs


# ---- binary `.data` files (full model output for importance sampling) -------------------------------------------------
# ParamSet_WriteModel (source/ParamSet.f90:32-73) + TCosmoTheoryPredictions_WriteTheory (source/CosmoTheory.f90:235-282) on a
# Fortran access='stream' unit (source/FileUtils.f90: no record markers; strings as int32 length + bytes, "sized" arrays as
# int32 extent(s) + values, 2-D arrays column-major).  The first record of the theory block is the memory image of the
# derived type TCosmoTheoryParams (source/CosmologyTypes.f90:30-71), which is processor-dependent: the layout below is the
# one gfortran and ifort produce on x86-64 (4-byte default logical / integer, 8-byte reals on 8-byte boundaries, 160 bytes;
# .true. written as 1).  No `.data` file ships with the reference, so this writer is pinned only by the reader next to it
# (the mirror of ParamSet_ReadModel / ReadTheory) - parity unpinned, stated in DESIGN.md.  The matter-power block
# (use_matterpower) is not written: the path has no P(k) arrays.
_THEORY_PARAMS_FMT = "<4i6i5d4i3i4xdi4x2d2i2i"
_THEORY_PARAMS_KEYS = ("get_sigma8", "use_LSS", "use_CMB", "use_nonlinear", "lmax", "num_cls", "lmax_tensor",
                       "lmax_computed_cl", "lmin_computed_cl", "lmin_store_all_cmb", "z_outputs", "CMB_lensing",
                       "use_lensing_potential", "use_nonlinear_lensing", "compute_tensors", "use_matterpower",
                       "use_Weylpower", "use_sigmaR", "power_kmax", "num_power_redshifts", "pivot_k", "tensor_pivot_k",
                       "inflation_consistency", "bbn_consistency", "num_massive_neutrinos", "neutrino_hierarchy")
THEORY_PARAMS_DEFAULT = dict(get_sigma8=1, use_LSS=0, use_CMB=1, use_nonlinear=0, lmax=2500, num_cls=4, lmax_tensor=600,
                             lmax_computed_cl=2500, lmin_computed_cl=2000, lmin_store_all_cmb=2500,
                             z_outputs=(0.15, 0.38, 0.51, 0.61, 2.33), CMB_lensing=1, use_lensing_potential=1,
                             use_nonlinear_lensing=1, compute_tensors=0, use_matterpower=0, use_Weylpower=0, use_sigmaR=0,
                             power_kmax=0.8, num_power_redshifts=0, pivot_k=0.05, tensor_pivot_k=0.05,
                             inflation_consistency=1, bbn_consistency=1, num_massive_neutrinos=-1, neutrino_hierarchy=1)


def _pack_theory_params(s):
    import struct
    v = []
    for k in _THEORY_PARAMS_KEYS:
        x = s[k]
        v.extend(x if k == "z_outputs" else [x])
    return struct.pack(_THEORY_PARAMS_FMT, *v)


def _unpack_theory_params(b):
    import struct
    v = list(struct.unpack(_THEORY_PARAMS_FMT, b))
    out = {}
    for k in _THEORY_PARAMS_KEYS:
        if k == "z_outputs":
            out[k] = tuple(v[:5]); v = v[5:]
        else:
            out[k] = v.pop(0)
    return out


def _wstr(f, s):
    import struct
    b = s.strip().encode()
    f.write(struct.pack("<i", len(b))); f.write(b)


def _wsized(f, a):
    import struct
    a = np.asarray(a)
    f.write(struct.pack("<%di" % a.ndim, *a.shape))
    f.write(np.asfortranarray(a).astype("<i4" if a.dtype.kind in "iu" else "<f8").tobytes(order="F"))


def write_data_model(f, first, mult, like, likelihoods, params_used, theory, param_names=None, like_names=(),
                     settings=None, cl_lmax=None):
    """One model of a `.data` file.  f: binary file object; first: write the file header too (ParamSet.f90:40-63,
    CosmoTheory.f90:243-250).  theory: dict(derived=[...], cls={(i, j): array over l = 1..cl_lmax[i, j]} with 1-based
    field indices T, E, B, Phi and i >= j, tensor_ratio_02, tensor_ratio_C10, tensor_ratio_BB, tensor_AT,
    lensing_rms_deflect, sigma_8).  cl_lmax: [num_cls][num_cls] int array (0 = spectrum not stored)."""
    import struct
    s = dict(THEORY_PARAMS_DEFAULT, **(settings or {}))
    if s["use_matterpower"] or s["use_LSS"]:
        raise NotImplementedError("matter-power block of .data files")
    cl_lmax = np.asarray(cl_lmax, dtype=np.int32)
    params_used = np.asarray(params_used, dtype="<f8")
    if first:
        f.write(struct.pack("<2i", 4, len(params_used)))                       # 4: double-precision mcp
        has_names = param_names is not None and all(n != "" for n in param_names)
        f.write(struct.pack("<i", 1 if has_names else 0))
        if has_names:
            for n in param_names:
                _wstr(f, n)
        f.write(struct.pack("<i", len(like_names)))
        for n in like_names:
            _wstr(f, n)
        f.write(struct.pack("<i", 0))                                          # "unused"
    f.write(struct.pack("<2d", mult, like))
    f.write(np.asarray(likelihoods, dtype="<f8").tobytes())
    f.write(params_used.tobytes())
    if first:
        f.write(_pack_theory_params(s))
        if s["use_CMB"]:
            _wsized(f, cl_lmax)
        _wsized(f, np.array([6], dtype=np.int32))                              # ArraySizes = size(valArray)
    d = np.asarray(theory.get("derived", ()), dtype="<f8")
    f.write(struct.pack("<i", len(d))); f.write(d.tobytes())
    for i in range(1, s["num_cls"] + 1):
        for j in range(i, 0, -1):
            if cl_lmax[i - 1, j - 1] > 0:
                cl = np.asarray(theory["cls"][(i, j)], dtype="<f8")
                assert len(cl) == cl_lmax[i - 1, j - 1], ((i, j), len(cl))
                _wsized(f, cl)
    _wsized(f, np.array([theory.get(k, 0.0) for k in ("tensor_ratio_02", "tensor_ratio_C10", "tensor_ratio_BB",
                                                      "tensor_AT", "lensing_rms_deflect", "sigma_8")]))


def read_data_models(path, n_likes_expected=None):
    """Mirror of ParamSet_ReadModel (ParamSet.f90:75-150) + ReadTheory (CosmoTheory.f90:297-372): every model of a `.data`
    file as a list of dicts, plus the header."""
    import struct
    b = open(path, "rb").read()
    pos = [0]

    def rd(fmt):
        n = struct.calcsize(fmt)
        v = struct.unpack(fmt, b[pos[0]:pos[0] + n]); pos[0] += n
        return v

    def rstr():
        n, = rd("<i")
        s = b[pos[0]:pos[0] + n].decode(); pos[0] += n
        return s

    def rarr(dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        a = np.frombuffer(b[pos[0]:pos[0] + n], dtype=dtype).reshape(shape, order="F"); pos[0] += n
        return a.copy()
    fmt, npar = rd("<2i")
    if fmt != 4:
        raise ValueError("ReadModel: wrong file format (single-precision or old cosmomc version)")
    has_names, = rd("<i")
    names = [rstr() for _ in range(npar)] if has_names else None
    nl, = rd("<i")
    like_names = [rstr() for _ in range(nl)]
    unused, = rd("<i")
    if unused != 0:
        raise ValueError("ReadModel: don't know what the extra info is")
    header = dict(param_names=names, like_names=like_names)
    models, first = [], True
    while pos[0] < len(b):
        mult, like = rd("<2d")
        likes = rarr("<f8", (nl,))
        P = rarr("<f8", (npar,))
        if first:
            s = _unpack_theory_params(b[pos[0]:pos[0] + 160]); pos[0] += 160
            cl_lmax = None
            if s["use_CMB"]:
                n1, n2 = rd("<2i")
                cl_lmax = rarr("<i4", (n1, n2))
            nas, = rd("<i")
            array_sizes = rarr("<i4", (nas,))
            header.update(settings=s, cl_lmax=cl_lmax, array_sizes=array_sizes)
            first = False
        nd, = rd("<i")
        derived = rarr("<f8", (nd,))
        cls = {}
        for i in range(1, s["num_cls"] + 1):
            for j in range(i, 0, -1):
                if cl_lmax[i - 1, j - 1] > 0:
                    n, = rd("<i")
                    cls[(i, j)] = rarr("<f8", (n,))
        nv, = rd("<i")
        val = rarr("<f8", (nv,))
        models.append(dict(mult=mult, like=like, likelihoods=likes, params=P, derived=derived, cls=cls,
                           tensor_ratio_02=val[0], tensor_ratio_C10=val[1], tensor_ratio_BB=val[2], tensor_AT=val[3],
                           lensing_rms_deflect=val[4], sigma_8=val[5]))
    return header, models
