"""Golden fixture for the supernova likelihood, produced by the reference's own Python port
python/planck/SN.py (port of source/supernovae_JLA.f90).  Build-container only (/root/reference needed).

The covariance blobs (six JLA blocks, Pantheon sys_full_long.txt) are MISSING from the reference checkout
(.MISSING_LARGE_BLOBS), so SN.py is run on the REAL light-curve tables with the documented synthetic stand-in
covariances of cosmomc_b200/synthetic.py (seed 2024) written to temporary files in the format SN.py reads.
Distances: the polynomial fit(z) of SN.py's own self-test (python/planck/SN.py:303-304).

Writes tests/golden/sn_py.npz: alpha, beta, -lnL for JLA (alpha,beta varied; two-scriptM) and Pantheon.
"""
import os
import shutil
import sys
import tempfile
import types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_port():
    for name in ["matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    for al, t in (("int", int), ("float", float), ("bool", bool), ("object", object)):
        if not hasattr(np, al):
            setattr(np, al, t)
    sys.path.insert(0, os.path.join(REF, "python"))
    sys.path.insert(0, os.path.join(REF, "python", "planck"))
    import SN
    return SN


def fit(z):
    return -338.65487197 * z ** 4 + 1972.59141641 * z ** 3 - 4310.60442428 * z ** 2 + 4357.72542145 * z


def main():
    from cosmomc_b200 import synthetic as syn
    SN = load_port()
    tmp = tempfile.mkdtemp()
    out = {}
    try:
        # ---- JLA
        shutil.copy(os.path.join(REF, "data/jla.dataset"), tmp)
        shutil.copy(os.path.join(REF, "data/jla_lcparams.txt"), tmp)
        z = np.loadtxt(os.path.join(tmp, "jla_lcparams.txt"), usecols=1)
        covs = syn.synthetic_sn_covs({"zcmb": z})
        fn = {"mag": "jla_v0_covmatrix.dat", "stretch": "jla_va_covmatrix.dat", "colour": "jla_vb_covmatrix.dat",
              "mag_stretch": "jla_v0a_covmatrix.dat", "mag_colour": "jla_v0b_covmatrix.dat",
              "stretch_colour": "jla_vab_covmatrix.dat"}
        for k, f in fn.items():
            np.savetxt(os.path.join(tmp, f), covs[k].reshape(-1), fmt="%.17e")
        like = SN.SN_likelihood(os.path.join(tmp, "jla.dataset"), marginalize=False, silent=True)
        zs = like.get_redshifts()
        ab = [(0.1325237, 2.959805), (0.14, 3.1), (0.11, 3.3)]
        out["jla_ab"] = np.array(ab)
        out["jla_lnl"] = np.array([like.loglike(fit(zs), {"alpha": a, "beta": b}) for a, b in ab])
        # ---- Pantheon
        os.makedirs(os.path.join(tmp, "P"))
        shutil.copy(os.path.join(REF, "data/Pantheon/full_long.dataset"), os.path.join(tmp, "P"))
        shutil.copy(os.path.join(REF, "data/Pantheon/lcparam_full_long_zhel.txt"), os.path.join(tmp, "P"))
        zp = np.loadtxt(os.path.join(tmp, "P", "lcparam_full_long_zhel.txt"), usecols=1)
        cp = syn.synthetic_sn_covs({"zcmb": zp}, names=("mag",), seed=2025)
        np.savetxt(os.path.join(tmp, "P", "sys_full_long.txt"), cp["mag"].reshape(-1), fmt="%.17e")
        likep = SN.SN_likelihood(os.path.join(tmp, "P", "full_long.dataset"), silent=True)
        out["pantheon_lnl"] = np.array([likep.loglike(fit(likep.get_redshifts()))])
        out["pantheon_lnl_scaled"] = np.array([likep.loglike(1.02 * fit(likep.get_redshifts()) + 3.0)])
    finally:
        shutil.rmtree(tmp)
    print({k: v for k, v in out.items()})
    np.savez(os.path.join(HERE, "sn_py.npz"), **out)


if __name__ == "__main__":
    main()
