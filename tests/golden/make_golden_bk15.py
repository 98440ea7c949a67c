"""BK15 fixture + golden value (build-container only: /root/reference needed).

1. Resolves data/BK15/BK15_dust.dataset with batch3/BK15.ini's maps_use (12 B-mode maps, 9 bins) into the compact pack
   tests/golden/bk15_pack.npz (band powers, noise, fiducial square roots, band-power windows of the 78 used spectra,
   bandpasses) so tests / bench on the GPU box need no reference tree.
2. data/BK15/BK15_covmat_dust.dat is MISSING from the reference checkout (.MISSING_LARGE_BLOBS): a documented synthetic
   stand-in (cosmomc_b200.synthetic.synthetic_bk15_cov, seed 15) over all 9 x 300 band powers is used, written to a
   temporary file in the format the reference reads.
3. Golden value: the reference's own Python port python/CMBlikes.py (port of source/CMBlikes.f90, Hamimeche-Lewis
   branch) evaluates chi^2 of that data set at the golden Planck best-fit C_l (no foregrounds: the port has no BK
   foreground model).  Stored as `port_chi2_nofg`; pins binning, noise, HL transform, vecp order, covariance selection.
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
MAPS_USE = "BK15_95_B BK15_150_B BK15_220_B W023_B P030_B W033_B P044_B P070_B P100_B P143_B P217_B P353_B"


def load_port():
    for name in ["matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for al, t in (("int", int), ("float", float), ("bool", bool), ("object", object)):
        if not hasattr(np, al):
            setattr(np, al, t)
    sys.path.insert(0, os.path.join(REF, "python"))
    import CMBlikes
    return CMBlikes


def main():
    from cosmomc_b200 import datasets as D, synthetic as syn
    ds = os.path.join(REF, "data/BK15/BK15_dust.dataset")
    fid = np.loadtxt(os.path.join(REF, "data/BK15/BK15_fiducial_dust.dat"))[:, 1:]
    noise = np.loadtxt(os.path.join(REF, "data/BK15/BK15_noise.dat"))[:, 1:]
    prov = syn.synthetic_bk15_cov(np.abs(fid) + np.abs(noise))     # [9 bins][300 spectra] -> 2700 entries, bin-major
    plan = D.BK15Plan(ds, overrides={"maps_use": MAPS_USE, "use_min": "1", "use_max": "9"}, cov_provider=prov)
    print("nmaps", plan.nmaps, "ncl_used", plan.ncl_used, "bins", plan.nbins_used, "HL", plan.like_approx)
    pack = os.path.join(HERE, "bk15_pack.npz")
    plan.save_pack(pack)
    # ---- reference python port on the same data, synthetic covariance written as the missing file
    CMBlikes = load_port()
    tmp = tempfile.mkdtemp()
    try:
        dst = os.path.join(tmp, "BK15")
        shutil.copytree(os.path.join(REF, "data/BK15"), dst)
        np.savetxt(os.path.join(dst, "BK15_covmat_dust.dat"), prov.full(), fmt="%.17e")
        like = CMBlikes.DatasetLikelihood(os.path.join(dst, "BK15_dust.dataset"),
                                          {"maps_use": MAPS_USE, "use_min": 1, "use_max": 9})
        cls = CMBlikes.ClsArray(os.path.join(REF, "data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl"))
        chi2 = like.chi_squared(cls, {})
    finally:
        shutil.rmtree(tmp)
    print("port chi2 (no foregrounds) =", repr(chi2))
    z = dict(np.load(pack))
    z["port_chi2_nofg"] = np.float64(chi2)
    np.savez_compressed(pack, **z)
    print("pack size", os.path.getsize(pack) / 1e6, "MB")


if __name__ == "__main__":
    main()
