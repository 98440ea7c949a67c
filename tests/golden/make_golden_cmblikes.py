"""Golden fixture for the binned CMBLikes path, produced by the reference's own Python port
python/CMBlikes.py (port of source/CMBlikes.f90).  Build-container only (/root/reference needed).

Writes tests/golden/lensing2018_cmblikes_py.npz holding
  chi2          chi_squared at data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl, calPlanck = 1.00061
  binned_theory the 9 binned PP band powers (after linear correction) at that point
  plus every dense array the .dataset resolves to (windows, correction windows, fiducial correction,
  inverse covariance, band powers), so tests on the GPU box need neither the reference tree nor the port.
"""
import os
import sys
import types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


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
    CMBlikes = load_port()
    ds = os.path.join(REF, "data/planck_lensing_2018/smicadx12_Dec5_ftl_mv2_ndclpp_p_teb_consext8.dataset")
    like = CMBlikes.DatasetLikelihood(ds)
    cls = CMBlikes.ClsArray(os.path.join(REF, "data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl"))
    params = {"calPlanck": 1.00061}
    chi2, binned = like.chi_squared(cls, params, return_binned_theory=True)
    print("chi2 =", repr(chi2), "binned", binned.ravel())
    out = dict(chi2=chi2, binned_theory=binned, calPlanck=1.00061,
               pcl_lmin=like.pcl_lmin, pcl_lmax=like.pcl_lmax, nbins=like.nbins_used, nmaps=like.nmaps,
               nmaps_required=like.nmaps_required, ncl_used=like.ncl_used,
               cl_used_index=np.asarray(like.cl_used_index), covinv=like.covinv,
               bandpowers=np.asarray(like.bandpower_matrix), log_calibration_prior=like.log_calibration_prior,
               bins_cols_in=like.bins.cols_in, bins_cols_out=np.asarray(like.bins.cols_out),
               bins_matrix=like.bins.binning_matrix,
               corr_cols_in=like.linear_correction.cols_in, corr_cols_out=np.asarray(like.linear_correction.cols_out),
               corr_matrix=like.linear_correction.binning_matrix, fid_correction=like.fid_correction,
               required_order=np.asarray(like.required_order), map_fields=np.asarray(like.map_fields))
    np.savez_compressed(os.path.join(HERE, "lensing2018_cmblikes_py.npz"), **out)


if __name__ == "__main__":
    main()
