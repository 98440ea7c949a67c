"""Generate golden fixtures by IMPORTING the reference's own Python ports (run in the build container only;
/root/reference does not exist on the GPU box).  Outputs are committed under tests/golden/.

  python tests/golden/make_golden.py            # writes *.npz next to this file

Fixtures
  lensing_correlations_py.npz : camb/pycamb/camb/correlations.py lensed_cls() on the shipped unlensed
      fiducial camb/HighLExtrapTemplate_lenspotentialCls.dat (lmax=6000 input, lensed to 3000).
      Reference's own tolerance for this cross-check is 1e-3 (camb_test.py:178-182).
  lensing2018_cmblikes_py.npz : python/CMBlikes.py (port of source/CMBlikes.f90) evaluated on
      data/planck_lensing_2018 at data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl, calPlanck=1.00061,
      plus the dense arrays (bin windows, correction windows, fiducial correction, covariance) the
      dataset resolves to, so the GPU box can test without the reference tree.
"""
import os
import sys
import types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def golden_lensing():
    import scipy.special as sp
    if not hasattr(sp, "lpn"):
        def lpn(n, x):
            p = sp.legendre_p_all(n, x, diff_n=1)
            return p[0], p[1]
        sp.lpn = lpn
    sys.path.insert(0, os.path.join(REF, "camb/pycamb/camb"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("correlations", os.path.join(REF, "camb/pycamb/camb/correlations.py"))
    corr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(corr)
    a = np.loadtxt(os.path.join(REF, "camb/HighLExtrapTemplate_lenspotentialCls.dat"))
    lmax_in = 6000
    cls = np.zeros((lmax_in + 1, 4))
    clpp = np.zeros(lmax_in + 1)
    L = a[:, 0].astype(int)
    m = L <= lmax_in
    cls[L[m], 0] = a[m, 1]
    cls[L[m], 1] = a[m, 2]
    cls[L[m], 2] = a[m, 3]
    cls[L[m], 3] = a[m, 4]
    clpp[L[m]] = a[m, 5]
    out = corr.lensed_cls(cls, clpp, lmax_lensed=3000)
    # same input band limit as CorrFuncFullSky uses at Max_l=2650 (lmax_extrap=3300, lensing.f90:98-101)
    out3300 = corr.lensed_cls(cls[:3301], clpp[:3301], lmax_lensed=2600)
    np.savez_compressed(os.path.join(HERE, "lensing_correlations_py.npz"), lensed=out, lmax_in=lmax_in,
                        lensed_lmax3300=out3300)
    print("lensing golden:", out.shape, out[2], out[2000])


if __name__ == "__main__":
    which = sys.argv[1:] or ["lensing", "cmblikes"]
    if "lensing" in which:
        golden_lensing()
    if "cmblikes" in which:
        import make_golden_cmblikes
        make_golden_cmblikes.main()
