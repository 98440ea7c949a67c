"""Pack the reference's shipped DATA files that the hot path reads at run time (not source code) into
compact fixtures, so tests/bench on the GPU box (no /root/reference there) can use them.

  templates.npz :
    highl_unlensed [4][8001]  camb/HighLExtrapTemplate_lenspotentialCls.dat  TT,EE,TE,PP (modules.f90:1162-1185)
    highl_unlensed_BB, _TP, _EP [8001]  remaining columns of the same file
    highl_lensed  [4][lmax+1] data/HighL_lensedCls.dat TT,EE,BB,TE muK^2 (Calculator_CAMB.f90:398-402)
    theory_cl     [2509][5]   data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl  TT,TE,EE,BB,PP by l
"""
import os
import numpy as np
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

def main():
    a = np.loadtxt(os.path.join(REF, "camb/HighLExtrapTemplate_lenspotentialCls.dat"))
    L = a[:, 0].astype(int); m = L <= 8000
    un = np.zeros((4, 8001)); bb = np.zeros(8001); tp = np.zeros(8001); ep = np.zeros(8001)
    un[0, L[m]] = a[m, 1]; un[1, L[m]] = a[m, 2]; un[2, L[m]] = a[m, 4]; un[3, L[m]] = a[m, 5]
    bb[L[m]] = a[m, 3]; tp[L[m]] = a[m, 6]; ep[L[m]] = a[m, 7]
    b = np.loadtxt(os.path.join(REF, "data/HighL_lensedCls.dat"))
    Lb = b[:, 0].astype(int)
    le = np.zeros((4, Lb.max() + 1))
    for i in range(4):
        le[i, Lb] = b[:, i + 1]
    t = np.loadtxt(os.path.join(REF, "data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl"))
    Lt = t[:, 0].astype(int)
    th = np.zeros((Lt.max() + 1, 5)); th[Lt] = t[:, 1:6]
    np.savez_compressed(os.path.join(HERE, "templates.npz"), highl_unlensed=un, highl_unlensed_BB=bb,
                        highl_unlensed_TP=tp, highl_unlensed_EP=ep, highl_lensed=le, theory_cl=th)
    print(un.shape, le.shape, th.shape)

if __name__ == "__main__":
    main()
