"""Host-side mirror of the reference's parameter plumbing for the background functions.

CosmoMC turns the sampled parameters into `CMBParams` (source/CosmologyParameterizations.f90:283-348 SetForH),
`CAMBCalc_CMBToCAMB` copies them into `CAMBparams` and resolves the neutrino hierarchy
(source/Calculator_CAMB.f90:84-129, camb/camb.f90:377-464 CAMB_SetNeutrinoHierarchy), and `CAMBParams_Set`
(camb/modules.f90:300-375) turns that into densities.  The C ABI takes the result of those three steps as a flat
`bg[16]` vector per parameter point (include/cosmob200.h, cb200_set_background):

    H0, omegab, omegac, omegan, omegav, w, tcmb, nu_massless_degeneracy, n_eigenstates,
    nu_mass_degeneracies[3], nu_mass_fractions[3], rdrag

Scalar bookkeeping only; all numerical work (Romberg integrals, distances, likelihoods) runs on the GPU.
"""
import math

import numpy as np

NBG = 16
neutrino_mass_fac = 94.07      # camb/modules.f90:1493
default_nnu = 3.046            # camb/constants.f90:57
delta_mnu21 = 7.54e-5          # camb/constants.f90:59
delta_mnu31 = 2.46e-3
mnu_min_normal = 0.06          # camb/constants.f90:62
COBE_CMBTemp = 2.7255
HIERARCHY = {"normal": 1, "inverted": 2, "degenerate": 3}


def _sum_mnu_for_m1(m1, targ, sgn):
    m2 = math.sqrt(m1 ** 2 + delta_mnu21)
    m3 = math.sqrt(m1 ** 2 + sgn * delta_mnu31)
    return m1 + m2 + m3 - targ, m1 / m2 + m1 / m3 + 1


def _newton_raphson(xxl, xxh, targ, sgn):
    """camb/subroutines.f90:1130-1183 (bracketed Newton-Raphson with bisection safeguard)."""
    xl, xh = xxl, xxh
    f, _ = _sum_mnu_for_m1(xl, targ, sgn)
    f2, _ = _sum_mnu_for_m1(xh, targ, sgn)
    if f * f2 > 0:
        raise ValueError("Newton_Raphson: root is not bracketed")
    if f > 0:
        xl, xh = xh, xl
    error = abs(xh - xl)
    xm = 0.5 * (xl + xh)
    k = 0
    while error > 1e-8 and k < 1000:
        k += 1
        f, df = _sum_mnu_for_m1(xm, targ, sgn)
        if f > 0:
            xh = xm
        else:
            xl = xm
        xn = xm - f / df
        if (xn - xl) * (xn - xh) > 0:
            xm = 0.5 * (xh + xl)
        else:
            xm = xn
        error = abs(xh - xl)
    return xm


def neutrino_hierarchy(omnuh2, omnuh2_sterile, nnu, hierarchy="normal", num_massive_neutrinos=3):
    """CAMB_SetNeutrinoHierarchy followed by the neutrino block of CAMBParams_Set (share_delta_neff = F).
    Returns (nu_massless_degeneracy, n_eigenstates, degeneracies[3], fractions[3])."""
    deg = [0.0, 0.0, 0.0]
    frac = [0.0, 0.0, 0.0]
    num_massless = nnu
    num_massive = 0
    n_eig = 0
    if omnuh2 == 0:
        return num_massless, 0, deg, frac
    hid = HIERARCHY[hierarchy] if isinstance(hierarchy, str) else int(hierarchy)
    if omnuh2 > omnuh2_sterile:
        normal_frac = (omnuh2 - omnuh2_sterile) / omnuh2
        if hid == 3:
            neff_massive_standard = num_massive_neutrinos * default_nnu / 3
            num_massive = num_massive_neutrinos
            n_eig = 1
            if nnu > neff_massive_standard:
                num_massless = nnu - neff_massive_standard
            else:
                num_massless = 0
                neff_massive_standard = nnu
            deg[0] = neff_massive_standard
            frac[0] = normal_frac
        else:
            mnu = (omnuh2 - omnuh2_sterile) * neutrino_mass_fac / (default_nnu / 3) ** 0.75
            m1 = 0.0
            if hid == 1:
                if mnu > mnu_min_normal + 1e-4:
                    m1 = _newton_raphson(0.0, mnu, mnu, 1.0)
                    num_massive = 3
                else:
                    num_massive = 1
            else:
                if mnu > math.sqrt(delta_mnu31) + math.sqrt(delta_mnu31 + delta_mnu21) + 1e-4:
                    m1 = _newton_raphson(math.sqrt(delta_mnu31), mnu, mnu, -1.0)
                    num_massive = 3
                else:
                    num_massive = 2
            neff_massive_standard = num_massive * default_nnu / 3
            if nnu > neff_massive_standard:
                num_massless = nnu - neff_massive_standard
            else:
                num_massless = 0
                neff_massive_standard = nnu
            if num_massive == 3:
                n_eig = 2
                deg[0] = neff_massive_standard * 2 / 3.0
                deg[1] = neff_massive_standard * 1 / 3.0
                m3 = mnu - 2 * m1
                frac[0] = 2 * m1 / mnu * normal_frac
                frac[1] = m3 / mnu * normal_frac
            else:
                deg[0] = neff_massive_standard
                n_eig = 1
                frac[0] = normal_frac
    else:
        neff_massive_standard = 0
    if omnuh2_sterile > 0:
        if nnu < default_nnu:
            raise ValueError("nnu < 3.046 with massive sterile")
        num_massless = default_nnu - neff_massive_standard
        num_massive += 1
        deg[n_eig] = max(1e-6, nnu - default_nnu)
        frac[n_eig] = omnuh2_sterile / omnuh2
        n_eig += 1
    return num_massless, n_eig, deg, frac


def cmb_to_background(ombh2, omch2, H0, omk=0.0, mnu=0.06, nnu=3.046, w=-1.0, meffsterile=0.0, hierarchy="normal",
                      num_massive_neutrinos=3, rdrag=0.0, tcmb=COBE_CMBTemp):
    """SetForH (theta parameterisation, H0 already solved) -> CMBToCAMB -> CAMBParams_Set densities: one bg[16] row."""
    hid = HIERARCHY[hierarchy] if isinstance(hierarchy, str) else int(hierarchy)
    if nnu > default_nnu or hid != 3:
        omnuh2 = mnu / neutrino_mass_fac * (default_nnu / 3) ** 0.75
    else:
        omnuh2 = mnu / neutrino_mass_fac * (nnu / 3) ** 0.75
    omnuh2_sterile = meffsterile / neutrino_mass_fac
    omnuh2 = omnuh2 + omnuh2_sterile
    omdmh2 = omch2 + omnuh2
    h2 = (H0 / 100) ** 2
    omb, omc, omnu, omdm = ombh2 / h2, omch2 / h2, omnuh2 / h2, omdmh2 / h2
    omv = 1 - omk - omb - omdm
    massless, n_eig, deg, frac = neutrino_hierarchy(omnuh2, omnuh2_sterile, nnu, hid, num_massive_neutrinos)
    if omnu == 0:
        n_eig = 0
    return np.array([H0, omb, omc, omnu, omv, w, tcmb, massless, n_eig] + deg + frac + [rdrag], dtype=np.float64)


def background_batch(ombh2, omch2, H0, rdrag=None, **kw):
    """Vectorised front end: arrays of (ombh2, omch2, H0[, rdrag]) -> bg [npts][16]."""
    ombh2, omch2, H0 = np.broadcast_arrays(np.atleast_1d(ombh2), np.atleast_1d(omch2), np.atleast_1d(H0))
    out = np.zeros((len(H0), NBG))
    for i in range(len(H0)):
        out[i] = cmb_to_background(ombh2[i], omch2[i], H0[i], rdrag=0.0 if rdrag is None else np.atleast_1d(rdrag)[i], **kw)
    return out
