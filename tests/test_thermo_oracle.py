"""CPU tests (no GPU): pin the ORACLE's thermal history (RECFAST, reionisation, inithermo, z_star / z_drag,
the theta -> H0 bisection) to numbers the REFERENCE produced: the derived-parameter block of
data/base_plikHM_TTTEEE_lowl_lowE.minimum (7 significant figures; inputs = the same file's parameters 1-4 and Y_He).
With this, GPU-vs-oracle parity of cb200_thermo means GPU-vs-reference parity (tests/test_gpu_thermo.py)."""
import numpy as np
import pytest

import helpers as H  # noqa: F401  (sets sys.path)

# data/base_plikHM_TTTEEE_lowl_lowE.minimum: parameters 1-4, derived 72, 85, 94, 97-109
BEST = dict(ombh2=0.2237737E-01, omch2=0.1201035E+00, theta100=0.1040920E+01, tau=0.5430138E-01, H0=0.6732178E+02,
            yhe=0.2453985E+00)
GOLD = dict(zrei=0.7679749E+01, age=0.1379731E+02, zstar=0.1089920E+04, rstar=0.1443990E+03, thetastar=0.1041097E+01,
            DAstar=0.1386989E+02, zdrag=0.1059971E+04, rdrag=0.1470552E+03, kd=0.1409104E+00, thetad=0.1607437E+00,
            zeq=0.3404856E+04, keq=0.1039196E-01, thetaeq=0.8128424E+00, thetarseq=0.4491390E+00)
# tolerance: the file rounds inputs and outputs to 7 figures (5e-8 each); kd / thetad are the most sensitive to the inputs
TOL = 2e-6


@pytest.fixture(scope="module")
def o():
    import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def best(o):
    from cosmomc_b200 import params as P
    bg = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])
    return bg, o.thermo(bg, BEST["yhe"], optical_depth=BEST["tau"], tables=True)


def test_derived_block_matches_the_reference_minimum(best):
    bg, r = best
    assert r["status"] == 0
    assert abs(r["zre"] / GOLD["zrei"] - 1) < TOL           # Reionization_zreFromOptDepth
    for k, v in r["derived"].items():
        assert abs(v / GOLD[k] - 1) < TOL, (k, v, GOLD[k])
    assert r["derived"]["zstar"] == r["z_star"] and r["derived"]["zdrag"] == r["z_drag"]


def test_time_grid_scalars_are_consistent(best, o):
    """The five scalars the projection takes per point (tau0, taurst, taurend, reionisation start / end) and the
    recombination time step feed SetTimeSteps (modules.f90:2994-3027): ordered, and the grid they give has the
    size SURVEY 8 quotes for this cosmology (n_tau ~ 600)."""
    bg, r = best
    assert 0 < r["taurst"] < r["tau_maxvis"] < r["taurend"] <= r["tau_start"] < r["tau_complete"] < r["tau0"]
    assert abs(r["dtaurec"] - min(4 / (14000 / r["tau0"]), r["taurst"] / 40)) < 1e-12
    tau, dtau = o.time_steps(r["taurst"], r["taurend"], r["tau0"], 14000, False, r["tau_start"], r["tau_complete"])
    assert 500 < len(tau) < 700
    # optical depth actually reached by the tanh model: the input tau to the bisection's tolerance
    assert abs(r["actual_opt_depth"] - BEST["tau"]) < 2e-4


def test_ionisation_history_limits(best, o):
    bg, r = best
    xe, dotmu, emmu, cs2 = r["tables"]
    fHe = BEST["yhe"] / (3.9715 * (1 - BEST["yhe"]))
    assert abs(xe[0] - (1 + 0.5 * BEST["yhe"] / (1 - BEST["yhe"]))) < 1e-12   # xe(1) of inithermo (mass ratio 4)
    assert abs(xe[1] - (1 + 2 * fHe)) < 1e-12                     # fully ionised H and He at the start (RECFAST)
    assert abs(xe[-1] - (1 + 2 * fHe)) < 1e-6                     # reionised, helium doubly ionised, today
    assert 1e-4 < xe.min() < 5e-4                                 # freeze-out residual before reionisation
    assert emmu[-1] == 1.0 and np.all(np.diff(emmu) >= 0)         # e^{-tau} grows monotonically to one
    # Recombination_xe interpolates the RECFAST table: Saha limits at high z, values within (0, 1 + 2 fHe]
    x = o.recfast_xe(bg, BEST["yhe"], 1 / (1 + np.array([9000.0, 6000.0, 4000.0, 1100.0, 500.0, 0.0])))
    assert abs(x[0] - (1 + 2 * fHe)) < 1e-12 and abs(x[2] - (1 + fHe)) < 1e-12 and 0.05 < x[3] < 0.3 and x[5] < 5e-4
    assert np.all(np.diff(x) < 1e-15)


def test_theta_to_h0_bisection(o):
    from cosmomc_b200 import params as P
    H0 = o.h0_from_theta(BEST["theta100"], lambda h: P.cmb_to_background(BEST["ombh2"], BEST["omch2"], h))
    assert abs(H0 / BEST["H0"] - 1) < 5e-6       # theta is given to 7 figures: dH0/H0 ~ 3 dtheta/theta
    assert o.h0_from_theta(2.0, lambda h: P.cmb_to_background(BEST["ombh2"], BEST["omch2"], h)) == 0.0


def test_no_reionisation_branch(o):
    from cosmomc_b200 import params as P
    bg = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])
    r = o.thermo(bg, BEST["yhe"], zre=0.0)       # Reion%redshift < 0.001 switches reionisation off
    assert r["status"] == 0 and r["zre"] == 0 and r["tau_start"] == r["tau0"] and r["actual_opt_depth"] == 0
    assert abs(r["derived"]["rdrag"] / GOLD["rdrag"] - 1) < TOL   # r_drag ignores reionisation
