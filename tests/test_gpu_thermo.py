"""GPU parity tests for the batched thermal history (cb200_thermo / cb200_theta_to_background, csrc/thermo.cuh),
through the C ABI: against the oracle (pinned to the reference's golden .minimum in tests/test_thermo_oracle.py) on a
spread of cosmologies, and directly against the golden derived-parameter block."""
import numpy as np
import pytest

import helpers as H  # noqa: F401
from test_thermo_oracle import BEST, GOLD, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from cosmomc_b200 import lib
    return lib.Handle(max_points=8, lmax_computed_cl=0)   # background-only handle: no CMB tables needed


def points():
    from cosmomc_b200 import params as P
    rng = np.random.default_rng(5)
    rows, yhe, tau = [P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])], [BEST["yhe"]], [BEST["tau"]]
    for i in range(5):
        rows.append(P.cmb_to_background(BEST["ombh2"] * (1 + 0.03 * rng.standard_normal()),
                                        BEST["omch2"] * (1 + 0.05 * rng.standard_normal()),
                                        BEST["H0"] * (1 + 0.04 * rng.standard_normal()),
                                        mnu=[0.06, 0.0, 0.3, 0.06, 0.12][i]))
        yhe.append(0.245 + 0.005 * rng.standard_normal())
        tau.append(max(0.02, 0.055 + 0.015 * rng.standard_normal()))
    return np.array(rows), np.array(yhe), np.array(tau)


def test_thermo_matches_golden_minimum_directly(h):
    from cosmomc_b200 import params as P
    bg = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])
    out, st = h.thermo(bg, BEST["yhe"], optical_depth=BEST["tau"])
    assert st[0] == 0 and out[0, 11] == 0
    assert abs(out[0, 7] / GOLD["zrei"] - 1) < TOL
    for i, k in enumerate(h.THERMO_DERIVED):
        assert abs(out[0, 12 + i] / GOLD[k] - 1) < TOL, (k, out[0, 12 + i], GOLD[k])


def test_thermo_matches_oracle_on_a_spread_of_points(h):
    import pyoracle as o
    bg, yhe, tau = points()
    out, st = h.thermo(bg, yhe, optical_depth=tau)
    assert np.all(st == 0)
    worst = 0.0
    for i in range(len(bg)):
        r = o.thermo(bg[i], yhe[i], optical_depth=tau[i])
        ref = np.array([r[k] for k in ("tau0", "taurst", "taurend", "tau_start", "tau_complete", "dtaurec", "tau_maxvis",
                                       "zre", "z_star", "z_drag", "actual_opt_depth")] + list(r["derived"].values()))
        got = np.concatenate([out[i, :11], out[i, 12:25]])
        worst = max(worst, np.abs(got / ref - 1).max())
    # the integrator's accept / reject decisions sit on libm-vs-CUDA last-bit differences: agreement is at the level of
    # dverk's own tolerance (1e-5 per unit step), far inside the 7 figures of the golden block
    assert worst < 2e-6, worst


def test_thermo_no_reionisation_and_fixed_redshift(h):
    import pyoracle as o
    from cosmomc_b200 import params as P
    bg = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])
    out, st = h.thermo(np.stack([bg, bg]), BEST["yhe"], zre=np.array([0.0, 9.5]))
    assert np.all(st == 0)
    assert out[0, 7] == 0 and out[0, 3] == out[0, 0] and out[0, 10] == 0          # reionisation switched off
    r = o.thermo(bg, BEST["yhe"], zre=9.5)
    assert abs(out[1, 7] - 9.5) < 1e-15 and abs(out[1, 3] / r["tau_start"] - 1) < 1e-9
    assert abs(out[1, 10] / r["actual_opt_depth"] - 1) < 2e-6


def test_theta_to_background_matches_reference_h0(h):
    import pyoracle as o
    from cosmomc_b200 import params as P
    bg0 = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], BEST["H0"])
    omnuh2 = bg0[3] * (BEST["H0"] / 100) ** 2
    th = np.array([BEST["theta100"], 1.0385, 2.0])
    bg = h.theta_to_background(np.full(3, BEST["ombh2"]), BEST["omch2"], th, omnuh2, bg0[7:15], rdrag=147.0)
    assert abs(bg[0, 0] / BEST["H0"] - 1) < 5e-6                 # golden H0 (theta rounded to 7 figures)
    want = o.h0_from_theta(1.0385, lambda x: P.cmb_to_background(BEST["ombh2"], BEST["omch2"], x))
    assert abs(bg[1, 0] / want - 1) < 1e-9
    assert np.all(bg[2, :15] == 0) and bg[2, 15] == 147.0        # theta out of range: H0 = 0, the point is rejected
    # the solved row is a complete bg row: same densities as the host mapping at that H0
    ref = P.cmb_to_background(BEST["ombh2"], BEST["omch2"], bg[0, 0], rdrag=147.0)
    assert np.allclose(bg[0], ref, rtol=1e-13)
