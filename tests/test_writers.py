"""CPU tests: the `.minimum` writer and the derived-parameter block against the reference's golden
data/base_plikHM_TTTEEE_lowl_lowE.minimum (fixture copy under tests/golden/data)."""
import os

import numpy as np

import helpers as H
from cosmomc_b200 import mcmc, params as P

GOLD = os.path.join(H.ROOT, "tests", "golden", "data", "base_plikHM_TTTEEE_lowl_lowE.minimum")


def test_minimum_writer_reproduces_the_golden_file_byte_for_byte(tmp_path):
    loglike, rows, contribs = mcmc.read_minimum(GOLD)
    base = sorted([r for r in rows if r[4] < 2])
    derived = [r for r in rows if r[4] == 2]
    assert [r[0] for r in base] == list(range(1, len(base) + 1)) and derived[0][0] == len(base) + 1
    out = tmp_path / "x.minimum"
    likes = []
    g = {r[2]: r[1] for r in rows}
    for v, typ, rest in contribs:
        tag, name = rest.split(" = ", 1)
        # the printed -lnL carries three decimals; the chi2_<tag> derived column of the same file has seven figures
        # (both are roundings of the same number: take the seven-figure one, kept inside the printed value's interval)
        likes.append((v + float(np.clip(g["chi2_" + tag] / 2 - v, -4.9e-4, 4.9e-4)), typ, tag, name, ""))
    mcmc.write_minimum(out, loglike, [r[1] for r in base], [r[4] == 0 for r in base], [r[2] for r in base],
                       [r[3] for r in base], [r[1] for r in derived], [r[2] for r in derived], [r[3] for r in derived], likes)
    got, want = open(out, "rb").read(), open(GOLD, "rb").read()
    assert got == want


def test_list_directed_real_matches_the_golden_header():
    assert mcmc._list_directed_real(1382.88563166157) == "   1382.88563166157     "
    assert mcmc._list_directed_real(2765.77126332314) == "   2765.77126332314     "
    assert mcmc._list_directed_real(0.5).strip() == "0.500000000000000" and len(mcmc._list_directed_real(-12.25)) == 24


def test_derived_block_order_and_values_match_the_golden_minimum():
    import pyoracle as o
    _, rows, _ = mcmc.read_minimum(GOLD)
    g = {r[2]: r[1] for r in rows}
    bg = P.cmb_to_background(g["omegabh2"], g["omegach2"], g["H0"])
    th = o.thermo(bg, g["yheused"], optical_depth=g["tau"])
    h = g["H0"] / 100
    omnuh2 = bg[3] * h * h
    cmb = dict(H0=g["H0"], h=h, omv=bg[4], omb=bg[1], omdm=bg[2] + bg[3], ombh2=g["omegabh2"], omdmh2=g["omegach2"] + omnuh2,
               omnuh2=omnuh2, zre=th["zre"], tau=g["tau"], logA=g["logA"], ns=g["ns"], yhe=g["yheused"])
    cl = np.zeros(2001)
    for L, k in zip(mcmc.DERIVED_CL, ("DL40", "DL220", "DL810", "DL1420", "DL2000")):
        cl[L] = g[k]
    bgout = [g[k] for k in ("Hubble015", "DM015", "Hubble038", "DM038", "Hubble051", "DM051", "Hubble061", "DM061",
                            "Hubble233", "DM233")]
    d = mcmc.calc_derived_params(cmb, list(th["derived"].values()), g["rmsdeflect"], cl_TT=cl, sigma8=g["sigma8"],
                                 bbn_dh=g["DHBBN"], background_outputs=bgout)
    names = [r[2] for r in rows if r[4] == 2]
    want = np.array([g[n] for n in names[:len(d)]])
    assert names[:3] == ["H0", "omegal", "omegam"] and names[len(d) - 1] == "DM233"
    assert np.abs(d / want - 1).max() < 2e-6, dict(zip(names, d / want - 1))


def test_inputparams_writer_lists_the_keys_read_in_order(tmp_path):
    """`<root>.inputparams` = Ini%SaveReadValues (source/IniObjects.f90:870-884): `name = value` per key read, defaults
    included, first-read order; keys never read are not listed."""
    from cosmomc_b200 import datasets as ds
    p = tmp_path / "a.ini"
    p.write_text("x = 3\nunused = 1\ny = hello # comment\n")
    ini = ds.IniFile(str(p))
    assert ini.int("x") == 3 and ini.string("z", "dflt") == "dflt" and ini.string("y") == "hello" and ini.int("x") == 3
    out = tmp_path / "run.inputparams"
    ini.save_read_values(str(out))
    assert out.read_text() == "x = 3\nz = dflt\ny = hello\n"


def test_data_file_layout_and_round_trip(tmp_path):
    """`.data` binary (ParamSet_WriteModel, ParamSet.f90:32-73 + WriteTheory, CosmoTheory.f90:235-282): header fields at the
    byte offsets a Fortran stream unit gives them, the 160-byte TCosmoTheoryParams image, and a round trip through the
    mirror of ReadModel / ReadTheory (no `.data` file ships with the reference: unpinned beyond this)."""
    import struct
    rng = np.random.default_rng(5)
    cl_lmax = np.zeros((4, 4), dtype=np.int32)
    cl_lmax[0, 0] = 2508; cl_lmax[1, 0] = 2508; cl_lmax[1, 1] = 2508; cl_lmax[3, 3] = 2500   # TT, ET, EE, PhiPhi
    names = ["omegabh2", "omegach2", "theta", "tau", "logA", "ns"]
    likes = ["lensing", "BAO"]
    models = []
    path = tmp_path / "chain_1.data"
    with open(path, "wb") as f:
        for m in range(3):
            th = dict(derived=rng.normal(size=44),
                      cls={(i + 1, j + 1): rng.normal(size=cl_lmax[i, j]) for i in range(4) for j in range(i + 1) if cl_lmax[i, j] > 0},
                      lensing_rms_deflect=2.5e-3 + m, sigma_8=0.81 + m)
            P, L = rng.normal(size=6), rng.normal(size=2)
            mcmc.write_data_model(f, m == 0, 1.0 + m, 1382.9 + m, L, P, th, param_names=names, like_names=likes, cl_lmax=cl_lmax)
            models.append((P, L, th))
    b = path.read_bytes()
    assert struct.unpack("<3i", b[:12]) == (4, 6, 1)                          # format 4 (double), num_params_used, has names
    assert struct.unpack("<i", b[12:16]) == (8,) and b[16:24] == b"omegabh2"  # WriteTrim: length then characters
    assert struct.calcsize(mcmc._THEORY_PARAMS_FMT) == 160
    s = mcmc._unpack_theory_params(mcmc._pack_theory_params(mcmc.THEORY_PARAMS_DEFAULT))
    assert s["z_outputs"] == (0.15, 0.38, 0.51, 0.61, 2.33) and s["pivot_k"] == 0.05 and s["num_massive_neutrinos"] == -1
    # power_kmax sits on the 8-byte boundary after three logicals + 4 bytes of padding (offset 112)
    assert struct.unpack("<d", mcmc._pack_theory_params(mcmc.THEORY_PARAMS_DEFAULT)[112:120]) == (0.8,)
    hdr, got = mcmc.read_data_models(str(path))
    assert hdr["param_names"] == names and hdr["like_names"] == likes and hdr["array_sizes"].tolist() == [6]
    assert np.array_equal(hdr["cl_lmax"], cl_lmax) and hdr["settings"]["lmax_computed_cl"] == 2500
    assert len(got) == 3
    for m, (P, L, th) in enumerate(models):
        g = got[m]
        assert g["mult"] == 1.0 + m and g["like"] == 1382.9 + m
        assert np.array_equal(g["params"], P) and np.array_equal(g["likelihoods"], L) and np.array_equal(g["derived"], th["derived"])
        assert list(g["cls"].keys()) == [(1, 1), (2, 2), (2, 1), (4, 4)]      # i = 1..4, j = i..1
        for k in th["cls"]:
            assert np.array_equal(g["cls"][k], th["cls"][k])
        assert g["sigma_8"] == th["sigma_8"] and g["lensing_rms_deflect"] == th["lensing_rms_deflect"] and g["tensor_AT"] == 0
