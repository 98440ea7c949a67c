"""CPU tests of the chain-statistics exchange: single-process logic against a direct numpy
evaluation of the reference formulas, and the world_size-2 gloo all-gather path."""
import os
import subprocess
import sys

import numpy as np

import helpers as H


def _chains(K, n, count, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    C = A @ A.T + n * np.eye(n)
    L = np.linalg.cholesky(C)
    return [(rng.normal(size=(count, n)) @ L.T + 0.05 * k) for k in range(K)], C


def test_chain_record_matches_reference_index_range():
    from cosmomc_b200 import chains
    s = np.arange(22, dtype=float).reshape(11, 2)      # Count = 10 -> items 5..10, m0 = 10 - 5 + 1 = 6
    m0, mean, cov = chains.chain_record(s)
    assert m0 == 6 and np.allclose(mean, s[5:11].mean(axis=0))
    assert np.allclose(cov, np.cov(s[5:11].T, bias=True))


def test_pooled_statistics_and_R():
    from cosmomc_b200 import chains
    K, n = 4, 5
    ch, C = _chains(K, n, 4001, seed=2)
    recs = [chains.chain_record(c) for c in ch]
    g = chains.pack(recs, n)
    st = chains.pooled_statistics(g, n)
    assert st["ready"] and st["cov"].shape == (n, n)
    assert np.abs(st["cov"] - C).max() < 0.15 * np.abs(C).max()
    # direct evaluation of SampleCollector.f90:262-277 + samples.f90:41-67
    m0 = np.array([r[0] for r in recs]); means = np.array([r[1] for r in recs]); covs = np.array([r[2] for r in recs])
    norm = m0.sum(); mean = (means * m0[:, None]).sum(0) / norm
    cov = covs.mean(0)
    mc = sum(m0[k] * np.outer(means[k] - mean, means[k] - mean) for k in range(K)) / norm * K / (K - 1)
    sc = np.sqrt(np.diag(cov)); rot = cov / np.outer(sc, sc); rm = mc / np.outer(sc, sc)
    Li = np.linalg.inv(np.linalg.cholesky(rot))
    R = np.linalg.eigvalsh(Li @ rm @ Li.T).max()
    assert abs(st["R"] - R) < 1e-12 and 0 < st["R"] < 0.1


def test_proposal_mapping_reproduces_covariance():
    from cosmomc_b200 import chains
    _, C = _chains(1, 6, 10)
    M = chains.proposal_mapping(C)
    assert np.allclose(M @ M.T, C, rtol=1e-12)
    assert np.allclose(M, np.tril(M))


def test_chain_row_format():
    from cosmomc_b200 import chains
    row = chains.format_chain_row(1.0, 1382.88563166157, [0.02236, -3.5e-5, 0.0])
    assert row == "   0.1000000E+01   0.1382886E+04   0.2236000E-01  -0.3500000E-04   0.0000000E+00"
    assert len(row) == 16 * 5


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from cosmomc_b200 import chains
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2,
                        init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"])
rank = dist.get_rank()
rng = np.random.default_rng(100 + rank)
local = [rng.normal(size=(2001, 3)) + 0.01 * (2 * rank + j) for j in range(2)]   # two chains per rank
st = chains.update_cov_and_check_converge(local, 3)
ll = chains.allgather_loglikes(np.arange(4.0) + 10 * rank)
if rank == 0:
    np.savez(os.environ["OUT"], cov=st["cov"], R=st["R"], ll=ll)
dist.destroy_process_group()
"""


def test_gloo_world_size_2(tmp_path):
    from cosmomc_b200 import chains
    out = str(tmp_path / "r0.npz")
    script = tmp_path / "w.py"
    script.write_text(WORKER % H.ROOT)
    env = dict(os.environ, PORT="29731", OUT=out)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    assert all(p.wait(timeout=120) == 0 for p in procs)
    got = np.load(out)
    # single-process evaluation over the same four chains, in rank order
    allc = []
    for rank in range(2):
        rng = np.random.default_rng(100 + rank)
        allc += [rng.normal(size=(2001, 3)) + 0.01 * (2 * rank + j) for j in range(2)]
    st = chains.pooled_statistics(chains.pack([chains.chain_record(c) for c in allc], 3), 3)
    assert np.allclose(got["cov"], st["cov"], rtol=1e-13) and abs(float(got["R"]) - st["R"]) < 1e-13
    assert np.array_equal(got["ll"], np.concatenate([np.arange(4.0), np.arange(4.0) + 10]))


def test_theory_cl_writer_reproduces_golden_bytes(tmp_path):
    """WriteTextCls (source/CosmoTheory.f90:197-232): the reference's own
    data/base_plikHM_TTTEEE_lowl_lowE.minimum.theory_cl, re-written from its parsed values, must come back byte for
    byte (header, I6 multipole column, E15.6 columns incl. the 0.000000E+00 BB / PP tail above lmax_lensed)."""
    import os
    import helpers as H
    from cosmomc_b200 import mcmc
    want = open(os.path.join(H.ROOT, "tests", "golden", "data", "theory_cl_excerpt.txt")).read().split("\n")
    T = H.load_templates()
    cls = T["theory_cl"].T            # [5][2509], parsed from the same file by tests/golden/make_golden_data.py
    path = str(tmp_path / "x.theory_cl")
    mcmc.write_theory_cl(path, cls, lmax=2508, digits=6)
    got = open(path).read().split("\n")
    assert len(got) == 2509 and got[-1] == ""          # header + L = 2..2508
    assert got[:31] == want[:31]
    assert got[1000:1006] == want[31:37]
    assert got[2498:2508] == want[37:47]
    L, back = mcmc.read_theory_cl(path)
    assert L[0] == 2 and L[-1] == 2508 and np.array_equal(back, cls[:, 2:])
    # the format statement of the checked-out source is E15.5
    mcmc.write_theory_cl(path, cls, lmax=4)
    assert open(path).read().split("\n")[1] == "     2    0.10175E+04    0.26191E+01    0.30894E-01    0.18185E-05    0.50156E-07"
    assert mcmc._fortran_e(-2.98549, 15, 6) == "  -0.298549E+01" and mcmc._fortran_e(9.99999999, 15, 5) == "    0.10000E+02"
