"""Host logic of the batched adaptive Metropolis driver (BASELINE configs[4] first stage: the
`test_likelihood = T` Gaussian, source/calclike.f90:180-199): CPU tests with a numpy likelihood, the world_size-2 gloo
path, chain files; the GPU tests run the same driver on the library's batched likelihood calls."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H


def _target(n, seed=3):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    C = A @ A.T / n + np.eye(n)
    s = np.linspace(0.5, 2.0, n)
    C = C * np.outer(s, s)
    return C, np.linalg.inv(C), rng.normal(size=n)


def _gauss(Ci, c):
    return lambda P: 0.5 * np.einsum("ki,ij,kj->k", P - c, Ci, P - c)


def test_metropolis_recovers_gaussian(tmp_path):
    from cosmomc_b200 import mcmc
    n, K = 4, 16
    C, Ci, c = _target(n)
    rng = np.random.default_rng(0)
    start = c + rng.normal(size=(K, n)) * np.sqrt(np.diag(C))
    root = str(tmp_path / "chains" / "test")
    names = ["p%d" % i for i in range(n)]
    m = mcmc.BatchedMetropolis(_gauss(Ci, c), start, np.diag(np.diag(C)) * 4.0, seed=1, update_every=400,
                               converge_test=0.02, chain_root=root, names=names)
    conv = m.run(max_steps=12000, min_steps=2000)
    assert conv and m.R_history[-1] < 0.02
    # the learned proposal covariance is the pooled sample covariance: close to the target
    assert np.abs(m.cov - C).max() < 0.15 * np.abs(C).max()
    acc = m.n_accept.sum() / (m.K * m.n_steps)
    assert 0.1 < acc < 0.6, acc
    # chain files: weights sum to steps + 1 per chain, rows re-read exactly at E16.7 precision, sample mean ~ centre
    w, ll, P = mcmc.read_chain(root + "_1.txt")
    assert abs(w.sum() - (m.n_steps + 1)) < 1e-9
    assert np.allclose(ll, _gauss(Ci, c)(P), rtol=2e-6, atol=2e-6)
    allP = np.concatenate([np.repeat(np.asarray(s), 1, axis=0) for s in m.samples])
    assert np.abs(allP[len(allP) // 2:].mean(axis=0) - c).max() < 0.15 * np.sqrt(np.diag(C)).max()
    lines = open(root + ".paramnames").read().split("\n")
    assert lines[0] == "p0\tp0" and len(open(root + ".ranges").read().split("\n")) == n + 1


def test_bounds_reject_and_logzero_never_accepted():
    from cosmomc_b200 import mcmc
    n, K = 2, 8
    C, Ci, c = np.eye(2), np.eye(2), np.zeros(2)
    m = mcmc.BatchedMetropolis(_gauss(Ci, c), np.zeros((K, n)) + 0.1, C, pmin=[-0.5, -10], pmax=[0.5, 10], seed=5)
    for _ in range(300):
        m.step()
    allP = np.concatenate([np.asarray(s) for s in m.samples])
    assert allP[:, 0].min() >= -0.5 and allP[:, 0].max() <= 0.5       # hard prior box (GetLogLikeBounds)
    assert np.abs(allP[:, 1]).max() > 0.6                            # the unbounded direction does move


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from cosmomc_b200 import mcmc
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2,
                        init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"])
rank = dist.get_rank()
n, K = 3, 4
C = np.array([[1.0, 0.3, 0.0], [0.3, 2.0, -0.4], [0.0, -0.4, 0.5]]); Ci = np.linalg.inv(C)
f = lambda P: 0.5 * np.einsum("ki,ij,kj->k", P, Ci, P)
rng = np.random.default_rng(7 + rank)
m = mcmc.BatchedMetropolis(f, rng.normal(size=(K, n)), np.eye(n), seed=11, rank=rank, update_every=300, converge_test=0.03)
conv = m.run(max_steps=9000, min_steps=1500)
np.savez(os.environ["OUT"] + str(rank) + ".npz", cov=m.cov, R=np.array(m.R_history), conv=conv, steps=m.n_steps)
dist.destroy_process_group()
"""


def test_gloo_two_ranks_share_the_learned_covariance(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % H.ROOT)
    env = dict(os.environ, PORT="29741", OUT=str(tmp_path / "r"))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    assert all(p.wait(timeout=240) == 0 for p in procs)
    a, b = np.load(str(tmp_path / "r0.npz")), np.load(str(tmp_path / "r1.npz"))
    # both ranks saw all 8 chains: identical pooled covariance and R-1 history, same stopping step
    assert np.array_equal(a["cov"], b["cov"]) and np.array_equal(a["R"], b["R"]) and int(a["steps"]) == int(b["steps"])
    C = np.array([[1.0, 0.3, 0.0], [0.3, 2.0, -0.4], [0.0, -0.4, 0.5]])
    assert bool(a["conv"]) and np.abs(a["cov"] - C).max() < 0.3


TRAJ_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from cosmomc_b200 import mcmc
world = int(os.environ["WORLD"])
if world > 1:
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=world,
                            init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"])
rank = int(os.environ["RANK"])
n, Ktot = 3, 8
K = Ktot // world
C = np.array([[1.0, 0.3, 0.0], [0.3, 2.0, -0.4], [0.0, -0.4, 0.5]]); Ci = np.linalg.inv(C)
f = lambda P: 0.5 * np.einsum("ki,ij,kj->k", P, Ci, P)
start = np.random.default_rng(7).normal(size=(Ktot, n))[rank * K:(rank + 1) * K]
m = mcmc.BatchedMetropolis(f, start, np.eye(n), seed=11, rank=rank, update_every=200, converge_test=1e-9)
m.run(max_steps=2400)
np.savez(os.environ["OUT"] + "w%%d_r%%d.npz" %% (world, rank), cov=m.cov, R=np.array(m.R_history), P=m.P)
if world > 1:
    dist.destroy_process_group()
"""


def test_R_trajectory_is_independent_of_the_rank_layout(tmp_path):
    """SURVEY 8d config 5's acceptance test on CPU (gloo): 8 chains on 2 ranks x 4 and in one process give the SAME R-1
    trajectory, learned covariance and chain positions, bit for bit - a chain's random stream depends on its global index
    only, and the all-gathered per-chain records are pooled in global chain order."""
    script = tmp_path / "w.py"
    script.write_text(TRAJ_WORKER % H.ROOT)
    env = dict(os.environ, PORT="29743", OUT=str(tmp_path) + os.sep)
    one = subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK="0", WORLD="1"))
    two = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), WORLD="2")) for r in range(2)]
    assert one.wait(timeout=240) == 0 and all(p.wait(timeout=240) == 0 for p in two)
    a = np.load(str(tmp_path / "w1_r0.npz"))
    b0, b1 = np.load(str(tmp_path / "w2_r0.npz")), np.load(str(tmp_path / "w2_r1.npz"))
    assert len(a["R"]) == 12 and np.array_equal(a["R"], b0["R"]) and np.array_equal(a["R"], b1["R"])
    assert np.array_equal(a["cov"], b0["cov"])
    assert np.array_equal(a["P"], np.concatenate([b0["P"], b1["P"]]))


def test_eval_batch_control_flow_restatement():
    """numpy restatement of GetLogLike (calclike.f90:97-151) used as the checker of cb200_eval_batch in the GPU tests."""
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import pyoracle as o
    P = np.array([[1.0, 2.0], [5.0, 2.0], [1.0, 2.5]])
    likes = np.array([[3.0, 4.0], [3.0, 4.0], [1e30, 4.0]])
    out, prior, st = o.get_loglike(P, likes, pmin=[0, 0], pmax=[2, 3], prior_mean=[1, 2], prior_std=[0, 0.5],
                                   lincomb=[[1.0, 1.0]], lincomb_mean=[3.0], lincomb_std=[0.1], temperature=2.0)
    assert st.tolist() == [0, 1, 0]
    assert out[0] == pytest.approx((7.0 + 0.0) / 2.0) and out[1] == 1e30 and out[2] == 1e30
    assert prior[2] == pytest.approx(0.5 * (1.0 + 25.0))


@pytest.mark.gpu
def test_gpu_test_likelihood_and_driver():
    from cosmomc_b200 import lib, mcmc
    n, K = 6, 64
    C, Ci, c = _target(n, seed=9)
    h = lib.Handle(lmax_computed_cl=0, max_points=K)  # background-only handle: no CMB tables needed
    rng = np.random.default_rng(2)
    X = c + rng.normal(size=(K, n))
    got = h.test_like_batch(X, c, Ci)
    assert np.allclose(got, _gauss(Ci, c)(X), rtol=1e-12, atol=1e-12)
    m = mcmc.BatchedMetropolis(lambda P: h.test_like_batch(P, c, Ci), X, np.eye(n), seed=3, update_every=300,
                               converge_test=0.02)
    assert m.run(max_steps=6000, min_steps=900)
    assert np.abs(m.cov - C).max() < 0.12 * np.abs(C).max()


@pytest.mark.gpu
def test_gpu_chains_on_the_theory_plus_likelihood_path():
    """BASELINE configs[4] in miniature: 64 chains in lockstep over (logA, n_s, calPlanck) at fixed cosmology - every
    step is ONE batched call of the hot path (shared transfer functions -> k-contraction GEMM -> lensing ->
    plik-lite-shaped chi^2) plus the Gaussian calibration prior; the proposal covariance is learned from the pooled
    chain statistics.  The posterior is close to Gaussian, so the learned covariance must approach the Fisher estimate
    obtained by finite differences of the same -lnL."""
    from cosmomc_b200 import lib, mcmc, synthetic as syn
    T = H.load_templates()
    K = 64
    h = lib.Handle(max_points=K + 8, chunk_points=K + 8, lmax_out=H.LMAX_OUT)
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    b = H.small_batch(1, seed=5, NT=h.info.n_tau_max, NK=h.info.n_k_max)
    h.upload_sources(b["thermo"], b["n_k"], b["k"], b["src"])
    ip0 = b["initpower"][0].copy()
    c0 = np.array([np.log(1e10 * ip0[0]), ip0[1], 1.0])

    def theory(P):
        ip = np.tile(ip0, (len(P), 1))
        ip[:, 0] = 1e-10 * np.exp(P[:, 0])
        ip[:, 1] = P[:, 1]
        return ip

    # data = the model's own band powers at the centre (no noise): the posterior peaks at c0
    cls0 = h.powers_shared(theory(c0[None, :]), np.ones(1), src_point=0, first=1)[0][0]
    fid = np.zeros((5, H.LMAX_OUT + 1))
    fid[:3] = cls0[:3]
    d = syn.synthetic_pliklite(H.LMAX_OUT, fiducial_cls=fid, seed=99)
    x_data = d["x_model"] if "x_model" in d else d["x_data"]
    h.add_pliklite(d["nb"], d["blmin"], d["blmax"], d["weights"], d["invcov"], x_data, 0)

    def nll(P):
        P = np.atleast_2d(P)
        h.powers_shared(theory(P), np.ones(len(P)), src_point=0, first=1, want_cls=False)
        ll, tot, st = h.loglike_batch(len(P), P[:, 2:3].copy(), first=1)
        out = tot + 0.5 * ((P[:, 2] - 1.0) / 0.0025) ** 2          # Gaussian prior on calPlanck (planck_calibration.ini)
        out[st != 0] = 1e30
        return out

    # Fisher matrix by central differences of -lnL around the best fit of this (noisy) data realisation
    steps = np.array([2e-3, 1e-3, 5e-4])
    f0 = nll(c0)[0]
    Hm = np.zeros((3, 3))
    for i in range(3):
        for j in range(i, 3):
            ei, ej = np.eye(3)[i] * steps[i], np.eye(3)[j] * steps[j]
            v = nll(np.stack([c0 + ei + ej, c0 + ei - ej, c0 - ei + ej, c0 - ei - ej]))
            Hm[i, j] = Hm[j, i] = (v[0] - v[1] - v[2] + v[3]) / (4 * steps[i] * steps[j])
    C_fisher = np.linalg.inv(Hm)
    assert np.all(np.linalg.eigvalsh(Hm) > 0) and np.isfinite(f0)
    rng = np.random.default_rng(1)
    start = c0 + rng.normal(size=(K, 3)) * np.sqrt(np.diag(C_fisher))
    m = mcmc.BatchedMetropolis(nll, start, np.diag(np.diag(C_fisher)), seed=4, update_every=100, converge_test=0.05)
    m.run(max_steps=600, min_steps=300)
    acc = m.n_accept.sum() / (m.K * m.n_steps)
    assert 0.08 < acc < 0.7, acc
    assert m.R_history and m.R_history[-1] < 0.2, m.R_history
    sd_f, sd_c = np.sqrt(np.diag(C_fisher)), np.sqrt(np.diag(m.cov))
    assert np.all(np.abs(sd_c / sd_f - 1) < 0.35), (sd_c, sd_f)


def test_likelihood_derived_columns_and_description(tmp_path):
    from cosmomc_b200 import mcmc
    likes = np.array([[3.0, 4.0, 5.0], [1.0, 2.0, 3.0]])
    ll = np.array([12.5, 6.25])                       # includes 0.5 / 0.25 of prior
    d = mcmc.likelihood_derived_params(likes, ll, type_indices=[[0, 1], [2]], derived=[[7.0], [8.0]])
    assert d.shape == (2, 1 + 3 + 1 + 2)
    assert d[0].tolist() == [7.0, 6.0, 8.0, 10.0, 1.0, 14.0, 10.0]
    assert d[1].tolist() == [8.0, 2.0, 4.0, 6.0, 0.5, 6.0, 6.0]
    p = str(tmp_path / "c.likelihoods")
    mcmc.write_likelihoods(p, [("CMB", "lensing", "smicadx12_Dec5", "v1"), dict(type="BAO", tag="DR12", name="DR12", version="")])
    assert open(p).read() == "1\tCMB\tlensing\tsmicadx12_Dec5\tv1\n1\tBAO\tDR12\tDR12\t\n"
