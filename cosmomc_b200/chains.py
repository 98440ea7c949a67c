"""Chain statistics exchange behind the adaptive proposal covariance.

Host-side mirror of
  * TMpiChainCollector_UpdateCovAndCheckConverge  (source/SampleCollector.f90:212-322): per-chain mean/covariance of
    the second half of the stored samples, gathered over all chains, pooled covariance (count-weighted), covariance of
    the chain means x K/(K-1), Gelman-Rubin R-1 = largest eigenvalue of L^-1 M L^-T;
  * GelmanRubinEvalues                             (source/samples.f90:41-67);
  * BlockedProposer_SetCovariance                  (source/propose.f90:210-244): Cholesky of the correlation matrix.
The reference's two MPI_ALLGATHERs (`SampleCollector.f90:248-251`) become ONE torch.distributed all_gather of the packed
per-chain record [count, mean(n), cov(n*n)] (NCCL over NVLink on the GPU box, gloo in CPU tests).  With several chains
per rank (batched points) every rank contributes all of its local chains.  The messages are < 8 KB per chain: latency
bound, nothing to fuse; the per-point hot path itself has no inter-GPU traffic.

chain rows `[weight, -lnL, params...]` are written with the reference's `(*(E16.7))` format (source/IO.f90:85-93).
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def chain_record(samples):
    """samples: [count][n] stored (thinned) samples of one chain.  Returns (m0, mean[n], cov[n][n]) over the second half,
    with the reference's index range i = Count/2 .. Count (0-based storage, SampleCollector.f90:232-243)."""
    s = np.asarray(samples, dtype=np.float64)
    count = s.shape[0] - 1                 # Samples%Count with items 0..Count
    lo = count // 2
    sel = s[lo:count + 1]
    m0 = float(count - count // 2 + 1)
    mean = sel.sum(axis=0) / m0
    d = sel - mean
    cov = d.T @ d / m0
    return m0, mean, cov


def pack(records, n):
    out = np.zeros((len(records), 1 + n + n * n))
    for i, (m0, mean, cov) in enumerate(records):
        out[i, 0] = m0
        out[i, 1:1 + n] = mean
        out[i, 1 + n:] = cov.ravel()
    return out


def allgather_records(local_records, n, device=None):
    """All ranks contribute [n_local][1+n+n*n]; returns the [K][...] array of every chain in rank order."""
    local = pack(local_records, n)
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return local
    backend = dist.get_backend()
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    t = torch.from_numpy(local).to(dev)
    world = dist.get_world_size()
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out.cpu().numpy()


def gelman_rubin_evalues(cov, meanscov):
    """source/samples.f90:41-67.  Returns (ok, eigenvalues ascending)."""
    n = cov.shape[0]
    sc = np.sqrt(np.diag(cov))
    rot = cov / sc[:, None] / sc[None, :]
    rotmeans = meanscov / sc[:, None] / sc[None, :]
    try:
        L = np.linalg.cholesky(rot)
    except np.linalg.LinAlgError:
        return False, np.zeros(n)
    Linv = np.linalg.inv(L)
    M = Linv @ rotmeans @ Linv.T
    return True, np.linalg.eigvalsh(0.5 * (M + M.T))


def pooled_statistics(gathered, n, min_samples=0):
    """SampleCollector.f90:253-277 on the gathered records.  Returns dict(cov=pooled covariance, R=R-1 or None)."""
    K = gathered.shape[0]
    m0 = gathered[:, 0]
    means = gathered[:, 1:1 + n]
    covs = gathered[:, 1 + n:].reshape(K, n, n)
    if not np.all(m0 > min_samples / 2 + 2):
        return dict(ready=False)
    norm = m0.sum()
    mean = (means * m0[:, None]).sum(axis=0) / norm
    pooled = (covs * m0[:, None, None]).sum(axis=0) / norm
    out = dict(ready=True, cov=pooled, mean=mean, R=None, evals=None)
    if K > 1:
        cov = covs.sum(axis=0) / K
        d = means - mean
        meanscov = (m0[:, None, None] * d[:, :, None] * d[:, None, :]).sum(axis=0) / norm
        meanscov = meanscov * K / (K - 1)
        ok, ev = gelman_rubin_evalues(cov, meanscov)
        out["R"] = float(ev.max()) if ok else 1e6
        out["evals"] = ev if ok else None
    return out


def proposal_mapping(propose_matrix, order=None):
    """source/propose.f90:210-244 for a single block: sigma_i * chol(corr)[i, :] in the (slow->fast) order."""
    C = np.asarray(propose_matrix, dtype=np.float64)
    n = C.shape[0]
    idx = np.arange(n) if order is None else np.asarray(order)
    sig = np.sqrt(np.diag(C))
    corr = C / sig[:, None] / sig[None, :]
    L = np.linalg.cholesky(corr[np.ix_(idx, idx)])
    return sig[idx][:, None] * L


def update_cov_and_check_converge(local_chain_samples, n, min_samples=0, device=None):
    """One call per update on every rank: local_chain_samples = list of [count][n] arrays (the rank's chains)."""
    recs = [chain_record(s) for s in local_chain_samples]
    g = allgather_records(recs, n, device)
    return pooled_statistics(g, n, min_samples)


def allgather_loglikes(total_local, device=None):
    """-lnL of every point of the sharded batch on every rank (bench.py does the same inline)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return np.asarray(total_local)
    backend = dist.get_backend()
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    t = torch.as_tensor(np.asarray(total_local, dtype=np.float64)).to(dev)
    out = torch.empty(dist.get_world_size() * t.numel(), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy()


def format_chain_row(weight, loglike, params):
    """source/IO.f90:85-93 / GeneralTypes.f90:254-274: `(*(E16.7))`, Fortran E-format (0.xxxxxxxE+yy)."""
    def e167(v):
        if v == 0 or not np.isfinite(v):
            return "   0.0000000E+00" if v == 0 else "%16s" % v
        exp = int(np.floor(np.log10(abs(v)))) + 1
        man = v / 10.0 ** exp
        if abs(round(man, 7)) >= 1.0:
            man /= 10.0
            exp += 1
        s = "%.7f" % abs(man)
        return ("%s%sE%+03d" % ("-" if v < 0 else " ", s, exp)).rjust(16)
    return "".join(e167(float(v)) for v in [weight, loglike] + list(params))
