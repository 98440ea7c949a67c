#!/usr/bin/env python
"""bench.py — lensed-C_l + lnL evaluations/sec at lmax=2500 (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[3]): base LCDM TTTEEE + lensing C_l at lmax = 2508 with a
plik-lite-shaped chi^2, 16 384 synthetic parameter points.  One "step" = one pass of the hot path over the batch:
  source spline -> line-of-sight projection -> k-contraction -> l-interpolation -> lensing -> units -> chi^2.
The Boltzmann source ODEs stay on the reference path (north_star) and are NOT part of the step: the step starts
from Src(k, source, tau) exactly where CAMB hands it to InitSourceInterpolation (camb/cmbmain.f90:238-263).

  value : whole-job evaluations/s, sources already resident in HBM when the timed region starts
  e2e   : same metric through the C-ABI calls a user makes (cb200_upload_sources from pinned HOST buffers - every
          block of points has its OWN buffer holding ITS sources -, cb200_powers returning the C_l to a host buffer,
          cb200_loglike_batch returning -lnL): H2D and D2H inside the timed region; the e2e -lnL must equal the
          resident-path -lnL bit for bit (checked before timing).

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1: under torchrun, one rank per GPU).
  --scaling strong (default): the 16 384 points of configs[3] are SHARDED over the ranks (no data-path collective;
                              NCCL all-gathers the log-likes);  --scaling weak: --points per GPU.
  --impl reference : the CPU restatement of the reference (oracle/, -O3 -march=native -fopenmp build; the Fortran
                     reference cannot be built: no Fortran compiler in this image or on the GPU box, see
                     profiles/r02_probe_gpu_host_no_fortran.log) on ALL host cores, rank 0 only.
  --config 2|3|5   : BASELINE configs[1] (background-only, JLA + BAO + HST) / configs[2] (BK15, shared transfers) /
                     configs[4] (64-chain adaptive MCMC, NCCL covariance learning; run it under torchrun for N > 1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lensed-C_l+lnL evaluations/sec at lmax=2500"
UNIT = "evaluations/s"
NT_MAX, NK_MAX, NQ_MAX = 576, 224, 3072
LMAX_OUT = 2508
FID_SEED, FID_POINTS = 999, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5],
                    help="SURVEY 8d config: 4 = BASELINE configs[3] (the metric's workload, default); 2, 3 = configs[1], [2]")
    ap.add_argument("--points", type=int, default=int(os.environ.get("CB200_BENCH_POINTS", 16384)),
                    help="parameter points per step: in total (strong scaling) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--chunk", type=int, default=2048, help="points per pass of the kernels (work buffers scale with it: 20 GB at 2048); 1024 -> 2048 -> 4096: 8 965 -> 9 018 -> 9 043 evals/s, the lensing GEMMs and element-wise kernels fill the GPU better")
    ap.add_argument("--block-points", type=int, default=0, help="points per pinned host block / upload (e2e); 0 = P/16 clamped to [128, 512]: at least 16 stages per GPU so that the fill and drain of the upload / evaluate / download pipeline stay small under strong scaling")
    ap.add_argument("--cpu-sample", type=int, default=48, help="points of the same workload timed on the CPU oracle")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cls-d2h", action="store_true", help="e2e: leave the C_l on the device (only -lnL comes back)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu_metrics():
    """Per-point DRAM bytes and shared-memory wavefronts of the dominant kernel, as tools/ncu_metrics.py extracted
    them from an `ncu --set full` capture (a profile, not part of the timed run): profiles/project_ncu_metrics.json."""
    p = os.path.join(ROOT, "profiles", "project_ncu_metrics.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def draw_workload(npts, seed):
    from cosmomc_b200 import synthetic as syn
    thermo = syn.draw_thermo(npts, seed)
    ip, alens, cal, pert = syn.draw_params(npts, seed)
    return dict(thermo=thermo, initpower=ip, alens=alens, cal=cal, pert=pert)


def add_grids(W, h):
    from cosmomc_b200 import synthetic as syn
    W["tau"], _, W["n_tau"], W["k"], W["n_k"] = syn.build_grids(h, W["thermo"])
    return W


def templates():
    return np.load(os.path.join(ROOT, "tests", "golden", "templates.npz"))


def pliklite_data(fid_cls):
    """plik-lite-shaped data set whose data vector is the binned MEAN C_l of FID_POINTS fixed-seed points of this
    very workload plus noise of the covariance, so that -lnL is O(10^2..10^4) as in a real chain (not 1e7)."""
    from cosmomc_b200 import synthetic as syn
    fid = np.zeros((5, LMAX_OUT + 1))
    fid[:3] = fid_cls[:3]
    return syn.synthetic_pliklite(LMAX_OUT, fiducial_cls=fid)


def oracle_batch(W, sel):
    """Host copy (sources generated on the CPU) of the selected points, in the layout BatchChain.run wants."""
    from cosmomc_b200 import synthetic as syn
    src = syn.make_sources(W["thermo"][sel], W["tau"][sel], W["k"][sel], W["pert"][sel]).numpy()
    return dict(thermo=W["thermo"][sel], initpower=W["initpower"][sel], alens=W["alens"][sel], cal=W["cal"][sel],
                n_tau=W["n_tau"][sel], n_k=W["n_k"][sel], k=W["k"][sel], src=src)


def oracle_grids_for(W):
    """tau / k grids from the ORACLE's own builders (reference arm: no GPU handle exists there)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    n = len(W["thermo"])
    W["tau"] = np.zeros((n, NT_MAX)); W["k"] = np.zeros((n, NK_MAX))
    W["n_tau"] = np.zeros(n, dtype=np.int32); W["n_k"] = np.zeros(n, dtype=np.int32)
    for i in range(n):
        t, dt, kk = H.oracle_grids(W["thermo"][i])
        W["n_tau"][i], W["n_k"][i] = len(t), len(kk)
        W["tau"][i, :len(t)] = t; W["tau"][i, len(t):] = t[-1]
        W["k"][i, :len(kk)] = kk; W["k"][i, len(kk):] = kk[-1]
    return W


def make_chain(T, data, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle as o
    import helpers as H
    ls = o.initlval(H.MAX_L)
    return o.BatchChain(ls, H.MAX_ETA_K, H.MAX_L, H.LMAX_COMPUTED, LMAX_OUT, T["highl_unlensed"], T["highl_lensed"],
                        data, fast=True, threads=threads or os.cpu_count())


def cpu_rates(chain, batch, reps=1):
    """evaluations/s of the CPU restatement on `batch`: one point per core (throughput-fair: what independent chains on
    every core would deliver) and OpenMP inside one point (how a single CosmoMC chain uses its cores)."""
    n = len(batch["thermo"])
    out = {}
    for name, mode in (("one_point_per_core", 1), ("openmp_inside_point", 0)):
        t0 = time.perf_counter()
        for _ in range(reps):
            ll = chain.run(batch, mode)
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"value": n / dt, "seconds": dt}
        out["lnl"] = ll
    return out


def workload_config(args):
    world = args.gpus
    per_gpu = args.points if args.scaling == "weak" else -(-args.points // world)
    return {"workload": "BASELINE configs[3]: base LCDM TTTEEE+lensing C_l, lmax_computed=2500 (Max_l 2650, "
                        "Max_eta_k 14000, lmax_out 2508), plik-lite-shaped chi2 (613 bins), synthetic sources",
            "points_per_step_total": per_gpu * world, "points_per_gpu_per_step": per_gpu, "scaling": args.scaling,
            "n_tau": "~555", "n_k": "~220", "n_q": "~2910", "n_lsamp": 88,
            "cache": "inputs (3.1 MB/point, %.1f GB/step/GPU) far larger than the 126 MB L2" %
                     (per_gpu * NT_MAX * 3 * NK_MAX * 8 / 1e9),
            "sharding": "points split over ranks, no data-path collective; NCCL all-gather of log-likes"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    T = templates()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import pyoracle as o
    # data vector: binned mean C_l of the fixed-seed fiducial points, from the oracle itself (no GPU in this arm)
    F = oracle_grids_for(draw_workload(FID_POINTS, FID_SEED))
    fb = oracle_batch(F, np.arange(FID_POINTS))
    ls = o.initlval(H.MAX_L)
    bes = o.Bessel(ls, H.MAX_ETA_K)
    fid = np.mean([H.oracle_point(fb, i, bes, ls, T["highl_unlensed"], T["highl_lensed"])["cls_out"]
                   for i in range(FID_POINTS)], axis=0)
    data = pliklite_data(fid)
    # torchrun exports OMP_NUM_THREADS=1 for N > 1: the thread count is set explicitly, from the core count
    chain = make_chain(T, data, threads=cores)
    npts = max(cores, args.cpu_sample)
    W = oracle_grids_for(draw_workload(npts, 1000))
    batch = oracle_batch(W, np.arange(npts))
    for _ in range(max(1, min(args.warmup, 2))):
        chain.run(batch, 1)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        chain.run(batch, 1)
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    chain.run(batch, 0)
    inside = npts / (time.perf_counter() - t0)
    value = float(npts * len(times) / sum(times))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": chain.threads, "kind": "port",
                            "openmp_inside_point": inside,
                            "sample": "%d points per step of the same synthetic workload; C++/OpenMP restatement of the "
                                      "reference (oracle/, built -O3 -march=native -fopenmp), one point per core on all %d "
                                      "cores (value) and, once, OpenMP inside one point (openmp_inside_point)"
                                      % (npts, chain.threads)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def bind_to_gpu_numa_node(local):
    """Several ranks on one host: run this rank (and so first-touch its pinned staging buffers) on the CPUs of the NUMA
    node its GPU hangs off, so that the per-block uploads do not cross the socket interconnect.  No-op on a single-node
    host or when sysfs does not say.  Returns what was done (reported in the e2e object)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return {"gpu_pci": bdf, "numa_node": node, "nodes": len(nodes), "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu_pci": bdf, "numa_node": node, "nodes": len(nodes), "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu_pci": bdf, "numa_node": node, "nodes": len(nodes), "bound": True, "cpus": len(cpus)}
    except Exception as e:   # never fatal: this is placement, not correctness
        return {"bound": False, "error": str(e)[:120]}


def main():
    args = parse()
    if args.config != 4:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        return bench_configs.run(args)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    from cosmomc_b200 import lib, synthetic as syn
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.gpus = world
    cfg = workload_config(args)
    P = cfg["points_per_gpu_per_step"]
    hbm_peak, peak_src = load_peaks()
    h = lib.Handle(device=local, max_points=P + FID_POINTS, chunk_points=min(args.chunk, P), lmax_out=LMAX_OUT,
                   n_tau_max=NT_MAX, n_k_max=NK_MAX, n_q_max=NQ_MAX)
    T = templates()
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    dev = "cuda:%d" % local

    # ---- data vector from the workload's own mean C_l (fixed seed: identical on every rank and in the CPU arm)
    F = add_grids(draw_workload(FID_POINTS, FID_SEED), h)
    s = syn.make_sources(F["thermo"], F["tau"], F["k"], F["pert"], device=dev)
    torch.cuda.synchronize()
    h.upload_sources(F["thermo"], F["n_k"], F["k"], None, first=P, src_device_ptr=s.data_ptr())
    fid_cls, _, _ = h.powers(F["initpower"], F["alens"], first=P)
    del s
    data = pliklite_data(fid_cls.mean(axis=0))
    h.add_pliklite(data["nb"], data["blmin"], data["blmax"], data["weights"], data["invcov"], data["x_data"], 0)

    W = add_grids(draw_workload(P, 1000 + rank), h)
    nuis = W["cal"].reshape(-1, 1)

    # ---- synthetic sources: generated on the device, made resident, and (e2e) copied to pinned host blocks - one
    #      block per upload, each holding ITS points' sources.  All outside every timed region.
    B = min(args.block_points if args.block_points > 0 else max(128, min(512, P // 16)), P)
    blocks = [(a, min(P, a + B)) for a in range(0, P, B)]
    if B >= 256 and P >= 4 * B:
        # the first upload has nothing to hide behind: start with small blocks (B/8, B/8, B/4, B/2) so that the pipeline
        # fills in an eighth of the time, then continue with blocks of B points
        ramp, a = [], 0
        for n in (B // 8, B // 8, B // 4, B // 2):
            ramp.append((a, a + n)); a += n
        blocks = ramp + [(x, min(P, x + B)) for x in range(a, P, B)]
    stage = []
    for a, b in blocks:
        s = syn.make_sources(W["thermo"][a:b], W["tau"][a:b], W["k"][a:b], W["pert"][a:b], device=dev)
        torch.cuda.synchronize()
        h.upload_sources(W["thermo"][a:b], W["n_k"][a:b], W["k"][a:b], None, first=a, src_device_ptr=s.data_ptr())
        if not args.no_e2e:
            # the block as a CosmoMC-side caller holds it: every point's Src(1:n_k, 1:3, 1:n_tau) back to back, no padding
            nt = torch.as_tensor(W["n_tau"][a:b], device=dev).view(-1, 1, 1, 1)
            nk = torch.as_tensor(W["n_k"][a:b], device=dev).view(-1, 1, 1, 1)
            keep = (torch.arange(NT_MAX, device=dev).view(1, -1, 1, 1) < nt) & \
                   (torch.arange(NK_MAX, device=dev).view(1, 1, 1, -1) < nk)
            packed = s[keep.expand_as(s)]
            hb = torch.empty(packed.numel(), dtype=torch.float64, pin_memory=True)
            hb.copy_(packed)
            stage.append(hb)
            del packed, keep
        del s
    torch.cuda.synchronize()
    h.sync()
    cls_host = None
    if not args.no_e2e and not args.no_cls_d2h:
        cls_host = torch.empty((P, 5, LMAX_OUT + 1), dtype=torch.float64, pin_memory=True)

    def step_resident():
        h.powers_resident(W["initpower"], W["alens"])
        ll, tot, st = h.loglike_batch(P, nuis)
        return tot

    def step_e2e():
        # the public calls a chain driver makes, block by block, with options "async_upload" / "async_results": the
        # packed upload of block i+1 (copy stream) overlaps the kernels of block i (compute stream) and the copy of
        # block i-1's C_l back to the host (result stream); nothing blocks the host until -lnL is asked for
        h.set_option("async_upload", 1)
        h.set_option("async_results", 1)
        for i, (a, b) in enumerate(blocks):
            h.upload_sources_packed(W["thermo"][a:b], W["n_tau"][a:b], W["n_k"][a:b], W["k"][a:b], first=a,
                                    src_host_ptr=stage[i].data_ptr())
            if cls_host is not None:
                h.powers_into(W["initpower"][a:b], W["alens"][a:b], first=a,
                              cls_ptr=cls_host.data_ptr() + a * 5 * (LMAX_OUT + 1) * 8)
            else:
                h.powers_resident(W["initpower"][a:b], W["alens"][a:b], first=a)
        h.set_option("async_upload", 0)
        h.set_option("async_results", 0)
        ll, tot, st = h.loglike_batch(P, nuis)   # returns once every upload, kernel and result copy has finished
        return tot

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        h.sync()

    def gather(tot):
        if dist is None:
            return tot
        t = torch.from_numpy(tot).cuda()
        out = torch.empty(world * len(tot), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(out, t)
        return out

    def timed(fn, steps):
        barrier()
        h.timing(reset=True)
        h.timer_start()
        for _ in range(steps):
            gather(fn())
        torch.cuda.synchronize()
        ms = h.timer_stop()
        tm = h.timing(reset=True)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, tm

    # ---- value: resident inputs
    tot_res = None
    for _ in range(args.warmup):
        tot_res = step_resident()
    if tot_res is None:
        tot_res = step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms, tm = timed(step_resident, args.steps)
    clocks = sampler.stop()
    value = world * P * args.steps / (ms * 1e-3)

    # ---- exact unit-of-work count of the dominant kernel (one untimed counting pass)
    h.set_option("count_triples", 1)
    h.timing(reset=True)
    h.powers_resident(W["initpower"], W["alens"])
    triples = h.timing(reset=True)["proj_triples"]
    h.set_option("count_triples", 0)
    dfma_peak, dmma_peak = h.measure_fp64_peaks()

    # ---- e2e: host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        tot_e2e = None
        for _ in range(max(1, min(args.warmup, 1))):
            tot_e2e = step_e2e()
        # every block was re-uploaded from ITS host buffer: the answer must be the resident path's, bit for bit
        if not np.array_equal(tot_e2e, tot_res):
            raise SystemExit("e2e -lnL differs from the resident-path -lnL: max |diff| = %g"
                             % np.abs(tot_e2e - tot_res).max())
        ms2, _ = timed(step_e2e, args.steps)
        h2d = int(sum(x.numel() for x in stage)) * 8 + \
            P * (5 * 8 + 16 + NK_MAX * 8 + 2 * NT_MAX * 8 + 2 * NQ_MAX * 8 + 8 + 344 + 10 * 8 + 8 + 8)
        d2h = P * (8 + 8 + 4) + (P * 5 * (LMAX_OUT + 1) * 8 if cls_host is not None else 0)
        # what the PCIe link of this GPU delivers for one of these very blocks (plain pinned -> device copy, best of 3):
        # the e2e rate cannot exceed link bandwidth / bytes per point
        link = 0.0
        dst = torch.empty(stage[0].numel(), dtype=torch.float64, device=dev)
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(stage[0], non_blocking=True)
            e1.record()
            e1.synchronize()
            link = max(link, stage[0].numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del dst
        e2e = {"value": world * P * args.steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "h2d_link_gbs": link, "h2d_gbs_achieved": h2d / (ms2 / args.steps * 1e-3) / 1e9,
               "h2d_bound_evals_per_s_per_gpu": link * 1e9 / (h2d / P),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2 / args.steps,
               "equals_resident_result": True,
               "note": "%d pinned host blocks of up to %d points per GPU (the first four smaller: the pipeline fills on a short upload), each holding its own points' sources packed at their "
                       "exact sizes (cb200_upload_sources_packed); upload of block i+1, kernels of block i and the "
                       "result copy of block i-1 overlap on three streams; C_l [pt][5][%d] %s"
                       % (len(blocks), B, LMAX_OUT + 1,
                          "copied back to a pinned host buffer" if cls_host is not None else "left on the device")}
        if numa is not None:
            e2e["host_placement_rank0"] = numa

    # ---- roofline of the dominant kernel (K1 projection; launched once per chunk of points)
    n_chunks = -(-P // min(args.chunk, P))
    n_launch_k1 = args.steps * n_chunks
    k1_ms = tm["ms_project"] / n_launch_k1
    pts_per_launch = P / n_chunks
    # algorithmic HBM bytes per point: Src + ddSrc read once, one partial-sum block [6][96] written per 24 wavenumbers
    # + the time integrals of a block on their way from the projection kernel to its finish kernel (6 groups x 33 values x 32
    #   lanes, written once and read once)
    bytes_pt = float(np.mean(W["n_tau"].astype(np.float64) * 3 * W["n_k"] * 8 * 2)) + (2910 / 24.0) * 6 * 96 * 8 \
        + (2910 / 24.0) * 6 * 33 * 32 * 8 * 2
    trip_pt = triples / P
    flop_per_triple = 13  # cubic j_l interpolation: 1 mul + 3 FMA; 3 FMA accumulations (T, E, lensing potential)
    sm_clock_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
    sec = k1_ms * 1e-3
    ncu = load_ncu_metrics()
    roof = {"kernel": "project4_kernel + project4_finish_kernel (K1+K2: line-of-sight projection, then Limber values and the "
                      "partial k-contraction; the phase time also holds the project3_kernel pass over the first wavenumber block)",
            "bound": "fp64", "achieved": trip_pt * flop_per_triple * pts_per_launch / sec / 1e12, "peak": dfma_peak,
            "unit": "TFLOP/s", "peak_source": "measured live (cb200_measure_fp64_peaks: DFMA micro-kernel; "
            "MEASURED_PEAKS.json has no FP64 entry)", "peak_dmma": dmma_peak, "flop_per_triple": flop_per_triple,
            "triples_per_point": trip_pt, "points_per_launch": pts_per_launch, "ms_per_launch": k1_ms,
            "us_per_point": 1e3 * k1_ms / pts_per_launch, "share_of_step": tm["ms_project"] / tm["ms_total"],
            "traffic": (ncu["dram_bytes_per_point"] * pts_per_launch) if ncu else None,
            "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum per point x points per launch, from %s "
                               "(%d-point capture, build %s)" % (ncu["source"], ncu["points"], ncu["build"])) if ncu else None,
            "hbm": {"achieved": bytes_pt * pts_per_launch / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": peak_src, "algorithmic_bytes_per_point": bytes_pt},
            "table_gather": {"achieved": trip_pt * 32 * pts_per_launch / sec / 1e9, "unit": "GB/s",
                             "bytes_per_triple": 32},
            "note": "SURVEY 8d: the projection is bound by FP64 issue and the on-chip table gather, not by HBM (DRAM traffic "
                    "= algorithmic bytes); `hbm` and `smem_lsu` are reported beside the FP64 fraction"}
    roof["frac"] = roof["achieved"] / max(dfma_peak, 1e-9)
    roof["hbm"]["frac"] = roof["hbm"]["achieved"] / hbm_peak
    if ncu:
        roof["smem_lsu"] = {"achieved": ncu["smem_wavefronts_per_point"] * pts_per_launch / sec / 1e9,
                            "peak": 148 * sm_clock_hz / 1e9, "unit": "Gwavefronts/s",
                            "wavefronts_per_point": ncu["smem_wavefronts_per_point"],
                            "note": "l1tex__data_pipe_lsu_wavefronts_mem_shared per point (same ncu capture) / launch time, "
                                    "against one 128-byte wavefront per clock per SM"}
        roof["smem_lsu"]["frac"] = roof["smem_lsu"]["achieved"] / roof["smem_lsu"]["peak"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        chain = make_chain(T, data)
        n = min(max(args.cpu_sample, chain.threads), P)
        batch = oracle_batch(W, np.arange(n))
        chain.run(batch, 1)  # warm-up
        r = cpu_rates(chain, batch)
        dl = float(np.abs(r["lnl"] - tot_res[:n]).max())
        cpu = {"value": r["one_point_per_core"]["value"], "unit": UNIT, "cores": chain.threads, "kind": "port",
               "openmp_inside_point": r["openmp_inside_point"]["value"],
               "max_abs_dlnl_vs_gpu": dl, "lnl_range": [float(tot_res[:n].min()), float(tot_res[:n].max())],
               "sample": "%d points of the same workload (%.1f s + %.1f s); C++/OpenMP restatement of the reference "
                         "(oracle/, -O3 -march=native -fopenmp): one point per core (value) / OpenMP inside one point; the "
                         "Fortran reference cannot be built (no Fortran compiler, profiles/r02_probe_gpu_host_no_fortran.log)"
                         % (n, r["one_point_per_core"]["seconds"], r["openmp_inside_point"]["seconds"])}
        if dl >= 0.01:
            raise SystemExit("GPU and CPU-oracle -lnL differ by %g (north_star: < 0.01)" % dl)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(tm["n_launches"]), "roofline": roof,
               "cpu_baseline": cpu,
               "phase_ms_per_step": {k: tm[k] / args.steps for k in
                                     ["ms_spline", "ms_project", "ms_contract", "ms_interp", "ms_lens", "ms_like"]}}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
