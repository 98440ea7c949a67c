#!/usr/bin/env python
"""bench.py — lensed-C_l + lnL evaluations/sec at lmax=2500 (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[3]): base LCDM TTTEEE + lensing C_l at lmax = 2508 with a
plik-lite-shaped chi^2, synthetic parameter points.  One "step" = one pass of the hot path over one batch:
  source spline -> line-of-sight projection -> k-contraction -> l-interpolation -> lensing -> units -> chi^2.
The Boltzmann source ODEs stay on the reference path (north_star) and are NOT part of the step: the step starts
from Src(k, source, tau) exactly where CAMB hands it to InitSourceInterpolation (camb/cmbmain.f90:238-263).

  value : whole-job evaluations/s, sources already resident in HBM when the timed region starts
  e2e   : same metric through the C-ABI calls a user makes (cb200_upload_sources from pinned HOST buffers,
          cb200_powers, cb200_loglike_batch returning host arrays): H2D and D2H inside the timed region.

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1: under torchrun, one rank per GPU; the batch is
sharded over ranks with no data-path collective; NCCL all-gathers the log-likes; scaling = weak).
        python bench.py --impl reference ...  times the CPU oracle (restated reference; the Fortran reference cannot
        be built: no Fortran compiler in this image or on the GPU box) on all host cores, rank 0 only.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=int(os.environ.get("CB200_BENCH_POINTS", 16384)),
                    help="parameter points per GPU per step")
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--stage-points", type=int, default=512, help="points in the pinned host staging buffer (e2e)")
    ap.add_argument("--cpu-sample", type=int, default=192, help="points of the same workload timed on the CPU oracle")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_workload(h, npts, seed):
    """Per-rank synthetic batch: thermo scalars, grids (library builders), init-power params, plik-lite-shaped data."""
    from cosmomc_b200 import synthetic as syn
    thermo = syn.draw_thermo(npts, seed)
    ip, alens, cal, pert = syn.draw_params(npts, seed)
    tau, dtau, n_tau, k, n_k = syn.build_grids(h, thermo)
    return dict(thermo=thermo, initpower=ip, alens=alens, cal=cal, pert=pert, tau=tau, k=k, n_tau=n_tau, n_k=n_k)


def pliklite_data():
    from cosmomc_b200 import synthetic as syn
    T = np.load(os.path.join(ROOT, "tests", "golden", "templates.npz"))
    fid = np.zeros((5, LMAX_OUT + 1))
    fid[:3] = T["theory_cl"][:, :3].T
    return T, syn.synthetic_pliklite(LMAX_OUT, fiducial_cls=fid)


def cpu_oracle_rate(W, T, data, npts, threads=None):
    """Time the CPU restatement (oracle, OpenMP over q / theta inside a point, points one at a time: CAMB is
    not thread-safe across models, camb/cmbmain.f90:7-8) on `npts` points of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle as o
    import helpers as H
    from cosmomc_b200 import synthetic as syn
    ls = o.initlval(H.MAX_L)
    bessel = o.Bessel(ls, H.MAX_ETA_K)
    sel = np.arange(npts) % len(W["thermo"])
    src = syn.make_sources(W["thermo"][sel], W["tau"][sel], W["k"][sel], W["pert"][sel]).numpy()
    batch = dict(thermo=W["thermo"][sel], initpower=W["initpower"][sel], alens=W["alens"][sel], cal=W["cal"][sel],
                 n_tau=W["n_tau"][sel], n_k=W["n_k"][sel], k=W["k"][sel], src=src)
    H.oracle_point(batch, 0, bessel, ls, T["highl_unlensed"], T["highl_lensed"])  # warm-up
    t0 = time.perf_counter()
    tot = 0.0
    for i in range(npts):
        r = H.oracle_point(batch, i, bessel, ls, T["highl_unlensed"], T["highl_lensed"])
        c = r["cls_out"]
        tot += o.pliklite(np.stack([c[0], c[1], c[2]]), data["nb"], data["blmin"], data["blmax"], data["weights"],
                          data["invcov"], data["x_data"], batch["cal"][i])
    dt = time.perf_counter() - t0
    return npts / dt, dt, tot


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the oracle only needs the grid builders of the library (host code) - no GPU handle here
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    from cosmomc_b200 import synthetic as syn
    T, data = pliklite_data()
    cores = os.cpu_count()
    npts = max(8, min(args.cpu_sample, 64))
    thermo = syn.draw_thermo(npts, 7)
    ip, alens, cal, pert = syn.draw_params(npts, 7)
    b = H.small_batch(npts, seed=7, NT=NT_MAX, NK=NK_MAX)
    W = dict(thermo=b["thermo"], initpower=b["initpower"], alens=b["alens"], cal=b["cal"], pert=pert, tau=b["tau"],
             k=b["k"], n_tau=b["n_tau"], n_k=b["n_k"])
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_oracle_rate(W, T, data, 8)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, _ = cpu_oracle_rate(W, T, data, npts)
        rates.append(r)
        times.append(dt)
    value = float(npts * len(times) / sum(times))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, npts_override=npts),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "%d points per step of the same synthetic workload, C++/OpenMP restatement "
                                      "(oracle/), one point at a time, OpenMP inside the point" % npts},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def workload_config(args, npts_override=None):
    return {"workload": "BASELINE configs[3]: base LCDM TTTEEE+lensing C_l, lmax_computed=2500 (Max_l 2650, "
                        "Max_eta_k 14000, lmax_out 2508), plik-lite-shaped chi2 (613 bins), synthetic sources",
            "points_per_gpu_per_step": npts_override or args.points, "n_tau": "~555", "n_k": "~220", "n_q": "~2910",
            "n_lsamp": 88, "cache": "inputs (3.1 MB/point, %.1f GB/step/GPU) far larger than the 126 MB L2"
                                    % ((npts_override or args.points) * NT_MAX * 3 * NK_MAX * 8 / 1e9),
            "sharding": "points split over ranks, no data-path collective; NCCL all-gather of log-likes"}


def main():
    args = parse()
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = args.points
    hbm_peak, peak_src = load_peaks()
    h = lib.Handle(device=local, max_points=P, chunk_points=min(args.chunk, P), lmax_out=LMAX_OUT, n_tau_max=NT_MAX,
                   n_k_max=NK_MAX, n_q_max=NQ_MAX)
    T, data = pliklite_data()
    h.set_templates(T["highl_unlensed"], T["highl_lensed"])
    h.add_pliklite(data["nb"], data["blmin"], data["blmax"], data["weights"], data["invcov"], data["x_data"], 0)
    W = make_workload(h, P, seed=1000 + rank)
    nuis = W["cal"].reshape(-1, 1)

    # ---- generate the synthetic sources on the device and make them resident (outside every timed region)
    gen = 512
    stage_n = min(args.stage_points, P)
    stage = None if args.no_e2e else torch.empty((stage_n, NT_MAX, 3, NK_MAX), dtype=torch.float64, pin_memory=True)
    for a in range(0, P, gen):
        b = min(P, a + gen)
        s = syn.make_sources(W["thermo"][a:b], W["tau"][a:b], W["k"][a:b], W["pert"][a:b], device="cuda:%d" % local)
        torch.cuda.synchronize()
        h.upload_sources(W["thermo"][a:b], W["n_k"][a:b], W["k"][a:b], None, first=a, src_device_ptr=s.data_ptr())
        if stage is not None and a < stage_n:
            stage[a:min(b, stage_n)].copy_(s[: min(b, stage_n) - a])
        del s
    torch.cuda.synchronize()
    h.sync()

    def step_resident():
        h.powers_resident(W["initpower"], W["alens"])
        ll, tot, st = h.loglike_batch(P, nuis)
        return tot

    def step_e2e():
        # block i+1 is uploaded (copy stream) while block i is being evaluated (compute stream): the public calls
        # are the same three - upload_sources / powers / loglike_batch - with the "async_upload" option on
        h.set_option("async_upload", 1)
        for a in range(0, P, stage_n):
            b = min(P, a + stage_n)
            h.upload_sources(W["thermo"][a:b], W["n_k"][a:b], W["k"][a:b], None, first=a,
                             src_host_ptr=stage.data_ptr())
            h.powers_resident(W["initpower"][a:b], W["alens"][a:b], first=a)
        h.set_option("async_upload", 0)
        ll, tot, st = h.loglike_batch(P, nuis)
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
    for _ in range(args.warmup):
        step_resident()
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
        for _ in range(max(1, min(args.warmup, 1))):
            step_e2e()
        ms2, _ = timed(step_e2e, args.steps)
        h2d = P * (NT_MAX * 3 * NK_MAX * 8 + 5 * 8 + 4 + NK_MAX * 8 + 2 * NT_MAX * 8 + 2 * NQ_MAX * 8 + 10 * 8 + 8 + 8)
        d2h = P * (8 + 8 + 4)
        e2e = {"value": world * P * args.steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2 / args.steps,
               "note": "pinned %d-point host staging buffer re-sent %d x per step per GPU; the upload of block i+1 "
                       "(copy stream) overlaps the evaluation of block i" % (stage_n, -(-P // stage_n))}

    # ---- roofline of the dominant kernel (K1 projection; launched once per %d-point chunk)
    n_launch_k1 = args.steps * (-(-P // min(args.chunk, P)))
    k1_ms = tm["ms_project"] / n_launch_k1
    pts_per_launch = P / (-(-P // min(args.chunk, P)))
    # algorithmic HBM bytes per point: Src + ddSrc read once, one partial-sum block [6][96] written per 24 wavenumbers
    bytes_pt = float(np.mean(W["n_tau"].astype(np.float64) * 3 * W["n_k"] * 8 * 2)) + (2910 / 24.0) * 6 * 96 * 8
    trip_pt = triples / P
    flop_per_triple = 13  # cubic j_l interpolation: 1 mul + 3 FMA; 3 FMA accumulations (T, E, lensing potential)
    # ncu --set full capture of this kernel (profiles/r01_project_v4_final_ncu_full.txt, 64 points): DRAM read + write and
    # shared-memory wavefronts per point; the kernel's own bound is the shared-memory pipe (1 wavefront / clock / SM)
    NCU_DRAM_BYTES_PER_POINT = (353.15e6 + 34.90e6) / 64
    NCU_SMEM_WAVEFRONTS_PER_POINT = 1396435361 / 64
    sm_clock_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
    roof = {"kernel": "project4_kernel (K1+K2 fused: line-of-sight projection + partial k-contraction)",
            "bound": "hbm", "achieved": bytes_pt * pts_per_launch / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "peak_source": peak_src, "traffic": NCU_DRAM_BYTES_PER_POINT * pts_per_launch,
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of a 64-point launch, scaled per point",
            "smem_lsu": {"achieved": NCU_SMEM_WAVEFRONTS_PER_POINT * pts_per_launch / (k1_ms * 1e-3) / 1e9,
                         "peak": 148 * sm_clock_hz / 1e9, "unit": "Gwavefronts/s",
                         "wavefronts_per_point": NCU_SMEM_WAVEFRONTS_PER_POINT,
                         "note": "l1tex__data_pipe_lsu_wavefronts_mem_shared per point (ncu) / launch time, against one "
                                 "128-byte wavefront per clock per SM"},
            "algorithmic_bytes_per_point": bytes_pt, "points_per_launch": pts_per_launch, "ms_per_launch": k1_ms,
            "share_of_step": tm["ms_project"] / tm["ms_total"],
            "note": "the kernel is NOT HBM-bound (SURVEY 8d): DRAM traffic equals the algorithmic bytes; its bounds are the "
                    "shared-memory pipe (table gather out of the ring) and FP64 issue; see smem_lsu, fp64, table_gather",
            "fp64": {"achieved": trip_pt * flop_per_triple * pts_per_launch / (k1_ms * 1e-3) / 1e12,
                     "peak": dfma_peak, "peak_dmma": dmma_peak, "unit": "TFLOP/s", "peak_source": "measured live "
                     "(cb200_measure_fp64_peaks: DFMA / DMMA micro-kernels)", "flop_per_triple": flop_per_triple,
                     "triples_per_point": trip_pt},
            "table_gather": {"achieved": trip_pt * 32 * pts_per_launch / (k1_ms * 1e-3) / 1e9, "unit": "GB/s",
                             "bytes_per_triple": 32}}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["smem_lsu"]["frac"] = roof["smem_lsu"]["achieved"] / roof["smem_lsu"]["peak"]
    roof["fp64"]["frac"] = roof["fp64"]["achieved"] / max(dfma_peak, 1e-9)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r, dt, _ = cpu_oracle_rate(W, T, data, args.cpu_sample)
        cpu = {"value": r, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d points of the same workload (%.1f s), C++/OpenMP restatement of the reference (oracle/); "
                         "the Fortran reference cannot be built (no Fortran compiler)" % (args.cpu_sample, dt)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(tm["n_launches"]), "roofline": roof,
               "cpu_baseline": cpu,
               "phase_ms_per_step": {k: tm[k] / args.steps for k in
                                     ["ms_spline", "ms_project", "ms_contract", "ms_interp", "ms_lens", "ms_like"]}}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
