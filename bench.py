#!/usr/bin/env python
"""Benchmark of the Sella saddle-search inner loop (BASELINE.json metric:
optimizer steps/sec, batched 3N-DOF Davidson + trust-region step).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]      # reference CPU path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU

Workload (config.workload): `--batch` (default 1024) independent order-1 saddle
searches PER GPU on the synthetic indefinite-quadratic surfaces of SURVEY.md 8d with
3N = `--n` (default 384) Cartesian degrees of freedom; Sella settings: quasi-Newton
step model, trust-radius restricted step, TS-BFGS updates, finite-difference
Jacobi-Davidson (jd0, gamma=0.1, eta=1e-4) capped at `--kdiag` (5) vectors and
re-run every `--diag-every` (3) steps through Sella's own `diag_every_n` switch (a
fixed quadratic surface never trips the "lowest mode went positive" test, so the
default policy would leave Davidson out of the timed region altogether).

A step = one Sella.step for every system of the batch: restricted-step solve,
surface evaluation, rho / trust-radius update, Hessian update and, when the policy
fires, a Davidson diagonalisation + block update.  value = system-steps per second
summed over all GPUs (weak scaling: per-GPU batch fixed).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "optimizer steps/sec (batched 3N-DOF Davidson+TR)"
UNIT = "system-steps/s"
# committed `ncu --set full` captures (raw page, csv) the roofline `traffic` figures are read from:
# (systems per GPU, 3N) -> file under profiles/
NCU_CAPTURES = {(1024, 384): dict(hv=("ncu_full_r2_hv.csv", "ncu_full_r1_a_hv.csv"),
                                  eigen=("ncu_full_r2_eigen_update.csv", "ncu_full_r1_h_eigen_update.csv"))}


def ncu_traffic(files, kernel_regex, per="launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernels matching `kernel_regex` in the first
    of `files` that exists under profiles/ (ncu --page raw --csv: a header row, a units row, one row
    per profiled launch).  per='launch': mean over the matching launches; per='group': summed over the
    distinct kernel names (one launch of each).  Returns (bytes, file) or (None, None)."""
    import csv
    import re
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for name in files:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.isfile(path):
            continue
        with open(path, newline="") as fh:
            rows = list(csv.reader(fh))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        try:
            kn, rd, wr = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        except ValueError:
            continue
        by_kernel = {}
        for r in rows[2:]:
            if len(r) <= max(kn, rd, wr) or not re.search(kernel_regex, r[kn]):
                continue
            tot = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
            by_kernel.setdefault(r[kn], []).append(tot)
        if not by_kernel:
            continue
        if per == "max":                    # the largest launch (the full n x n pass among rectangular ones)
            return max(max(v) for v in by_kernel.values()), name
        means = [sum(v) / len(v) for v in by_kernel.values()]
        return (sum(means) if per == "group" else sum(means) / len(means)), name
    return None, None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="quadratic", choices=["quadratic", "emt-slab", "emt-cluster"],
                    help="quadratic: synthetic indefinite-quadratic PES (SURVEY 8d; the default, C4/C5 family); "
                         "emt-slab: C3-style 128-atom Cu(111) slabs, bottom half fixed, EMT-form surface; "
                         "emt-cluster: C2-style 64-atom Cu clusters, translation + rotation projection, EMT-form surface")
    ap.add_argument("--batch", type=int, default=None, help="systems per GPU (default 1024; 256 for emt-cluster)")
    ap.add_argument("--n", type=int, default=None, help="3N degrees of freedom (default 384; 192 for emt-cluster)")
    ap.add_argument("--rs", default="tr")
    ap.add_argument("--method", default="prfo", help="step model: prfo (Sella's default for saddles), rfo, qn")
    ap.add_argument("--kdiag", type=int, default=5)
    ap.add_argument("--diag-every", type=int, default=3)
    ap.add_argument("--proj-rot", dest="proj_rot", action="store_true", default=True,
                    help="emt-cluster only (default on): also hold the three rotation coordinates, the reference's "
                         "default projection for non-periodic systems (peswrapper.py:246-253)")
    ap.add_argument("--no-proj-rot", dest="proj_rot", action="store_false",
                    help="emt-cluster only: centre of mass held, rotations left free (linear constraints only)")
    ap.add_argument("--cpu-systems", type=int, default=0, help="reference sample size (0 = 2 per core)")
    ap.add_argument("--parity-systems", type=int, default=4,
                    help="systems of the timed batch re-run on the host for the `parity` key (0 = off)")
    ap.add_argument("--long-steps", type=int, default=100,
                    help="length of the extra run whose per-decile step times go into `long_run` (0 = off)")
    ap.add_argument("--spectrum", default=None, choices=[None, "compact", "dense"],
                    help="engine representation of the Hessian spectrum (default: the engine's own choice)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--internal", action="store_true",
                    help="emt-slab only: BASELINE config C3 as named -- the search runs in internal coordinates "
                         "(nearest-neighbour bonds + the fixed atoms' Cartesian coordinates, MaxInternalStep, geodesic "
                         "steps) on sella_b200.batched_internal.BatchedInternalSella")
    ap.add_argument("--inexact-geodesic", action="store_true",
                    help="--internal only: the reference's exact_geodesic=False (B+ frozen along the geodesic)")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 256 if a.workload == "emt-cluster" else 1024
    if a.n is None:
        a.n = 192 if a.workload == "emt-cluster" else 384
    if a.workload == "emt-cluster" and "--kdiag" not in sys.argv:
        a.kdiag = 2                                  # BASELINE.json C2: Davidson k=2
    return a


def emt_problem(args, first, count):
    """Geometries and linear constraints of the EMT workloads (host arrays; shared by both arms)."""
    from sella_b200.synthetic import fcc_cluster, fcc111_slab
    nat = args.n // 3
    if args.workload == "emt-cluster":
        x0 = np.stack([fcc_cluster(nat, seed=first + i).ravel() for i in range(count)])
        C = np.zeros((3, args.n))
        for d in range(3):
            C[d, d::3] = 1.0 / nat                   # the reference's default translation projection
        return x0, C, None, (False, False, False)
    ny = 2
    nl = 8
    nx = nat // (2 * ny * nl)
    if 2 * nx * ny * nl != nat:
        raise SystemExit("emt-slab needs 3N = 96 * k (k surface cells x 2 x 8 layers)")
    geo = [fcc111_slab(nx, ny, nl, seed=first + i) for i in range(count)]
    ideal = fcc111_slab(nx, ny, nl)[0]
    fixed = np.nonzero(ideal[:, 2] < ideal[:, 2].mean())[0]            # bottom half, Constraints.fix_translation
    C = np.zeros((3 * len(fixed), args.n))
    for r, i in enumerate(fixed):
        for d in range(3):
            C[3 * r + d, 3 * i + d] = 1.0
    return np.stack([g[0].ravel() for g in geo]), C, geo[0][1], geo[0][2]


def workload(args):
    common = dict(batch_per_gpu=args.batch, dof=args.n, rs=args.rs, method=args.method, davidson_maxiter=args.kdiag,
                  diag_every_n=args.diag_every, eta=1e-4, gamma=0.1)
    tail = ("Cartesian, order=1, %s + %s restricted step, TS-BFGS, jd0 Davidson gamma=0.1 maxiter=%d, diag_every_n=%d"
            % (args.method, args.rs, args.kdiag, args.diag_every))
    if args.workload == "quadratic":
        return dict(workload="batch=%d/GPU x 3N=%d synthetic indefinite-quadratic PES (SURVEY 8d), %s"
                             % (args.batch, args.n, tail),
                    l2_policy="working set %.1f GB per GPU >> 126 MB L2 (no flush needed)"
                              % (args.batch * args.n * args.n * 8 * 4 / 1e9), **common)
    what = ("%d-atom Cu(111) slabs (rattled 0.05 A), bottom half held by fix_translation (%d linear constraints)"
            % (args.n // 3, args.n // 2)) if args.workload == "emt-slab" else \
           ("%d-atom Cu clusters (fcc ball + 0.05 A rattle), centre of mass held (3 linear constraints)%s"
            % (args.n // 3, " + the three rotation coordinates held (position-dependent; the reference's default "
                            "projection)" if args.proj_rot else ", rotations free (--no-proj-rot)"))
    return dict(workload="batch=%d/GPU x 3N=%d EMT-form surface on the device, %s, %s" % (args.batch, args.n, what, tail),
                l2_policy="working set %.1f GB per GPU >> 126 MB L2 (no flush needed)"
                          % (args.batch * args.n * args.n * 8 * 4 / 1e9), **common)


# ----------------------------------------------------------------------------- CPU reference
def _limit_threads(threads):
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["OPENBLAS_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=threads)
    except Exception:
        return None


def _gen_chunk(job):
    """Host generation of quadratic systems first..first+count-1 (numpy recipe of SURVEY 8d): the
    SAME function and seeds feed the GPU arm, the CPU arm and the in-run parity check."""
    first, count, n = job
    _limit_threads(1)
    from sella_b200.synthetic import quadratic_batch
    return quadratic_batch(count, n, first=first)


class _CpuSearch:
    """One saddle search on the host with the reference's algorithm: the reference's OWN Sella + PES
    classes (kind 'reference'; loaded piecewise by oracle/ref_loader.py from /root/reference or from its
    staged copy baseline/_ref/) when they can serve the workload, else the oracle port (kind 'port')."""

    def __init__(self, wl, index, n, rs, method, kdiag, diag_every, proj_rot, prefer_reference=True):
        from sella_b200.synthetic import quadratic_system, quadratic_func
        C = c = None
        if wl == "quadratic":
            A, xs, x0 = quadratic_system(index, n)
            func = quadratic_func(A, xs)
        else:
            from oracle.emt import emt_func
            ns = argparse.Namespace(workload=wl, n=n)
            X0, C, cell, pbc = emt_problem(ns, index, 1)
            x0 = X0[0]
            func = emt_func(cell, pbc)
            c = C @ x0
        self.kind = "port"
        nonlinear = proj_rot and wl == "emt-cluster"
        ref = None
        if prefer_reference and not nonlinear:
            try:
                from oracle import ref_loader, ref_harness
                if ref_loader.available():
                    ref = ref_loader.load()
            except Exception:
                ref = None
        if ref is not None:
            self.dyn = ref_harness.make_reference_sella(ref, func, x0, C, c, method=method, rs=rs,
                                                        diag_every_n=diag_every)
            self.dyn.diagkwargs["maxiter"] = kdiag      # PES.diag(maxiter=...), peswrapper.py:508
            self.pes = self.dyn.pes
            self.kind = "reference"
        else:
            from oracle.driver import SaddleSearch
            if nonlinear:
                from oracle.pes import NonlinearPES
                self.pes = NonlinearPES(func, x0, dict(rotation_ref=x0.reshape(-1, 3)), np.zeros(3), C, c)
            else:
                from oracle.pes import CartesianPES
                self.pes = CartesianPES(func, x0, C, c)
            self.dyn = SaddleSearch(self.pes, method=method, rs=rs, diag_maxiter=kdiag, diag_every_n=diag_every)

    def step(self):
        self.dyn.step()

    def x(self):
        return np.asarray(self.pes.get_x(), dtype=float).copy()

    def lowest_eval(self):
        ev = self.pes.H.evals
        return None if ev is None else float(ev[0])


def _cpu_worker(job):
    """Runs the searches `indices` for warm + steps steps with `threads` BLAS threads; returns
    (steps done, seconds in the timed part, kind)."""
    indices, n, rs, kdiag, diag_every, warm, steps, threads, method, wl, proj_rot = job
    limiter = _limit_threads(threads)
    runs = [_CpuSearch(wl, i, n, rs, method, kdiag, diag_every, proj_rot) for i in indices]
    done = 0
    alive = []
    for o in runs:
        try:
            for _ in range(warm):
                o.step()
            alive.append(o)
        except Exception:           # "Restricted step failed to converge!" / LAPACK failure inside RFO:
            pass                    # the reference would abort this search; it is dropped from the sample
    t0 = time.perf_counter()
    for o in alive:
        try:
            for _ in range(steps):
                o.step()
                done += 1
        except Exception:
            pass
    dt = time.perf_counter() - t0
    del limiter
    return done, dt, (runs[0].kind if runs else "port")


def _parity_worker(job):
    """In-run parity: system `index` of the timed batch taken `nsteps` steps on the host; returns
    (final positions, lowest eigenvalue of the approximate Hessian, kind, error text)."""
    index, n, rs, kdiag, diag_every, nsteps, method, wl, proj_rot = job
    _limit_threads(1)
    try:
        o = _CpuSearch(wl, index, n, rs, method, kdiag, diag_every, proj_rot)
        for _ in range(nsteps):
            o.step()
        return o.x(), o.lowest_eval(), o.kind, None
    except Exception as exc:
        return None, None, "port", repr(exc)


def cpu_reference(args, warm, steps, per_proc=2):
    """The reference's algorithm on the host cores, timed on a FIXED sample: systems 0 .. cores*per_proc-1
    of rank 0's batch (the same inputs the GPU arm runs), one single-threaded process per core, rate =
    median per-process rate x processes; beside it one process with all BLAS threads on system 0.  The
    better of the two is reported."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    if args.cpu_systems:
        per_proc = max(1, args.cpu_systems // cores)
    common = (args.n, args.rs, args.kdiag, args.diag_every, warm, steps)
    tail = (args.method, args.workload, args.proj_rot)
    done, dt, kind = _cpu_worker(([0], ) + common + (cores,) + tail)
    rate_mt = done / dt if dt > 0 else 0.0
    ctx = mp.get_context("spawn")
    jobs = [(list(range(i * per_proc, (i + 1) * per_proc)),) + common + (1,) + tail for i in range(cores)]
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    rates = sorted(r[0] / r[1] for r in res if r[1] > 0 and r[0] > 0)
    rate_mp = (rates[len(rates) // 2] if len(rates) % 2 else 0.5 * (rates[len(rates) // 2 - 1] + rates[len(rates) // 2])) \
        * cores if rates else 0.0
    best = max(rate_mt, rate_mp)
    mode = "%d procs x 1 BLAS thread" % cores if rate_mp >= rate_mt else "1 proc x %d BLAS threads" % cores
    return dict(value=best, unit=UNIT, cores=cores, kind=kind,
                sample="systems 0..%d of the GPU arm's batch x (%d warm-up + %d timed) steps, %s (median per-process "
                       "rate x processes); other mode: %.3g"
                       % (cores * per_proc - 1 if rate_mp >= rate_mt else 0, warm, steps, mode, min(rate_mt, rate_mp)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm, steps = min(args.warmup, 2), min(args.steps, 4)     # bounded sample of the same workload
    t0 = time.perf_counter()
    if args.internal:
        return _run_reference_internal(args, t0)
    cb = cpu_reference(args, warm, steps)
    wall = time.perf_counter() - t0
    what = ("the reference's own Sella + PES classes (unmodified files, oracle/ref_loader.py)" if cb["kind"] == "reference"
            else "CPU restatement (oracle/) of the reference path")
    out = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm,
               ms_per_step=1e3 * args.batch / cb["value"], higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f64", data="synthetic", impl="reference", config=workload(args),
               cpu_baseline=cb,
               e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               note="%s on host cores; ms_per_step is the extrapolated time for one step of the whole "
                    "%d-system batch; wall %.1fs" % (what, args.batch, wall))
    print(json.dumps(out))


def _internal_cpu_baseline(args, first, exact, warm=1, steps=2):
    """C3 as named on the host cores: one system per core, the oracle InternalPES loop with the reference's
    integrator (scipy LSODA); rate = median per-process rate x processes."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    jobs = [([first + i], args.n, args.kdiag, args.diag_every, warm, steps, args.method, "lsoda", exact)
            for i in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_internal_cpu_worker, jobs)
    rates = sorted(r[0] / r[1] for r in res if r[1] > 0)
    med = rates[len(rates) // 2]
    return dict(value=med * cores, unit=UNIT, cores=cores, kind="port",
                sample="systems 0..%d of the GPU arm's batch x (%d warm-up + %d timed) steps of the oracle InternalPES "
                       "loop (reference algorithm, scipy LSODA geodesic), %d procs x 1 BLAS thread, median per-process "
                       "rate x processes" % (cores - 1, warm, steps, cores))


def _run_reference_internal(args, t0):
    cb = _internal_cpu_baseline(args, 0, not args.inexact_geodesic)
    wall = time.perf_counter() - t0
    cfg = dict(workload="batch=%d/GPU x 3N=%d EMT-form surface, Cu(111) slabs in INTERNAL coordinates (nearest-neighbour "
                        "bonds + the fixed atoms' Cartesian coordinates), %s + MaxInternalStep, TS-BFGS, jd0 Davidson "
                        "maxiter=%d, diag_every_n=%d" % (args.batch, args.n, args.method, args.kdiag, args.diag_every),
               batch_per_gpu=args.batch, dof=args.n, rs="mis", method=args.method, davidson_maxiter=args.kdiag,
               diag_every_n=args.diag_every, coordinates="internal")
    print(json.dumps(dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=2, warmup=1,
                          ms_per_step=1e3 * args.batch / cb["value"], higher_is_better=True, scaling="weak",
                          vs_baseline=None, dtype="f64", data="synthetic", impl="reference", config=cfg, cpu_baseline=cb,
                          e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                          note="CPU restatement (oracle/internal_pes.py) of the reference's InternalPES path on host "
                               "cores; wall %.1fs" % wall)))


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi polled every 200 ms (B200_PROFILING.md recipe) from BEFORE the warm-up: its start-up
    and every query stall kernel launches for a millisecond or more (polling at 20 ms cost 25 % of the
    measured rate), which must not dominate a timed region of a few tens of ms.  Samples are
    time-stamped; those within 0.25 s of the timed region are reported."""

    def __init__(self, index):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        import datetime
        rows = []
        for line in out.splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.25 <= r[0] <= self.t1 + 0.25]
        window = "timed region +- 0.25 s (200 ms polling)"
        if not inside:                      # clock skew between time.time() and nvidia-smi: fall back to the last samples
            inside, window = rows[-5:], "last samples (no time-stamp inside the timed region)"
        sm = [r[1] for r in inside]
        mx = [r[2] for r in inside]
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), window=window, reasons=sorted(reasons))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_generated_batch(first, count, n, dev, world=1, chunk=64):
    """Systems first..first+count-1 of the synthetic family, generated on the HOST by the numpy recipe
    (seed 1000 + index, sella_b200/synthetic.py) in a pool of workers and uploaded chunk by chunk: the
    CPU arm and the in-run parity check run the very same systems."""
    import multiprocessing as mp
    import torch
    A = torch.empty((count, n, n), dtype=torch.float64, device=dev)
    xs = torch.empty((count, n), dtype=torch.float64, device=dev)
    x0 = torch.empty((count, n), dtype=torch.float64, device=dev)
    workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    per = max(1, min(chunk, (count + workers - 1) // workers))
    jobs = [(first + lo, min(per, count - lo), n) for lo in range(0, count, per)]
    with mp.get_context("spawn").Pool(workers) as pool:
        for (f0, cnt, _), (Ac, xc, x0c) in zip(jobs, pool.imap(_gen_chunk, jobs)):
            lo = f0 - first
            A[lo:lo + cnt] = torch.from_numpy(Ac).to(dev)
            xs[lo:lo + cnt] = torch.from_numpy(xc).to(dev)
            x0[lo:lo + cnt] = torch.from_numpy(x0c).to(dev)
    return A, xs, x0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sella_b200 import _lib, kernels as K
    from sella_b200.batched import BatchedSella, QuadraticSurface

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the single JSON line: NCCL's own log (whatever level the driver asked for via
        # NCCL_DEBUG; WARN if it did not) goes to a file next to the bench, one per rank
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        if "NCCL_DEBUG_FILE" not in os.environ:
            logdir = os.path.join(ROOT, "gpurun_out")
            try:
                os.makedirs(logdir, exist_ok=True)
                os.environ["NCCL_DEBUG_FILE"] = os.path.join(logdir, "nccl_bench.%h.%p.log")
            except OSError:
                pass
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.get_lib()
    lib.sb_launch_count.restype = __import__("ctypes").c_longlong

    b, n = args.batch, args.n
    cons = None
    from sella_b200.sharding import shard_range, max_over_ranks
    # weak scaling: the global batch is world * b systems, rank r owns the contiguous block shard_range gives it
    # (system index = seed index: both arms and the parity check draw system i from RandomState(1000 + i))
    first, last = shard_range(world * b, rank, world)
    assert last - first == b
    if args.workload == "quadratic":
        A, xs, x0 = host_generated_batch(first, b, n, dev, world)
        surf = QuadraticSurface(A, xs)
    else:
        from sella_b200.emt import EMTSurface
        X0, C, cell, pbc = emt_problem(args, first, b)
        x0 = torch.from_numpy(X0).to(dev)
        surf = EMTSurface(b, n // 3, dev, cell=cell, pbc=pbc)
        cons = (C, None)
        if args.proj_rot and args.workload == "emt-cluster":     # the reference's default for molecules
            from sella_b200.internal import BatchedInternals
            cons = (C, None, BatchedInternals(n // 3, rotation_ref=X0.reshape(b, n // 3, 3)), None)

    extra = {} if args.spectrum is None else dict(spectrum=args.spectrum)

    def make():
        return BatchedSella(surf, x0, method=args.method, rs=args.rs, diag_maxiter=args.kdiag,
                            diag_every_n=args.diag_every, kcap=max(8, args.kdiag + 1), constraints=cons, **extra)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one-off, untimed: a throw-away engine takes 7 steps so that every kernel variant of a step and of a
    # re-diagonalisation is loaded (CUDA loads kernels lazily on first use; a re-diagonalisation first
    # happens at step 5, i.e. inside the timed region, where it cost ~20 % of the measured rate)
    pre = make()
    for _ in range(7):
        pre.step()
    del pre
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    # ---------------- device-resident run: `value`
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("SB_NO_SAMPLER")) else None
    eng = make()
    for _ in range(args.warmup):
        eng.step()
    barrier()
    if sampler:
        sampler.begin()
    l0 = lib.sb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = bool(os.environ.get("SB_PROFILER_RANGE"))   # ncu --profile-from-start off: timed loop only
    if prof_range:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for _ in range(args.steps):
        eng.step()
    ev1.record()
    barrier()
    if prof_range:
        torch.cuda.cudart().cudaProfilerStop()
    if sampler:
        sampler.end()
    ms = ev0.elapsed_time(ev1)
    launches = lib.sb_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    st = eng.status.cpu().numpy()
    flagged = {name: int(((st & bit) != 0).sum()) for name, bit in
               (("mgs_maxiter", 1), ("restricted_step_noconv", 2), ("eigh_noconv", 4), ("davidson_cap", 8),
                ("singular", 16), ("davidson_stall", 32)) if ((st & bit) != 0).any()}
    steps_done = int(eng.nsteps.sum().item()) - b * args.warmup
    assert steps_done == b * args.steps
    ms_max = max_over_ranks(ms, device=dev)
    value = world * b * args.steps / (ms_max / 1e3)

    # ---------------- in-run parity: systems 0..P-1 of THIS run's timed batch vs the reference algorithm on
    # the host for the same warm-up + timed steps (same seeds -> same A, x*, x0)
    parity = None
    if rank == 0 and args.parity_systems > 0:
        import multiprocessing as mp
        npar = min(args.parity_systems, b)
        nst = args.warmup + args.steps
        engp, horizon = eng, "the timed run itself"
        if args.workload != "quadratic" and nst > 12:
            # a non-quadratic surface amplifies rounding differences step by step (a lowest eigenvalue crossing
            # zero turns 1e-9 into 1e-5 within a few steps, in the reference against its own restatement too):
            # the EMT workloads are compared over the first 12 steps of a fresh engine on the same systems
            nst, horizon = 12, "a fresh engine on the same systems, first 12 steps"
            engp = make()
            for _ in range(nst):
                engp.step()
        jobs = [(first + i, n, args.rs, args.kdiag, args.diag_every, nst, args.method, args.workload, args.proj_rot)
                for i in range(npar)]
        with mp.get_context("spawn").Pool(min(npar, os.cpu_count() or 1)) as pool:
            res = pool.map(_parity_worker, jobs)
        xg = engp.x[:npar].cpu().numpy()
        lg = engp.lowest_evals()[:npar].cpu().numpy()
        dxs, dls, errs = [], [], []
        for i, (xr, lr, kind, err) in enumerate(res):
            if xr is None:
                errs.append("system %d: %s" % (i, err))
                continue
            dxs.append(float(np.abs(xg[i] - xr).max()))
            if lr is not None:
                dls.append(float(abs(lg[i] - lr) / max(abs(lr), 1e-300)))
        parity = dict(systems=npar, steps=nst, max_dx=max(dxs) if dxs else None,
                      max_rel_lam=max(dls) if dls else None, checker=res[0][2] if res else None,
                      compared=len(dxs), errors=errs, horizon=horizon,
                      note="max |x_gpu - x_cpu| and relative difference of the lowest eigenvalue of the approximate "
                           "Hessian after `steps` steps, systems 0..%d of the timed batch against the reference "
                           "algorithm on the host (checker: the reference's own classes, or the oracle port)" % (npar - 1))
        if engp is not eng:
            del engp

    # ---------------- long run: ms per batch step by decile of a >= 100-step search (fresh engine; the
    # headline above stays the driver's K/W) -- shows whether the step cost drifts as the Hessian model
    # accumulates rank
    long_run = None
    if args.long_steps > 0:
        engl = make()
        nl_ = max(10, args.long_steps // 10 * 10)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        barrier()
        evs[0].record()
        rows = []
        for d in range(10):
            for _ in range(nl_ // 10):
                engl.step()
            evs[d + 1].record()
            rows.append(engl.rank_bound())
        barrier()
        dec = [evs[d].elapsed_time(evs[d + 1]) / (nl_ // 10) for d in range(10)]
        long_run = dict(steps=nl_, decile_ms_per_step=dec, rank_rows_max=engl.rank_bound(), rows_at_decile_end=rows,
                        note="ms per step of the whole batch, mean over each tenth of the run; rows_at_decile_end = "
                             "largest number of explicit eigenpairs r any system holds (one or two more per step, up "
                             "to 2k more per diagonalisation, capped at 3N); every pass over the eigenvectors costs "
                             "O(r n) and the rotation of the eigenvector rows 2 r^2 n flops per rank-one term (fp64 "
                             "tensor-core GEMM), so the step cost climbs until r reaches 3N and is flat afterwards")
        del engl
        torch.cuda.empty_cache()

    # ---------------- end-to-end through host buffers: `e2e`
    # the caller owns positions on the host (as ASE does): every step uploads the
    # batch of positions from pinned memory and reads back new positions, energies and
    # the convergence measure.
    eng2 = make()
    hx = torch.empty((b, n), dtype=torch.float64).pin_memory()
    hx.copy_(x0.cpu())
    hf = torch.empty(b, dtype=torch.float64).pin_memory()
    hfmax = torch.empty(b, dtype=torch.float64).pin_memory()

    def host_step():
        eng2.x.copy_(hx, non_blocking=True)
        eng2.step()
        eng2.converged(0.0)
        hx.copy_(eng2.x, non_blocking=True)
        hf.copy_(eng2.f, non_blocking=True)
        hfmax.copy_(eng2.fmax, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        host_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        host_step()
    e1.record()
    barrier()
    e2e_value = world * b * args.steps / (max_over_ranks(e0.elapsed_time(e1), device=dev) / 1e3)
    e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=world * b * n * 8,
               d2h_bytes_per_step=world * (b * n * 8 + 2 * b * 8))

    # ---------------- per-kernel timings (CUDA events on the launching stream)
    peak, peak_src = peaks()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        z.record()
        torch.cuda.synchronize()
        return a.elapsed_time(z) / reps

    xv = eng.g.view(b, 1, n)
    yv = torch.empty_like(xv)
    Mhv = surf.A if args.workload == "quadratic" else eng.B.contiguous()      # any resident [b, n, n] matrix
    hv_ms = timed(lambda: K.hv_ld(Mhv, xv, yv, 1), 20)
    hv_bytes = b * 8 * (n * n + 2 * n)
    hv_gbs = hv_bytes / (hv_ms * 1e-3) / 1e9
    # profiled pass (not part of `value`): CUDA events around the eigen-update of the plain steps
    import ctypes
    eng.prof = {}
    compact = bool(getattr(eng, "compact", False))
    parts = []
    if not compact:
        lib.sb_secular_timing(None, 1)
    t3 = (ctypes.c_float * 3)()
    rows_before = float(eng.mrows.double().mean().item()) if compact else float(n)
    for _ in range(6):
        nd0 = eng.ndiag
        eng.step()
        if not compact and eng.ndiag == nd0 and lib.sb_secular_timing(t3, -1) == 0:   # plain step: the rank-2 update
            parts.append(tuple(t3))
    if not compact:
        lib.sb_secular_timing(t3, 0)
    rows_after = float(eng.mrows.double().mean().item()) if compact else float(n)
    prof = eng.prof_summary()
    eng.prof = None
    step_ms = ms_max / args.steps
    sec_cnt, sec_ms = prof.get("eigen_update_k1" if compact else "secular_update_k1", (0, 0.0))
    part_ms = [sum(p[i] for p in parts) / len(parts) for i in range(3)] if parts else [0.0, 0.0, 0.0]
    caps = NCU_CAPTURES.get((b, n), {})
    traffic, traffic_src = ncu_traffic(caps.get("eigen", ()), r"cluster_reflect|secular_update|secular_apply|cluster_qr|append_", per="group")
    hv_traffic, hv_traffic_src = ncu_traffic(caps.get("hv", ()), r"hv_tma_kernel<1>", per="max")
    if compact:
        # one rank-2 eigen-update on r explicit rows: Z = VR P, W1 = VR^T Z, VR Qc, VR^T D2 (4 reads of the
        # r x n block) + the secular rotation (read + write of the rows it changes, <= r)
        r_mean = 0.5 * (rows_before + rows_after)
        sec_bytes = b * 8 * n * 6.0 * r_mean
        what = ("eigen-update of the compact spectrum after the rank-2 secant update (r = %.0f explicit rows of %d): "
                "lowrank_factor + 4 rectangular H.V passes + append_a/b + secular_update_kernel" % (r_mean, n))
        note = ("algorithmic bytes = 8 n (4 r + 2 r): four reads of the r x n block of explicit eigenvectors and one "
                "read + write by the rotation; the rotation itself costs 2 r^2 n flops per rank-one term (fp64 FMA "
                "bound once r is a few hundred), see DESIGN.md section 5")
    else:
        sec_bytes = b * 8 * 2 * n * n               # every eigenvector read once and written once
        what = ("eigen-update of (evals, Vt) after the rank-2 secant update: cluster_qr_kernel + "
                "cluster_reflect_kernel<2> + secular_update_kernel<4> (dense representation)")
        note = ("algorithmic bytes = read+write the eigenvector matrix once (2*n^2*8 per system); cluster_reflect "
                "streams the degenerate cluster's rows; secular_update_kernel is latency-bound, see DESIGN.md section 5")
    sec_gbs = sec_bytes / (sec_ms * 1e-3) / 1e9 if sec_ms else 0.0
    roofline_eig = dict(kernel=what, bound="hbm", achieved=sec_gbs, peak=peak, unit="GB/s", frac=sec_gbs / peak,
                        traffic=traffic, traffic_source=("profiles/" + traffic_src) if traffic_src else None,
                        ms_per_launch=sec_ms, share_of_step=min(1.0, sec_ms / step_ms) if step_ms else None,
                        bytes_per_launch=sec_bytes, peak_source=peak_src,
                        explicit_rows=dict(before=rows_before, after=rows_after, of=n),
                        kernels_ms=dict(cluster_qr_kernel=part_ms[0], cluster_reflect_kernel=part_ms[1],
                                        secular_update_kernel=part_ms[2]) if not compact else None,
                        note=note)
    # the dominant kernel family of a step by GPU time (profiles/launches_r1_j.txt: hv_tma / hvt_tma
    # variants = 41 %): the batched H.V pass, used for V^T g, V c, B s, the surface and Z = Vt P
    roofline = dict(kernel="hv_tma_kernel<1> (batched H.V: TMA bulk-copy pipeline, one pass over a [b, n, n] matrix; "
                           "~9 such passes per step, 41 % of GPU time with its <2>/<4>/transposed variants)",
                    bound="hbm", achieved=hv_gbs, peak=peak, unit="GB/s", frac=hv_gbs / peak,
                    frac_of_8TBs_nominal=hv_gbs / 8000.0, traffic=hv_traffic,
                    traffic_source=("profiles/" + hv_traffic_src) if hv_traffic_src else None, ms_per_launch=hv_ms,
                    bytes_per_launch=hv_bytes, peak_source=peak_src,
                    note="read-dominated (the measured copy peak is a read+write figure, hence fractions close to "
                         "1); algorithmic bytes = 8 (n^2 + 2 n) per system")
    # ---------------- fp64 compute rooflines: measured DFMA / DMMA peaks, and the rotation GEMM against them
    fp64 = None
    roofline_rot = None
    try:
        tf = ctypes.c_double(0.0)
        scratch = torch.empty(lib.sb_device_sms() * 8 * 256, dtype=torch.float64, device=dev)
        peaks64 = {}
        for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops")):
            best = 0.0
            for cps in (4, 8):
                if lib.sb_fp64_peak(kind, 20000, cps, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(tf)) == 0:
                    best = max(best, tf.value)
            peaks64[name] = best
        fp64 = dict(peaks64, how="register-resident DFMA / DMMA m8n8k4 loops on all SMs (sb_fp64_peak), best of 4 and 8 "
                                 "CTAs of 256 threads per SM")
        if compact:
            rr = max(16, int(round(rows_after)))
            msf = ctypes.c_float(0.0)
            sp = eng.spB
            if lib.sb_secular_apply_bench(ctypes.c_void_p(sp.Vt.data_ptr()), ctypes.c_void_p(eng.qwork.data_ptr()),
                                          ctypes.c_void_p(eng.eig_ws.work.data_ptr()),
                                          ctypes.c_void_p(eng.sec_aux.data_ptr()), rr, n, ctypes.c_longlong(n * n), b, 10,
                                          ctypes.byref(msf), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0:
                fl = 2.0 * rr * rr * n * b
                ach = fl / (msf.value * 1e-3) / 1e12
                pk = max(peaks64.values())
                roofline_rot = dict(kernel="secular_apply_kernel: rotation of the eigenvector rows of one rank-one term, "
                                           "batched fp64 tensor-core GEMM (DMMA m8n8k4), r = %d rows of %d" % (rr, n),
                                    bound="tensor", achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk if pk else None,
                                    traffic=None, ms_per_launch=msf.value, flops_per_launch=fl,
                                    peak_source="measured here (fp64_peak: max of DFMA and DMMA)")
    except Exception as exc:
        fp64 = dict(error=repr(exc))
    kernel_ms = {k: v[1] for k, v in prof.items()}
    if b * n * n <= 1024 * 768 * 768:             # a full batched eigensolve: seconds beyond this size
        kernel_ms["sb_eigh_full (direct mode only; not on the default path)"] = timed(lambda: K.eigh(Mhv), 2)

    out = None
    if rank == 0:
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                   data="synthetic", config=dict(workload(args), spectrum="compact" if compact else "dense"),
                   clocks=clocks, e2e=e2e,
                   gpu_launches=int(launches), parity=parity, long_run=long_run, roofline=roofline,
                   roofline_eigen_update=roofline_eig, roofline_rotation=roofline_rot, fp64_peak=fp64,
                   kernel_ms=kernel_ms, systems_flagged=flagged, diagonalisations=eng.ndiag,
                   note="systems_flagged: per-system status words (the batched analogue of the reference's "
                        "exceptions); restricted_step_noconv reproduces the reference's own 'Restricted step "
                        "failed to converge!' on the same inputs (see DESIGN.md)")
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_reference(args, min(args.warmup, 2), min(args.steps, 4))
            except Exception as exc:       # never lose the GPU line over the baseline leg
                out["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port",
                                           sample="failed: %r" % (exc,))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ----------------------------------------------------------------------------- internal coordinates (C3 as named)
class _SlabAtoms:
    def __init__(self, pos, cell, pbc):
        self.positions, self.cell, self.pbc = np.array(pos, dtype=float), cell, np.array(pbc)
        self.numbers = np.full(len(pos), 29)

    def __len__(self):
        return len(self.positions)


def internal_problem(args, first, count):
    """C3 as named: slabs of emt_problem, one coordinate list for the batch (nearest-neighbour bonds
    of the ideal slab + the Cartesian coordinates of the fixed atoms), constraint rows, per-system model Hessian."""
    from sella_b200.constraints import Constraints
    from sella_b200.topology import Internals
    from sella_b200.synthetic import fcc111_slab
    from oracle.intcoords import CoordinateSet
    X0, C, cell, pbc = emt_problem(args, first, count)
    nat = args.n // 3
    nx = nat // (2 * 2 * 8)
    ideal = _SlabAtoms(fcc111_slab(nx, 2, 8)[0], cell, pbc)
    cons = Constraints(ideal)
    for i in np.nonzero(ideal.positions[:, 2] < ideal.positions[:, 2].mean())[0]:
        cons.fix_translation(int(i))
    ints = Internals(ideal, cons=cons)
    ints.find_all_bonds()
    rows, _ = ints.constraint_rows()
    tr, bd, an, dh, tv = ints.lists()
    cs = CoordinateSet(nat, tr, bd, an, dh, tvecs=tv, numbers=ideal.numbers)      # host twin (model Hessian, parity)
    h0 = np.stack([cs.guess_hessian(x) for x in X0])
    return X0, cell, pbc, ints, rows, cs, h0


def _internal_cpu_worker(job):
    """Systems `indices` of the C3 batch on the host: the oracle InternalPES loop (reference algorithm,
    scipy LSODA geodesic) for warm + steps steps; returns (steps done, seconds, final positions of the first)."""
    indices, n, kdiag, diag_every, warm, steps, method, integrator, exact = job
    _limit_threads(1)
    from oracle.emt import emt_func
    from oracle.internal_pes import InternalPES
    from oracle.intcoords import CoordinateSet
    from oracle.driver import SaddleSearch
    ns = argparse.Namespace(workload="emt-slab", n=n)
    done, dt, xs = 0, 0.0, []
    for idx in indices:
        X0, cell, pbc, ints, rows, cs, h0 = internal_problem(ns, idx, 1)
        tr = ints.lists()[0]
        csc = CoordinateSet(n // 3, [tr[r] for r in rows])
        p = InternalPES(emt_func(cell, pbc), X0[0], cs, csc, integrator=integrator, exact_geodesic=exact)
        o = SaddleSearch(p, rs="mis", method=method, diag_maxiter=kdiag, diag_every_n=diag_every)
        for _ in range(warm):
            o.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step()
            done += 1
        dt += time.perf_counter() - t0
        xs.append(p.pos.copy())
    return done, dt, xs


def run_internal(args):
    import torch
    import torch.distributed as dist
    import multiprocessing as mp
    from sella_b200 import _lib, kernels as K
    from sella_b200.batched_internal import BatchedInternalSella
    from sella_b200.emt import EMTSurface
    from sella_b200.sharding import shard_range, max_over_ranks
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.get_lib()
    lib.sb_launch_count.restype = __import__("ctypes").c_longlong
    b, n = args.batch, args.n
    first, last = shard_range(world * b, rank, world)
    X0, cell, pbc, ints, rows, cs, h0 = internal_problem(args, first, b)
    exact = not args.inexact_geodesic
    surf = EMTSurface(b, n // 3, dev, cell=cell, pbc=pbc)
    x0 = torch.from_numpy(X0).to(dev)

    def make(count=None):
        # count: an engine over the first `count` systems only (the parity check); one full-batch engine holds
        # ~10 nint x nint sets per system, two of them do not fit next to each other at 1024 systems
        sf = surf if count is None else EMTSurface(count, n // 3, dev, cell=cell, pbc=pbc)
        xs = x0 if count is None else x0[:count].contiguous()
        return BatchedInternalSella(sf, xs, ints.device_coordinates(), cons_rows=rows,
                                    h0=h0 if count is None else h0[:count], method=args.method,
                                    diag_maxiter=args.kdiag, diag_every_n=args.diag_every, kcap=max(8, args.kdiag + 1),
                                    exact_geodesic=exact)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("SB_NO_SAMPLER")) else None
    eng = make()
    nint = eng.n
    for _ in range(args.warmup):
        eng.step()
    barrier()
    if sampler:
        sampler.begin()
    l0 = lib.sb_launch_count()
    ode0 = eng.ode_steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        eng.step()
    ev1.record()
    barrier()
    if sampler:
        sampler.end()
    ms = ev0.elapsed_time(ev1)
    launches = lib.sb_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    st = eng.status.cpu().numpy()
    flagged = {name: int(((st & bit) != 0).sum()) for name, bit in
               (("mgs_maxiter", 1), ("restricted_step_noconv", 2), ("eigh_noconv", 4), ("davidson_cap", 8),
                ("singular", 16), ("davidson_stall", 32), ("wilson_rank", 256), ("geodesic_noconv", 512))
               if ((st & bit) != 0).any()}
    ms_max = max_over_ranks(ms, device=dev)
    value = world * b * args.steps / (ms_max / 1e3)
    step_ms = ms_max / args.steps

    # ---------------- in-run parity: the first systems of the batch, first steps, against the oracle loop run
    # with the SAME integrator (Dormand-Prince); the LSODA run of the reference algorithm is the CPU baseline
    parity = None
    if rank == 0 and args.parity_systems > 0:
        npar, nst = min(args.parity_systems, b), min(6, args.warmup + args.steps)
        engp = make(npar)
        for _ in range(nst):
            engp.step()
        jobs = [([first + i], n, args.kdiag, args.diag_every, 0, nst, args.method, "rk", exact) for i in range(npar)]
        with mp.get_context("spawn").Pool(min(npar, os.cpu_count() or 1)) as pool:
            res = pool.map(_internal_cpu_worker, jobs)
        xg = engp.pos[:npar].cpu().numpy()
        parity = dict(systems=npar, steps=nst, max_dx=max(float(np.abs(xg[i] - r[2][0]).max()) for i, r in enumerate(res)),
                      checker="port", horizon="a fresh engine on systems 0..%d of the batch, first %d steps" % (npar - 1, nst),
                      note="max |x_gpu - x_cpu| (Cartesian positions, Angstrom) against oracle/internal_pes.py with the "
                           "engine's Dormand-Prince geodesic integrator")
        del engp

    # ---------------- end-to-end through host buffers (the timed engine goes on: a second full-batch engine does not
    # fit next to it; the caller's positions travel host -> device -> host every step)
    eng2 = eng
    hx = torch.empty((b, n), dtype=torch.float64).pin_memory()
    hx.copy_(eng.pos.cpu())
    hf = torch.empty(b, dtype=torch.float64).pin_memory()
    hfmax = torch.empty(b, dtype=torch.float64).pin_memory()

    def host_step():
        eng2.pos.copy_(hx, non_blocking=True)
        eng2.step()
        eng2.converged(0.0)
        hx.copy_(eng2.pos, non_blocking=True)
        hf.copy_(eng2.f, non_blocking=True)
        hfmax.copy_(eng2.fmax, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ne2e = max(1, min(args.steps, 4))
    ne2e = 3 * max(1, ne2e // 3)             # whole re-diagonalisation periods (every 3rd step carries a Davidson run)
    host_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ne2e):
        host_step()
    e1.record()
    barrier()
    e2e_value = world * b * ne2e / (max_over_ranks(e0.elapsed_time(e1), device=dev) / 1e3)
    e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=world * b * n * 8,
               d2h_bytes_per_step=world * (b * n * 8 + 2 * b * 8), steps=ne2e)

    # ---------------- phases of a step (CUDA events) and the rooflines of its two dominant operators
    peak, peak_src = peaks()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        z.record()
        torch.cuda.synchronize()
        return a.elapsed_time(z) / reps

    geo = eng.geo
    Bw = geo["Bw"]
    Rw = K.qr(Bw)[1]
    Gw = K.gemm(Bw, Bw, transA=True)
    Gw = (0.5 * (Gw + Gw.transpose(1, 2))).contiguous()
    ncart = n
    phases = dict(
        wilson_qB=timed(lambda: eng.ints.calc(eng.pos, jacobian=True), 3),
        wilson_qr=timed(lambda: K.qr(Bw), 3),
        wilson_trtri=timed(lambda: K.trtri(Rw), 3),
        rdot=timed(lambda: eng.ints.rdot(eng.pos, eng._gc), 3),
        geometry_total=timed(lambda: eng._geometry(eng.pos), 2),
        model_total=timed(lambda: eng._model(), 2),
        geodesic_total=timed(lambda: eng._set_x(eng.x + eng.s), 1),
        eigh_ncart=timed(lambda: K.eigh(geo["HLr"].contiguous()), 2),
        eigvalsh_ncart=timed(lambda: K.eigvalsh(geo["HLr"].contiguous()), 2),
        wilson_gram_gemm=timed(lambda: K.gemm(Bw, Bw, transA=True), 3),
        wilson_potrf=timed(lambda: K.potrf(Gw), 3),
        wilson_factor_R=timed(lambda: eng._factor(Bw, want_q=False), 3),
        wilson_factor_QR=timed(lambda: eng._factor(Bw, want_q=True), 3))
    xv = eng.g.view(b, 1, nint)
    yv = torch.empty_like(xv)
    hv_ms = timed(lambda: K.hv_ld(eng.B, xv, yv, 1), 10)
    hv_bytes = b * 8 * (nint * nint + 2 * nint)
    hv_gbs = hv_bytes / (hv_ms * 1e-3) / 1e9
    roofline = dict(kernel="hv_tma_kernel<1> (batched H.V on the nint x nint approximate Hessian: B s, |B| s, H scons)",
                    bound="hbm", achieved=hv_gbs, peak=peak, unit="GB/s", frac=hv_gbs / peak, traffic=None,
                    ms_per_launch=hv_ms, bytes_per_launch=hv_bytes, peak_source=peak_src,
                    note="the H.V kernel of the headline; in this configuration the step is dominated by the dense "
                         "factorisations of the Wilson matrix along the geodesic (phase_ms, roofline_qr)")
    qr_flops = b * (4.0 * nint * ncart * ncart - 4.0 / 3.0 * ncart ** 3)         # factorisation + explicit Q
    fp64 = None
    try:
        import ctypes
        tf = ctypes.c_double(0.0)
        scratch = torch.empty(lib.sb_device_sms() * 8 * 256, dtype=torch.float64, device=dev)
        best = 0.0
        for kind in (0, 1):
            for cps in (4, 8):
                if lib.sb_fp64_peak(kind, 20000, cps, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(tf)) == 0:
                    best = max(best, tf.value)
        fp64 = best
    except Exception:
        fp64 = None
    gram_flops = 2.0 * b * nint * ncart * ncart
    gach = gram_flops / (phases["wilson_gram_gemm"] * 1e-3) / 1e12
    gtraffic, gsrc = (ncu_traffic(("ncu_full_r2_wilson.csv",), r"gemm_kernel", per="max")
                      if (b, n) == (1024, 384) else (None, None))
    roofline_gemm = dict(kernel="gemm_kernel (sb_gemm): G = Bw^T Bw [%d x %d x %d], fp64 tensor-core tiles (DMMA m8n8k4, "
                                "64 x 64 x 16); feeds sb_potrf, the R factor of every Wilson-matrix factorisation" % (ncart, ncart, nint),
                         bound="tensor", achieved=gach, peak=fp64, unit="TFLOP/s", frac=(gach / fp64) if fp64 else None,
                         traffic=gtraffic, traffic_source=("profiles/" + gsrc) if gsrc else None,
                         algorithmic_bytes=b * 8.0 * (nint * ncart + ncart * ncart),
                         ms_per_launch=phases["wilson_gram_gemm"], flops_per_launch=gram_flops,
                         peak_source="measured here (sb_fp64_peak: max of DFMA and DMMA loops)")
    ach = qr_flops / (phases["wilson_qr"] * 1e-3) / 1e12
    roofline_qr = dict(kernel="sb_qr: blocked Householder QR of the Wilson matrix [%d x %d] (16-column panels in shared "
                              "memory, compact WY, DMMA trailing updates) incl. the explicit Q" % (nint, ncart),
                       bound="tensor", achieved=ach, peak=fp64, unit="TFLOP/s", frac=(ach / fp64) if fp64 else None,
                       traffic=None, ms_per_launch=phases["wilson_qr"], flops_per_launch=qr_flops,
                       calls_per_step="not on the geodesic path any more (R comes from sb_potrf of Bw^T Bw: phase_ms "
                                      "wilson_factor_R / _QR); still factors the constraint rows per geometry",
                       peak_source="measured here (sb_fp64_peak: max of DFMA and DMMA loops)")
    out = None
    if rank == 0:
        cfg = dict(workload="batch=%d/GPU x 3N=%d EMT-form surface on the device, %d-atom Cu(111) slabs (rattled 0.05 A), "
                            "INTERNAL coordinates: %d nearest-neighbour bonds + the %d Cartesian coordinates of the fixed "
                            "bottom half (fix_translation constraints), nint=%d; order=1, %s + MaxInternalStep, TS-BFGS, "
                            "geodesic steps (Dormand-Prince 5(4), %s B+), jd0 Davidson gamma=0.1 maxiter=%d, diag_every_n=%d"
                            % (b, n, n // 3, ints.nbonds, ints.ntrans, nint, args.method,
                               "exact" if exact else "frozen", args.kdiag, args.diag_every),
                   l2_policy="working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (b * nint * nint * 8 * 4 / 1e9),
                   batch_per_gpu=b, dof=n, nint=nint, rs="mis", method=args.method, davidson_maxiter=args.kdiag,
                   diag_every_n=args.diag_every, eta=1e-4, gamma=0.1, coordinates="internal")
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                   data="synthetic", config=cfg, clocks=clocks, e2e=e2e, gpu_launches=int(launches), parity=parity,
                   roofline=roofline, roofline_gemm=roofline_gemm, roofline_qr=roofline_qr, phase_ms=phases,
                   geodesic_steps_per_call=(eng.ode_steps - ode0) / max(1, args.steps), systems_flagged=flagged,
                   diagonalisations=eng.ndiag)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                out["cpu_baseline"] = _internal_cpu_baseline(args, first, exact)
            except Exception as exc:
                out["cpu_baseline"] = dict(value=None, unit=UNIT, cores=cores, kind="port", sample="failed: %r" % (exc,))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def main():
    args = parse()
    if args.internal and args.workload != "emt-slab":
        raise SystemExit("--internal goes with --workload emt-slab")
    if args.impl == "reference":
        run_reference(args)
    elif args.internal:
        run_internal(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
