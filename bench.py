#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native h-NUMO hot path.

Metric (BASELINE.json): DG DOF-updates/s per RK stage on the synthetic double gyre, nop=4, 1000x1000 elements,
3 layers (dt=12 s, dt_btp=0.6 s -> 20 barotropic substeps x 5 SSPRK stages, twice per step).
A "step" is one call of ti_rk_bcl = one baroclinic predictor-corrector step = 200 barotropic stages + the layer work.
  value  = 3 * npoin * (barotropic stages in the timed steps) / (time of the timed steps), whole job, state resident
  e2e    = same metric through hnumo_ti_rk_bcl() with HOST buffers (H2D + step + D2H inside the timed region)
  roofline: dominant kernel = fused barotropic stage kernel; algorithmic bytes per launch = 953.3 B per node
           (visc>0) x npoin, divided by the mean stage time from CUDA events on the library's
           compute stream; peak = MEASURED_PEAKS.json hbm_gbs
  check  = diagnostics of the state after warm-up + timed steps (per-layer mass, extrema, Courant numbers from
           hnumo_diagnostics, reduced over the ranks like the reference's mpi_reduce) so that the lines of different N can be
           compared, and -- for N > 1 -- `cross_n`: one step of a 64x64 brick on the N ranks (NCCL halo exchange) against
           the same step on one GPU
  configs: one more leg in the same process, BASELINE config 5 (nop 8, 10 layers, 500x500 elements, 1000 stages per
           step) with its own roofline
  cpu_baseline: the C++ CPU oracle (port of the reference algorithm, dense tables as in the reference) run the way the
           reference is run, `mpirun -np C`: C single-threaded instances, one pinned to each host core, each advancing its
           own 1/C of the sample brick; rates summed.  Rank 0 at N=1 only.

`--impl reference` times that same CPU arm with the bench's own --steps/--warmup (the Fortran build needs
gfortran+MPI+p4est+NetCDF, none of which exist here); under torchrun only rank 0 runs it.
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

BYTES_PER_NODE_STAGE = {4: {True: 953.3, False: 766.1}, 8: {True: 895.8, False: 895.8}}  # BASELINE.md section 3 (visc>0 / visc==0)
CPU_BRICK = 32   # edge of the brick one CPU instance advances (1024 elements: 77 MB of the reference's operator tables, far beyond its cache)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nelx", type=int, default=1000)
    ap.add_argument("--nely", type=int, default=1000)
    ap.add_argument("--nop", type=int, default=4)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--variant", type=int, default=0, help="0 = element-record stage kernel (default), 1 = simple reference-form kernel")
    ap.add_argument("--cpu-cores", type=int, default=0, help="CPU instances of the baseline (default: every core this process may run on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-host", default="caller", choices=["caller", "torch-pinned"], help="host arrays of the e2e leg: ordinary allocations that the "
                    "library page-locks on first use (default; what the Fortran shim passes), or torch pinned buffers")
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE config 5 leg (nop 8, 10 layers, 500x500)")
    ap.add_argument("--no-cross-check", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (hnumo_set_option)")
    ap.add_argument("--partition", default=None, help="element partition for --gpus > 1: morton (default: chunks of the z-order curve, the shape of "
                                                      "p4est's uniform partition), rows, blocks:PXxPY (decks.element_owner)")
    return ap.parse_args()


def workload(nelx, nely, nop, layers):
    from hnumo_loader import hnumo_b200 as hn
    n = max(nelx, nely)
    if nop == 8:
        # BASELINE config 5: dt = 35 s, dt_btp = 0.35 s at 500x500 (CFL-scaled with the grid)
        return hn.decks.synthetic_double_gyre(nelx, nely, nop=8, nlayers=layers, dt=35.0 * 500.0 / n, dt_btp=0.35 * (1 + 1e-9) * 500.0 / n)
    if nop == 4:
        # BASELINE config 4.  dt_btp is nudged up by 1e-9 so that N_btp = ceiling(dt/dt_btp) = 20 and not 21 (12/0.6 > 20 in binary FP);
        # the reference re-derives dt_btp = dt/N_btp = 0.6 anyway (mod_initial.F90:176-177)
        return hn.decks.synthetic_double_gyre(nelx, nely, nop=4, nlayers=layers, dt=12.0 * 1000.0 / n, dt_btp=0.6 * (1 + 1e-9) * 1000.0 / n)
    return hn.decks.synthetic_double_gyre(nelx, nely, nop=nop, nlayers=layers)


def workload_name(nelx, nely, nop, layers):
    cfg = "4" if (nop, layers) == (4, 3) else "5" if (nop, layers) == (8, 10) else "-"
    return "synthetic double gyre %dx%d elements nop=%d %d layers (BASELINE config %s)" % (nelx, nely, nop, layers, cfg)


def base_config(args, world):
    """`config` of the JSON line: the same dictionary on the b200 arm and on the reference arm"""
    part = args.partition or ("morton" if world > 1 else "single")
    rec_doubles = {4: 2576, 8: 9322}.get(args.nop)   # element record of the stage kernel (PairRec<G,Q>::REC)
    nelem = args.nelx * args.nely
    l2 = "inputs larger than L2"
    if rec_doubles:
        l2 += ": every stage launch streams the rank's element records, %.1f GB per rank (%d B per element) vs 126 MB of L2" % (
            rec_doubles * 8.0 * nelem / world / 1e9, rec_doubles * 8)
    return {"workload": workload_name(args.nelx, args.nely, args.nop, args.layers), "nelem": nelem,
            "npoin": nelem * (args.nop + 1) ** 2, "partition": "%s, %d rank(s)" % (part, world), "l2_policy": l2,
            "cpu_arms": "cpu_baseline and --impl reference advance a bounded sample of this workload: C independent %dx%d-element bricks of it, "
                        "one single-threaded oracle instance pinned to each of the C host cores (the way the reference runs: mpirun -np C)" % (CPU_BRICK, CPU_BRICK)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---- CPU arm: C pinned single-threaded oracle instances, like `mpirun -np C ./numo3d` ---------------------------------
def _cpu_worker(core, nop, layers, steps, warmup, barrier, out):
    """one "MPI rank" of the CPU arm: pinned to `core`, 1 OpenMP thread, its own CPU_BRICK x CPU_BRICK piece of the workload"""
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.set_threads(1)
    n = CPU_BRICK if nop <= 4 else max(8, CPU_BRICK // 4)
    # the piece keeps the element size, time step and physics of the full workload (dt scales with the element size only)
    p = workload(n, n, nop, layers)
    o = oracle_lib.Oracle(p)
    for _ in range(warmup):
        o.step(1)
    _, st0 = o.timing()
    barrier.wait()
    t0 = time.perf_counter()
    o.step(steps)
    t1 = time.perf_counter()
    _, st1 = o.timing()
    out.put((core, 3.0 * o.npoin * (st1 - st0), t1 - t0, n, o.npoin, st1 - st0))


def cpu_reference(args, steps, warmup, cores=None):
    """DOF-updates/s of the CPU oracle run as C independent single-threaded instances (sum of the instances' rates while
    all of them run), and the rate of one instance running alone (so that the scaling over the cores is visible)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    avail = sorted(os.sched_getaffinity(0))
    if cores is None:
        cores = args.cpu_cores or len(avail)
    cores = max(1, min(cores, len(avail)))

    def run(core_list, nsteps, nwarm):
        barrier, out = ctx.Barrier(len(core_list)), ctx.Queue()
        procs = [ctx.Process(target=_cpu_worker, args=(c, args.nop, args.layers, nsteps, nwarm, barrier, out)) for c in core_list]
        [p.start() for p in procs]
        res = [out.get(timeout=3600) for _ in procs]
        [p.join() for p in procs]
        return res

    alone = run(avail[:1], 1, 0)[0]
    one_core = alone[1] / alone[2]
    res = run(avail[:cores], steps, warmup)
    wall = max(r[2] for r in res)
    value = sum(r[1] / r[2] for r in res)
    n, stages = res[0][3], res[0][5]
    return dict(value=value, cores=cores, one_core_value=one_core, scaling_vs_linear=value / (cores * one_core), wall=wall, steps=steps,
                sample="%d independent %dx%d-element bricks (nop=%d, %d layers; one single-threaded oracle instance pinned per host core), %d "
                       "baroclinic step(s) = %d barotropic stages each" % (cores, n, n, args.nop, args.layers, steps, stages))


def reference_build_probe():
    """the preferred CPU baseline is the Fortran reference itself (`make hnumo`, mpirun -np C ./numo3d); it needs
    an MPI Fortran compiler, p4est and NetCDF-Fortran on the box.  Say explicitly what is missing."""
    import shutil
    missing = [t for t in ("mpif90", "gfortran", "mpirun", "nf-config") if shutil.which(t) is None]
    if missing:
        return "reference build unavailable on this box: no " + ", ".join(missing) + " (p4est 2.8 is not vendored either)"
    return "toolchain present; the reference still needs its un-vendored p4est 2.8 download (no network)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    r = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    unit = "DOF-updates/s"
    line = {
        "impl": "reference", "metric": "DG DOF-updates/s per RK stage", "value": r["value"], "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args, world),
        "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "one_core_value": r["one_core_value"], "scaling_vs_linear": r["scaling_vs_linear"],
                         "note": "CPU port of the reference algorithm (C++ oracle, the reference's dense operator tables); " + reference_build_probe()},
        "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local_rank):
    """Bind this rank's host threads -- and with them the first-touch placement of its page-locked state arrays -- to the NUMA node
    its GPU hangs off, the way `mpirun --bind-to numa` places the ranks of the Fortran program.  Without it the N ranks of a
    multi-GPU run stream their state over PCIe through whichever socket the kernel happened to start them on.  Returns the node
    number, or None when the topology is not visible (then nothing is changed)."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class Comm:
    """torch.distributed plumbing of one run (barrier, reductions over the ranks)"""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.numa_node = bind_to_gpu_numa_node(torch, self.local_rank) if self.world > 1 else None
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        if self.dist is None:
            return x
        t = self.torch.tensor(x, dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return t.cpu().numpy()

    def max(self, x):
        return float(self.reduce([x], "max")[0])


def make_solver(hn, comm, params, args, variant=0):
    deck = hn.decks.build_deck(params, comm.rank, comm.world)
    S = hn.Solver(deck, device=comm.local_rank + 1, variant=variant)
    if comm.world > 1:
        ids = [hn.nccl_unique_id() if comm.rank == 0 else None]
        comm.dist.broadcast_object_list(ids, src=0)
        S.comm_init(ids[0])
    for kv in args.opt:
        k, v = kv.split("=")
        S.set_option(k, float(v))
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    return deck, S


def check_block(comm, S, deck):
    """diagnostics of the resident state reduced over the ranks (mass: sum; extrema: max/min), print_diagnostics.F90:59-128"""
    d = S.diagnostics()
    out = {"layer_mass": [float(x) for x in comm.reduce(d["mass"], "sum")]}
    for f in ("h", "u", "v"):
        out[f + "_max"] = [float(x) for x in comm.reduce(d[f][:, 0], "max")]
        out[f + "_min"] = [float(x) for x in comm.reduce(d[f][:, 1], "min")]
    out["qb_max"] = [float(x) for x in comm.reduce(d["qb"][:, 0], "max")]
    out["qb_min"] = [float(x) for x in comm.reduce(d["qb"][:, 1], "min")]
    out["cfl_btp"], out["cfl"] = comm.max(d["cfl_b"]), comm.max(d["cfl"])
    return out


def cross_n_check(hn, comm, args):
    """N > 1: one baroclinic step (200 barotropic stages, every one with an NCCL halo exchange) of a 64x64 brick on the N
    ranks of this job vs the same step on rank 0's GPU alone.  Without viscosity every face term is antisymmetric and the
    two must agree to round-off (asserted: 1e-12 of the field's scale).  With viscosity the reference's as-written LDG face flux
    (mod_laplacian_quad.F90:485-486) takes the rank's own element as the left one on processor faces, so the reference
    itself depends on the partition at O(visc); that difference is reported."""
    out = {}
    for label, visc in (("inviscid", 0.0), ("viscous", 50.0)):
        p = dict(workload(64, 64, 4, 3), visc_mlswe=visc, partition=args.partition or "morton")
        deck, S = make_solver(hn, comm, p, args)
        assert S.step(1) == 0
        q, qb, qp = S.download_state()
        S.close()
        parts = [None] * comm.world
        comm.dist.all_gather_object(parts, (deck["elem_global"], q, qb))
        if comm.rank == 0:
            single = hn.decks.build_deck(p)
            S1 = hn.Solver(single, device=comm.local_rank + 1)
            S1.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
            assert S1.step(1) == 0
            q1, qb1, _ = S1.download_state()
            S1.close()
            npts = single["npts"]
            c = float(np.sqrt(single["gravity"] * 9928.0))
            worst = {"qb.pbpert": 0.0, "qb.momentum": 0.0, "q.dp": 0.0, "q.momentum": 0.0}
            for eg, qr, qbr in parts:
                idx = (np.asarray(eg)[:, None] * npts + np.arange(npts)[None, :]).ravel()
                pn = np.linalg.norm(qb1[idx, 0])
                worst["qb.pbpert"] = max(worst["qb.pbpert"], float(np.linalg.norm(qbr[:, 1] - qb1[idx, 1]) / pn))
                worst["qb.momentum"] = max(worst["qb.momentum"], float(np.linalg.norm(qbr[:, 2:] - qb1[idx, 2:]) / (c * pn)))
                dn = np.linalg.norm(q1[:, idx, 0])
                worst["q.dp"] = max(worst["q.dp"], float(np.linalg.norm(qr[:, :, 0] - q1[:, idx, 0]) / dn))
                worst["q.momentum"] = max(worst["q.momentum"], float(np.linalg.norm(qr[:, :, 1:] - q1[:, idx, 1:]) / (c * dn)))
            out[label] = worst
            if visc == 0.0:
                assert max(worst.values()) <= 1e-12, ("N-way result differs from the 1-way result", worst)
    return out if comm.rank == 0 else None


def run_leg(hn, comm, args, nelx, nely, nop, layers, steps, warmup, do_e2e):
    """one workload on the ranks of this job: warm-up, `steps` timed steps (CUDA events on the library's stream, max over the
    ranks), roofline of the stage kernel, optional end-to-end arm through host buffers"""
    torch = comm.torch
    params = workload(nelx, nely, nop, layers)
    params["partition"] = args.partition or "morton"
    deck, S = make_solver(hn, comm, params, args, variant=args.variant)
    stages_per_step = 2 * deck["N_btp"] * deck["kstages"]
    npoin_global = nelx * nely * deck["npts"]
    has_visc = deck["visc_mlswe"] != 0.0
    for _ in range(warmup):
        assert S.step(1) == 0, "physics error during warm-up"
    S.timing(reset=True)
    sampler = ClockSampler(comm.local_rank)
    comm.barrier()
    if comm.rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    rc = S.step(steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    comm.barrier()
    clocks = sampler.stop() if comm.rank == 0 else None
    assert rc == 0, "physics error during the timed steps"
    tm = S.timing(reset=True)
    # device time of the steps (CUDA events on the library's stream), max over ranks; wall clock as a cross-check
    dev_s, wall_s, btp_s = comm.max(tm["ms_step"] * 1e-3), comm.max(t1 - t0), comm.max(tm["ms_btp"] * 1e-3)
    launches = tm["launches"]
    stages = stages_per_step * steps
    value = 3.0 * npoin_global * stages / dev_s
    stage_ms = 1e3 * btp_s / stages
    bpn = BYTES_PER_NODE_STAGE.get(nop, BYTES_PER_NODE_STAGE[4])[has_visc]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    nodes_per_launch = comm.max(deck["npoin"])            # the stage time is the slowest rank's: its nodes per launch
    achieved = bpn * nodes_per_launch / (stage_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")))
        # ncu capture of one launch over prof["nelem"] elements (profiles/README.md says of which kernel build)
        key = "nop%d" % nop
        if args.variant == 0 and has_visc and key in prof:
            traffic = prof[key]["dram_bytes_per_launch"] * deck["nelem"] / prof[key]["nelem"]
    except Exception:
        pass
    check = check_block(comm, S, deck)
    # resident bytes of this rank (state + records + work arrays), for the L2 policy statement
    free_b, total_b = torch.cuda.mem_get_info()
    resident = total_b - free_b
    e2e = None
    if do_e2e:
        nl, npn = deck["nlayers"], deck["npoin"]
        if args.e2e_host == "torch-pinned":
            q = torch.from_numpy(deck["q_df"].copy()).pin_memory()
            qb = torch.from_numpy(deck["qb_df"].copy()).pin_memory()
            qp = torch.from_numpy(deck["qprime_df"].copy()).pin_memory()
            host_note = "pinned (torch)"
        else:
            # ordinary host arrays, as the Fortran driver's allocatables are: the library page-locks the caller's arrays once
            # (cudaHostRegister in hnumo_ti_rk_bcl, first call = the warm-up call below) and copies from / to them directly
            q = torch.from_numpy(deck["q_df"].copy()); qb = torch.from_numpy(deck["qb_df"].copy()); qp = torch.from_numpy(deck["qprime_df"].copy())
            host_note = "caller's ordinary arrays, page-locked once by the library (cudaHostRegister)"
        S.download_state((q.numpy(), qb.numpy(), qp.numpy()))
        n_e2e = max(1, min(steps, 2))
        S.ti_rk_bcl(q.numpy(), qb.numpy(), qp.numpy())  # warm the path
        comm.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rc = S.ti_rk_bcl(q.numpy(), qb.numpy(), qp.numpy())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        comm.barrier()
        e2e_s = comm.max(t1 - t0)
        nbytes = (3 * nl + 4 + 3 * nl) * npn * 8
        e2e = {"value": 3.0 * npoin_global * stages_per_step * n_e2e / e2e_s, "unit": "DOF-updates/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e, "host_buffers": host_note}
        launches += S.timing(reset=True)["launches"]
    nfaces_proc = int(comm.max(len(deck["nbh_send_recv"])))
    nnbh = int(comm.max(len(deck["nbh_proc"])))
    S.close()
    leg = {
        "value": value, "ms_per_step": 1e3 * dev_s / steps, "steps": steps, "warmup": warmup,
        "config": {"nelem": nelx * nely, "npoin": npoin_global, "stages_per_step": stages_per_step, "dt": params["dt"], "dt_btp": deck["dt_btp"],
                   "processor_faces_per_rank_max": nfaces_proc, "neighbour_ranks_max": nnbh,
                   "resident_gb_per_rank": resident / 1e9,
                   "host_numa_node_rank0": comm.numa_node,   # ranks are bound to the NUMA node of their GPU (None: topology not visible)
                   "stage_kernel_variant": args.variant, "options": args.opt},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_btp_stage_pair (fused barotropic SSPRK stage, element records)" if args.variant == 0 else "k_btp_stage_simple",
                     "algorithmic_bytes_per_node_stage": bpn, "stage_ms": stage_ms, "peak_source": peak_src, "per_gpu": True},
        "stage_only": {"value": 3.0 * npoin_global * stages / btp_s, "unit": "DOF-updates/s", "share_of_step": btp_s / dev_s},
        "wall_ms_per_step": 1e3 * wall_s / steps, "check": check, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    return leg


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    from hnumo_loader import hnumo_b200 as hn
    comm = Comm()
    hn.build_library()
    leg = run_leg(hn, comm, args, args.nelx, args.nely, args.nop, args.layers, args.steps, args.warmup, do_e2e=not args.no_e2e)
    configs = []
    if not args.no_config5 and (args.nop, args.layers) == (4, 3):
        # BASELINE config 5 at the same N: nop 8, 10 layers, 500x500 elements, 1000 barotropic stages per step (one timed step)
        l5 = run_leg(hn, comm, args, 500, 500, 8, 10, 1, 1, do_e2e=False)
        l5["config"]["workload"] = workload_name(500, 500, 8, 10)
        configs.append({k: l5[k] for k in ("value", "ms_per_step", "steps", "warmup", "config", "roofline", "stage_only", "check", "gpu_launches", "clocks")})
    cross = None
    if comm.world > 1 and not args.no_cross_check:
        cross = cross_n_check(hn, comm, args)
    cpu = None
    if comm.rank == 0 and comm.world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args, steps=1, warmup=0)
        cpu = {"value": r["value"], "unit": "DOF-updates/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "one_core_value": r["one_core_value"], "scaling_vs_linear": r["scaling_vs_linear"]}
    if comm.rank == 0:
        cfg = base_config(args, comm.world)   # identical on the reference arm
        check = leg["check"]
        if cross is not None:
            check["cross_n"] = cross
        line = {
            "metric": "DG DOF-updates/s per RK stage", "value": leg["value"], "unit": "DOF-updates/s", "n_gpus": comm.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "run": leg["config"],
            "roofline": leg["roofline"], "stage_only": leg["stage_only"], "wall_ms_per_step": leg["wall_ms_per_step"],
            "cpu_baseline": cpu, "e2e": leg["e2e"], "gpu_launches": leg["gpu_launches"], "clocks": leg["clocks"], "check": check,
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if comm.dist is not None:
        comm.dist.destroy_process_group()


if __name__ == "__main__":
    main()
