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
  cpu_baseline: the C++ CPU oracle (port of the reference algorithm, dense tables as in the reference) on the host
           cores, bounded sample (smaller brick, same physics/nop/layers), rank 0 at N=1 only

`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the Fortran build needs
gfortran+MPI+p4est+NetCDF, none of which exist here) with all host threads on a bounded sample per step.
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

BYTES_PER_NODE_STAGE = {True: 953.3, False: 766.1}  # visc>0 / visc==0 at nop=4
BYTES_PER_NODE_STAGE_NOP8 = {True: 895.8, False: 895.8}


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
    ap.add_argument("--variant", type=int, default=0, help="0 = element-record stage kernel (default), 1 = simple kernel, 2/3 = record-layout TMA kernels, 5 = warp-per-element kernel")
    ap.add_argument("--cpu-sample", type=int, default=128, help="edge (elements) of the CPU baseline sample brick")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (hnumo_set_option)")
    ap.add_argument("--partition", default="rows", help="element partition for --gpus > 1: rows (default), blocks:PXxPY, morton (decks.element_owner)")
    return ap.parse_args()


def workload(args):
    from hnumo_loader import hnumo_b200 as hn
    if args.nop == 4 and args.layers == 3:
        # dt_btp is nudged up by 1e-9 so that N_btp = ceiling(dt/dt_btp) = 20 and not 21 (12/0.6 > 20 in binary FP);
        # the reference re-derives dt_btp = dt/N_btp = 0.6 anyway (mod_initial.F90:176-177)
        p = hn.decks.synthetic_double_gyre(args.nelx, args.nely, nop=4, nlayers=3, dt=12.0 * 1000.0 / max(args.nelx, args.nely),
                                           dt_btp=0.6 * (1 + 1e-9) * 1000.0 / max(args.nelx, args.nely))
    elif args.nop == 8:
        p = hn.decks.synthetic_double_gyre(args.nelx, args.nely, nop=8, nlayers=args.layers, dt=35.0 * 500.0 / max(args.nelx, args.nely),
                                           dt_btp=0.35 * (1 + 1e-9) * 500.0 / max(args.nelx, args.nely))
    else:
        p = hn.decks.synthetic_double_gyre(args.nelx, args.nely, nop=args.nop, nlayers=args.layers)
    return p


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


def cpu_oracle_rate(args, nsteps=1, warmup=0):
    """DOF-updates/s of the CPU oracle on a bounded sample of the same workload (smaller brick)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from hnumo_loader import hnumo_b200 as hn
    n = args.cpu_sample
    if args.nop == 8:
        n = max(16, n // 4)
    p = hn.decks.synthetic_double_gyre(n, n, nop=args.nop, nlayers=args.layers, dt=12.0 * 1000.0 / n, dt_btp=0.6 * (1 + 1e-9) * 1000.0 / n)
    o = oracle_lib.Oracle(p)
    for _ in range(warmup):
        o.step(1)
    t_btp0, st0 = o.timing()
    t0 = time.time()
    per_step = []
    for _ in range(nsteps):
        t1 = time.time()
        o.step(1)
        per_step.append(time.time() - t1)
    wall = time.time() - t0
    t_btp1, st1 = o.timing()
    stages = st1 - st0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return dict(value=3.0 * o.npoin * stages / wall, stage_only_value=3.0 * o.npoin * stages / max(t_btp1 - t_btp0, 1e-9), wall=wall,
                per_step=per_step, stages=stages, npoin=o.npoin, cores=cores,
                sample="%dx%d elements, nop=%d, %d layers, %d baroclinic step(s) = %d barotropic stages" % (n, n, args.nop, args.layers, nsteps, stages))


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
    if rank != 0:
        return
    r = cpu_oracle_rate(args, nsteps=args.steps, warmup=args.warmup)
    unit = "DOF-updates/s"
    line = {
        "impl": "reference", "metric": "DG DOF-updates/s per RK stage", "value": r["value"], "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic double gyre %dx%d elements nop=%d %d layers (BASELINE config %s)" % (args.nelx, args.nely, args.nop, args.layers, "5" if args.nop == 8 else "4"),
                   "note": "reference arm = CPU port of the reference algorithm (C++ oracle, OpenMP); the Fortran/MPI build cannot be "
                           "compiled here (no gfortran/MPI/p4est/NetCDF); each step is a bounded sample: " + r["sample"],
                   "reference_build": reference_build_probe()},
        "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    from hnumo_loader import hnumo_b200 as hn
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hn.build_library()
    params = workload(args)
    params["partition"] = args.partition
    deck = hn.decks.build_deck(params, rank, world)
    S = hn.Solver(deck, device=local_rank + 1, variant=args.variant)
    if world > 1:
        ids = [hn.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        S.comm_init(ids[0])
    for kv in args.opt:
        k, v = kv.split("=")
        S.set_option(k, float(v))
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    stages_per_step = 2 * deck["N_btp"] * deck["kstages"]
    npoin_global = params["nelx"] * params["nely"] * deck["npts"]
    has_visc = deck["visc_mlswe"] != 0.0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm
    for _ in range(args.warmup):
        rc = S.step(1)
        assert rc == 0, "physics error during warm-up"
    S.timing(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    rc = S.step(args.steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    assert rc == 0, "physics error during the timed steps"
    tm = S.timing(reset=True)
    # device time of the steps (CUDA events on the library's stream), max over ranks; wall clock as a cross-check
    dev_s = max_over_ranks(tm["ms_step"] * 1e-3)
    wall_s = max_over_ranks(t1 - t0)
    btp_s = max_over_ranks(tm["ms_btp"] * 1e-3)
    launches = tm["launches"]
    stages = stages_per_step * args.steps
    value = 3.0 * npoin_global * stages / dev_s
    stage_ms = 1e3 * btp_s / stages
    bpn = (BYTES_PER_NODE_STAGE if args.nop == 4 else BYTES_PER_NODE_STAGE_NOP8)[has_visc]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = bpn * deck["npoin"] / (stage_ms * 1e-3) / 1e9  # per GPU: this rank's nodes per launch
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")))
        # ncu capture of one launch over prof["nelem"] elements; a launch of this rank covers deck["nelem"] elements
        if args.variant in (0, 4) and args.nop == 4 and has_visc:
            traffic = prof.get("dram_bytes_per_launch") * deck["nelem"] / prof.get("nelem")
    except Exception:
        pass
    # ---- end-to-end arm: host buffers through the drop-in entry point
    e2e = None
    if not args.no_e2e:
        nl, npn = deck["nlayers"], deck["npoin"]
        q = torch.from_numpy(deck["q_df"].copy()).pin_memory()
        qb = torch.from_numpy(deck["qb_df"].copy()).pin_memory()
        qp = torch.from_numpy(deck["qprime_df"].copy()).pin_memory()
        S.download_state((q.numpy(), qb.numpy(), qp.numpy()))
        n_e2e = max(1, min(args.steps, 2))
        S.ti_rk_bcl(q.numpy(), qb.numpy(), qp.numpy())  # warm the path
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rc = S.ti_rk_bcl(q.numpy(), qb.numpy(), qp.numpy())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        e2e_s = max_over_ranks(t1 - t0)
        nbytes = (3 * nl + 4 + 3 * nl) * npn * 8
        e2e = {"value": 3.0 * npoin_global * stages_per_step * n_e2e / e2e_s, "unit": "DOF-updates/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e}
        launches += S.timing(reset=True)["launches"]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_rate(args, nsteps=1, warmup=0)
        cpu = {"value": r["value"], "unit": "DOF-updates/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "stage_only_value": r["stage_only_value"]}
    if rank == 0:
        line = {
            "metric": "DG DOF-updates/s per RK stage", "value": value, "unit": "DOF-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic double gyre %dx%d elements nop=%d %d layers (BASELINE config %s)" % (args.nelx, args.nely, args.nop, args.layers, "5" if args.nop == 8 else "4"),
                       "nelem": params["nelx"] * params["nely"], "npoin": npoin_global, "stages_per_step": stages_per_step,
                       "dt": params["dt"], "dt_btp": deck["dt_btp"], "partition": "%s, %d rank(s)" % ("row blocks" if args.partition == "rows" else args.partition, world),
                       "l2_policy": "inputs larger than L2 (>= 60 GB of resident state per job vs 126 MB L2)",
                       "stage_kernel_variant": args.variant, "options": args.opt},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_btp_stage_pair (fused barotropic SSPRK stage, element records)" if args.variant in (0, 4) else "k_btp_stage (fused barotropic SSPRK stage)", "algorithmic_bytes_per_node_stage": bpn,
                         "stage_ms": stage_ms, "peak_source": peak_src, "per_gpu": True},
            "stage_only": {"value": 3.0 * npoin_global * stages / btp_s, "unit": "DOF-updates/s", "share_of_step": btp_s / dev_s},
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    S.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
