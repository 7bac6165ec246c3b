#!/usr/bin/env python
"""bench.py -- gstrf GFLOP/s and wall time of PanguLU's numeric factorisation on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--size S] [--nb NB]

Default workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one GPU): synthetic 3D 7-point
Poisson matrix on a 100^3 grid (1 000 000 rows), R64, METIS_NodeND ordering, nb = 256.  One "step" is one pangulu_gstrf
over the whole matrix.  GFLOP/s uses the reference's own flop model (src/pangulu_kernel_interface.c:4-176), which
depends on the structure only.  The other BASELINE configurations are `--workload laplace2d_1000 | stencil27_64 |
stencil27_64_r32 | circuit_cr64` (same JSON schema).

Our arm prints one JSON line with
  value      : GFLOP/s with the blocks already resident in HBM (device-resident static plan, ONE launch of the
               persistent kernel per step), timed with CUDA events on the module's stream, max over ranks
  e2e        : the same metric through pangulu_gstrf() on HOST block buffers: the value region is uploaded, the plan runs,
               the finished factors are downloaded (overlapped with the factorisation) -- all inside the timed region (wall clock)
  e2e_resident : upload of nnz(A) values + value scatter on the device + gstrf, factors stay in HBM; pangulu_gstrs on them timed
               beside it (one GPU)
  roofline   : the persistent kernel against the measured FP64 tensor (DMMA) peak, its algorithmic HBM rate, measured DRAM
               traffic (ncu, profiles/) and per-kernel-class fractions from a traced step
  parity     : ||L(U1) - A1|| / ||A1|| of the factors just timed (several ranks: partial products summed over the ranks,
               as the reference's own check does, src/pangulu_numeric.c:1082-1341)
  same_config_sample : value and e2e of THIS arm on exactly the matrix the reference arm factors (like-for-like ratio)
  cpu_baseline : the unmodified reference CPU platform (oracle/_ref) on that bounded sample
`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host ranks it can use) on the bounded
sample; its flop count comes from the reference's own -DPANGULU_PERF build, not from this repo's library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gstrf_gflops"
UNIT = "GFLOP/s"
NOMINAL_FP64_TFLOPS = 40.0  # B200 FP64 / FP64-tensor nominal peak
# MEASURED_PEAKS.json carries HBM and bf16 numbers only; the FP64 tensor (DMMA.8x8x4) peak was measured on this pool with
# scripts/micro/dmma_bench.cu (profiles/r01_dmma_microbench.txt) and again with the compact tile routine's own tensor loop
# (profiles/r02_tensor_loop_microbench.txt): 37.0 TFLOP/s at 1965 MHz = 16.1 cycles per DMMA per SM sub-partition
MEASURED_DMMA_TFLOPS = 37.0
DMMA_CYCLES = 16.1
SM_HZ = 1.965e9
TASK_KINDS = ["ssssm_tile", "ssssm_sparse", "getrf_smem", "getrf_dense", "getrf_global", "panel_sparse", "ssssm_compact", "panel_compact", "diag_densify"]

# name -> (generator, value type, default size, ordering text, sizes of the bounded CPU sample (largest that fits first))
WORKLOADS = {
    "poisson3d_100": ("poisson3d", "r64", 100, [64, 56, 48, 40, 32, 24]),
    "laplace2d_1000": ("laplace2d", "r64", 1000, [1000, 700, 500, 300]),
    "stencil27_64": ("stencil27", "r64", 64, [32, 28, 24, 20, 16]),
    "stencil27_64_r32": ("stencil27", "r32", 64, [32, 28, 24, 20, 16]),
    # BASELINE config 5 names 2M rows; a random power-law pattern has no small separators, so its LU fill grows like n^2
    # (n = 50 000 already holds 6.3e8 stored entries = 10 GB of CR64 values): the size is a parameter, default 40 000
    "circuit_cr64": ("circuit_like", "cr64", 40000, [8000, 6000, 4000]),
}
MULTI_LOOKAHEAD = 2   # look-ahead window (waves) with several ranks; measured sweep in profiles/r02_multi_lookahead_sweep.txt
DTYPES = {"r64": "float64", "r32": "float32", "cr64": "complex128", "cr32": "complex64"}
# measured seconds of ONE reference gstrf (8 ranks over the shared-memory MPI shim) by sample size, used to pick the largest
# sample whose warmup + steps fit the reference arm's budget; anything not listed is estimated by flop count
REF_STEP_SECONDS = {("poisson3d", 48): 4.5, ("poisson3d", 56): 11.0, ("poisson3d", 64): 25.0, ("poisson3d", 40): 1.6, ("poisson3d", 32): 0.5,
                    ("poisson3d", 24): 0.15}
REF_BUDGET_SECONDS = 240.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson3d_100", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="grid points per dimension / rows (default: the named configuration)")
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--ref-size", type=int, default=0, help="size of the bounded CPU sample (default: the largest that fits the budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sample", action="store_true", help="skip the same_config_sample block")
    ap.add_argument("--schedule", type=int, default=-1, help="placement of the updates: 0 lazy, 1 eager, 2 windowed (default: the driver's, 2)")
    ap.add_argument("--lookahead", type=int, default=-1, help="window of the windowed schedule in waves (default: 2 on one GPU, MULTI_LOOKAHEAD on several)")
    ap.add_argument("--defer", type=int, default=-1, help="persistent executor: tasks a CTA may set aside while their dependencies are outstanding (default: the library's)")
    ap.add_argument("--no-classes", action="store_true", help="skip the traced run that attributes device time to kernel classes")
    return ap.parse_args()


def make_matrix(gen, size, vtype):
    import numpy as np

    from pangulu_b200 import matrices

    dt = np.dtype(DTYPES[vtype])
    if gen == "circuit_like":
        return matrices.circuit_like(size, dtype=dt)
    return getattr(matrices, gen)(size).astype(dt)


def workload_name(gen, size, vtype, nb):
    shape = {"poisson3d": f"poisson3d_7pt_{size}^3", "laplace2d": f"laplace2d_5pt_{size}^2", "stencil27": f"stencil27_{size}^3",
             "circuit_like": f"circuit_like_powerlaw_n{size}"}[gen]
    return f"{shape}_{vtype}_metis_nb{nb}"


# ---------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
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
        sm, smmax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference CPU platform on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def reference_ranks():
    """Ranks of the reference CPU run: its numeric phase is parallel over MPI ranks only (one compute thread plus one mostly
    sleeping receiver thread per rank); the image has no MPI, so the ranks are processes talking through the
    shared-memory shim oracle/shims/mpi_local.c.  Up to 8 ranks (the largest grid the benchmark names), one per two cores."""
    cores = os.cpu_count() or 1
    return max(1, min(8, cores // 2))


def pick_sample_size(a, gen, sizes):
    """Largest sample whose warmup + steps reference factorisations fit REF_BUDGET_SECONDS (both arms use the same rule, so
    they agree on the sample without talking to each other)."""
    if a.ref_size:
        return a.ref_size
    per_step = REF_BUDGET_SECONDS / max(1, a.steps + a.warmup)
    for s in sizes:
        est = REF_STEP_SECONDS.get((gen, s))
        if est is None:
            est = {"laplace2d": 10.0 * (s / 1000.0) ** 3, "stencil27": 20.0 * (s / 32.0) ** 6, "circuit_like": 30.0 * (s / 8000.0) ** 3}.get(gen, 1e9)
        if est <= per_step:
            return s
    return sizes[-1]


def reference_step(A, nb, vtype, ranks=1, perf=False):
    """One reference pangulu_gstrf (fresh processes; the reference is single-shot). Returns (seconds, blas, flop or None)."""
    from oracle.ref_run import run_reference

    z = run_reference(A, nb, vtype=vtype, perf=perf, solve=False, blocks=False, nthread=1, ranks=ranks)
    return float(z["t_gstrf"]), str(z["blas"]), (float(z["flop"]) if perf else None)


def reference_flop(A, nb, ranks):
    """Flop count of the sample by the reference's OWN counter: its -DPANGULU_PERF build (src/pangulu_kernel_interface.c:4-176,
    summed over the ranks by the MPI_Reduce in pangulu_gstrf).  R64 only (the PERF flavour is built for r64); the count
    depends on the structure alone, so the other value types of the same pattern share it."""
    from oracle.ref import ref_available

    if not ref_available("r64", perf=True):
        return None
    import numpy as np

    Ar = A.astype(np.float64) if A.dtype.kind != "c" else abs(A).astype(np.float64)
    return reference_step(Ar, nb, "r64", ranks, perf=True)[2]


def cpu_baseline(A, nb, vtype, label):
    from oracle.ref import ref_available

    if not ref_available(vtype):
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref missing on this box"}
    ranks = reference_ranks()
    flop = reference_flop(A, nb, ranks)
    t, blas, _ = reference_step(A, nb, vtype, ranks)
    return {"value": (flop / t / 1e9) if flop else None, "unit": UNIT, "cores": ranks, "kind": "reference", "seconds": t, "flop": flop,
            "sample": f"unmodified reference CPU platform 0100000, {ranks} ranks (processes over the shared-memory MPI shim; one compute "
                      f"thread per rank as the reference runs its kernels with tid 0; BLAS {blas.split('/')[-1] or 'builtin'}) on {label}; "
                      f"flop count from the reference's own -DPANGULU_PERF build"}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gen, vtype, full_size, sizes = WORKLOADS[a.workload]
    size = pick_sample_size(a, gen, sizes)
    A = make_matrix(gen, size, vtype)
    ranks = reference_ranks()
    label = workload_name(gen, size, vtype, a.nb)
    flop = reference_flop(A, a.nb, ranks)
    times, blas = [], ""
    for i in range(a.warmup + a.steps):
        t, blas, _ = reference_step(A, a.nb, vtype, ranks)
        if i >= a.warmup:
            times.append(t)
    t = sum(times) / len(times)
    v = flop / t / 1e9 if flop else None
    full = workload_name(gen, a.size or full_size, vtype, a.nb)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if "64" in vtype else "f32",
            "data": "synthetic",
            "config": {"workload": label, "n": int(A.shape[0]), "nnz_A": int(A.nnz), "flop_per_step": flop, "nb": a.nb,
                       "bounded_sample_of": full,
                       "note": "the reference CPU platform cannot factor the full configuration inside the run's budget; this line and the "
                               "`same_config_sample` block of the other arm are measured on exactly this matrix"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ranks, "kind": "reference",
                             "sample": f"reference CPU platform, {ranks} ranks (shared-memory MPI shim) x 1 compute thread, BLAS {blas.split('/')[-1] or 'builtin'}, {label}; "
                                       f"flop count from the reference's own -DPANGULU_PERF build"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def measured_traffic(label):
    """DRAM traffic of one launch of the persistent kernel from an ncu pass of this workload (profiles/r02_ncu_dram_traffic.json,
    written from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); None when no capture exists for it."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_dram_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(label)


def recorded_full_size_reference(label):
    """NOT measured by this run: the one-off run of the unmodified reference CPU platform on the FULL configuration `label`
    (scripts/ref_cpu_full_size.py; profiles/r02_reference_cpu_full_*.json name host, ranks and date), quoted so that the
    same-configuration CPU figure stands next to the bounded sample the reference arm can afford per step."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_cpu_full_100cubed.json")) as f:
            d = json.load(f)
        runs = [{"workload": "poisson3d_7pt_100^3_r64_metis_nb256", "t_gstrf_s": d["t_gstrf_s"], "gflops": d["gflops"], "ranks": d["ranks"], "host": d["host"]}]
        with open(os.path.join(ROOT, "profiles", "r02_reference_cpu_full_other_configs.json")) as f:
            o = json.load(f)
        runs += [{"workload": r["workload"], "t_gstrf_s": r["t_gstrf_s"], "gflops": r["gflops"], "ranks": r["ranks"], "host": d["host"]} for r in o["runs"]]
        for r in runs:
            if r["workload"] == label:
                r["note"] = "recorded one-off run (profiles/r02_reference_cpu_full_*.json), not timed by this bench run"
                return r
    except Exception:
        pass
    return None


def time_resident(s, L, steps, warmup, world, dist, stream, torch, sampler=None):
    """W untimed + K timed factorisations on values resident in HBM; CUDA events on the module's stream, max over ranks."""
    times = []
    for it in range(warmup + steps):
        if it > 0:
            s.reset_values(True)           # untimed: fresh values of A in the HBM block buffers
        if it == warmup:
            if sampler:
                sampler.start()
            L.pangulu_sm100_stats_reset()
            L.pangulu_sm100_set_option(b"profile", 1)   # async event pairs around every launch, no host sync
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            L.pangulu_sm100_driver_gstrf_reset(s._h)   # clear this rank's dependency flags ...
            torch.cuda.synchronize()
            dist.barrier()                             # ... on every rank before any rank may push into them
            e0.record(stream)
            L.pangulu_sm100_driver_gstrf_launch(s._h)
        else:
            torch.cuda.synchronize()
            e0.record(stream)
            s.gstrf_enqueue()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms_it = e0.elapsed_time(e1)
        if world > 1:  # device time, max over ranks
            mt = torch.tensor([ms_it], dtype=torch.float64, device="cuda")
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
            ms_it = float(mt.item())
            dist.barrier()
        if it >= warmup:
            times.append(ms_it)
    return sum(times) / len(times)


def time_e2e(s, L, reps, world, dist, torch):
    """pangulu_gstrf() on HOST block buffers (pinned): upload the value region, run the resident plan, download the finished
    factors -- what a PanguLU user sees from the call.  Wall clock, max over ranks."""
    et = []
    for it in range(1 + max(1, reps)):
        s.reset_values(False)          # host block buffers <- A (untimed)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            t0 = time.perf_counter()
            # owned blocks only: the mirrors of remote blocks are written by the peers' kernels
            L.platform("memcpy")(s.slot_ptr(0).contents.d_value, s.slot_ptr(0).contents.value, int(s.info("own_value_bytes")), 0)
            s.gstrf()
            s.download()
            torch.cuda.synchronize()
            dist.barrier()
        else:
            t0 = time.perf_counter()
            s.gstrf()
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 1:
            et.append(dt)
    return sum(et) / len(et)


def traced_classes(s, L, world, peaks, np):
    """One extra, traced step (per-task timestamps from %globaltimer; not part of the timing): device time per task class of
    THIS rank and the rate of every class while one of its tasks holds an SM."""
    L.pangulu_sm100_set_option(b"trace", 1)
    L.pangulu_sm100_driver_rebuild_plan(s._h)
    if world > 1:
        s.connect()
    s.set_mode(1)
    s.reset_values(True)
    s.gstrf()
    nt = int(L.pangulu_sm100_driver_trace(s._h, None, 0))
    buf = np.zeros((max(nt, 1), 4), dtype=np.uint64)
    L.pangulu_sm100_driver_trace(s._h, buf.ctypes.data, nt)
    buf = buf[:nt]
    kind = (buf[:, 0] & 0xff).astype(np.int64)
    run = (buf[:, 3] - buf[:, 2]).astype(np.float64) * 1e-9
    wait = (buf[:, 2] - buf[:, 1]).astype(np.float64) * 1e-9
    span = float(buf[:, 3].max() - buf[:, 1].min()) * 1e-9
    classes = {"span_s": span, "sm_count": 148, "capacity_cta_s": 148 * span, "busy_cta_s": float(run.sum()), "wait_cta_s": float(wait.sum())}
    for k, name in enumerate(TASK_KINDS):
        m = kind == k
        if m.any():
            classes[name] = {"tasks": int(m.sum()), "busy_cta_s": float(run[m].sum()), "wait_cta_s": float(wait[m].sum())}
    busy = lambda names: sum(classes[n]["busy_cta_s"] for n in names if n in classes)
    per_sm_tf = MEASURED_DMMA_TFLOPS / 148
    per_sm_gbs = peaks["hbm_gbs"] / 148
    wa = {}
    b_ct = busy(("ssssm_compact",))
    if b_ct > 0:
        dmma = s.info("plan_ct_dmma")
        wa["ssssm_compact"] = {"dmma_instructions": dmma, "dmma_pipe_frac_while_active": dmma * DMMA_CYCLES / 4 / SM_HZ / b_ct,
                               "note": "DMMA.8x8x4 instructions issued x 16.1 cycles / 4 sub-partitions, over the SM time of the class"}
    b_ss = busy(("ssssm_tile", "ssssm_sparse", "ssssm_compact"))
    if b_ss > 0:
        wa["ssssm"] = {"tflops_per_sm": s.info("flop_ssssm") / b_ss / 1e12, "frac_of_dmma_peak": s.info("flop_ssssm") / b_ss / 1e12 / per_sm_tf,
                       "algorithmic_gbs_per_sm": s.info("bytes_ssssm") / b_ss / 1e9, "frac_of_hbm_peak": s.info("bytes_ssssm") / b_ss / 1e9 / per_sm_gbs}
    b_pn = busy(("panel_sparse", "panel_compact"))
    if b_pn > 0:
        fl, by = s.info("flop_tstrf") + s.info("flop_gessm"), s.info("bytes_tstrf") + s.info("bytes_gessm")
        wa["tstrf_gessm"] = {"tflops_per_sm": fl / b_pn / 1e12, "frac_of_dmma_peak": fl / b_pn / 1e12 / per_sm_tf,
                             "algorithmic_gbs_per_sm": by / b_pn / 1e9, "frac_of_hbm_peak": by / b_pn / 1e9 / per_sm_gbs}
    b_gt = busy(("getrf_smem", "getrf_dense", "getrf_global"))
    if b_gt > 0:
        wa["getrf"] = {"algorithmic_gbs_per_sm": s.info("bytes_getrf") / b_gt / 1e9, "frac_of_hbm_peak": s.info("bytes_getrf") / b_gt / 1e9 / per_sm_gbs}
    classes["while_active"] = wa
    L.pangulu_sm100_set_option(b"trace", 0)
    return classes


def run_ours(a):
    import numpy as np
    import torch

    from pangulu_b200 import PanguLU, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the numeric factorisation has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gen, vtype, full_size, sample_sizes = WORKLOADS[a.workload]
    size = a.size or full_size
    A = make_matrix(gen, size, vtype)
    n = A.shape[0]
    label = workload_name(gen, size, vtype, a.nb)
    L = _lib.load(vtype)
    t0 = time.time()
    if a.defer >= 0:
        assert L.pangulu_sm100_set_option(b"defer", a.defer) == 0
    sched_kw = {"schedule": a.schedule if a.schedule >= 0 else 2,
                "lookahead": a.lookahead if a.lookahead >= 0 else (2 if world == 1 else MULTI_LOOKAHEAD)}
    s = PanguLU(A, nb=a.nb, mode=1, rank=rank, nproc=world, **sched_kw)
    init_host = time.time() - t0
    t0 = time.time()
    if world > 1:
        s.connect()  # CUDA-IPC peer mappings + push lists: finished blocks travel over NVLink inside the kernel
    init_wire = time.time() - t0
    # pangulu_init of this driver: reordering, symbolic factorisation, block storage + upload, static schedule, device plan
    init = {"total": init_host + init_wire, "peer_wiring": init_wire}
    for k in ("reorder", "symbolic", "blocks", "schedule", "work_model", "plan"):
        init[k] = s.info("time_" + k)
    flop = s.info("flop")
    if world > 1:  # whole-job flop count = sum over ranks (every rank counts the tasks it owns)
        ft = torch.tensor([flop], dtype=torch.float64, device="cuda")
        dist.all_reduce(ft)
        flop = float(ft.item())
    stream = torch.cuda.ExternalStream(L.pangulu_sm100_stream())

    # ---- resident (value): device-resident plan, CUDA events on the module's stream -----------------
    s.set_mode(1)
    sampler = ClockSampler(local_rank)
    ms = time_resident(s, L, a.steps, a.warmup, world, dist, stream, torch, sampler)
    clocks = sampler.stop()
    class_launches = [L.pangulu_sm100_stats_launches(c) / a.steps for c in range(9)]
    mega_ms = L.pangulu_sm100_stats_ms(6) / a.steps              # (read here: the e2e runs below launch the same kernel again)
    mega_launches = class_launches[6]
    launches = int(round(sum(class_launches[:5]) + class_launches[6] + class_launches[7] + class_launches[8]))
    L.pangulu_sm100_set_option(b"profile", 0)
    value = flop / (ms * 1e-3) / 1e9

    # ---- parity of what was just timed: ||L(U1) - A1|| / ||A1|| (size-independent property) --------------------
    s.download()
    if world > 1:
        check = s.factor_check_distributed()
    else:
        check = s.factor_check()

    # ---- e2e: pangulu_gstrf with host buffers ---------------------------------------------------------------------
    e2e, check_e2e = None, None
    if not a.no_e2e:
        if world == 1:
            s.set_mode(2)
        e2e_t = time_e2e(s, L, min(a.steps, 3), world, dist, torch)
        if world > 1:
            mt = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
            e2e_t = float(mt.item())
        h2d = int(s.info("value_bytes" if world == 1 else "own_value_bytes"))
        d2h = int(s.info("own_value_bytes"))
        if world > 1:
            bt = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(bt)
            h2d, d2h = int(bt[0].item()), int(bt[1].item())
        e2e = {"value": flop / e2e_t / 1e9, "unit": UNIT, "seconds": e2e_t, "path": "pangulu_gstrf, pinned host block buffers in/out, resident plan; the factors travel back while the factorisation runs (pieces of consecutive blocks as their done flags come up)",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        check_e2e = s.factor_check_distributed() if world > 1 else s.factor_check()

    # ---- e2e, device-resident flavour: CSC values in, x out ------------------------------------------------------------
    # What a Newton / time-stepping loop pays per step with the preprocessing and the solve on the device:
    # nnz(A) values travel to the GPU and are scattered into the block value region by a kernel, the plan runs, pangulu_gstrs
    # runs on the device-resident factors and only x comes back -- the factors never leave HBM.  Reported NEXT to `e2e` (which
    # keeps the reference's contract: host block buffers in, host factors out); the rate is the gstrf flop over the WHOLE time.
    e2e_resident = None
    if not a.no_e2e and world == 1:
        s.set_mode(1)
        rhs = (A @ np.ones(n)).astype(A.dtype)
        ts, tsolve, resid = [], [], None
        for it in range(1 + min(a.steps, 3)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.update_values(s._value)
            s.gstrf()
            t1 = time.perf_counter()
            x = s.gstrs(rhs)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if it >= 1:
                ts.append(t1 - t0)
                tsolve.append(t2 - t1)
            resid = float(np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs))
        tt = sum(ts) / len(ts)
        # value: the gstrf flop over upload + scatter + factorisation; the solve that reads the resident factors is timed beside it
        e2e_resident = {"value": flop / tt / 1e9, "unit": UNIT, "seconds": tt, "seconds_gstrs": sum(tsolve) / len(tsolve),
                        "path": "pangulu_sm100_driver_update_values (CSC values host -> device, scatter kernel) + pangulu_gstrf (resident plan) + "
                                "pangulu_gstrs on the device-resident factors, x back to the host",
                        "h2d_bytes_per_step": int(A.nnz * A.dtype.itemsize + n * A.dtype.itemsize), "d2h_bytes_per_step": int(n * A.dtype.itemsize),
                        "solve_residual": resid}

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    # ONE persistent kernel per step runs every task class, so it is the dominant kernel by construction; its algorithmic
    # bytes / flops are the per-task figures of SURVEY 8(d) summed over all tasks of the step (driver's work model); its
    # duration comes from CUDA events recorded around the launch on the module's stream.
    peaks, peak_kind = measured_peaks()
    all_bytes = sum(s.info("bytes_" + k) for k in ("getrf", "tstrf", "gessm", "ssssm"))
    all_flop = s.info("flop")
    tf = all_flop / (mega_ms * 1e-3) / 1e12 if mega_ms > 0 else None
    ach_gbs = all_bytes / (mega_ms * 1e-3) / 1e9 if mega_ms > 0 else None
    roof = {"kernel": "mega_kernel (persistent executor: GETRF + TSTRF/GESSM + SSSSM tasks of this rank in one launch)", "bound": "tensor",
            "achieved": tf, "peak": MEASURED_DMMA_TFLOPS, "unit": "TFLOP/s", "frac": tf / MEASURED_DMMA_TFLOPS if tf else None,
            "traffic": measured_traffic(label) if world == 1 else None,
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r02_ncu_dram_traffic.json",
            "peak_kind": "FP64 tensor (DMMA.8x8x4) issue rate measured on this pool (scripts/micro/dmma_bench.cu, profiles/r01_dmma_microbench.txt; "
                         "the compact tile routine's own tensor loop reaches it, profiles/r02_tensor_loop_microbench.txt); MEASURED_PEAKS.json has no "
                         "FP64 entry; nominal 40 TFLOP/s",
            "launches_per_step": mega_launches, "ms_per_step": mega_ms,
            "algorithmic_flops_per_launch": all_flop, "algorithmic_bytes_per_launch": all_bytes,
            "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": peaks["hbm_gbs"], "frac": ach_gbs / peaks["hbm_gbs"] if ach_gbs else None,
                    "peak_kind": f"{peak_kind} (MEASURED_PEAKS.json)"},
            "fp64_frac_of_nominal_40tf": tf / NOMINAL_FP64_TFLOPS if tf else None}

    # ---- device time per task class (every rank traces its own step; rank 0 reports its own and the spread) ----------
    classes = None
    if mega_launches > 0 and not a.no_classes:
        classes = traced_classes(s, L, world, peaks, np)
        if world > 1:
            mine = torch.tensor([classes["busy_cta_s"], classes["wait_cta_s"], classes["capacity_cta_s"]], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            classes["per_rank"] = [{"rank": r, "busy_cta_s": float(t[0]), "wait_cta_s": float(t[1]), "capacity_cta_s": float(t[2]),
                                    "idle_or_waiting_frac": 1.0 - float(t[0]) / float(t[2])} for r, t in enumerate(allr)]
        roof["per_class"] = classes.get("while_active")

    # ---- the reference arm's bounded sample on this arm: like-for-like numbers ----------------------------------------
    sample, cpu = None, None
    if world == 1 and rank == 0 and not a.no_sample:
        ssize = pick_sample_size(a, gen, sample_sizes)
        As = make_matrix(gen, ssize, vtype)
        slabel = workload_name(gen, ssize, vtype, a.nb)
        s2 = PanguLU(As, nb=a.nb, mode=1)
        sms = time_resident(s2, L, 3, 3, 1, None, stream, torch)
        L.pangulu_sm100_set_option(b"profile", 0)
        s2.download()
        scheck = s2.factor_check()
        s2.set_mode(2)
        se2e = time_e2e(s2, L, 3, 1, None, torch)
        sflop = s2.info("flop")
        sample = {"workload": slabel, "n": int(As.shape[0]), "flop_per_step": sflop, "value": sflop / (sms * 1e-3) / 1e9, "ms_per_step": sms,
                  "e2e": {"value": sflop / se2e / 1e9, "seconds": se2e, "h2d_bytes_per_step": int(s2.info("value_bytes")),
                          "d2h_bytes_per_step": int(s2.info("own_value_bytes"))},
                  "unit": UNIT, "parity": {"factor_check": scheck},
                  "note": "this arm on exactly the matrix `--impl reference` factors with the same --steps/--warmup (same sizing rule)"}
        s2.finalize()
        if not a.no_cpu_baseline:
            cpu = cpu_baseline(As, a.nb, vtype, slabel)
            if cpu.get("flop"):
                sample["flop_reference_counter"] = cpu["flop"]

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": {"r64": "f64", "r32": "f32", "cr64": "c128", "cr32": "c64"}[vtype],
            "data": "synthetic",
            "config": {"workload": label, "n": n, "nnz_A": int(A.nnz), "nnz_LU": int(s.info("symbolic_nnz")),
                       "flop_per_step": flop, "nb": a.nb, "ordering": "METIS_NodeND (CUDA toolkit libmetis_static.a)",
                       "waves": int(s.info("nwave")), "tasks": int(s.info("ntask")),
                       "l2": "inputs (factor values %.1f GB) exceed the 126 MB L2; values are rewritten between steps" % (s.info("value_bytes") / 1e9),
                       "init_seconds": init},
            "clocks": clocks, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": launches, "roofline": roof, "classes": classes, "cpu_baseline": cpu,
            "same_config_sample": sample, "reference_cpu_full_size_recorded": recorded_full_size_reference(label),
            "parity": {"factor_check_resident": check, "factor_check_e2e": check_e2e,
                       "what": "||L(U 1) - A 1|| / ||A 1|| of the factors just timed; several ranks: partial products summed over the ranks"}}
    s.finalize()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
