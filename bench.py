#!/usr/bin/env python
"""bench.py -- gstrf GFLOP/s and wall time of PanguLU's numeric factorisation on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--nb NB]

Workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one GPU): synthetic 3D 7-point
Poisson matrix on a 100^3 grid (1 000 000 rows), R64, METIS_NodeND ordering, nb = 256.  One "step" is one
pangulu_gstrf over the whole matrix.  GFLOP/s uses the reference's own flop model
(src/pangulu_kernel_interface.c:4-176), which depends on the structure only.

Our arm prints one JSON line with
  value      : GFLOP/s with the blocks already resident in HBM (device-resident static plan), timed with CUDA
               events on the module's stream, max over ranks
  e2e        : the same metric through the reference-facing C ABI call path: factor values start in HOST
               block buffers, are uploaded, every wave goes through pangulu_platform_0201001_hybrid_batched
               with a HOST task array, finished factors are downloaded -- all inside the timed region
  roofline   : dominant kernel class (by device time), algorithmic bytes or flops / device time
  cpu_baseline : the unmodified reference CPU platform (oracle/_ref) on a bounded sample of the same workload
`--impl reference` times the reference's own CPU implementation (oracle/_ref) instead.
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
# scripts/micro/dmma_bench.cu (profiles/r01_dmma_microbench.txt): 37.0 TFLOP/s at 1965 MHz, 16.1 cycles per DMMA per SM sub-partition
MEASURED_DMMA_TFLOPS = 37.0
TASK_KINDS = ["ssssm_tile", "ssssm_sparse", "getrf_smem", "getrf_dense", "getrf_global", "panel_sparse", "ssssm_compact", "panel_compact", "diag_densify"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=100, help="grid points per dimension (default: the named config, 100)")
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--ref-size", type=int, default=48, help="grid size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--abi-e2e", action="store_true", help="also time the per-wave hybrid_batched path with host task arrays")
    ap.add_argument("--no-classes", action="store_true", help="skip the traced run that attributes device time to kernel classes")
    return ap.parse_args()


def workload_name(size, nb):
    return f"poisson3d_7pt_{size}^3_r64_metis_nb{nb}"


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
        sm, smmax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference CPU platform on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def reference_ranks():
    """Ranks of the reference CPU run: its numeric phase is parallel over MPI ranks only (one compute thread plus one mostly
    sleeping receiver thread per rank); the image has no MPI, so the ranks are processes talking through the
    shared-memory shim oracle/shims/mpi_local.c.  Up to 8 ranks (the largest grid the benchmark names), one per two cores."""
    cores = os.cpu_count() or 1
    return max(1, min(8, cores // 2))


def reference_step(size, nb, ranks=1):
    """One reference pangulu_gstrf (fresh processes; the reference is single-shot). Returns (seconds, blas)."""
    from oracle.ref_run import run_reference
    from pangulu_b200 import matrices

    A = matrices.poisson3d(size)
    z = run_reference(A, nb, vtype="r64", perf=False, solve=False, blocks=False, nthread=1, ranks=ranks)
    return float(z["t_gstrf"]), str(z["blas"])


def sample_flop(size, nb):
    """Reference-model flop count of the sample (structure only) from our host-only structure pass."""
    from pangulu_b200 import PanguLU, matrices

    s = PanguLU(matrices.poisson3d(size), nb=nb, host_only=True)
    f = s.info("flop")
    s.finalize()
    return f


def cpu_baseline(size, nb, reps=1):
    from oracle.ref import ref_available

    if not ref_available("r64"):
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref missing on this box"}
    flop = sample_flop(size, nb)
    ranks = reference_ranks()
    best, blas = None, ""
    for _ in range(reps):
        t, blas = reference_step(size, nb, ranks)
        best = t if best is None else min(best, t)
    return {"value": flop / best / 1e9, "unit": UNIT, "cores": ranks, "kind": "reference", "seconds": best,
            "sample": f"unmodified reference CPU platform 0100000, {ranks} ranks (processes over the shared-memory MPI shim; one compute "
                      f"thread per rank as the reference runs its kernels with tid 0; BLAS {blas.split('/')[-1] or 'builtin'}) on "
                      f"poisson3d {size}^3 nb={nb}, {flop:.3e} flop"}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size, nb = a.ref_size, a.nb
    flop = sample_flop(size, nb)
    ranks = reference_ranks()
    times = []
    blas = ""
    for i in range(a.warmup + a.steps):
        t, blas = reference_step(size, nb, ranks)
        if i >= a.warmup:
            times.append(t)
    t = sum(times) / len(times)
    v = flop / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(a.size, a.nb), "sample": f"poisson3d {size}^3 nb={nb} (bounded CPU sample)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ranks, "kind": "reference",
                             "sample": f"reference CPU platform, {ranks} ranks (shared-memory MPI shim) x 1 compute thread, BLAS {blas.split('/')[-1] or 'builtin'}, poisson3d {size}^3 nb={nb}"},
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
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def run_ours(a):
    import numpy as np
    import torch

    from pangulu_b200 import PanguLU, matrices

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

    A = matrices.poisson3d(a.size)
    n = A.shape[0]
    t0 = time.time()
    s = PanguLU(A, nb=a.nb, mode=1, rank=rank, nproc=world)
    if world > 1:
        s.connect()  # CUDA-IPC peer mappings + push lists: finished blocks travel over NVLink inside the kernel
    t_init = time.time() - t0
    L = s.L
    flop = s.info("flop")
    if world > 1:  # whole-job flop count = sum over ranks (every rank counts the tasks it owns)
        ft = torch.tensor([flop], dtype=torch.float64, device="cuda")
        dist.all_reduce(ft)
        flop = float(ft.item())
    stream = torch.cuda.ExternalStream(L.pangulu_sm100_stream())

    # ---- resident (value): device-resident plan, CUDA events on the module's stream -----------------
    s.set_mode(1)
    times = []
    sampler = ClockSampler(local_rank)
    class_ms = None
    launches = 0
    for it in range(a.warmup + a.steps):
        if it > 0:
            s.reset_values(True)           # untimed: fresh values of A in the HBM block buffers
        timed = it >= a.warmup
        if it == a.warmup:
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
        if timed:
            times.append(ms_it)
    clocks = sampler.stop()
    class_ms = [L.pangulu_sm100_stats_ms(c) / a.steps for c in range(6)]
    class_launches = [L.pangulu_sm100_stats_launches(c) / a.steps for c in range(6)]
    launches = int(sum(class_launches[:5])) + int(sum(L.pangulu_sm100_stats_launches(c) for c in (7, 8)) / a.steps)
    mega_ms = L.pangulu_sm100_stats_ms(6) / a.steps              # (read here: the e2e runs below launch the same kernel again)
    mega_launches = L.pangulu_sm100_stats_launches(6) / a.steps
    L.pangulu_sm100_set_option(b"profile", 0)
    ms = sum(times) / len(times)
    value = flop / (ms * 1e-3) / 1e9

    # parity spot-check of what was just timed (size-independent property: ||L(U1) - A1|| / ||A1||; single rank only --
    # with several ranks the factors are spread over the GPUs and tests/test_multigpu_gpu.py holds the parity gate)
    s.download()
    check = s.factor_check() if world == 1 else None

    # ---- e2e: reference-facing call path with host buffers --------------------------------------------------
    e2e = None
    if not a.no_e2e and world == 1:
        # pangulu_gstrf() on HOST block buffers (pinned): upload the value region, run the resident plan, download the
        # finished factors -- what a PanguLU user sees from the call
        s.set_mode(2)
        et = []
        for it in range(1 + max(1, a.steps)):
            s.reset_values(False)          # host block buffers <- A (untimed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.gstrf()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= 1:
                et.append(dt)
        e2e_t = sum(et) / len(et)
        e2e = {"value": flop / e2e_t / 1e9, "unit": UNIT, "seconds": e2e_t, "path": "pangulu_gstrf, host buffers in/out, resident plan",
               "h2d_bytes_per_step": int(s.info("value_bytes")), "d2h_bytes_per_step": int(s.info("own_value_bytes"))}
        check_e2e = s.factor_check()
        if a.abi_e2e:
            # the same through the platform entry point PanguLU's scheduler calls: one <id>_hybrid_batched per wave with HOST
            # task arrays (first-generation per-wave kernels, no plan)
            s.set_mode(0)
            s.reset_values(False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.gstrf()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e["hybrid_batched_per_wave"] = {"value": flop / dt / 1e9, "seconds": dt,
                                              "h2d_bytes_per_step": int(s.info("value_bytes")) + int(s.info("ntask")) * 48}
    else:
        check_e2e = None

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    # Default executor: ONE persistent kernel per step runs every task class, so it is the dominant kernel by
    # construction; its algorithmic bytes are the per-task figures of SURVEY 8(d) summed over all tasks of the step
    # (driver's work model), its duration comes from CUDA events recorded around the launch on the module's stream.
    peaks, peak_kind = measured_peaks()
    names = ["getrf", "tstrf+gessm", "-", "ssssm_sparse", "ssssm_tile", "aux"]
    class_ms = class_ms + [0.0] * (7 - len(class_ms))
    launches += int(mega_launches)
    all_bytes = sum(s.info("bytes_" + k) for k in ("getrf", "tstrf", "gessm", "ssssm"))
    all_flop = s.info("flop")
    if mega_launches > 0:
        # 96 % of the step's flops are SSSSM updates executed as compact dense products on the FP64 tensor pipe (DMMA.8x8x4):
        # the kernel is bound by that pipe, not by HBM (its algorithmic byte rate is reported next to it)
        tf = all_flop / (mega_ms * 1e-3) / 1e12
        ach_gbs = all_bytes / (mega_ms * 1e-3) / 1e9
        roof = {"kernel": "mega_kernel (persistent executor: GETRF + TSTRF/GESSM + SSSSM tasks of this rank in one launch)", "bound": "tensor",
                "achieved": tf, "peak": MEASURED_DMMA_TFLOPS, "unit": "TFLOP/s", "frac": tf / MEASURED_DMMA_TFLOPS, "traffic": None,
                "peak_kind": "FP64 tensor (DMMA.8x8x4) issue rate measured on this pool (scripts/micro/dmma_bench.cu, profiles/r01_dmma_microbench.txt); "
                             "MEASURED_PEAKS.json has no FP64 entry; nominal 40 TFLOP/s",
                "launches_per_step": mega_launches, "ms_per_step": mega_ms,
                "algorithmic_flops_per_launch": all_flop, "algorithmic_bytes_per_launch": all_bytes,
                "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": peaks["hbm_gbs"], "frac": ach_gbs / peaks["hbm_gbs"], "peak_kind": f"{peak_kind} (MEASURED_PEAKS.json)"},
                "fp64_frac_of_nominal_40tf": tf / NOMINAL_FP64_TFLOPS}
    else:
        dom = max(range(5), key=lambda c: class_ms[c])
        if dom in (3, 4):
            ssssm_ms = class_ms[3] + class_ms[4]
            by = s.info("bytes_ssssm")
            ach = by / (ssssm_ms * 1e-3) / 1e9
            roof = {"kernel": "ssssm (sparse + tile kernels)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_kind": f"hbm copy bandwidth, {peak_kind}",
                    "launches_per_step": class_launches[3] + class_launches[4], "ms_per_step": ssssm_ms,
                    "fp64_tflops": s.info("flop_ssssm") / (ssssm_ms * 1e-3) / 1e12}
        else:
            by = s.info("bytes_getrf") if dom == 0 else s.info("bytes_tstrf") + s.info("bytes_gessm")
            ach = by / (class_ms[dom] * 1e-3) / 1e9
            roof = {"kernel": names[dom], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_kind": f"hbm copy bandwidth, {peak_kind}",
                    "launches_per_step": class_launches[dom], "ms_per_step": class_ms[dom]}
        roof["class_ms"] = {names[c]: class_ms[c] for c in range(5) if c != 2}

    # ---- device time per task class: one extra, traced step (per-task timestamps from %globaltimer; not part of the timing)
    classes = None
    if world == 1 and mega_launches > 0 and not a.no_classes:
        L.pangulu_sm100_set_option(b"trace", 1)
        L.pangulu_sm100_driver_rebuild_plan(s._h)
        s.set_mode(1)
        s.reset_values(True)
        s.gstrf()
        nt = int(L.pangulu_sm100_driver_trace(s._h, None, 0))
        buf = np.zeros((nt, 4), dtype=np.uint64)
        L.pangulu_sm100_driver_trace(s._h, buf.ctypes.data, nt)
        kind = (buf[:, 0] & 0xff).astype(np.int64)
        run = (buf[:, 3] - buf[:, 2]).astype(np.float64) * 1e-9
        wait = (buf[:, 2] - buf[:, 1]).astype(np.float64) * 1e-9
        span = float(buf[:, 3].max() - buf[:, 1].min()) * 1e-9
        classes = {"span_s": span, "sm_count": 148, "capacity_cta_s": 148 * span}
        for k, name in enumerate(TASK_KINDS):
            m = kind == k
            if m.any():
                classes[name] = {"tasks": int(m.sum()), "busy_cta_s": float(run[m].sum()), "wait_cta_s": float(wait[m].sum())}
        busy_ssssm = sum(classes[n]["busy_cta_s"] for n in ("ssssm_tile", "ssssm_sparse", "ssssm_compact") if n in classes)
        busy_panel = sum(classes[n]["busy_cta_s"] for n in ("panel_sparse", "panel_compact") if n in classes)
        busy_getrf = sum(classes[n]["busy_cta_s"] for n in ("getrf_smem", "getrf_dense", "getrf_global") if n in classes)
        per_sm_tf = MEASURED_DMMA_TFLOPS / 148
        per_sm_gbs = peaks["hbm_gbs"] / 148
        # rate of a class WHILE one of its tasks occupies an SM, against one SM's share of the peak
        classes["while_active"] = {
            "ssssm": {"tflops_per_sm": s.info("flop_ssssm") / busy_ssssm / 1e12, "frac_of_dmma_peak": s.info("flop_ssssm") / busy_ssssm / 1e12 / per_sm_tf,
                      "algorithmic_gbs_per_sm": s.info("bytes_ssssm") / busy_ssssm / 1e9, "frac_of_hbm_peak": s.info("bytes_ssssm") / busy_ssssm / 1e9 / per_sm_gbs},
            "tstrf_gessm": {"tflops_per_sm": (s.info("flop_tstrf") + s.info("flop_gessm")) / busy_panel / 1e12,
                            "frac_of_dmma_peak": (s.info("flop_tstrf") + s.info("flop_gessm")) / busy_panel / 1e12 / per_sm_tf,
                            "algorithmic_gbs_per_sm": (s.info("bytes_tstrf") + s.info("bytes_gessm")) / busy_panel / 1e9,
                            "frac_of_hbm_peak": (s.info("bytes_tstrf") + s.info("bytes_gessm")) / busy_panel / 1e9 / per_sm_gbs},
            "getrf": {"algorithmic_gbs_per_sm": s.info("bytes_getrf") / busy_getrf / 1e9, "frac_of_hbm_peak": s.info("bytes_getrf") / busy_getrf / 1e9 / per_sm_gbs},
        }
        L.pangulu_sm100_set_option(b"trace", 0)

    cpu = None
    if not a.no_cpu_baseline and rank == 0 and world == 1:
        cpu = cpu_baseline(a.ref_size, a.nb)
    if world > 1 and not a.no_e2e:
        # end to end at N ranks: host block buffers -> HBM, wired persistent kernels, factors back to the host
        et = []
        for it in range(2):
            s.reset_values(False)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            L.platform("memcpy")(s.slot_ptr(0).contents.d_value, s.slot_ptr(0).contents.value, int(s.info("value_bytes")), 0)
            s.gstrf()
            s.download()
            torch.cuda.synchronize()
            dist.barrier()
            et.append(time.perf_counter() - t0)
        e2e = {"value": flop / et[-1] / 1e9, "unit": UNIT, "seconds": et[-1], "h2d_bytes_per_step": int(s.info("value_bytes")),
               "d2h_bytes_per_step": int(s.info("own_value_bytes"))}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a.size, a.nb), "n": n, "nnz_A": int(A.nnz), "nnz_LU": int(s.info("symbolic_nnz")),
                       "flop_per_step": flop, "nb": a.nb, "ordering": "METIS_NodeND (CUDA toolkit libmetis_static.a)",
                       "waves": int(s.info("nwave")), "tasks": int(s.info("ntask")),
                       "l2": "inputs (factor values %.1f GB) exceed the 126 MB L2; values are rewritten between steps" % (s.info("value_bytes") / 1e9),
                       "init_seconds": t_init},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "classes": classes, "cpu_baseline": cpu,
            "parity": {"factor_check_resident": check, "factor_check_e2e": check_e2e}}
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
