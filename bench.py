#!/usr/bin/env python
"""bench.py -- instances/sec of the ppx hot path on 1..8 B200s, next to its roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--n INSTANCES] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

A "step" is one pass of the hot path over one batch of synthetic inputs: one kernel launch over n instances per
GPU, inputs resident in HBM when the timed region starts.  The headline workload is BASELINE.json's configs[1]
(UR5 product-of-exponentials forward kinematics + space Jacobian, 16M joint configurations per GPU); with N GPUs
every rank runs its own 16M-instance shard (weak scaling, no data-path collective: the instances are independent,
SURVEY.md section 8e).  Rank 0 prints ONE JSON line (the task contract); extra keys:

  roofline      dominant (only) kernel of the step: algorithmic bytes per launch / its CUDA-event duration, against
                the measured HBM copy bandwidth in MEASURED_PEAKS.json
  e2e           the same metric through the host ABI (ppx_cuda_host_*): pinned HOST arrays of reference records in
                and out, H2D + kernels + D2H inside the timed region.  N = 1: the device of the rank.  N > 1: ONE
                call of rank 0 spread over all N devices by the library (ppx_cuda_host_set_devices), N x 16M
                instances -- what a C++ caller of ppx::cuda::batch_* gets; the per-rank form (every rank its own
                device and buffers, max over ranks) is reported beside it, and `ceiling` holds the aggregate
                host<->device copy rate the same devices sustain when they only copy (no kernels), measured with
                all of them running at once
  cpu_baseline  the reference's own CPU routines (oracle/_ref, OpenMP loop over ppx calls) or, if that was not
                built, the C restatement, on a bounded sample on this box's host cores (N = 1 only)
  configs       every BASELINE config measured the same way: kernel-only value + roofline (HBM and fp64 fractions),
                end to end through the host ABI, CPU baseline (N = 1 only); at N > 1 each rank runs its shard

--impl reference times only the CPU implementation (the reference arm).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
ARENA_BYTES = (1 << 24) * 464 + (64 << 20)  # pinned host arena of the end-to-end legs: the headline's 16M x 464 B + slack


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def config_for(workload, n, world):
    """the `config` key, identical in both arms (ours and --impl reference) for the same workload, size and N"""
    from matrix_b200.workloads_meta import META
    m = META[workload]
    per = m["bytes"] * n / 1e6
    return {"workload": workload, "instances_per_gpu": n,
            "device_layout": {"records": "arrays of the reference's dense records (std::vector<MatrixS>)"}.get(
                m.get("layout"), "component planes"),
            "chain": "UR5 (test/lie_test.cpp:187-196)" if m.get("chain") else None,
            "l2": f"inputs+outputs {per:.0f} MB per step >> 126 MB L2, every byte touched once per step",
            "parallelism": f"instance-sharded x{world}, no collective"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def time_cpu(workload, sample, min_seconds, min_passes=2):
    """-> cpu_baseline dict: the CPU implementation on `sample` instances, all host threads, passes repeated until
    `min_seconds` of CPU work have been timed; value = best pass"""
    from bench_cpu import CPU, NTHREADS, load_cpu_oracle
    oracle, kind = load_cpu_oracle()
    fn = CPU[workload](oracle, sample)
    fn()  # warm-up: page faults, thread pool
    times = []
    while len(times) < min_passes or sum(times) < min_seconds:
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
        if len(times) >= 50:
            break
    return {"value": sample / min(times), "unit": "instances/s", "cores": NTHREADS, "kind": kind,
            "mean_value": sample / statistics.mean(times),
            "sample": f"{sample} instances of {workload} per pass, best of {len(times)} passes ({sum(times):.1f} s of CPU "
                      "work), OpenMP static loop over the ppx call chain on host arrays of records"}


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path, every step one pass over the SAME number of
    instances as a step of the GPU arm on one GPU (the CPU throughput does not depend on the batch size; N ranks of the
    GPU arm each run that many)."""
    from matrix_b200.workloads_meta import META
    if rank != 0:
        return
    from bench_cpu import CPU, NTHREADS, load_cpu_oracle
    meta = META[args.workload]
    n = args.n if args.n else meta["n"]
    oracle, kind = load_cpu_oracle()
    fn = CPU[args.workload](oracle, n)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    desc = (f"{n} instances of {args.workload} per step, {args.steps} steps ({dt:.1f} s of CPU work): OpenMP static loop "
            "over the ppx call chain (oracle/_ref: the unmodified reference headers), host arrays of records")
    print(json.dumps({
        "impl": "reference", "metric": "instances/sec", "value": value, "unit": "instances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args.workload, n, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": NTHREADS, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--n", type=int, default=0, help="instances per GPU per step (default: the BASELINE size)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-others", action="store_true", help="skip the per-config table")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated workloads of the per-config table")
    ap.add_argument("--layout", choices=["soa", "aos"], default="soa", help="record layout of --all-ops")
    ap.add_argument("--all-ops", action="store_true",
                    help="time every batched entry point (kernel only) next to its CPU reference; one JSON line")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    from matrix_b200.workloads_meta import META, HEADLINE
    args.workload = args.workload or HEADLINE
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from matrix_b200 import _lib, host
    from matrix_b200.workloads import WORKLOADS, Arena

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_pg = dist.new_group(backend="gloo") if world > 1 else None

    def host_barrier():
        """CPU-side barrier (gloo): unlike an NCCL barrier it parks no kernel on the GPUs while the ranks wait"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_pg)

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def min_over_ranks(x):
        return -max_over_ranks(-x)

    def time_steps(w, steps, warmup):
        """-> (ms for all `steps` steps: CUDA events on the launching stream, max over ranks; per-step ms list)"""
        for _ in range(warmup):
            w.step()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            w.step()
            ev[i + 1].record()
        barrier()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        return max_over_ranks(ev[0].elapsed_time(ev[steps])), per

    hbm_peak, peak_src = measured_peaks()
    if args.all_ops:
        all_ops(args, device, hbm_peak)
        return

    # per-launch DRAM traffic and per-instance fp64 counts measured by ncu on the same seeded workloads
    # (tools/make_counters.py -> profiles/counters.json); not re-measured in this run
    try:
        with open(os.path.join(ROOT, "profiles", "counters.json")) as f:
            counters = json.load(f)
    except Exception:
        counters = {}
    fp64_peak = None
    if rank == 0:
        from matrix_b200 import batch as mb
        fp64_peak = mb.fp64_peak_tflops()
    fp64_peak = max_over_ranks(fp64_peak or 0.0) or None  # rank 0's probe, shared with everybody
    arena = None

    def roofline_of(cls, name, n_launch, kernel_ms):
        """HBM: algorithmic bytes / launch time vs the measured copy peak.  fp64: ncu-counted flops per instance
        (2 per DFMA, 1 per DADD / DMUL, profiles/counters.json) x instances/s vs the DFMA probe of this run."""
        gbs = n_launch * cls.bytes_per_instance() / (kernel_ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
             "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "peak_source": peak_src, "kernel": cls.kernel,
             "kernel_ms": kernel_ms, "algorithmic_bytes_per_instance": cls.bytes_per_instance(),
             "instances_per_launch": n_launch, "expected_bound": cls.bound}
        c = counters.get(name) or counters.get(META[name].get("counters_as", ""), {})
        r["traffic"] = c.get("dram_bytes_per_launch") if c.get("instances") == n_launch else None
        r["traffic_source"] = (f"ncu --set full capture of the same kernel and batch size, {c.get('report', '?')} "
                               "(profiles/counters.json); not re-measured in this run") if r["traffic"] else None
        if c.get("fp64_flops_per_instance") and fp64_peak:
            tf = c["fp64_flops_per_instance"] * n_launch / (kernel_ms * 1e-3) / 1e12
            r.update({"fp64_tflops": tf, "fp64_frac": tf / fp64_peak, "fp64_peak_tflops": fp64_peak,
                      "fp64_flops_per_instance_ncu": c["fp64_flops_per_instance"],
                      "fp64_counters_from": c.get("report"), "fp64_pipe_pct_ncu": c.get("fp64_pipe_pct")})
            if cls.bound.startswith("fp64") and tf / fp64_peak > r["hbm_frac"]:
                r.update({"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                          "peak_source": "measured in this run (ppx_cuda_fp64_peak_probe: dependent-free DFMA on every SM)"})
        return r

    def e2e_leg(w, cls, n, steps):
        """the same step through the host ABI on this rank's device: pinned host arrays, copies inside the timed region;
        every rank runs its own shard at once, time = max over ranks"""
        nonlocal arena
        if arena is None:
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = 64 << 30
            # every rank page-locks its own arena: leave the host 60 % of what is free
            arena = Arena(max(min(ARENA_BYTES, int(0.4 * avail / world)), 64 << 20))
        arena.reset()
        n_e2e = n
        while n_e2e * cls.bytes_per_instance() + (1 << 20) > arena.buf.size and n_e2e > 1024:
            n_e2e //= 2  # e.g. svd_6x6: 912 B per instance, 8M instances per call instead of 16M
        h2d, d2h = w.host_setup(n_e2e, arena)
        w.host_step()  # warm-up (stream-ordered pool growth)
        barrier()
        t0 = time.perf_counter()
        per = []
        for _ in range(steps):
            ts = time.perf_counter()
            probe = w.host_step()  # blocking: returns when the outputs are in host memory
            per.append((time.perf_counter() - ts) * 1e3)
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        w.host_release()
        return {"value": world * n_e2e * steps / dt, "unit": "instances/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "instances_per_step": n_e2e * world, "steps": steps,
                "ms_per_step": dt / steps * 1e3, "api": f"{cls.host_api} (pinned host arrays of records)",
                "mode": "one call per rank on its own device, max over ranks" if world > 1 else "one call, one device",
                "pcie_gbs": (h2d + d2h) * world * steps / dt / 1e9, "result_probe": probe,
                "step_ms": {"min": min(per), "median": statistics.median(per), "max": max(per)}}

    def run_config(name, n, steps, warmup, cpu_seconds, e2e_steps, headline=False, sampler=None):
        cls = WORKLOADS[name]
        w = cls(n, device, first=rank * n * getattr(cls, "NBUF", 1))
        w.setup()
        torch.cuda.synchronize()
        per_step_launches = getattr(w, "launches_per_step", 1)  # > 1: a CUDA graph of that many launches per step
        per_step_instances = getattr(w, "instances_per_step", n)
        if sampler is not None:
            sampler.__enter__()
        launches0 = _lib.launch_count()
        total_ms, per = time_steps(w, steps, warmup)
        counted = _lib.launch_count() - launches0  # launches the library itself issued (graph replays are not seen)
        launches = steps * per_step_launches if per_step_launches > 1 else counted - warmup
        if sampler is not None:
            # keep the sampler alive for a minimum window so short runs still get a few samples under load
            t_end = time.time() + max(0.0, 1.0 - total_ms / 1e3)
            while time.time() < t_end:
                w.step()
            torch.cuda.synchronize()
            sampler.__exit__(None, None, None)
        kernel_ms = max_over_ranks(statistics.mean(per)) / per_step_launches
        entry = {"workload": name, "instances_per_gpu": per_step_instances, "n_gpus": world,
                 "value": world * per_step_instances * steps / (total_ms * 1e-3), "unit": "instances/s",
                 "steps": steps, "ms_per_step": total_ms / steps, "gpu_launches": int(launches),
                 "roofline": roofline_of(cls, name, n, kernel_ms),
                 "step_ms": {"min": min(per), "median": statistics.median(per), "max": max(per)}}
        keep = (w, total_ms, per, launches) if headline else None
        e2e = None
        if not args.no_e2e and cls.host_api:
            w.free_outputs()
            torch.cuda.empty_cache()
            e2e = e2e_leg(w, cls, n, e2e_steps)
        entry["e2e"] = e2e
        cpu = None
        if world == 1 and rank == 0 and not args.no_cpu:
            cpu = time_cpu(name, META[name]["cpu_sample"], cpu_seconds)
        entry["cpu_baseline"] = cpu
        if not headline:
            del w
            torch.cuda.empty_cache()
        return entry, keep

    # ---- headline ------------------------------------------------------------------------------------------------
    cls = WORKLOADS[args.workload]
    n = args.n or cls.n_default
    clocks = ClockSampler(local)  # nvidia-smi samples during the timed kernel region of the headline only
    head, (w, total_ms, per_step, launches) = run_config(args.workload, n, args.steps, args.warmup, 12.0,
                                                         args.e2e_steps, headline=True, sampler=clocks)
    e2e = head["e2e"]

    # ---- N > 1: the host-buffer face of the library itself over all N devices, one call from rank 0 ---------------
    if world > 1 and e2e is not None:
        per_rank = dict(e2e)
        arena = None  # every rank gives its pinned arena back before rank 0 page-locks the big one
        del w
        torch.cuda.empty_cache()
        host_barrier()
        multi = None
        if rank == 0:
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = 64 << 30
            n_all = n * world
            while n_all * cls.bytes_per_instance() * 1.25 > avail and n_all > (1 << 20):
                n_all //= 2
            try:
                multi = multi_device_e2e(cls, n, n_all, world, args.e2e_steps, device, host, Arena)
            except Exception as e:  # report, never hide: the per-rank number stays the e2e of the line
                per_rank["multi_device_error"] = repr(e)
        host_barrier()
        if rank == 0 and multi is not None:
            e2e = multi
            e2e["per_rank"] = {k: per_rank[k] for k in ("value", "ms_per_step", "pcie_gbs", "mode", "instances_per_step")}
    elif e2e is not None and rank == 0:
        # the ceiling of the one-device host path: what this device's link carries when it only copies
        arena.reset()
        buf = arena.take((min(1 << 30, arena.buf.size - (1 << 20)),), "uint8")
        rates = {d: host.copy_probe(buf, d, 0.5)[0] for d in ("h2d", "d2h", "both")}
        e2e["ceiling"] = {"probe": "ppx_cuda_host_copy_probe: the same pinned arena, 48 MiB pieces, no kernels, 0.5 s each",
                          "h2d_gbs": rates["h2d"], "d2h_gbs": rates["d2h"], "both_gbs": rates["both"],
                          "pipeline_gbs": e2e["pcie_gbs"],
                          "frac_of_d2h": (e2e["d2h_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9) / rates["d2h"]}

    # ---- every other BASELINE config, measured the same way --------------------------------------------------------
    table = [head]
    if not args.no_others:
        want = [s for s in args.only.split(",") if s] or [k for k in WORKLOADS if k != args.workload]
        for name in want:
            try:
                entry, _ = run_config(name, WORKLOADS[name].n_default, 5, 3, 3.0, 3)
                table.append(entry)
            except Exception as e:  # report, never hide
                table.append({"workload": name, "error": repr(e)})
                torch.cuda.empty_cache()

    if rank == 0:
        value = world * n * args.steps / (total_ms * 1e-3)
        for t in table:  # GPU / CPU ratios where both legs exist (N = 1)
            c = t.get("cpu_baseline")
            if c:
                t["kernel_vs_cpu"] = t["value"] / c["value"]
                if t.get("e2e"):
                    t["e2e_vs_cpu"] = t["e2e"]["value"] / c["value"]
        out = {
            "metric": "instances/sec", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.workload, n, world),
            "layout": f"{cls.layout} (device resident)",
            "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks.summary(), "fp64_peak_tflops_measured": fp64_peak,
            "step_ms": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step)},
            "configs": table,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def multi_device_e2e(cls, n_per_gpu, n_all, world, steps, device, host, Arena):
    """rank 0 only: ONE ppx_cuda_host_* call over `world` devices on n_all instances of pinned host records, and the
    aggregate copy ceiling of the same devices (all copying at once, no kernels)"""
    import statistics as st
    import torch
    t_pin = time.perf_counter()
    probe_bytes = world * (256 << 20)  # scratch of the copy probes and the tuner: they overwrite what they are given
    big = Arena(n_all * cls.bytes_per_instance() + (64 << 20) + probe_bytes)
    t_pin = time.perf_counter() - t_pin
    scratch = big.take((probe_bytes,), "uint8")
    # the inputs of all shards: regenerated shard by shard on this device (counter-based generator), copied to the host
    hq, w, first = None, None, 0
    while first < n_all:
        cnt = min(n_per_gpu, n_all - first)
        w = cls(cnt, device, first=first)
        w.setup()
        w.free_outputs()
        ins = w.host_inputs()
        if hq is None:
            hq = [big.take((n_all, t.shape[1])) for t in ins]
        for a, t in zip(hq, ins):
            torch.from_numpy(a[first:first + cnt]).copy_(t)
        torch.cuda.synchronize()
        del ins
        first += cnt
    torch.cuda.empty_cache()
    h_out = [big.take((n_all, wd) if wd > 1 else (n_all,), dt) for wd, dt in cls.host_out]
    h2d, d2h = sum(a.nbytes for a in hq), sum(a.nbytes for a in h_out)
    def timed(devices):
        host.set_devices(devices)
        w.host_call(n_all, hq, h_out)  # warm-up: contexts, streams and scratch pools of these devices
        per = []
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            w.host_call(n_all, hq, h_out)
            per.append((time.perf_counter() - ts) * 1e3)
        return time.perf_counter() - t0, per

    everyone = list(range(world))
    share = d2h / (h2d + d2h)
    direction = "d2h" if share > 2 / 3 else ("h2d" if share < 1 / 3 else "both")
    try:
        dt_all, per_all = timed(everyone)
        rates_all = {d: host.copy_probe(scratch, d, 1.0) for d in sorted({"d2h", direction})}
        # let the library pick the devices whose links carry the most together (ppx_cuda_host_tune_devices)
        t_tune = time.perf_counter()
        chosen, tuned_gbs = host.tune_devices(everyone, scratch, direction, 0.15)
        t_tune = time.perf_counter() - t_tune
        if chosen != everyone:
            dt_sel, per_sel = timed(chosen)
            rates_sel = host.copy_probe(scratch, direction, 1.0)
        else:
            dt_sel, per_sel, rates_sel = dt_all, per_all, rates_all[direction]
        probe = float(h_out[0].reshape(n_all, -1)[-1, -1])  # a value of the last instance's result: it arrived
        if cls.name.startswith("ur5_fk") and probe != 1.0:
            raise RuntimeError(f"end-to-end result not written: T[-1][15] = {probe}")
    finally:
        host.set_devices(None)
    use_sel = dt_sel < dt_all
    dt, per, devices = (dt_sel, per_sel, chosen) if use_sel else (dt_all, per_all, everyone)
    pipeline_gbs = (h2d + d2h) * steps / dt / 1e9
    moved = {"d2h": d2h, "h2d": h2d, "both": h2d + d2h}[direction] * steps / dt / 1e9
    ceil_used = sum(rates_sel) if use_sel else sum(rates_all[direction])
    return {"value": n_all * steps / dt, "unit": "instances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "instances_per_step": n_all, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "api": f"{cls.host_api} (pinned host arrays of records)",
            "mode": f"ONE call from one process spread by the library over devices {devices} of the {world} of this run "
                    "(worker thread per device, chunks from a shared queue); the devices were chosen by "
                    "ppx_cuda_host_tune_devices from measured copy rates",
            "devices_used": devices, "pcie_gbs": pipeline_gbs, "result_probe": probe, "pin_seconds": t_pin,
            "tune_seconds": t_tune,
            "step_ms": {"min": min(per), "median": st.median(per), "max": max(per)},
            "all_devices": {"value": n_all * steps / dt_all, "ms_per_step": dt_all / steps * 1e3,
                            "pcie_gbs": (h2d + d2h) * steps / dt_all / 1e9},
            "ceiling": {"probe": "ppx_cuda_host_copy_probe: the listed devices copying at once through the same pinned "
                                 f"buffer ({direction}), 48 MiB pieces, no kernels, 1 s",
                        "direction": direction,
                        "all_devices_gbs_per_device": rates_all[direction], "all_devices_gbs": sum(rates_all[direction]),
                        "all_devices_d2h_gbs": sum(rates_all["d2h"]),
                        "chosen_devices": chosen, "chosen_devices_gbs_per_device": rates_sel,
                        "chosen_devices_gbs": sum(rates_sel), "tuner_estimate_gbs": tuned_gbs,
                        "pipeline_gbs_same_direction": moved, "frac_of_ceiling": moved / max(ceil_used, 1e-9),
                        "all_devices_pipeline_frac": ({"d2h": d2h, "h2d": h2d, "both": h2d + d2h}[direction] * steps
                                                      / dt_all / 1e9) / max(sum(rates_all[direction]), 1e-9)}}


def all_ops(args, device, hbm_peak):
    """one line per C-ABI entry point that is not already a BASELINE workload: kernel-only inst/s and GB/s on
    device-resident data, with the CPU reference (all host threads, 2^20-instance sample) beside it"""
    import time as _t
    import torch
    from matrix_b200 import batch as _mb, ops_table
    from bench_cpu import NTHREADS, cpu_op, load_cpu_oracle
    oracle, kind = load_cpu_oracle()
    rows = []
    op_layout = _mb.AOS if args.layout == "aos" else _mb.SOA
    for op in ops_table.make_ops(op_layout):
        op.setup(device)
        for _ in range(3):
            op.step()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(5):
            op.step()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = statistics.mean(ev[i].elapsed_time(ev[i + 1]) for i in range(5))
        sample = 1 << 20 if "ik_" not in op.name else 1 << 18
        fn = cpu_op(op.name, oracle, sample)
        fn()
        t0 = _t.perf_counter()
        fn()
        cpu_s = _t.perf_counter() - t0
        gbs = op.n * op.bytes / (ms * 1e-3) / 1e9
        rows.append({"op": op.name, "reference": op.ref, "instances": op.n, "kernel_ms": ms,
                     "value": op.n / (ms * 1e-3), "unit": "instances/s", "algorithmic_bytes_per_instance": op.bytes,
                     "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "touched_bytes_per_instance": op.touched,
                     "hbm_frac_touched": gbs / hbm_peak * op.touched / op.bytes,
                     "cpu_value": sample / cpu_s, "cpu_cores": NTHREADS, "cpu_kind": kind})
        op.step = None
        torch.cuda.empty_cache()
    print(json.dumps({"all_ops": rows, "hbm_peak_gbs": hbm_peak, "layout": f"{args.layout} (device resident)",
                      "dtype": "f64"}))


if __name__ == "__main__":
    main()
