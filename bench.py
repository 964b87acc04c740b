#!/usr/bin/env python
"""bench.py -- instances/sec of the ppx hot path on 1..8 B200s, next to its roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--n INSTANCES] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

A "step" is one pass of the hot path over one batch of synthetic inputs: one kernel launch over n instances
per GPU, inputs resident in HBM when the timed region starts.  The headline workload is BASELINE.json's
configs[1] (UR5 product-of-exponentials forward kinematics + space Jacobian, 16M joint configurations per GPU);
with N GPUs every rank runs its own 16M-instance shard (weak scaling, no data-path collective: the instances are
independent).  Rank 0 prints ONE JSON line (see the task contract); extra keys:

  roofline      dominant (only) kernel of the step: algorithmic bytes per launch / its CUDA-event duration,
                against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  e2e           the same metric through the host ABI (ppx_cuda_host_*): pinned HOST AoS buffers in and out,
                H2D + kernels + D2H inside the timed region
  cpu_baseline  the reference's own CPU routines (oracle/_ref, OpenMP loop over ppx calls) or, if that was not
                built, the C restatement, on a bounded sample on this box's host cores (N = 1 only)
  other_workloads  kernel-only throughput + roofline fraction of the other BASELINE configs (N = 1 only)

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


def gpu_local_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None: pinned buffers allocated by a thread running there land
    in memory one PCIe hop from the device -- a remote node can cost the host-buffer path most of its PCIe rate."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id  # not available on every torch build
    except Exception:
        bdf = None
    try:
        if bdf is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                                  str(device_index)], capture_output=True, text=True).stdout.strip()
            bdf = out.lower().replace("00000000:", "0000:")
        node = int(open(f"/sys/bus/pci/devices/{str(bdf).lower()}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


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


def time_cpu(cpu_factory, sample, passes, warm=1):
    """best-of-`passes` instances/s of the CPU implementation on `sample` instances, all host threads"""
    from bench_cpu import load_cpu_oracle
    oracle, kind = load_cpu_oracle()
    fn = cpu_factory(oracle, sample)
    for _ in range(warm):
        fn()
    times = []
    for _ in range(passes):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    from bench_cpu import NTHREADS
    return sample / min(times), sample / statistics.mean(times), times, kind, NTHREADS


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path, bounded sample per step."""
    from matrix_b200.workloads_meta import META
    if rank != 0:
        return
    from bench_cpu import CPU, NTHREADS, load_cpu_oracle
    meta = META[args.workload]
    sample = args.n if args.n else meta["cpu_sample"]
    oracle, kind = load_cpu_oracle()
    fn = CPU[args.workload](oracle, sample)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} instances of {args.workload} per step (OpenMP static loop over the ppx call chain, AoS host arrays)"
    print(json.dumps({
        "impl": "reference", "metric": "instances/sec", "value": value, "unit": "instances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "instances_per_step": sample, "layout": "aos (host)",
                   "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": NTHREADS, "kind": kind,
                         "sample": desc},
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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--layout", choices=["soa", "aos"], default="soa", help="record layout of --all-ops")
    ap.add_argument("--all-ops", action="store_true",
                    help="time every batched entry point (kernel only, SoA) next to its CPU reference; one JSON line")
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

    from matrix_b200 import _lib
    from matrix_b200.workloads import WORKLOADS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def time_steps(w, steps, warmup):
        """-> (ms for all `steps` steps, CUDA events on the launching stream, max over ranks; per-step ms list)"""
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
        # one line per C-ABI entry point that is not already a BASELINE workload: kernel-only inst/s and GB/s on
        # device-resident SoA data, with the CPU reference (all host threads, 2^20-instance sample) beside it
        import time as _t
        from matrix_b200 import ops_table
        from bench_cpu import NTHREADS, cpu_op, load_cpu_oracle
        oracle, kind = load_cpu_oracle()
        rows = []
        from matrix_b200 import batch as _mb
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
                         "cpu_value": sample / cpu_s, "cpu_cores": NTHREADS,
                         "cpu_kind": kind})
            op.step = None
            torch.cuda.empty_cache()
        print(json.dumps({"all_ops": rows, "hbm_peak_gbs": hbm_peak, "layout": f"{args.layout} (device resident)",
                          "dtype": "f64"}))
        return
    cls = WORKLOADS[args.workload]
    n = args.n or cls.n_default
    w = cls(n, device, first=rank * n)
    w.setup()
    torch.cuda.synchronize()

    with ClockSampler(local) as clocks:
        launches0 = _lib.launch_count()
        total_ms, per_step = time_steps(w, args.steps, args.warmup)
        launches = _lib.launch_count() - launches0 - args.warmup
        # keep the sampler alive for a minimum window so short runs still get a few samples under load
        t_end = time.time() + max(0.0, 1.0 - total_ms / 1e3)
        while time.time() < t_end:
            w.step()
        torch.cuda.synchronize()
    value = world * n * args.steps / (total_ms * 1e-3)
    kernel_ms = statistics.mean(per_step)
    achieved = n * cls.bytes_per_instance() / (kernel_ms * 1e-3) / 1e9
    # per-launch DRAM traffic and per-instance fp64 flop counts measured once by ncu on the same seeded workload
    # (tools/make_counters.py -> profiles/counters.json); null when this n was not profiled
    try:
        with open(os.path.join(ROOT, "profiles", "counters.json")) as f:
            counters = json.load(f)
    except Exception:
        counters = {}
    c_head = counters.get(args.workload, {})
    traffic = c_head.get("dram_bytes_per_launch") if c_head.get("instances") == n else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": cls.kernel,
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_instance": cls.bytes_per_instance(),
                "instances_per_launch": n}

    # ---- e2e: host AoS buffers through the host ABI, copies inside the timed region -----------------------------
    e2e = None
    if not args.no_e2e and hasattr(w, "host_setup"):
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        per_inst = cls.bytes_per_instance()
        n_e2e = n
        while n_e2e * per_inst * 3 * max(world, 1) > avail and n_e2e > (1 << 18):
            n_e2e //= 2
        del w.T, w.Js  # free device outputs of the kernel-only leg
        torch.cuda.empty_cache()
        all_cpus = os.sched_getaffinity(0)
        near = gpu_local_cpus(local)
        if near:
            os.sched_setaffinity(0, near)  # allocate (first-touch) the pinned buffers on the GPU's NUMA node
        h2d, d2h = w.host_setup(n_e2e)
        # raw pinned-copy rates of this box: the ceiling of any host-buffer path
        probe_h = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True)
        probe_d = torch.empty(1 << 28, dtype=torch.uint8, device=device)
        pcie = {}
        for name, (dst, src) in {"h2d": (probe_d, probe_h), "d2h": (probe_h, probe_d)}.items():
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            pcie[name] = 4 * probe_h.numel() / (time.perf_counter() - t0) / 1e9
        del probe_h, probe_d
        w.host_step()  # warm-up (page-locks, stream-ordered pool growth)
        barrier()
        t0 = time.perf_counter()
        e2e_steps_ms = []
        for _ in range(args.e2e_steps):
            ts = time.perf_counter()
            loss = w.host_step()  # blocking: returns when the outputs are in host memory
            e2e_steps_ms.append((time.perf_counter() - ts) * 1e3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        dt = max_over_ranks(dt)
        e2e = {"value": world * n_e2e * args.e2e_steps / dt, "unit": "instances/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "instances_per_step": n_e2e, "steps": args.e2e_steps,
               "ms_per_step": dt / args.e2e_steps * 1e3, "api": "ppx_cuda_host_poe_fk_jacobian (pinned host AoS)",
               "pcie_gbs": (h2d + d2h) * args.e2e_steps / dt / 1e9, "pcie_h2d_gbs_measured": pcie["h2d"],
               "pcie_d2h_gbs_measured": pcie["d2h"], "result_probe": loss}
        e2e["numa_bound"] = bool(near)
        e2e["step_ms"] = {"min": min(e2e_steps_ms), "median": statistics.median(e2e_steps_ms), "max": max(e2e_steps_ms)}
        os.sched_setaffinity(0, all_cpus)
        for k in ("hq", "hT", "hJ"):
            if hasattr(w, k):
                delattr(w, k)
    del w
    torch.cuda.empty_cache()

    # ---- secondary workloads + fp64 peak + CPU baseline (rank 0 of a single-GPU run only) -----------------------
    others, fp64_peak, cpu = [], None, None
    if world == 1 and rank == 0:
        from matrix_b200 import batch as mb
        fp64_peak = mb.fp64_peak_tflops()
        if not args.no_others:
            for name, c in WORKLOADS.items():
                if name == args.workload:
                    continue
                try:
                    ww = c(c.n_default, device)
                    ww.setup()
                    ms, per = time_steps(ww, 5, 3)
                    k_ms = statistics.mean(per)
                    gbs = c.n_default * c.bytes_per_instance() / (k_ms * 1e-3) / 1e9
                    entry = {"workload": name, "instances": c.n_default, "value": c.n_default / (k_ms * 1e-3),
                             "unit": "instances/s", "kernel_ms": k_ms, "hbm_gbs": gbs,
                             "hbm_frac": gbs / hbm_peak, "expected_bound": c.bound,
                             "algorithmic_bytes_per_instance": c.bytes_per_instance()}
                    cc = counters.get(name, {})
                    if cc.get("fp64_flops_per_instance") and fp64_peak:
                        tf = cc["fp64_flops_per_instance"] * entry["value"] / 1e12
                        entry.update({"fp64_flops_per_instance_ncu": cc["fp64_flops_per_instance"],
                                      "fp64_tflops": tf, "fp64_frac_of_measured_peak": tf / fp64_peak,
                                      "fp64_pipe_pct_ncu": cc.get("fp64_pipe_pct")})
                    others.append(entry)
                    del ww
                    torch.cuda.empty_cache()
                except Exception as e:  # report, never hide
                    others.append({"workload": name, "error": repr(e)})
        if not args.no_cpu:
            from bench_cpu import CPU
            meta = META[args.workload]
            best, mean, times, kind, cores = time_cpu(CPU[args.workload], meta["cpu_sample"], passes=3)
            cpu = {"value": best, "unit": "instances/s", "cores": cores, "kind": kind,
                   "sample": f"{meta['cpu_sample']} instances of {args.workload}, best of 3 passes "
                             f"({sum(times):.1f} s of CPU work), OpenMP static loop over the ppx call chain"}

    if rank == 0:
        out = {
            "metric": "instances/sec", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "instances_per_gpu": n, "layout": f"{cls.layout} (device resident)",
                       "chain": "UR5 (test/lie_test.cpp:187-196)" if "ur5" in args.workload or "ik" in args.workload or "inverse" in args.workload
                       else None,
                       "l2": f"inputs+outputs {n * cls.bytes_per_instance() / 1e6:.0f} MB per step >> 126 MB L2, "
                             "evict-first loads/stores", "parallelism": f"instance-sharded x{world}, no collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(), "fp64_peak_tflops_measured": fp64_peak,
            "step_ms": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step)},
            "other_workloads": others,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
