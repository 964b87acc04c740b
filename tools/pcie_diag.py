"""Where the host-buffer path's time goes on a multi-GPU box: one process, all visible devices.

    python tools/pcie_diag.py [--gb 2] [--devices all|0,1] [--e2e-n INSTANCES_PER_DEVICE]

Prints one JSON object: the box (cores, memory, THP, topology), how long page-locking takes with and without huge
pages, the D2H / H2D / bidirectional rate of every device alone and of all devices at once (ppx_cuda_host_copy_probe:
aligned start, same pinned buffer, no kernels), and the end-to-end rate of ppx_cuda_host_poe_fk_jacobian over the same
devices.  PPX_HOST_SLOTS / PPX_HOST_CHUNK_MIB are read once per process: sweep them from the shell.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return repr(e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0, help="pinned GB per device for the copy probes")
    ap.add_argument("--devices", default="all")
    ap.add_argument("--e2e-n", type=int, default=1 << 22, help="instances per device of the end-to-end call")
    ap.add_argument("--skip-box", action="store_true")
    ap.add_argument("--skip-solo", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch
    from matrix_b200 import host, synth
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if args.devices == "all" else [int(x) for x in args.devices.split(",")]
    out = {"devices": devs, "env": {k: os.environ.get(k) for k in ("PPX_HOST_SLOTS", "PPX_HOST_CHUNK_MIB",
                                                                   "PPX_HOST_HUGEPAGES", "PPX_HOST_ALLOC")}}
    if not args.skip_box:
        out["box"] = {"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                      "meminfo": sh("grep -E 'MemTotal|MemAvailable|Hugepagesize|AnonHugePages|HugePages_Total' /proc/meminfo"),
                      "thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled; cat /sys/kernel/mm/transparent_hugepage/defrag"),
                      "numa": sh("lscpu | grep -i -E 'numa|socket|model name|thread'"),
                      "ulimit_l": sh("ulimit -l"),
                      "topo": sh("nvidia-smi topo -m | head -14"),
                      "pcie": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current "
                                 "--format=csv,noheader")}
    nbytes = int(args.gb * len(devs) * (1 << 30))
    torch.zeros(1, device="cuda").item()  # CUDA context up before anything is timed
    warm = host.pinned_empty((1 << 20,), np.uint8)
    del warm
    pin = {}
    for hp in ("1", "0"):
        os.environ["PPX_HOST_HUGEPAGES"] = hp
        t0 = time.perf_counter()
        buf = host.pinned_empty((nbytes,), np.uint8)
        pin[f"hugepages={hp}"] = {"seconds": time.perf_counter() - t0, "gb": nbytes / 1e9}
        host.set_devices(devs)
        pin[f"hugepages={hp}"]["d2h_all"] = host.copy_probe(buf, "d2h", 0.7)
        host.set_devices(None)
        if hp == "0":
            del buf
    os.environ["PPX_HOST_HUGEPAGES"] = "1"
    out["pin"] = pin
    buf = host.pinned_empty((nbytes,), np.uint8)
    probes = {}
    if not args.skip_solo:
        solo = {}
        for d in devs:
            host.set_devices([d])
            solo[d] = {k: host.copy_probe(buf, k, 0.4)[0] for k in ("d2h", "h2d")}
        probes["solo"] = solo
    host.set_devices(devs)
    for k in ("d2h", "h2d", "both"):
        r = host.copy_probe(buf, k, 1.0)
        probes[f"all_{k}"] = {"per_device": r, "aggregate": sum(r)}
    if len(devs) >= 4:
        half = devs[: len(devs) // 2]
        host.set_devices(half)
        r = host.copy_probe(buf, "d2h", 0.7)
        probes["first_half_d2h"] = {"per_device": r, "aggregate": sum(r)}
        host.set_devices(devs[len(devs) // 2:])
        r = host.copy_probe(buf, "d2h", 0.7)
        probes["second_half_d2h"] = {"per_device": r, "aggregate": sum(r)}
        host.set_devices(devs[::2])
        r = host.copy_probe(buf, "d2h", 0.7)
        probes["even_d2h"] = {"per_device": r, "aggregate": sum(r)}
    out["probes"] = probes
    del buf
    # end to end: one call over all devices
    n = args.e2e_n * len(devs)
    q = host.pinned_empty((n, 6))
    blk = 1 << 22
    for lo in range(0, n, blk):
        q[lo:lo + blk] = synth.joint_angles(min(blk, n - lo), first=lo)
    T, Js = host.pinned_empty((n, 16)), host.pinned_empty((n, 36))
    kin = host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    e2e = {}
    for label, dl in (("all", devs), ("first", devs[:1])):
        host.set_devices(dl)
        m = n if label == "all" else args.e2e_n
        kin.fk_jacobian(q[:m], out_T=T[:m], out_Js=Js[:m])
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            kin.fk_jacobian(q[:m], out_T=T[:m], out_Js=Js[:m])
            ts.append(time.perf_counter() - t0)
        e2e[label] = {"instances": m, "best_s": min(ts), "minst_s": m / min(ts) / 1e6, "gbs": m * 464 / min(ts) / 1e9}
    host.set_devices(None)
    out["e2e"] = e2e
    print(json.dumps(out))


if __name__ == "__main__":
    main()
