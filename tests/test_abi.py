"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/ppx_cuda.h
declares, the C++14 host header compiles and links against it, compute entry points fail loudly without a
GPU (no CPU fallback), and the host logic around the kernels (synthetic generator, instance sharding across
ranks -- exercised with a world_size-2 gloo group) behaves."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from matrix_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppx_cuda.h")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol():
    import ctypes
    from matrix_b200 import _lib
    declared = set(re.findall(r"\b(ppx_cuda_\w+)\s*\(", open(HEADER).read()))
    assert len(declared) >= 50
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in ppx_cuda.h but not exported by libppx_cuda.so"
    assert declared == set(_lib.EXPORTS), "Python binding and header out of sync"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ppx_cuda_\w+)", out))
    assert declared <= exported
    assert "sm_100a" in _lib.version()
    assert ctypes.c_char_p(_lib.lib.ppx_cuda_error_string(-2)).value.startswith(b"shape")
    assert ctypes.c_char_p(_lib.lib.ppx_cuda_error_string(-5)).value.startswith(b"host-side")  # PPX_ERR_HOST


def test_library_carries_sm_100a_code_only():
    from matrix_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_cpp14_host_header_compiles_and_links(tmp_path):
    exe = tmp_path / "host_api_test"
    cmd = ["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-L", os.path.join(ROOT, "matrix_b200"),
           "-lppx_cuda", f"-Wl,-rpath,{os.path.join(ROOT, 'matrix_b200')}", "-o", str(exe)]
    subprocess.check_call(cmd)
    if _no_gpu():
        # no device: the API must fail loudly (exception -> abort), never compute on the CPU
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "libppx_cuda" in (r.stderr + r.stdout)


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_compute_entry_points_fail_loudly_without_a_gpu():
    from matrix_b200 import host
    from matrix_b200._lib import PpxCudaError
    with pytest.raises(PpxCudaError):
        host.se3_exp_log(synth.se3_twists(16))
    with pytest.raises(PpxCudaError):
        host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME).fk_jacobian(synth.joint_angles(16))


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "matrix_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_lib" not in text and "ppx_oracle" not in text and "ppxo_" not in text, \
                    f"{f} references the oracle"
    for f in ("ppx_cuda.h", os.path.join("ppx", "ppx.hpp")):
        text = open(os.path.join(ROOT, "include", f)).read()
        assert "oracle" not in text.lower()


def test_generator_is_counter_based():
    a = synth.joint_angles(1000)
    b = np.vstack([synth.joint_angles(300), synth.joint_angles(700, first=300)])
    assert np.array_equal(a, b)  # any shard can be regenerated anywhere
    u = synth.u01(1, 0, 200000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    xi = synth.se3_twists(5000)
    th = np.linalg.norm(xi[:, :3], axis=1)
    assert th.min() >= 0.05 - 1e-12 and th.max() <= np.pi - 0.05 + 1e-12
    S = synth.symmetric(100, 4).reshape(100, 4, 4)
    assert np.array_equal(S, S.transpose(0, 2, 1))


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 7, 16, 1 << 24, (1 << 28) + 3):
        for world in (1, 2, 3, 4, 8):
            r = shard.shard_ranges(n, world)
            assert sum(c for _, c in r) == n
            pos = 0
            for first, count in r:
                assert first == pos or count == 0
                pos += count
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _rank_main(rank, world, port, n_total, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.shard_range(n_total, rank, world)
    q = torch.from_numpy(synth.joint_angles(count, first=first))  # this rank regenerates only its block
    per = shard.shard_range(n_total, 0, world)[1]
    pad = torch.zeros((per, 6), dtype=torch.float64)
    pad[:count] = q
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.cat([g[:shard.shard_range(n_total, r, world)[1]] for r, g in enumerate(gathered)])
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
        np.save(os.path.join(tmp, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reassembles_the_batch(tmp_path):
    import torch.multiprocessing as mp
    n_total, world = 100003, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(full, synth.joint_angles(n_total))  # bitwise the single-process batch
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0


def test_reference_arm_of_the_bench_runs_on_the_host():
    """`bench.py --impl reference` (the arm the driver times beside the GPU arm) needs no GPU: one JSON line with the
    contract's keys, metric and config of the GPU arm, zero GPU launches."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--n", "65536"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "instances/sec" and d["unit"] == "instances/s"
    assert d["config"]["workload"] == "ur5_fk_jacobian" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d.get("gpu_launches", 0) == 0
    # both arms describe the run with the same `config` (workload, instances per GPU and step, chain, N)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_for("ur5_fk_jacobian", 65536, 1)
    from matrix_b200.workloads_meta import HEADLINE, META
    assert META[HEADLINE]["n"] == 1 << 24  # the default step of the reference arm = the GPU arm's 16M instances


def test_bench_tables_cover_every_baseline_config():
    """every BASELINE.json config has a workload with its size, algorithmic bytes and a CPU leg (bench.py's `configs`)"""
    import bench_cpu
    from matrix_b200.workloads_meta import META
    want = {"se3_exp_log_1M": (1 << 20, 96), "se3_exp_log": (1 << 26, 96), "ur5_fk_jacobian": (1 << 24, 464),
            "linsolve_qr_4x3": (1 << 26, 153), "svd_6x6": (1 << 24, 912), "eig_sym_4": (1 << 24, 288),
            "ik_newton_step": (1 << 25, 224)}
    for name, (n, nbytes) in want.items():
        assert META[name]["n"] == n and META[name]["bytes"] == nbytes and name in bench_cpu.CPU, name
    assert 8 * META["ik_newton_step"]["n"] == 1 << 28  # config 5: 256M instances over 8 GPUs
    # the headline is benched on arrays of the reference's records (the kernel the host face launches as well), and
    # what META says about a workload's device layout is what its class does
    from matrix_b200 import workloads
    from matrix_b200.workloads_meta import HEADLINE
    assert META[HEADLINE]["layout"] == "records" and "AOS" in workloads.WORKLOADS[HEADLINE].kernel
    for name, cls in workloads.WORKLOADS.items():
        assert (cls.layout == "aos") == (META[name].get("layout") == "records"), name
