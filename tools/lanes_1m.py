import sys, statistics, torch
sys.path.insert(0, ".")
from matrix_b200 import workloads as W
dev = torch.device("cuda", 0)
for lanes, nbuf in ((1, 8), (2, 8), (4, 8), (8, 8), (8, 16)):
    W.Se3ExpLog1M.LANES, W.Se3ExpLog1M.NBUF = lanes, nbuf
    w = W.Se3ExpLog1M(1 << 20, dev); w.setup()
    for _ in range(5): w.step()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    torch.cuda.synchronize(); ev[0].record()
    for i in range(20): w.step(); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = statistics.mean(ev[i].elapsed_time(ev[i + 1]) for i in range(20)) / nbuf
    print(f"lanes {lanes} nbuf {nbuf}: {ms*1e3:.2f} us per 1M launch, {(1<<20)/ms/1e6:.2f} G inst/s, {(1<<20)*96/ms/1e6/6550.4:.3f} of HBM peak")
    del w; torch.cuda.empty_cache()
