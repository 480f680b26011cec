#!/bin/bash
# Phase breakdown of the one-shot call + the other single-GPU workloads.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
SPARTA_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
grep -E "sparta" gpurun_out/bench_e2e.err | tail -12
timeout 600 python bench.py --workload er14_fixed --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_er14.json 2> gpurun_out/bench_er14.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_er14.json",):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"])
PY
timeout 900 python bench.py --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_tf32.json").read().strip().splitlines()[-1])
print("tf32", d["value"], d["ms_per_step"], d["check"])
PY
python - <<'PY'
# raw H2D / D2H bandwidth of this box from pinned memory, for the e2e floor
import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, x), ("d2h", x, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, "GB/s", 4 * (1<<30) / dt / 1e9)
t = time.perf_counter(); y = torch.empty(4<<30, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); print("cudaMalloc 4GiB ms", (time.perf_counter()-t)*1e3)
PY
