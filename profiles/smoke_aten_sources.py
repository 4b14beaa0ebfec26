"""Which Python lines launch the library (ATen / cuBLAS / MAGMA) kernels that show up in smoke()'s launch list: runs
__graft_entry__.smoke() under torch.profiler with stacks and prints the ops that own device time, grouped by call site.
Usage (GPU box): python profiles/smoke_aten_sources.py > gpurun_out/smoke_aten_sources.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.smoke()                                                   # warm: library handles, lazy module state
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    G.smoke()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_stack_n=8):
    dev_us = getattr(ev, "self_device_time_total", None)
    if dev_us is None:
        dev_us = ev.self_cuda_time_total
    if dev_us <= 0:
        continue
    stack = [s for s in ev.stack if "/repo/" in s or "flowk" in s][:3]
    rows.append((dev_us, ev.count, ev.key, " <- ".join(s.strip().split("/")[-1] for s in stack)))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print("device time owned by torch ops: %.0f us over %d call sites" % (total, len(rows)))
for us, n, key, where in rows[:70]:
    print("%8.0f us %5d  %-34s %s" % (us, n, key[:34], where))
