"""One forward step of the bench workload inside a cudaProfilerStart/Stop window, for ncu:

  python profiles/profile_step.py --workload cfg2 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py --workload cfg2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--mode", default="eager", choices=["eager", "graph", "inverse"])
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()

dev = torch.device("cuda:0")
model = bench.build_model(args.workload, dev)
coupling, image, L, K, hidden, batch = bench.WORKLOADS[args.workload]
x = bench.synthetic_batches(1, batch, image, seed=5)[0].to(dev)
with torch.no_grad():
    if args.mode == "graph":
        from flowk.graphs import GraphedDensity
        g = GraphedDensity(model, x)
        step = lambda: g.run(x)
    elif args.mode == "inverse":
        z, outs, _ = model.flow.encode_latents(x, x.new_zeros(batch))
        step = lambda: model.flow.decode_latents(z, outs)
    else:
        step = lambda: model(x)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled", args.mode, args.workload)
