"""Config 4 on the GPU alone: beam widths 1/10/50/200 at T=4000 (bench.bench_cfg4) -> one JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = argparse.Namespace(steps=5, warmup=3, scaling="weak")
print(json.dumps(bench.bench_cfg4(args, 0, 1, torch.device("cuda", 0))["results"]))
