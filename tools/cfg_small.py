"""BASELINE configs[0] / configs[1] on the GPU from the committed fixture (bench.py: small_configs), one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fawkes_crypto_b200 as fb
import bench
ctx = fb.Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
print(json.dumps(bench.small_configs(fb, ctx, torch, 0, flush, "--cpu" in sys.argv)))
