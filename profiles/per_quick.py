import sys, json, torch
sys.path.insert(0, '.')
import bench
hbm, _ = bench.peaks()
out = bench.per_microbench("cuda:0", torch, hbm)
print(json.dumps(out, indent=1))
