#!/usr/bin/env python
"""The GP-predict side metric of bench.py alone (n = 512, d = 6, 65536 queries): us/pt mean + variance and mean only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda', 0)
out = bench.gp_predict_line(dev)
out['peaks'] = bench.measured_peaks_in_run(dev)
print(json.dumps(out))
