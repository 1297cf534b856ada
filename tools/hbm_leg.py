"""The HBM-bound CG leg of bench.py alone (503 k-node mesh, plain streaming kernel): python tools/hbm_leg.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
peaks, _ = bench.measured_peaks()
print(json.dumps(bench.hbm_regime_leg(0, peaks)))
