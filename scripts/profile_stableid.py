"""Timing / ncu driver for the StableIdentification sweep (config 3 as a batch): python scripts/profile_stableid.py [pairs=2048] [reps=2]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import riptrm_b200 as rb
import bench


class A:
    pass


pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out = []
print(json.dumps(bench.stableid_leg(A(), rb, out, pairs=pairs)))
