"""GPU box: the secondary corridor measurement of bench.py on its own.

    python scripts/corridor_bench.py [agents] [steps]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
line, gbs = bench.corridor_measure(n, steps)
line["achieved_gbs"] = gbs
bench.emit(line)
