"""GPU box: the secondary reference-trajectory measurement of bench.py on its own.
    python scripts/reftraj_bench.py [agents] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bench.emit(bench.reftraj_measure(n, steps))
