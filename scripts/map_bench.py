"""GPU box: the secondary local-map measurement of bench.py on its own.   python scripts/map_bench.py [grids] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bench.emit(bench.map_measure(n, steps))
