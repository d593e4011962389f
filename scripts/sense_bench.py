"""Stand-alone run of bench.py's secondary local-map acquisition measurement (prints one JSON object)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    print(json.dumps(bench.sense_measure(n, int(sys.argv[2]) if len(sys.argv) > 2 else 10)), file=sys.stderr)
