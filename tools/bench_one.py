#!/usr/bin/env python3
"""One case of tools/bench_configs.py, for profiling: bench_one.py W H Q FRAMES [noise_y noise_c]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import bench_configs as B  # noqa: E402

w, h, q, f = (int(x) for x in sys.argv[1:5])
ny = int(sys.argv[5]) if len(sys.argv) > 5 else 2
nc = int(sys.argv[6]) if len(sys.argv) > 6 else 0
B.run("case %dx%d Q%d" % (w, h, q), w, h, q, f, noise_y=ny, noise_c=nc, steps=int(os.environ.get("BENCH_ONE_STEPS", "2")))
