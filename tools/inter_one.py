#!/usr/bin/env python3
"""One inter-coded case of tools/bench_configs.py, for profiling K3 / K2 on skip-heavy streams: inter_one.py [lm]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import bench_configs as B  # noqa: E402
lm = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B.run("inter lm=cm=%d" % lm, 720, 576, 128, 2048, key_rate=29, lm=lm, cm=lm, noise_y=2, steps=int(os.environ.get("BENCH_ONE_STEPS", "3")))
