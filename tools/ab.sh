#!/bin/bash
# Development aid (GPU box): stage times of configs[1]-shaped batches for each library variant given.
for v in "$@"; do
    for rep in 1 2; do
        RTJPEG_B200_LIBFILE=$PWD/gmerlin-avdecoder_b200/lib/variant_$v.so python tools/bench_one.py 720 576 128 ${AB_FRAMES:-2048} | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stage_ms'].items()})"
    done
done
