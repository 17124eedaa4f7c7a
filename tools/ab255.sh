#!/bin/bash
# Development aid (GPU box): like tools/ab.sh, at a quality of 255 (frames with a raw prefix) and at configs[1].
for v in "$@"; do
  for cfg in "255 1024" "255 2048" "128 4096"; do
    set -- $cfg
    RTJPEG_B200_LIBFILE=$PWD/gmerlin-avdecoder_b200/lib/variant_$v.so python tools/bench_one.py 720 576 $1 $2 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'Q$1', $2, round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stage_ms'].items()})"
  done
done
