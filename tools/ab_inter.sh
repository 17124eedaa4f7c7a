for v in "$@"; do for rep in 1 2; do RTJPEG_B200_LIBFILE=$PWD/gmerlin-avdecoder_b200/lib/variant_$v.so python - <<PY
import sys,os,json
sys.path.insert(0,os.getcwd())
import tools.bench_configs as B
import io,contextlib
buf=io.StringIO()
with contextlib.redirect_stdout(buf):
    B.run("inter", 720, 576, 128, 2048, key_rate=29, lm=1, cm=1, noise_y=2)
d=json.loads(buf.getvalue()); print("$v", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["stage_ms"].items()})
PY
done; done
