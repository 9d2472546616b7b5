set -u
O=gpurun_out
OLD=$PWD/lzma-sdk-zliblike_b200/lib/variants/old/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none --workload lzma2_8g"
run zz2_l2_new $B
run zz2_l2_old env LZB_LIB=$OLD $B
run zz2_l2_new2 $B
