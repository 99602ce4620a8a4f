set -x
mkdir -p gpurun_out
for cfg in "56 30" "56 100" "33 30"; do set -- $cfg
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_bench -s 1 -c 1 -f -o gpurun_out/s3_v$1_pl$2 python tools/interp_one.py $1 $2 > gpurun_out/s3_v$1_pl$2.log 2>&1; tail -2 gpurun_out/s3_v$1_pl$2.log
done
for aw in 4 8; do python tools/interp_one.py 56 30 $aw; python tools/interp_one.py 56 100 $aw;  python tools/interp_one.py 56 12 $aw; done
