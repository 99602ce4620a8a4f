# measurement only: ESCORT_TM_SKIP bit 1 = no window fill, 2 = no taps, 4 = no input loads (results are wrong on purpose)
for L in "alexnet 1" "alexnet 0" "resnet50 3"; do
for v in ${VARIANTS:-59}; do
for s in 0 1 2 3 4 5 6 7; do
echo "== $L v$v skip=$s"
ESCORT_TM_SKIP=$s python tools/run_layer.py $L $v 4 2>&1 | grep RESULT
done; done; done
