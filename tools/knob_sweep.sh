# ESCORT_TM_CHS (channels per slot group) x ESCORT_TM_NSMIN (min. shared-memory stages) sweep of TMEM variants
for spec in ${LAYERS:-"alexnet 1" "alexnet 0" "alexnet 3" "resnet50 3" "resnet50 7" "resnet50 13"}; do
for v in ${VARIANTS:-58 51}; do for chs in ${CHSS:-2 3 4 5 6 8}; do for ns in ${NSS:-4 3}; do
  ESCORT_TM_CHS=$chs ESCORT_TM_NSMIN=$ns python tools/run_layer.py $spec $v 5 > /tmp/ks.log 2>&1
  echo "$spec v$v CHS=$chs NSMIN=$ns | $(grep -o 'CHS=[0-9]* NSLOT=[0-9]*' /tmp/ks.log) $(grep -o 'CI=[0-9]* nchunks=[0-9]*' /tmp/ks.log) $(grep -o ' NS=[0-9]*' /tmp/ks.log) | $(grep RESULT /tmp/ks.log | sed 's/.*: //')"
done; done; done; done
