#!/bin/bash
# usage: scripts/gpurun_multi.sh <N gpus> <log> <timeout>   -- multi-GPU bench check with a slimmer snapshot (a multi-GPU call is charged
# N x the push time too): files only the tests / the reference arm need are left out of THIS call's snapshot, then .gpurunignore is put back.
N=$1; LOG=$2; TO=${3:-1200}
cp .gpurunignore .gpurunignore.keep
cat >> .gpurunignore <<'IGN'
oracle/_ref/models/offline/translator.onnx
oracle/_ref/models/streaming/translator.onnx
oracle/_ref/models/punc
oracle/_ref/models/vad
tests/golden
IGN
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $TO -- "bash scripts/gpu_r02_ngpu.sh $N" > $LOG 2>&1
  if grep -q "status=ok\|status=fail" $LOG; then break; fi
  sleep 120
done
mv .gpurunignore.keep .gpurunignore
tail -20 $LOG | cut -c1-400
