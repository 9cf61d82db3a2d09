#!/usr/bin/env bash
# tools/gpurun_retry.sh SCRIPT OUT [TIMEOUT]: run a call script through gpurun, retrying while the pod answers busy (exit 3)
s=$1; out=$2; t=${3:-2400}
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $t -- bash $s > $out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
