#!/bin/bash
# retry a gpurun call while the pod answers "transient" / busy (exit code 3); usage: tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  st=$(python3 -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i: status=$st rc=$rc; sleeping 120 s"
  sleep 120
done
exit 3
