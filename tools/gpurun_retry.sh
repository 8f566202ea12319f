#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...> ; retries while the pod answers busy (nothing charged)
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt busy; sleeping 90 s"
  sleep 90
done
exit 3
