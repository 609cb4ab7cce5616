#!/bin/bash
# usage: profiles/gpurun_retry.sh LOGFILE TIMEOUT 'command'  -- retries while the pod answers busy/transient (nothing charged)
log="$1"; tmo="$2"; shift 2
cd /root/repo
for attempt in $(seq 1 40); do
  gpurun --timeout "$tmo" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|rc=3\|answers busy\|retry in a few minutes" "$log"; then sleep 45; continue; fi
  break
done
echo "[retry] finished after $attempt attempt(s)" >> "$log"
