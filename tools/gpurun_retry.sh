#!/usr/bin/env bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those).  usage: tools/gpurun_retry.sh <log> <timeout> <command...>
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" -- "$@" > "$LOG" 2>&1
  if grep -q "status=ok\|status=failed\|status=timeout" "$LOG"; then break; fi
  if ! grep -q "transient\|busy\|retry" "$LOG"; then break; fi
  sleep 90
done
tail -40 "$LOG"
