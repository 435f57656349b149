#!/bin/bash
# gpurun with retries on "no slot right now" (exit code 3; nothing is charged for those)
for try in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
