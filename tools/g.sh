#!/bin/bash
# build the library here (CPU box), then run the given command on a B200 through gpurun; retries while the pod is busy
cd "$(dirname "$0")/.."
python -c "import crfr_b200; crfr_b200.build()" || exit 1
for i in $(seq 1 ${GPU_RETRIES:-30}); do
  out=$(gpurun --timeout "${GPU_TIMEOUT:-600}" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
