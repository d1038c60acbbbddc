#!/bin/bash
# build the library here (CPU box), then run the given command on a B200 through gpurun
set -e
cd "$(dirname "$0")/.."
python -c "import crfr_b200; crfr_b200.build()"
exec gpurun --timeout "${GPU_TIMEOUT:-600}" -- "$@"
