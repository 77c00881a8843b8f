#!/bin/bash
# quick confirmation of HEAD on a B200: full GPU parity suite + the driver's smoke entry
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
