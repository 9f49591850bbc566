#!/bin/bash
# Where does the reference's sm_100 rebuild stop on an input that takes its NSPARSE step-1 path?
# usage: tools/ref_probe.sh <seconds>   (run on the GPU box)
set -u
cd "$(dirname "$0")/.."
mkdir -p /tmp/rp gpurun_out
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import pem_spgemm_b200 as pem
from pem_spgemm_b200 import synth
rows, cols, I, J, V = synth.laplacian2d(600)
pem.mtx_write("/tmp/rp/lap600.mtx", rows, cols, I, J, V)
PY
cd /tmp/rp
timeout "${1:-60}" stdbuf -o0 -e0 "$OLDPWD/oracle/_ref/pemspgemm_ref" /tmp/rp/lap600.mtx 0 > "$OLDPWD/gpurun_out/ref_probe_lap600.log" 2>&1
echo "exit code $?" >> "$OLDPWD/gpurun_out/ref_probe_lap600.log"
tail -25 "$OLDPWD/gpurun_out/ref_probe_lap600.log"
