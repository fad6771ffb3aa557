#!/bin/bash
# debug aid: run the probe over a matrix, each in its own process with a hard kill
cd "$(dirname "$0")/.."
for n in ramp_1080p random_odd solid alpha natural narrow one_px; do
  timeout -s KILL 60 python tools/gpu_probe3.py $n 2>&1 | grep -E "start|returned|ok|rror|ssert" ; echo "   rc=${PIPESTATUS[0]}"
done
