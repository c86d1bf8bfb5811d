#!/bin/bash
# Developer tool: parity gate + short bench + phase cycles in one gpurun call (about 1.5 GPU-minutes).
#   python tools/phase_timing.py --build     # here (no GPU), optional
#   gpurun --timeout 400 -- 'bash tools/quick_gpu.sh r02b'
set -u
p=${1:?prefix}
o=gpurun_out
mkdir -p $o
timeout 200 python -m pytest tests -m gpu -x -q > $o/${p}_gpu_tests.log 2>&1
tail -3 $o/${p}_gpu_tests.log
timeout 150 python bench.py --no-cpu-baseline > $o/${p}_bench.json 2> $o/${p}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$o/${p}_bench.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms")}, d["e2e"]["ms_per_step"])
except Exception as e:
    print("bench failed", e); print(open("$o/${p}_bench.err").read()[-2000:])
PY
if [ -f hypercraft_b200/csrc/libhcm_timing.so ]; then
	timeout 90 python tools/phase_timing.py > $o/${p}_phase_timing.txt 2>&1
	cat $o/${p}_phase_timing.txt
fi
