#!/bin/bash
# Developer tool: ONE gpurun call that produces everything tools/refresh_profiles.py turns into profiles/*.
#
#   python tools/phase_timing.py --build        # here (no GPU): the -DHCM_PHASE_TIMING library travels with the snapshot
#   gpurun --timeout 420 -- 'bash tools/capture_profiles.sh r02a'
#   python tools/refresh_profiles.py r02a       # here: reads gpurun_out/r02a_*, rewrites the tracked summaries
#
# Outputs under gpurun_out/ (prefix = $1):
#   <p>_gpu_tests.log     pytest -m gpu (the parity gate comes first: a profile of a wrong kernel is worthless)
#   <p>_bench_full.json   python bench.py                                  -- the bench line (never taken under a profiler)
#   <p>_launches_r20.csv  ncu launch list of bench.py --steps 2 --warmup 1 -- per-launch duration + DRAM bytes
#   <p>_prof_mesh.ncu-rep ncu --set full, mesh_kernel, r = 11 sphere       -- details / source / raw pages are read from it
#   <p>_prof_derive.ncu-rep  the same for derive_kernel
#   <p>_phase_timing.txt  tools/phase_timing.py (clock64 per phase)
# Budget: about 3.5 GPU-minutes of box time (tests 30 s, bench 20 s, launch list 75 s, two captures 25 s each, phases 15 s).
set -u
p=${1:?prefix}
o=gpurun_out
mkdir -p $o
NCU="ncu --clock-control none"

timeout 150 python -m pytest tests -m gpu -x -q > $o/${p}_gpu_tests.log 2>&1
tail -1 $o/${p}_gpu_tests.log
grep -q " passed" $o/${p}_gpu_tests.log && ! grep -q "failed" $o/${p}_gpu_tests.log || { echo "parity gate failed: no profiles taken"; exit 1; }

timeout 120 python bench.py > $o/${p}_bench_full.json 2> $o/${p}_bench_full.err
cut -c1-300 $o/${p}_bench_full.json

timeout 150 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $o/${p}_launches_r20.csv \
	python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/${p}_ncu_bench.log 2>&1
wc -l $o/${p}_launches_r20.csv

# one launch of each kernel on the r = 11 sphere (the application's default load radius): -k picks the kernel, -s skips the
# warm-up launch, -c 1 keeps the replay count down (ncu --set full replays a kernel about 40 times)
R=11 timeout 120 $NCU --set full --import-source on -k regex:mesh_kernel -s 1 -c 1 -f -o $o/${p}_prof_mesh \
	python tools/rederive_once.py > $o/${p}_prof_mesh.log 2>&1
R=11 timeout 120 $NCU --set full --import-source on -k regex:derive_kernel -s 1 -c 1 -f -o $o/${p}_prof_derive \
	python tools/rederive_once.py > $o/${p}_prof_derive.log 2>&1
ls -la $o/${p}_prof_mesh.ncu-rep $o/${p}_prof_derive.ncu-rep

if [ -f hypercraft_b200/csrc/libhcm_timing.so ]; then
	timeout 90 python tools/phase_timing.py > $o/${p}_phase_timing.txt 2>&1
	head -3 $o/${p}_phase_timing.txt
else
	echo "no libhcm_timing.so in the snapshot (python tools/phase_timing.py --build before the gpurun call)" | tee $o/${p}_phase_timing.txt
fi
