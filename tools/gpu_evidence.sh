#!/bin/bash
# One GPU session that produces the evidence kept under profiles/ for a round:
# the GPU test suite, one bench line per workload, the reference arm, the ncu launch list
# of the default bench command, a full-size `ncu --set full` capture of the sweep kernel,
# and compute-sanitizer memcheck / racecheck on a small solve.
# Usage: tools/gpu_evidence.sh <tag> [skip-list, e.g. "pytest sanit"]
tag=$1; skip=" $2 "
o=gpurun_out; mkdir -p $o
export PYTHONPATH=.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.log 2>&1
if [[ "$skip" != *" pytest "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1
  echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
  tail -3 $o/${tag}_pytest.log
fi
if [[ "$skip" != *" bench "* ]]; then
  for w in c5 c4 c3 c2 c1; do
    timeout 600 python bench.py --workload $w > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
    echo "bench $w exit $?"
  done
  timeout 600 python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
  echo "bench reference exit $?"
fi
if [[ "$skip" != *" launches "* ]]; then
  timeout 600 python bench.py --steps 2 --warmup 3 > $o/${tag}_plain_bench.json 2> $o/${tag}_plain_bench.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $o/${tag}_launches_ncu.log 2>&1
  echo "launch list exit $?"
fi
if [[ "$skip" != *" ncu "* ]]; then
  timeout 300 python tools/profile_sweep.py 100000 400 > $o/${tag}_sweep_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intr_sweep -s 2 -c 1 -f -o $o/${tag}_sweep \
    python tools/profile_sweep.py 100000 400 > $o/${tag}_sweep_ncu.log 2>&1
  echo "ncu full exit $?"; tail -2 $o/${tag}_sweep_plain.log
fi
if [[ "$skip" != *" ubench "* ]] && [ -x tools/bin/ubench_fp64block ]; then
  timeout 120 tools/bin/ubench_fp64block > $o/${tag}_ubench_fp64block.txt 2>&1
  echo "ubench exit $?"; tail -3 $o/${tag}_ubench_fp64block.txt
fi
if [[ "$skip" != *" sanit "* ]]; then
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py > $o/${tag}_sanitizer_$tool.log 2>&1
    echo "$tool exit $?"; tail -4 $o/${tag}_sanitizer_$tool.log
  done
fi
