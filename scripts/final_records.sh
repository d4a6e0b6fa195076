#!/bin/bash
# Final records of a tree on one B200 (run under gpurun from the repo root): device suite, bench line,
# ncu launch list of the bench, one full ncu capture of the moment kernel.  $1: tag of the output files.
tag=${1:-r2_zf}
out=gpurun_out
timeout 300 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
tail -1 $out/${tag}_pytest.log
timeout 200 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cut -c1-300 $out/${tag}_bench_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no_cpu_baseline > $out/${tag}_ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k lag_moment_warp_kernel -s 2 -c 1 -f \
  -o $out/${tag}_warp python scripts/profile_one.py 1.1e6 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
