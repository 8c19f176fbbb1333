#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the bench command, full ncu capture of the dominant kernel on cfg2 and on
# a cfg4-shaped launch, compute-sanitizer passes.  Outputs land in gpurun_out/ (summaries are copied to profiles/ by hand).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|screen_|prep_rows|ovf_|merge_|absmax|fix_scale|init_aux|diff_small|select_|remap" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches_bench.err
ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 2 -c 1 -o gpurun_out/r2_screen_cfg2 -f \
    python scripts/one_search.py cfg2 0 3 > gpurun_out/r2_ncu_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:screen_finalize -s 2 -c 1 -o gpurun_out/r2_finalize_cfg2 -f \
    python scripts/one_search.py cfg2 0 3 > gpurun_out/r2_ncu_cfg2_fin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 6 -c 1 -o gpurun_out/r2_screen_cfg4 -f \
    python scripts/one_search.py cfg4 0 2 > gpurun_out/r2_ncu_cfg4.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_shapes.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_shapes.py > gpurun_out/r2_sanitizer_synccheck.log 2>&1
for f in gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_synccheck.log gpurun_out/r2_ncu_cfg2.log gpurun_out/r2_ncu_cfg4.log; do tail -n 3 $f; done
ls -la gpurun_out | tail -20
