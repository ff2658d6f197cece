#!/bin/bash
# Round-1 (second session) evidence run on ONE B200: bench lines, ncu launch list, ncu full-set of the four pass kernels.
# Usage (from the repo root, under gpurun):  bash profiles/collect_r1b.sh
set -x
OUT=gpurun_out/r1b
mkdir -p $OUT
python bench.py > $OUT/bench_1gpu_batch256x2048.json 2> $OUT/bench_1gpu.err
python bench.py --impl reference > $OUT/bench_reference_arm.json 2>> $OUT/bench_1gpu.err
python bench.py --workload rgb4096 --flush-l2 --steps 10 --no-cpu-baseline > $OUT/bench_1gpu_rgb4096.json 2>> $OUT/bench_1gpu.err
python bench.py --workload rgb16384 --steps 3 --warmup 3 --no-cpu-baseline --no-check > $OUT/bench_1gpu_rgb16384.json 2>> $OUT/bench_1gpu.err
python bench.py --workload car --flush-l2 --steps 20 > $OUT/bench_1gpu_car.json 2>> $OUT/bench_1gpu.err
python bench.py --workload cat --flush-l2 --steps 20 > $OUT/bench_1gpu_cat.json 2>> $OUT/bench_1gpu.err
python profiles/check_col_variants.py > $OUT/check_col_variants.txt 2>&1
python profiles/time_passes.py 2048 3 12 48 > $OUT/time_passes_2048.txt 2>&1
python profiles/probe_pipe.py 2048 12 102,104,108,116,132,7,8,6 > $OUT/tma_copy_probe.txt 2>&1
# launch list of the bench command itself (one timed step of 32 images, cold caches, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launch_list_bench_images32.csv \
    python bench.py --steps 1 --warmup 3 --images 32 --no-e2e --no-cpu-baseline --no-check > $OUT/ncu_ll.log 2>&1
# full set of the four pass kernels on 12 plane pairs of 2048^2 (one launch each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:"row_pass_kernel|col_wiener_wide_kernel|pack_u8" --launch-skip 12 -c 4 \
    -o $OUT/prof_r1b_passes -f python bench.py --steps 1 --warmup 3 --images 8 --chunk-images 8 --no-e2e --no-cpu-baseline --no-check > $OUT/ncu_full.log 2>&1
ls -la $OUT
