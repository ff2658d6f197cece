#!/bin/bash
# Round-2 evidence run on ONE B200 (from the repo root, under gpurun):  bash profiles/collect_r2.sh
set -x
OUT=gpurun_out/r2
mkdir -p $OUT
python bench.py > $OUT/bench_1gpu_batch256x2048.json 2> $OUT/bench_1gpu.err
python bench.py --impl reference > $OUT/bench_reference_arm.json 2>> $OUT/bench_1gpu.err
python bench.py --workload rgb4096 --flush-l2 --steps 20 --no-cpu-baseline --no-side > $OUT/bench_1gpu_rgb4096.json 2>> $OUT/bench_1gpu.err
python bench.py --workload rgb16384 --steps 10 --warmup 3 --no-cpu-baseline --no-side --no-check > $OUT/bench_1gpu_rgb16384.json 2>> $OUT/bench_1gpu.err
python bench.py --workload car --flush-l2 --steps 30 --no-side > $OUT/bench_1gpu_car.json 2>> $OUT/bench_1gpu.err
python bench.py --workload cat --flush-l2 --steps 30 --no-side > $OUT/bench_1gpu_cat.json 2>> $OUT/bench_1gpu.err
FDR_BENCH_NO_KTIMING=1 python bench.py --workload car --flush-l2 --steps 30 --no-side --no-cpu-baseline --no-check --no-e2e > $OUT/bench_1gpu_car_noevents.json 2>> $OUT/bench_1gpu.err
FDR_BENCH_NO_KTIMING=1 python bench.py --workload cat --flush-l2 --steps 30 --no-side --no-cpu-baseline --no-check --no-e2e > $OUT/bench_1gpu_cat_noevents.json 2>> $OUT/bench_1gpu.err
# launch list of the bench command itself (one timed step of 32 images, cold caches, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launch_list_bench_images32.csv \
    python bench.py --steps 1 --warmup 3 --images 32 --no-e2e --no-cpu-baseline --no-check --no-side > $OUT/ncu_ll.log 2>&1
# full set of the four pass kernels on 12 plane pairs of 2048^2 (one launch each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:"row_pass_kernel|col_wiener_wide_kernel|pack_u8" --launch-skip 10 -c 4 \
    -o $OUT/prof_r2_passes -f python bench.py --steps 1 --warmup 3 --images 8 --chunk-images 8 --no-e2e --no-cpu-baseline --no-check --no-side > $OUT/ncu_full.log 2>&1
python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
ls -la $OUT
