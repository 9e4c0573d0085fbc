#!/bin/bash
# Evidence of round 2 on one B200: bench line, launch list, full ncu captures (MLP + HBM kernels), configs 2/4/5.
mkdir -p gpurun_out/r02
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_r02_1gpu.json 2> gpurun_out/r02/bench_r02_1gpu.err
python bench.py --cfg1 --steps 20 2>/dev/null | tail -1 > gpurun_out/r02/cfg1_r02.json
python tools/run_configs.py > gpurun_out/r02/run_configs.log 2>&1; cp gpurun_out/configs_r02.json gpurun_out/r02/
python tools/texture_bench.py > gpurun_out/r02/texture_bench.log 2>&1
python tools/resample_bench.py > gpurun_out/r02/resample_bench.log 2>&1
python tools/composite_bench.py > gpurun_out/r02/composite_bench.log 2>&1
python tools/raygen_bench.py > gpurun_out/r02/raygen_bench.log 2>&1
python tools/clock_probe.py > gpurun_out/r02/clock_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_r02_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
CTXNERF_GRAPH=0 CTXNERF_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"mlp_dgrad|mlp_fwd|mlp_wgrad_k" \
    -s 12 -c 6 -o gpurun_out/r02/ncu_full_mlp_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"composite|resample|raygen|posenc|viewmask|faceview" -c 16 \
    -o gpurun_out/r02/ncu_full_hbm_r02 python tools/hbm_once.py > /dev/null 2>&1
# the .ncu-rep files exceed what gpurun brings back: export the raw pages here and drop them
for f in ncu_full_mlp_r02 ncu_full_hbm_r02; do
  ncu -i gpurun_out/r02/$f.ncu-rep --page raw --csv > gpurun_out/r02/${f}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02/$f.ncu-rep
done
# per-source-line instruction counts of the issue-bound HBM kernels (text summaries of the ncu source page)
bash tools/source_profile.sh resample_fast resample_det 1 > /dev/null 2>&1
bash tools/source_profile.sh raygen_kernel raygen 0 > /dev/null 2>&1
bash tools/source_profile.sh composite_bwd composite_bwd 0 > /dev/null 2>&1
for k in resample_det raygen composite_bwd; do
  python tools/source_report.py gpurun_out/src/${k}_source.csv 40 > gpurun_out/r02/source_lines_$k.txt 2>&1
done
ls -la gpurun_out/r02
