#!/bin/bash
# Round-end evidence on one B200.  Usage (under gpurun): bash profiles/tools/final_evidence.sh <tag> [bench|ncu]
#   bench: GPU test log, bench lines of every workload, ncu launch list of the default bench command
#   ncu:   ncu --set full captures of the dominant kernels, summarised on the box (the .ncu-rep files are too large to bring back)
TAG=${1:-final}; WHAT=${2:-bench}
if [ "$WHAT" = bench ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
  python bench.py > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
  for wl in cfg1 default4k default1080 cfg3 cfg4 cfg5; do
    python bench.py --workload $wl --steps 3 --no-cpu --no-also > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  done
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-also"
  $CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_cfg2_$TAG.csv $CMD > /dev/null 2>&1
  cat gpurun_out/pytest_gpu_$TAG.log
else
  export CRT_SHARDS=1
  cap() {   # workload, kernel regex, name, frames, skip, count
    profiles/tools/ncu_capture.sh "$1" "$2" "$3_$TAG" "$4" "$5" "${6:-2}"
    python profiles/summarize_ncu.py gpurun_out/prof_$3_$TAG.ncu-rep gpurun_out/ncu_$3_$TAG.md
    rm -f gpurun_out/prof_$3_$TAG.ncu-rep
  }
  # (CRT_SHARDS=1: clip mode — per step one launch for frame 0 and one for the other frames of the clip)
  cap cfg2 k_fused_gauss_ps2 cfg2 24 5 2
  cap default4k k_fused_ps2_pipe default4k 12 3 1
  CRT_CLIP=0 cap cfg2 k_fused_gauss_ps2 cfg2_perframe 24 40 2
  cap cfg3 "k_gather_box|k_fused_ps2_pipe" cfg3 12 30
  cap cfg4 "k_gather_box|k_fused_gauss_ps2|k_noise_gen|k_glitch_gen" cfg4 8 40
fi
