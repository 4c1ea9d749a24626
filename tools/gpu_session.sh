#!/bin/bash
# One GPU-box session (run under gpurun from the repository root):  tools/gpu_session.sh <what> [...]
#   tests [pytest args]  GPU parity tests (default: the whole -m gpu suite)
#   bench [bench args]   one bench line (default flags) -> gpurun_out/bench_<tag>.json
#   sizes                bench lines at the shard sizes of 1 / 4 / 8 GPUs (1 M, 250 k, 125 k patterns)
#   launches [args]      ncu launch list of the bench command -> gpurun_out/launches_<tag>.csv
#   prof <regex> <skip> <tag> [bench args]   ncu --set full of one kernel -> gpurun_out/prof_<tag>.ncu-rep
# TAG (environment) names the outputs; several commands can be chained with "--".
mkdir -p gpurun_out
TAG=${TAG:-r2}
run_one() {
  what=$1; shift
  case "$what" in
    tests)
      if [ $# -eq 0 ]; then set -- tests -m gpu -q; fi
      timeout 1500 python -m pytest "$@" > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests_$TAG.log | cut -c1-400 ;;
    bench)
      timeout 900 python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.json ;;
    sizes)
      for n in 1000000 250000 125000; do
        timeout 300 python bench.py --patterns $n --steps 100 --warmup 5 --no-cpu-baseline --no-configs "$@" 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_$n.json
        python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${TAG}_$n.json'))
print($n, 'evals/s %.1f ms %.4f cl_ms %.4f e2e %.1f' % (d['value'], d['ms_per_step'], d['roofline'].get('cl_ms_per_eval', 0), d['e2e']['value']), d.get('clocks'))"
      done ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs "$@" > /dev/null 2>&1; echo "launches rc=$?" ;;
    prof)
      regex=$1; skip=$2; tag=$3; shift 3
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -o gpurun_out/prof_$tag "$@" > gpurun_out/prof_$tag.log 2>&1; echo "prof $tag rc=$?" ;;
    *) echo "unknown command $what"; return 1 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" == "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
