#!/bin/bash
# One parametrised GPU job for `gpurun` (replaces the per-capture run_*.sh scripts of round 1).
#   tools/gpu/run.sh TAG STEP [STEP ...]
# Steps (run in the order given; every artefact goes to gpurun_out/TAG_*):
#   tests            python -m pytest tests -m gpu -q
#   tests:EXPR       ... restricted with -k EXPR
#   bench[:ARGS]     python bench.py ARGS                 -> TAG_bench[_n].json  (ARGS: comma separated, e.g. --workload,cfg5)
#   benchn:N[:ARGS]  the same under torchrun on N GPUs (gpurun --gpus N)
#   ref[:ARGS]       python bench.py --impl reference ARGS
#   launches[:ARGS]  ncu launch list (gpu__time_duration + dram bytes) of a short bench run
#   ncu:REGEX[:SKIP[:COUNT[:ARGS]]]  ncu --set full of the kernels matching REGEX
#   py:SCRIPT[:ARGS] python SCRIPT ARGS (a probe under tools/gpu/)
#   env:K=V          export K=V for the following steps
set -u
TAG=$1; shift
mkdir -p gpurun_out
n=0
for step in "$@"; do
  n=$((n+1))
  kind=${step%%:*}; rest=""; [[ "$step" == *:* ]] && rest=${step#*:}
  case $kind in
    env) export "$rest"; echo "[run] export $rest";;
    tests)
      if [ -n "$rest" ]; then (time timeout 1500 python -m pytest tests -m gpu -q -x -k "$rest") 2>&1 | tail -15
      else (time timeout 1800 python -m pytest tests -m gpu -q -x) 2>&1 | tail -15; fi;;
    bench)
      out=gpurun_out/${TAG}_bench_$n.json
      python bench.py ${rest//,/ } > $out 2> gpurun_out/${TAG}_bench_$n.err; echo "[run] bench $rest rc=$?"; cut -c1-1500 $out; tail -3 gpurun_out/${TAG}_bench_$n.err;;
    benchn)    # benchn:N[:ARGS]  the bench under torchrun on N GPUs (use with gpurun --gpus N)
      ng=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
      out=gpurun_out/${TAG}_bench_n${ng}.json
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $ng ${args//,/ } \
          > $out 2> gpurun_out/${TAG}_bench_n${ng}.err; echo "[run] bench x$ng $args rc=$?"; grep '^{' $out | cut -c1-600; tail -3 gpurun_out/${TAG}_bench_n${ng}.err;;
    ref)
      out=gpurun_out/${TAG}_ref_$n.json
      python bench.py --impl reference ${rest//,/ } > $out 2>/dev/null; cut -c1-400 $out;;
    launches)
      ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
          --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${rest//,/ } \
          > gpurun_out/${TAG}_launches.log 2>&1; echo "[run] launches rc=$?"; wc -l gpurun_out/${TAG}_launches.csv;;
    ncu)
      IFS=: read -r regex skip count args <<< "$rest"
      ncu --set full --clock-control none --import-source on -k regex:$regex -s ${skip:-0} -c ${count:-3} -f -o gpurun_out/${TAG}_ncu_$n \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${args//,/ } > gpurun_out/${TAG}_ncu_$n.log 2>&1
      echo "[run] ncu rc=$?"; ls -la gpurun_out/${TAG}_ncu_$n.ncu-rep;;
    nculist)   # nculist:SCRIPT[:ARGS]  per-launch durations of any probe script
      script=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_nculist_$n.csv \
          python $script ${args//,/ } > gpurun_out/${TAG}_nculist_$n.log 2>&1; echo "[run] nculist rc=$?"; wc -l gpurun_out/${TAG}_nculist_$n.csv;;
    fp64)      # measured FP64 throughput (roofline denominator of k_superpose)
      nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_peak tools/gpu/fp64_peak.cu && /tmp/fp64_peak | tee gpurun_out/${TAG}_fp64.json;;
    py)
      script=${rest%%:*}; args=""; [[ "$rest" == *:* ]] && args=${rest#*:}
      python $script ${args//,/ } > gpurun_out/${TAG}_py_$n.log 2>&1; echo "[run] $script rc=$?"; tail -40 gpurun_out/${TAG}_py_$n.log;;
    *) echo "unknown step $step";;
  esac
done
