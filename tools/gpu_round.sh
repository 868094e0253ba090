# One round of GPU evidence (run through gpurun).  Usage: tools/gpu_round.sh TAG [phase ...]
# phases (default: all): tests bench launches full timing configs filter dropin
# Outputs in gpurun_out/ (TAG = suffix).  Summaries for profiles/: tools/ncu_summary.py, ncu_src.py, ncu_kernels.py, ncu_traffic.py.
TAG=${1:-round}; shift
PH=" ${*:-tests bench launches full timing configs filter dropin} "
has() { case "$PH" in *" $1 "*) return 0;; *) return 1;; esac; }
T0=$(date +%s); mark() { echo "== $1 at +$(( $(date +%s) - T0 )) s"; }
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
if has tests; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5; mark tests
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; mark smoke
fi
if has bench; then
  timeout 900 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; tail -3 gpurun_out/bench_${TAG}_default.err; cat gpurun_out/bench_${TAG}_default.json; mark bench
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; cat gpurun_out/bench_${TAG}_reference.json; mark reference
fi
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum"
if has launches; then
  timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --pairs 262144 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1; mark launches
fi
if has full; then
  for k in ${FULL_KERNELS:-k_seed k_deferred k_passes k_align k_pair k_em_colsum k_em_rowsum}; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_${k}_${TAG} -f python bench.py --pairs 50000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_${k}_${TAG}.log 2>&1; mark full_$k
  done
fi
if has timing; then
  T1K_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_timing.json 2> gpurun_out/bench_${TAG}_timing.err; grep "t1k timing" gpurun_out/bench_${TAG}_timing.err | grep -v "assign: \(launch\|input\|store\|tail\)" | tail -24; mark timing
fi
if has configs; then
  for c in 3 4; do timeout 900 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err; cat gpurun_out/bench_${TAG}_config$c.json; mark config$c; done
fi
if has filter; then
  timeout 600 python bench.py --stage filter --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_filter.json 2> gpurun_out/bench_${TAG}_filter.err; cat gpurun_out/bench_${TAG}_filter.json; mark filter
fi
if has dropin; then
  timeout 900 python bench.py --stage dropin --steps 1 --warmup 1 > gpurun_out/bench_${TAG}_dropin.json 2> gpurun_out/bench_${TAG}_dropin.err; cat gpurun_out/bench_${TAG}_dropin.json; mark dropin
fi
