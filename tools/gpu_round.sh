#!/bin/bash
# One gpurun call: GPU tests, every BASELINE config through bench.py, then (optionally) ncu captures.
#   tools/gpu_round.sh [tests] [bench] [configs] [ncu_hbm] [ncu_net] [ncu_full <regex>]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case $what in
    tests)
      timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
      tail -5 gpurun_out/pytest.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log ;;
    bench)
      timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/bench_ref.json ;;
    configs)
      for c in 1 3 4 0; do
        timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/summary.txt
        cat gpurun_out/bench_c$c.json; tail -3 gpurun_out/bench_c$c.err
      done ;;
    trace)
      TIK_PLAN_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-hbm --no-c3 > gpurun_out/trace.json 2> gpurun_out/trace.err; grep "tik trace" gpurun_out/trace.err | tail -24 ;;
    hbm)
      timeout 600 python tools/hbm_bench.py > gpurun_out/hbm_bench.log 2>&1; echo "hbm exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/hbm_bench.log ;;
    hbm_fk_sweep)
      for c in -1 1 2 3; do
        echo "TIK_FK_BULK=$c" | tee -a gpurun_out/hbm_fk_sweep.log
        TIK_FK_BULK=$c timeout 600 python tools/hbm_bench.py 2>&1 | grep fk_body | tee -a gpurun_out/hbm_fk_sweep.log
      done ;;
    ncu_hbm)
      timeout 300 python tools/hbm_bench.py 2097152 --once > gpurun_out/hbm_plain.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rot6d|aa_kornia|rodrigues|rotmat_to_aa|fk_' \
          -c 16 -f -o gpurun_out/prof_hbm python tools/hbm_bench.py 2097152 --once > gpurun_out/ncu_hbm.log 2>&1
      echo "ncu_hbm exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_net)
      timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-hbm --no-c3 > gpurun_out/net_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
          --clock-control none -k regex:'rowgemm|tcn_halo|gcn_fused|gcn_wide|stem_|block_fused|fk_|aggregate' -c 120 --csv --log-file gpurun_out/launches.csv \
          python bench.py --steps 1 --warmup 3 --no-cpu --no-hbm --no-c3 --no-graph > gpurun_out/ncu_net.log 2>&1
      echo "ncu_net exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_c1)
      timeout 300 python bench.py --config 1 --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/c1_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
          --clock-control none -k regex:'rowgemm_tf32|rowgemm_f32|aggregate|stem_' -s 75 -c 25 --csv --log-file gpurun_out/launches_c1.csv \
          python bench.py --config 1 --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_c1.log 2>&1
      echo "ncu_c1 exit $?" | tee -a gpurun_out/summary.txt ;;
    ncu_full)
      timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-hbm --no-c3 --no-graph > gpurun_out/full_plain.log 2>&1 &&
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'rowgemm|tcn_halo|gcn_fused|gcn_wide|stem_|block_fused' -s 51 -c 17 -f -o gpurun_out/prof_net \
          python bench.py --steps 1 --warmup 3 --no-cpu --no-hbm --no-c3 --no-graph > gpurun_out/ncu_full.log 2>&1
      echo "ncu_full exit $?" | tee -a gpurun_out/summary.txt ;;
  esac
done
cat gpurun_out/summary.txt
