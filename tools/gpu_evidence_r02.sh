#!/bin/bash
# Evidence run of round 2 on the GPU box: bench lines, launch lists, ncu captures (summaries only: the .ncu-rep files stay in /tmp), phase cycles.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py > $O/bench_r02.json 2> $O/bench_r02.err; echo "bench rc $?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_r02.json 2>/dev/null; echo "ref rc $?"
timeout 500 python bench.py --config 5 --steps 3 --warmup 1 --no-extras > $O/bench_config5_r02.json 2> $O/bench_config5_r02.err; echo "config5 rc $?"; tail -2 $O/bench_config5_r02.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-fast > $O/b_under_ncu.log 2>&1; echo "launch list rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file /tmp/tp_launches.csv python tools/ncu_tp.py 32 1024 > /dev/null 2>&1; python tools/ncu_agg.py /tmp/tp_launches.csv > $O/tp_launches_r02.txt; echo "tp launch list rc $?"
timeout 500 ncu --set full --clock-control none -k regex:solve_kernel -c 4 -o /tmp/solve_r02 python tools/ncu_bench_batch.py > $O/ncu_solve.log 2>&1; echo "solve ncu rc $?"
python tools/ncu_extract.py /tmp/solve_r02.ncu-rep > $O/ncu_solve_kernel_r02.json
timeout 200 ncu --set full --clock-control none -k regex:penalty_only -c 2 -o /tmp/penalty_r02 python tools/ncu_bench_batch.py > /dev/null 2>&1; python tools/ncu_extract.py /tmp/penalty_r02.ncu-rep > $O/ncu_penalty_only_r02.json
timeout 200 ncu --set full --clock-control none -k regex:kb_kernel -s 2 -c 1 -o /tmp/kb_r02 python tools/ncu_kb.py 32 8192 > $O/ncu_kb.log 2>&1; python tools/ncu_extract.py /tmp/kb_r02.ncu-rep > $O/ncu_kb_kernel_r02.json
timeout 200 ncu --set full --clock-control none -k regex:kb_kernel -s 2 -c 1 -o /tmp/kb_tma_r02 python tools/ncu_kb.py 32 8192 tma > /dev/null 2>&1; python tools/ncu_extract.py /tmp/kb_tma_r02.ncu-rep > $O/ncu_kb_kernel_tma_r02.json
timeout 300 ncu --set full --clock-control none -k regex:ka_kernel -s 300 -c 1 -o /tmp/ka_r02 python tools/ncu_tp.py 32 4096 > /dev/null 2>&1; python tools/ncu_extract.py /tmp/ka_r02.ncu-rep > $O/ncu_ka_kernel_r02.json
python tools/gpu_profile_dev.py 1024 > $O/phase_cycles_r02.txt 2>&1
UALM_TP_PROFILE=1 python tools/gpu_tp_perf.py 32 1024 8 16 > $O/tp_phase_cycles_r02.txt 2>&1
timeout 400 python tools/config5_sweep.py forest 4096 > $O/config5_sweep_r02.json 2> $O/config5.err; echo "sweep rc $?"; tail -2 $O/config5.err
timeout 200 python tools/config5_sweep.py hill 1024 > $O/config2_sweep_r02.json 2>> $O/config5.err; echo "sweep2 rc $?"
du -sh $O
