#!/bin/bash
# Final evidence run of round 2 (after the table-based MINCO of the throughput path, the MPC hand-over and the front-end): GPU tests, ncu extracts
# of the throughput kernels, bench lines of configs 2-5 + the reference arm, launch list, phase cycles, population sweeps.  .ncu-rep files stay in /tmp.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_r02.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu_r02.log
timeout 200 ncu --set full --clock-control none -k regex:kb_kernel -s 2 -c 1 -o /tmp/kb1024 python tools/ncu_kb.py 32 1024 > /dev/null 2>&1; python tools/ncu_extract.py /tmp/kb1024.ncu-rep > $O/ncu_kb_kernel_b1024_r02.json
timeout 200 ncu --set full --clock-control none -k regex:kb_kernel -s 2 -c 1 -o /tmp/kb8192 python tools/ncu_kb.py 32 8192 > $O/ncu_kb.log 2>&1; python tools/ncu_extract.py /tmp/kb8192.ncu-rep > $O/ncu_kb_kernel_r02.json
timeout 300 ncu --set full --clock-control none -k regex:ka_kernel -s 300 -c 1 -o /tmp/ka python tools/ncu_tp.py 32 4096 > /dev/null 2>&1; python tools/ncu_extract.py /tmp/ka.ncu-rep > $O/ncu_ka_kernel_r02.json
python tools/make_traffic.py
timeout 400 python bench.py > $O/bench_r02.json 2> $O/bench_r02.err; echo "bench rc $?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_r02.json 2>/dev/null; echo "ref rc $?"
timeout 400 python bench.py --config 3 --steps 4 --warmup 3 --no-extras > $O/bench_config3_r02.json 2> $O/bench_config3_r02.err; echo "config3 rc $?"
timeout 500 python bench.py --config 4 --steps 3 --warmup 3 --no-extras > $O/bench_config4_r02.json 2> $O/bench_config4_r02.err; echo "config4 rc $?"
timeout 500 python bench.py --config 5 --steps 3 --warmup 3 --no-extras > $O/bench_config5_r02.json 2> $O/bench_config5_r02.err; echo "config5 rc $?"
timeout 300 python bench.py --front-end astar --steps 4 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_astar_r02.json 2> $O/bench_astar_r02.err; echo "astar rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file /tmp/tp_launches.csv python tools/ncu_tp.py 32 1024 > /dev/null 2>&1; python tools/ncu_agg.py /tmp/tp_launches.csv > $O/tp_launches_r02.txt; echo "tp launch list rc $?"
UALM_TP_PROFILE=1 python tools/gpu_tp_perf.py 32 1024 8 16 > $O/tp_phase_cycles_r02.txt 2>&1
timeout 200 python tools/config5_sweep.py hill 1024 > $O/config2_sweep_r02.json 2> $O/config5.err; echo "sweep2 rc $?"
timeout 400 python tools/config5_sweep.py forest 4096 > $O/config5_sweep_r02.json 2>> $O/config5.err; echo "sweep rc $?"; tail -2 $O/config5.err
du -sh $O
