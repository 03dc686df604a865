#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): default bench, per-config benches, the CPU reference arm,
# the ncu launch list of the bench command and one full capture of its critic kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
cd /root/repo
python bench.py > gpurun_out/f_bench_2p20.json 2> gpurun_out/f_bench_2p20.err; echo "bench rc=$?"
for w in lqr_d20 lqr_d5 vdp_d10 ekn_d20 lqr_var_d20; do python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/f_bench_cfg_$w.json 2>/dev/null; echo "$w rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_reference.json 2>/dev/null; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_list.log 2>&1; echo "list rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:critic_tc_kernel -s 3 -c 1 -f -o gpurun_out/f_critic_bench_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/f_*
