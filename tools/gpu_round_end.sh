#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): default bench, per-config benches, the CPU reference arm,
# the ncu launch list of the bench command and one full capture of each rollout kernel.  Outputs under gpurun_out/ with prefix $1.
p=${1:-f}
mkdir -p gpurun_out
cd /root/repo
python bench.py > gpurun_out/${p}_bench_2p20.json 2> gpurun_out/${p}_bench_2p20.err; echo "bench rc=$?"
for w in lqr_d20 lqr_d5 vdp_d10 ekn_d20 lqr_var_d20; do python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/${p}_bench_cfg_$w.json 2>/dev/null; echo "$w rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${p}_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${p}_plain_for_ncu.json 2> gpurun_out/${p}_plain_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${p}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${p}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:critic_tc_kernel -s 3 -c 1 -f -o gpurun_out/${p}_critic_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${p}_ncu_critic.log 2>&1; echo "critic full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_tc_kernel -s 3 -c 1 -f -o gpurun_out/${p}_actor_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${p}_ncu_actor.log 2>&1; echo "actor full rc=$?"
ls -la gpurun_out/${p}_*
