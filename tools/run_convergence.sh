#!/bin/bash
# Full-length training runs of several configs AT THE SAME TIME on one B200 (run through gpurun).  A config-size batch
# (2048 paths = 16 tiles) occupies 16 of the 148 SMs, so the runs share the GPU through the CUDA MPS daemon when it can
# be started (otherwise they time-slice).  Usage: tools/run_convergence.sh <tag> <num_iterations|0=config> cfg1 cfg2 ...
# Outputs: gpurun_out/conv_<tag>_<cfg>.log and the reference-format CSVs under gpurun_out/conv_<tag>_csv/.
tag=$1; iters=$2; shift 2
mkdir -p gpurun_out/conv_${tag}_csv
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
if nvidia-cuda-mps-control -d 2>/tmp/mps_start.err; then echo "MPS daemon started"; sleep 1; else echo "MPS not available: $(cat /tmp/mps_start.err)"; unset CUDA_MPS_PIPE_DIRECTORY CUDA_MPS_LOG_DIRECTORY; fi
pids=()
names=()
for spec in "$@"; do                                   # spec = config[@extra-flag[@extra-flag...]]  e.g. configs/vdp_d4.json@--compute_dtype=float64
  cfg=${spec%%@*}; flags=""; [ "$spec" != "$cfg" ] && flags=$(echo "${spec#*@}" | tr '@' ' ')
  name=$(basename $cfg .json)$(echo "$flags" | tr -d ' -' | tr '=' '_')
  names+=($name)
  extra=""; [ "$iters" != "0" ] && extra="--num_iterations=$iters"
  ( cd gpurun_out/conv_${tag}_csv && python ../../main.py --config_path=../../$cfg --exp_name=$name --seed=1 $extra $flags > ../conv_${tag}_${name}.log 2>&1 ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; echo "run $p exit $?"; done
[ -n "$CUDA_MPS_PIPE_DIRECTORY" ] && echo quit | nvidia-cuda-mps-control
for name in "${names[@]}"; do echo "== $name"; grep "step:" gpurun_out/conv_${tag}_${name}.log | tail -1; done
mv gpurun_out/conv_${tag}_csv/logs/* gpurun_out/conv_${tag}_csv/ 2>/dev/null; rmdir gpurun_out/conv_${tag}_csv/logs 2>/dev/null
ls gpurun_out/conv_${tag}_csv | head -30
