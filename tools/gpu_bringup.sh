#!/bin/bash
# Kernel bring-up on a GPU box: every stage in its own process under a timeout, logs into gpurun_out/.
# usage: tools/gpu_bringup.sh [stage-list]      (default: all stages)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
run() {
  echo "=== $*" | tee -a $LOG
  timeout 180 python tests/gpu_stage_check.py "$@" >> $LOG 2>&1
  echo "exit $?" | tee -a $LOG
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee -a $LOG
STAGES=${1:-"small fwd fg bwd"}
for s in $STAGES; do
  case $s in
    small) run small 3 40 6 300 128 ;;
    fwd)
      run fwd 2 40 6 300 64
      run fwd 2 40 6 300 128
      run fwd 2 40 6 1000 512 ;;
    fg)
      run fg 2 40 6 300 128
      run fg 2 40 6 300 256
      run fg 3 50 9 1000 512 ;;
    bwd)
      TTX_BWD=da run bwd 2 40 6 300 64
      TTX_BWD=dw run bwd 2 40 6 300 64
      TTX_BWD=both run bwd 2 40 6 300 128
      TTX_BWD=both run bwd 2 40 6 300 256
      TTX_BWD=both run bwd 2 40 6 300 384
      TTX_BWD=both TTX_SPLITS=3 run bwd 3 50 9 1000 512
      TTX_BWD=both TTX_SPLITS=3 run bwd 3 50 9 1000 384 ;;
  esac
done
grep -E "^===|STAGE|FAIL|Error|error|timed out" $LOG | tail -60
