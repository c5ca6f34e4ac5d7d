#!/bin/bash
# A/B timing of builds of libmpcb in one GPU session:
#   scripts/ab.sh "<libA> <libB> ..." [reps] [profile_step args]
# default workload: 2368 scenarios x 8 starts at the reference settings (18 944 solves, 10.7 waves of warps)
D=$PWD/dyobav_mpcnwta_warehouse_b200/csrc
REPS=${2:-2}
ARGS=${3:-"2368 8"}
for i in $(seq $REPS); do
  for L in $1; do
    echo -n "$L: "; MPCB_LIB_PATH=$D/$L python scripts/profile_step.py $ARGS | tail -1 | cut -c1-60
  done
done
