#!/bin/bash
# A/B timing of two builds of libmpcb in one GPU session: scripts/ab.sh <libA> <libB> [reps]
D=$PWD/dyobav_mpcnwta_warehouse_b200/csrc
REPS=${3:-3}
for i in $(seq $REPS); do
  for L in $1 $2; do
    echo -n "$L: "; MPCB_LIB_PATH=$D/$L python scripts/profile_step.py 4736 8 100 3 | tail -1 | cut -c1-48
  done
done
