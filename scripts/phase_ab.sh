#!/bin/bash
# A/B of kernel variants on one box (development aid): scripts/phase_ab.sh <variant> ...   (libraries under multi_camera_calibration_b200/_variants)
for v in "$@"; do
  echo -n "$v: "
  if [ $v = default ]; then python scripts/phase_times.py 2>&1 | tail -1
  else MCCBA_LIB=$PWD/multi_camera_calibration_b200/_variants/libmccba_$v.so python scripts/phase_times.py 2>&1 | tail -1; fi
done
