#!/bin/bash
# A/B of kernel variants on one box (development aid): scripts/k1_ab.sh <variant> ...   (libraries under multi_camera_calibration_b200/_variants)
for v in "$@"; do
  if [ $v = default ]; then python scripts/k1_time.py 2>&1 | grep mixed
  else MCCBA_LIB=$PWD/multi_camera_calibration_b200/_variants/libmccba_$v.so python scripts/k1_time.py 2>&1 | grep mixed; fi
done
