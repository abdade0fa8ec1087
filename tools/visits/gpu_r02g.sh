#!/bin/bash
for v in bt2r0 bt2r1 bt2r2 bt2r2w32; do
echo "== $v"
export SWB_LIB=build/libswb200_$v.so
timeout 300 python tools/bt_time.py 2>&1 | tail -4
done
