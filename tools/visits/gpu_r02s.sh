#!/bin/bash
# round 2, visit s: what bounds the full fill -- the same kernel without the global stores / without the writers' work
set -u
for v in nostg nowriter; do echo "== $v"; SWB_LIB=build/libswb200_$v.so timeout 600 python tools/bench_configs.py --configs square,big 2>&1 | cut -c1-150; done
