#!/bin/bash
set -u
timeout 300 python tools/overlap_fills.py --streams 1 --sets 2
timeout 300 python tools/overlap_fills.py --streams 2 --sets 3
timeout 300 python tools/overlap_fills.py --streams 2 --sets 4
timeout 300 python tools/overlap_fills.py --streams 3 --sets 4
