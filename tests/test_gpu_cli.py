"""The CLI clone (smith-waterman_b200/swb, csrc/swb_cli.cpp) against the reference program's surface
(omp_smithW.c:87-253): same argv, same stdout lines in the same order, results equal to the oracle's."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
SWB = ROOT / "smith-waterman_b200" / "swb"


def run(args, **env):
    out = subprocess.run([str(SWB), *args], capture_output=True, text=True, timeout=120, env={**os.environ, **env})
    return out.returncode, out.stdout


def test_builtin_case_prints_the_reference_lines_in_order():
    rc, out = run([])
    assert rc == 0
    wanted = ["Using built-in data for testing ..",                       # omp_smithW.c:100
              "Problem size: Matrix[9][8], FACTOR=128 CUTOFF=1024",      # :101
              "Using 1 out of max 1 threads...",                          # :194
              "Elapsed time for scoring matrix computation:",             # :220
              "Elapsed time for backtracking:",                           # :228
              "Verifying results using the builtinIn data: true"]         # :232
    pos = -1
    for w in wanted:
        nxt = out.find(w, pos + 1)
        assert nxt > pos, (w, out)
        pos = nxt


def test_sized_run_matches_the_oracle_and_the_run_scripts_can_parse_it(oracle, swb):
    rc, out = run(["300", "200"], SWB_SEED="42")
    assert rc == 0 and "Problem size: Matrix[200][300], FACTOR=128 CUTOFF=1024" in out
    # readme.liao:12 -- grep "Elapsed time for scoring matrix computation" | cut -d: -f2
    line = [ln for ln in out.splitlines() if "Elapsed time for scoring matrix computation" in ln][0]
    assert float(line.split(":")[1]) >= 0.0
    a, b = swb.generate(42, 300, 200)
    H, P, mp = oracle.fill(a, b, order="wavefront")
    plen = oracle.backtrack(P, mp)
    m = re.search(r"maxPos: (\d+)\s+path length: (\d+)", out)
    assert (int(m.group(1)), int(m.group(2))) == (mp, plen)


def test_debug_dump_has_the_matrices():
    rc, out = run([], SWB_DEBUG="1")
    assert rc == 0 and "Similarity Matrix:" in out and "Predecessor Matrix:" in out
    # last row of H of the built-in case ends with 7 (omp_smithW.c:232: H[n*m-1] == 7)
    sim = out.split("Similarity Matrix:")[1].split("Predecessor Matrix:")[0].strip().splitlines()
    assert sim[-1].split()[-1] == "7"
