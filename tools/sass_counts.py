#!/usr/bin/env python3
"""Developer tool: instruction mix of the compute warp's fast interior group (8 steps) of each fill-kernel
instantiation, from the SASS of the shipped library -> profiles/r02_sass_counts.{txt,json}.

The fast interior group is the largest basic block without WARPSYNC.COLLECTIVE (the divergent copies of a group
carry one per shuffle) and with the fewest SELs (the head/tail variants select per cell).  Pipe classes follow
B300_MICROARCH.md "Pipe rates": IMAD* -> FMA pipe, integer/logic/min-max -> ALU pipe (one warp instruction per
2 clk per SM sub-partition = 16 lanes/clk), SHFL/LDS/STS/LDG/STG -> LSU/MIO.
  python tools/sass_counts.py [lib.so]"""
import collections, json, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "smith-waterman_b200" / "libswb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, name = {}, None
for l in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        name = m.group(1); funcs[name] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and name:
        funcs[name].append(m.group(2))
FMA = ("IMAD", "IDP")
ALU = ("VIADDMNMX", "VIMNMX3", "VIMNMX", "LOP3", "SEL", "ISETP", "IADD3", "VIADD", "SHF", "PRMT", "LEA", "IABS", "FLO", "POPC", "MOV", "CS2R", "R2P", "P2R", "PLOP3")
LSU = ("SHFL", "LDS", "STS", "LDG", "STG", "LD", "ST", "CCTL", "ATOMS", "RED", "ATOMG")
# (single pairs: one kernel per form of the cell arithmetic; batches: one kernel holds both forms -- the look-up form is
#  the one reported there: its fast group is the one with IDP.4A)
names = {"swb_tall16fill_kernel_formILi64ELb1ELb1": "swb_tall::fill_kernel_form<64,store,look-up>  (single pair, full fill, 3 rows per lane, half skew)",
         "swb_tall16fill_kernel_formILi64ELb1ELb0": "swb_tall::fill_kernel_form<64,store,compare>  (single pair, > 7 letters)",
         "swb_tall16fill_kernel_formILi64ELb0ELb1": "swb_tall::fill_kernel_form<64,score-only,look-up>  (single pair)",
         "3swb11fill_kernelILi64ELb0EEE": "swb::fill_kernel<64,score-only>, look-up form  (batch, 2 rows per lane)",
         "3swb11fill_kernelILi32ELb1EEE": "swb::fill_kernel<32,store>, look-up form  (batch, 2 rows per lane)"}
ROWS = {"swb_tall": 3, "3swb": 2}
res, lines = {}, []
for fn, ins in funcs.items():
    key = next((k for k in names if k in fn), None)
    if not key:
        continue
    blocks, cur = [], []
    for t in ins:
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        cur.append(op.split(".")[0])
        if op.startswith(("BRA", "EXIT", "BSYNC", "BSSY", "RET")):
            blocks.append(cur); cur = []
    lookup = "Lb0EEEvNS" not in fn or "fill_kernelI" in fn         # (the compare form has no IDP.4A)
    cand = [b for b in blocks if b.count("SHFL") >= 12 and "WARPSYNC" not in b and (("IDP" in b) == lookup)]
    if not cand:
        continue
    b = min(cand, key=lambda b: (b.count("SEL") / max(b.count("SHFL"), 1), -len(b)))
    c = collections.Counter(b)
    # cells in the block: one VIMNMX3 per cell (profile) or three VIADDMNMX per cell (compare); a block holds whole
    # steps of cell arithmetic plus the shuffles / stores of the step that straddles its first branch
    rows = 3 if "swb_tall" in fn else 2
    ncell = c["IDP"] if lookup else c["VIADDMNMX"] / 3.0       # one IDP.4A per cell (look-up) / three VIADDMNMX per cell (compare)
    steps = ncell / (4.0 * rows)
    alu = sum(v for k, v in c.items() if k in ALU); fma = sum(v for k, v in c.items() if k in FMA)
    lsu = sum(v for k, v in c.items() if k in LSU)
    cells = ncell
    rec = {"block_instructions": len(b), "steps": steps, "alu_per_step": alu / steps, "fma_per_step": fma / steps,
           "lsu_per_step": lsu / steps, "total_per_step": len(b) / steps, "alu_ops_per_cell": alu / cells,
           "all_ops_per_cell": len(b) / cells, "mix": dict(c.most_common())}
    res[names[key]] = rec
    lines.append(f"{names[key]}\n  fast interior block: {len(b)} instructions ~ {steps:.2f} steps\n"
                 f"  per step: ALU pipe {alu / steps:.1f}  FMA pipe {fma / steps:.1f}  LSU/MIO {lsu / steps:.1f}  total {len(b) / steps:.1f}\n"
                 f"  per cell: ALU {alu / cells:.2f}  all {len(b) / cells:.2f}\n  mix: {dict(c.most_common(14))}\n")
so = res.get(names["swb_tall16fill_kernel_formILi64ELb0ELb1"])
summary = {"score_only_alu_ops_per_cell": so["alu_ops_per_cell"] if so else None, "kernels": res,
           "source": "cuobjdump -sass of smith-waterman_b200/libswb200.so (tools/sass_counts.py)"}
(ROOT / "profiles" / "r02_sass_counts.json").write_text(json.dumps(summary, indent=1))
(ROOT / "profiles" / "r02_sass_counts.txt").write_text("\n".join(lines))
print("\n".join(lines))
