#!/usr/bin/env python3
"""Column-strip fill of ONE pair across the GPUs of a box (BASELINE config 3), one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bench_strip.py --cols 100000 --rows 100000 [--steps 5 --warmup 2] [--check]
Prints one JSON line on rank 0: GCUPS of the fill (max over ranks of the device time of a step, barrier before
each step), the per-rank fill-kernel times and the share of the HBM-write roofline per GPU."""
import argparse, importlib, json, os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
swb = importlib.import_module("smith-waterman_b200")
strips = importlib.import_module("smith-waterman_b200.strips")

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=100000); ap.add_argument("--rows", type=int, default=100000)
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--seed", type=int, default=42); ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    peak = 6522.7
a, b = swb.generate(args.seed, args.cols, args.rows)

if world == 1:
    # the 1-GPU figure of the same pair, through the ordinary fill
    a_d = torch.frombuffer(bytearray(a), dtype=torch.uint8).to(dev); b_d = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    cells = (args.rows + 1) * (args.cols + 1)
    dH = torch.empty(cells, dtype=torch.int32, device=dev); dP = torch.empty(cells, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(1, dtype=torch.int64, device=dev); d_sc = torch.zeros(1, dtype=torch.int32, device=dev)
    timer = swb.KernelTimer(local)
    ts, ks = [], []
    for it in range(args.warmup + args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        swb.fill_async(a_d, args.cols, b_d, args.rows, dH, dP, args.cols + 1, d_pos, d_sc, device=local,
                       stream=torch.cuda.current_stream(), timer=timer)
        e1.record(); torch.cuda.synchronize()
        if it >= args.warmup: ts.append(e0.elapsed_time(e1)); ks.append(timer.elapsed_ms())
    ms = sum(ts) / len(ts)
    print(json.dumps({"workload": f"{args.cols}x{args.rows} single pair, 1 GPU (ordinary fill)", "n_gpus": 1, "ms_per_step": ms,
                      "gcups": args.cols * args.rows / ms / 1e6, "kernel_ms": ks, "maxPos": int(d_pos.item()),
                      "hbm_frac_per_gpu": 8.0 * cells / (min(ks) * 1e-3) / 1e9 / peak}))
    sys.exit(0)

pipe = strips.StripPipeline(a, b, local)
timer = swb.KernelTimer(local)
ts, ks = [], []
mp = 0
for it in range(args.warmup + args.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    pipe.fill_async(stream=torch.cuda.current_stream(), timer=timer)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mp = pipe.maxpos()                       # the NCCL all-gather that also separates consecutive calls
    if it >= args.warmup:
        ts.append(float(t.item())); ks.append(timer.elapsed_ms())
kt = torch.tensor([min(ks)], device=dev); allk = [torch.zeros_like(kt) for _ in range(world)]; dist.all_gather(allk, kt)
t0 = time.time(); plen = pipe.backtrack(mp); bt_s = time.time() - t0
res = {"workload": f"{args.cols}x{args.rows} single pair in {world} column strips", "n_gpus": world,
       "ms_per_step": sum(ts) / len(ts), "ms_best": min(ts), "gcups": args.cols * args.rows / (sum(ts) / len(ts)) / 1e6,
       "kernel_ms_per_rank": [round(float(x.item()), 3) for x in allk], "maxPos": mp, "path_len": plen, "backtrack_s": round(bt_s, 4),
       "hbm_frac_per_gpu": 8.0 * (args.rows + 1) * (args.cols / world + 1) / (max(float(x.item()) for x in allk) * 1e-3) / 1e9 / peak}
if args.check:
    # size-independent parity: the recurrence re-evaluated on the GPU from the finished local matrices, and the
    # boundary column against the left neighbour's copy
    H = pipe.strip.dH.view(args.rows + 1, pipe.strip.pitch); P = pipe.strip.dP.view(args.rows + 1, pipe.strip.pitch)
    a_t = pipe.strip.a_d.to(torch.int32); b_t = pipe.strip.b_d.to(torch.int32)
    bad = 0
    for r0 in range(1, args.rows + 1, 2048):
        r1 = min(args.rows + 1, r0 + 2048)
        s = torch.where(a_t[None, :] == b_t[r0 - 1:r1 - 1, None], 3, -3)
        diag = H[r0 - 1:r1 - 1, :-1] + s; up = H[r0 - 1:r1 - 1, 1:] - 2; left = H[r0:r1, :-1] - 2
        want = torch.clamp(torch.maximum(torch.maximum(diag, up), left), min=0)
        bad += int((want != H[r0:r1, 1:]).sum())
        pred = torch.zeros_like(want); mx = torch.zeros_like(want)
        for cand, code in ((diag, 3), (up, 1), (left, 2)):
            upd = cand > mx; pred = torch.where(upd, code, pred); mx = torch.where(upd, cand, mx)
        bad += int((pred != P[r0:r1, 1:].abs()).sum())
    t = torch.tensor([bad], device=dev); dist.all_reduce(t)
    res["recurrence_mismatches"] = int(t.item())
if rank == 0:
    print(json.dumps(res), flush=True)
pipe.close()
dist.destroy_process_group()
