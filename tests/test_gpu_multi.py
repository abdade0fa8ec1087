"""The C++ multi-GPU host driver (swb_multi_*, SURVEY 8(b) swb_fill_multi / 8(e)), variable-length pair batches
(swb_fill_pairs_async) and the CLI forms built on them, bit-exact against the oracle.
On the single-GPU box the strips share device 0 (they then run left to right in stream order: the boundary
stores, flags, hand-off marker and the strip-hopping backtrack are still exercised); with more devices they
run concurrently over peer access."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_pair(seed, m, n):
    rng = np.random.default_rng(seed)
    a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
    k = min(m, n) // 2
    b[n // 5: n // 5 + k] = a[m // 4: m // 4 + k]             # a long path crossing strip boundaries
    return a.tobytes(), b.tobytes()


def run_multi(swb, oracle, a, b, devices):
    m, n = len(a), len(b)
    Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
    with swb.MultiGpuPair(m, n, devices) as mg:
        for _ in range(2):                                         # second call: epoch / double buffering
            mp, ms = mg.fill(a, b)
            assert mp == mpo and ms == (Ho.reshape(-1)[mpo] if mpo else 0)
            H = np.full((n + 1, m + 1), -7, np.int32); P = np.full((n + 1, m + 1), -7, np.int32)
            mg.gather_host(H, P)
            assert (H == Ho).all() and (P == Po).all()
        plen = mg.backtrack(mp)
        Pb = Po.copy()
        assert plen == oracle.backtrack(Pb, mpo)
        mg.gather_host(None, P)
        assert (P == Pb).all()
        # the strips' own slabs: local column 0 of strip g > 0 is the left neighbour's last column / the marker
        for g in range(1, len(devices)):
            s = mg.strip(g)
            assert s["col0"] + s["m"] <= m and s["pitch"] == s["m"] + 1


@pytest.mark.parametrize("m,n,nstrips", [(300, 200, 2), (1027, 700, 3), (4100, 260, 8), (40, 300, 8), (9, 70, 3)])
def test_multi_driver_one_device(swb, oracle, m, n, nstrips):
    a, b = make_pair(m + n, m, n)
    run_multi(swb, oracle, a, b, [0] * nstrips)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_multi_driver_peer_devices(swb, oracle):
    ndev = min(torch.cuda.device_count(), 4)
    a, b = make_pair(5, 6000, 5000)
    run_multi(swb, oracle, a, b, list(range(ndev)))


def test_variable_length_pairs(swb, oracle):
    """swb_fill_batch with per-pair m[], n[]: a run of equal shapes (one batched launch) between
    pairs of other shapes."""
    shapes = [(37, 301), (64, 64), (64, 64), (64, 64), (64, 64), (500, 130), (1, 50), (50, 1), (256, 256), (256, 256)]
    rng = np.random.default_rng(11)
    A, B, a_off, b_off, hp_off, cells = b"", b"", [], [], [], 0
    for (m, n) in shapes:
        a_off.append(len(A)); b_off.append(len(B)); hp_off.append(cells)
        A += rng.choice(ACGT, m).tobytes(); B += rng.choice(ACGT, n).tobytes()
        cells += ((m + 1) * (n + 1) + 3) // 4 * 4
    dev = torch.device("cuda:0")
    dH = torch.full((cells,), -9, dtype=torch.int32, device=dev); dP = torch.full((cells,), -9, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(len(shapes), dtype=torch.int64, device=dev); d_sc = torch.zeros(len(shapes), dtype=torch.int32, device=dev)
    A_d = torch.frombuffer(bytearray(A), dtype=torch.uint8).to(dev); B_d = torch.frombuffer(bytearray(B), dtype=torch.uint8).to(dev)
    for (sa, sb) in ((A, B), (A_d, B_d)):                          # host and device concatenations
        swb.fill_pairs_async(sa, a_off, [s[0] for s in shapes], sb, b_off, [s[1] for s in shapes], hp_off, dH, dP, d_pos, d_sc,
                             stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        for k, (m, n) in enumerate(shapes):
            Ho, Po, mpo = oracle.fill(A[a_off[k]:a_off[k] + m], B[b_off[k]:b_off[k] + n], order="wavefront")
            H = dH[hp_off[k]:hp_off[k] + (m + 1) * (n + 1)].view(n + 1, m + 1).cpu().numpy()
            P = dP[hp_off[k]:hp_off[k] + (m + 1) * (n + 1)].view(n + 1, m + 1).cpu().numpy()
            assert (H == Ho).all() and (P == Po).all(), k
            assert int(d_pos[k]) == mpo and int(d_sc[k]) == (Ho.reshape(-1)[mpo] if mpo else 0)


def test_large_pairs_overlap_on_two_streams(swb, oracle):
    # pairs with n >= 14208 and m >= 4096 leave the batched launch: single-pair kernel, two internal streams that
    # fork from / join into the caller's stream.  Mixed with small pairs; results against the oracle (maxPos, score,
    # digests of H and P per pair) and against a plain single fill of the same pair.
    from oracle.digest import digest_np, digest_torch
    dev = torch.device("cuda:0")
    shapes = [(4100, 14300), (300, 200), (4096, 14208), (300, 200), (5000, 15000)]
    seqs = [make_pair(40 + k, m, n) for k, (m, n) in enumerate(shapes)]
    A = b"".join(s[0] for s in seqs); B = b"".join(s[1] for s in seqs)
    a_off = np.cumsum([0] + [s[0] for s in shapes[:-1]]).tolist(); b_off = np.cumsum([0] + [s[1] for s in shapes[:-1]]).tolist()
    sizes = [((m + 1) * (n + 1) + 3) // 4 * 4 for (m, n) in shapes]
    hp_off = np.cumsum([0] + sizes[:-1]).tolist()
    dH = torch.full((sum(sizes),), -9, dtype=torch.int32, device=dev); dP = torch.full_like(dH, -9)
    d_pos = torch.zeros(len(shapes), dtype=torch.int64, device=dev); d_sc = torch.zeros(len(shapes), dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    for rnd in range(2):
        swb.fill_pairs_async(A, a_off, [x[0] for x in shapes], B, b_off, [x[1] for x in shapes], hp_off, dH, dP, d_pos, d_sc, stream=s)
        # work enqueued on the caller's stream AFTER the call must see the results (the side streams have joined)
        with torch.cuda.stream(s):
            pos = d_pos.clone(); last = [dH[hp_off[k] + (x[0] + 1) * (x[1] + 1) - 1].clone() for k, x in enumerate(shapes)]
        s.synchronize()
        for k, (m, n) in enumerate(shapes):
            a, b = seqs[k]
            if rnd == 0:
                mso, mpo = oracle.score_only(a, b)
                assert int(pos[k]) == mpo and int(d_sc[k]) == mso, k
            mpo = int(pos[k])
            assert int(last[k]) != -9, k                          # the last cell of the pair was written before the join
            Hk = dH[hp_off[k]:hp_off[k] + (m + 1) * (n + 1)].view(n + 1, m + 1); Pk = dP[hp_off[k]:hp_off[k] + (m + 1) * (n + 1)].view(n + 1, m + 1)
            H1 = torch.empty((n + 1) * (m + 1), dtype=torch.int32, device=dev); P1 = torch.empty_like(H1)
            assert swb.fill(a, m, b, n, H1, P1) == mpo
            assert bool((Hk.reshape(-1) == H1).all()) and bool((Pk.reshape(-1) == P1).all()), k
            if rnd == 1:
                continue
            if n < 1000:
                Ho, Po, _ = oracle.fill(a, b)
                assert (Hk.cpu().numpy() == Ho).all() and (Pk.cpu().numpy() == Po).all()
            else:
                for i0, i1, Hb, Pb in oracle.fill_blocks(a, b, 4096):
                    if i0 is None:
                        break
                    assert digest_torch(Hk[i0:i1], i0, 0) == digest_np(Hb, i0, 0), (k, i0)
                    assert digest_torch(Pk[i0:i1], i0, 0) == digest_np(Pb, i0, 0), (k, i0)
        dH.fill_(-9); dP.fill_(-9)


def cli(args, env_extra=None):
    env = dict(os.environ); env.update(env_extra or {})
    p = subprocess.run([str(ROOT / "smith-waterman_b200" / "swb")] + args, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_cli_files_manifest_devices(swb, oracle, tmp_path):
    a, b = make_pair(3, 700, 500)
    (tmp_path / "a.fa").write_text(">a\n" + a.decode() + "\n>a2\n" + a.decode()[:100] + "\n")
    (tmp_path / "b.fa").write_text(">b\n" + b.decode().lower() + "\n")
    Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
    plen = oracle.backtrack(Po.copy(), mpo)
    out = cli(["--fasta", str(tmp_path / "a.fa"), str(tmp_path / "b.fa")])
    assert "Problem size: Matrix[500][700], FACTOR=128 CUTOFF=1024" in out
    assert f"maxPos: {mpo}  path length: {plen}" in out
    # one pair over 3 column strips through the C++ driver
    out = cli(["--fasta", str(tmp_path / "a.fa"), str(tmp_path / "b.fa")], {"SWB_DEVICES": "0,0,0"})
    assert f"maxPos: {mpo}  path length: {plen}  (3 column strips)" in out
    assert "Elapsed time for scoring matrix computation:" in out and "Elapsed time for backtracking:" in out
    # built-in case over 2 strips keeps the self-check (omp_smithW.c:230-234)
    out = cli([], {"SWB_DEVICES": "0,0"})
    assert "Verifying results using the builtinIn data: true" in out
    # manifest: two pairs of different shapes
    (tmp_path / "m.txt").write_text("a.fa b.fa\na.fa:1 b.fa\n")
    out = cli(["--pairs", str(tmp_path / "m.txt")])
    H2, P2, mp2 = oracle.fill(a[:100], b, order="wavefront")
    assert f"pair 0: 700 x 500  maxScore {Ho.reshape(-1)[mpo]}  maxPos {mpo}  path length {plen}" in out
    assert f"pair 1: 100 x 500  maxScore {H2.reshape(-1)[mp2]}  maxPos {mp2}  path length {oracle.backtrack(P2.copy(), mp2)}" in out
    # v1's extra lines and the SKIP_BACKTRACK configuration (omp_smithW-v1-refinedOrig.cpp:138-142,190-192)
    out = cli(["700", "500"], {"SWB_SEED": "42", "SWB_V1_LINES": "1", "SWB_SKIP_BACKTRACK": "1"})
    lines = out.splitlines()
    assert lines[1] == "Total memory footprint is:2 MB" and "Skipping backtrack ..." in lines
    a42, b42 = swb.generate(42, 700, 500)
    ms, mp = oracle.score_only(a42, b42)
    assert f"maxScore: {ms}  maxPos: {mp}" in out


def test_traceback_moves_cigar(swb, oracle):
    """the backtrack that also emits the path's moves (swb_traceback_async, the jump-table kernel)
    against a host walk of the oracle's P (omp_smithW.c:405-420); CIGAR / gapped strings from the moves."""
    for (m, n, seed) in [(8, 9, None), (700, 500, 3), (3000, 2600, 5), (257, 4000, 6)]:
        if seed is None:
            a, b = b"TGTTACGG", b"GGTTGACTA"
        else:
            a, b = make_pair(seed, m, n)
        Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
        want, pos = bytearray(), mpo
        while Po.reshape(-1)[pos] != 0:
            code = int(Po.reshape(-1)[pos]); want.append(code)
            pos -= {3: m + 2, 1: m + 1, 2: 1}[code]
        dev = torch.device("cuda:0")
        dH = torch.empty((n + 1) * (m + 1), dtype=torch.int32, device=dev); dP = torch.empty_like(dH)
        d_pos = torch.zeros(1, dtype=torch.int64, device=dev)
        swb.fill_async(a, m, b, n, dH, dP, m + 1, d_pos, None, stream=torch.cuda.current_stream())
        d_len = torch.zeros(1, dtype=torch.int64, device=dev); d_end = torch.zeros(1, dtype=torch.int64, device=dev)
        d_moves = torch.zeros(m + n + 2, dtype=torch.uint8, device=dev)
        swb.traceback_async(dP, m + 1, d_startPos=d_pos, d_pathLen=d_len, d_endPos=d_end, d_moves=d_moves,
                            stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        plen = int(d_len.item())
        assert int(d_pos.item()) == mpo and plen == len(want) and int(d_end.item()) == pos
        moves = bytes(d_moves[:plen].cpu().numpy())
        assert moves == bytes(want)
        Pb = Po.copy(); oracle.backtrack(Pb, mpo)
        assert (dP.view(n + 1, m + 1).cpu().numpy() == Pb).all()
        ga, gb = swb.alignment_from_moves(moves, a, b, mpo, m + 1)
        score = sum(-2 if (x == 45 or y == 45) else (3 if x == y else -3) for x, y in zip(ga, gb))
        assert score == Ho.reshape(-1)[mpo]
        if seed is None:
            assert swb.cigar_from_moves(moves) == "3M1I2M"
