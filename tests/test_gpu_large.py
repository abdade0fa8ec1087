"""Oracle-backed parity at BASELINE.json's full sizes (configs 2-5), through committed digests.

tests/golden/large_digests.json holds what the CPU oracle (oracle/sw_oracle.c, pinned against the unmodified
reference in tests/test_oracle.py) computed for the same generate()-seeded sequences: position-weighted digests
(oracle/digest.py) of H and P per 1024-row block and column chunk for a fixed sample of blocks, maxScore, maxPos
with the reference's tie-break, the backtrack's path length and a digest of the path cells
(oracle/make_golden_large.py wrote it; ~6 minutes of CPU).  The GPU results are hashed on the GPU with the
same formula.  Bit-exact: int32/int64 arithmetic only (omp_smithW.c:331-420).
"""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.digest import digest_torch, path_digest     # noqa: E402


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.loads((golden_dir / "large_digests.json").read_text())


def check_blocks(g, Hv, Pv, col0=0, m_local=None):
    """Hv, Pv: (rows+1, pitch) views whose local column j is global column col0 + j."""
    cols, chunk, B = g["cols"], g["chunk_cols"], g["block_rows"]
    m_local = m_local if m_local is not None else cols
    nchecked = 0
    for key, d in g["blocks"].items():
        i0 = int(key); i1 = min(i0 + B, g["rows"] + 1)
        for c, c0 in enumerate(range(1, cols + 1, chunk)):
            c1 = min(c0 + chunk, cols + 1)
            if c0 < col0 + 1 or c1 > col0 + m_local + 1:
                continue                                            # chunk not (wholly) in this strip
            hd = digest_torch(Hv[i0:i1, c0 - col0:c1 - col0], i0, c0)
            pd = digest_torch(Pv[i0:i1, c0 - col0:c1 - col0], i0, c0)
            assert list(hd) == d["H"][c], f"H rows {i0}..{i1} cols {c0}..{c1}"
            assert list(pd) == d["P"][c], f"P rows {i0}..{i1} cols {c0}..{c1}"
            nchecked += 1
    return nchecked


@pytest.mark.parametrize("name", ["45000x45000", "1000x2000000", "2000000x1000", "100000x100000"])
def test_single_pair_against_oracle_digests(swb, golden, name):
    g = golden["single"][name]
    cols, rows = g["cols"], g["rows"]
    dev = torch.device("cuda:0")
    need = 8 * (rows + 1) * (cols + 1) + (12 << 30)
    if torch.cuda.mem_get_info(0)[0] < need:
        pytest.skip(f"{name} needs {need >> 30} GB of device memory")
    a, b = swb.generate(g["seed"], cols, rows)
    cells = (rows + 1) * (cols + 1)
    dH = torch.empty(cells, dtype=torch.int32, device=dev); dP = torch.empty(cells, dtype=torch.int32, device=dev)
    d_sc = torch.zeros(1, dtype=torch.int32, device=dev); d_pos = torch.zeros(1, dtype=torch.int64, device=dev)
    swb.fill_async(a, cols, b, rows, dH, dP, cols + 1, d_pos, d_sc, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert (int(d_sc.item()), int(d_pos.item())) == (g["maxScore"], g["maxPos"])
    Hv, Pv = dH.view(rows + 1, cols + 1), dP.view(rows + 1, cols + 1)
    assert check_blocks(g, Hv, Pv) >= len(g["blocks"])
    assert int(Hv[:, 0].abs().max()) == 0 and int(Hv[0].abs().max()) == 0 and int(Pv[:, 0].abs().max()) == 0
    # backtrack (omp_smithW.c:405-420): length, the path cells, and nothing else touched in the sampled blocks
    plen = swb.backtrack(dP, cols + 1, g["maxPos"])
    assert plen == g["path_len"]
    neg = []
    for r0 in range(0, rows + 1, 8192):                              # (a 10^10-element mask at once would be 10 GB)
        idx = torch.nonzero(Pv[r0:r0 + 8192] < 0)
        if idx.numel():
            neg.append((idx[:, 0] + r0) * (cols + 1) + idx[:, 1])
    neg = torch.cat(neg).cpu().numpy()
    assert neg.size == plen and path_digest(neg) == g["path_digest"]
    # score-only kernel on the same pair (no H/P stores): same maximum, same maxPos
    del dH, dP, Hv, Pv
    swb.score_only_async(a, cols, b, rows, 1, d_pos, d_sc, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert (int(d_sc.item()), int(d_pos.item())) == (g["maxScore"], g["maxPos"])


def test_batch_65536_pairs_sample(swb, golden):
    """BASELINE config 5: 65536 independent 256x256 pairs (pair k seeded 1000+k) in one launch; every 16th pair
    (4096 pairs) against the oracle."""
    g = golden["batch"]
    m, n, npairs = g["cols"], g["rows"], g["pairs"]
    dev = torch.device("cuda:0")
    A = bytearray(); B = bytearray()
    for k in range(npairs):
        a, b = swb.generate(g["seed0"] + k, m, n)
        A += a; B += b
    A_d = torch.frombuffer(A, dtype=torch.uint8).to(dev); B_d = torch.frombuffer(B, dtype=torch.uint8).to(dev)
    pitch = m + 1
    stride = ((n + 1) * pitch + 3) // 4 * 4
    dH = torch.empty(npairs * stride, dtype=torch.int32, device=dev); dP = torch.empty(npairs * stride, dtype=torch.int32, device=dev)
    d_pos = torch.zeros(npairs, dtype=torch.int64, device=dev); d_sc = torch.zeros(npairs, dtype=torch.int32, device=dev)
    swb.fill_batch_async(A_d, m, B_d, n, npairs, dH, dP, pitch, stride, d_pos, d_sc, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    pos, sc = d_pos.cpu().numpy(), d_sc.cpu().numpy()
    Hv = dH.view(npairs, stride)[:, :(n + 1) * pitch].view(npairs, n + 1, pitch)
    Pv = dP.view(npairs, stride)[:, :(n + 1) * pitch].view(npairs, n + 1, pitch)
    for key, want in g["sample"].items():
        k = int(key)
        got = [int(sc[k]), int(pos[k])]
        assert got == want[:2], k
        assert list(digest_torch(Hv[k], 0, 0)) == want[3:5] and list(digest_torch(Pv[k], 0, 0)) == want[5:7], k
    # backtrack of a subset (one launch per pair)
    for key in list(g["sample"])[::64]:
        k = int(key)
        assert swb.backtrack(dP[k * stride:(k + 1) * stride], pitch, int(pos[k])) == g["sample"][key][2]
    # the same batch, score only
    swb.score_only_async(A_d, m, B_d, n, npairs, d_pos, d_sc, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert (d_pos.cpu().numpy() == pos).all() and (d_sc.cpu().numpy() == sc).all()
