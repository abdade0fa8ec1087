"""Column-strip mode on the GPU (swb_fill_strip_async, swb_backtrack_from_async): one pair cut
into column strips must reproduce the oracle's H, P, maxPos and backtrack bit for bit.
 * strips sharing ONE device run left to right in stream order: exercises the boundary injection, the peer
   stores / flags, the hand-off marker and the backtrack chain on the single-GPU box;
 * with >= 2 devices the strips run concurrently (one device each, peer access), and the process-per-GPU
   driver (torch.distributed over NCCL, CUDA IPC boundary buffers) is run with world_size 2."""
import importlib
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def strips(swb):
    return importlib.import_module("smith-waterman_b200.strips")


def make_pair(seed, m, n):
    rng = np.random.default_rng(seed)
    a, b = rng.choice(ACGT, m), rng.choice(ACGT, n)
    k = min(m, n) // 2
    b[n // 5: n // 5 + k] = a[m // 4: m // 4 + k]             # a long path crossing strip boundaries
    return a.tobytes(), b.tobytes()


def check_against_oracle(sset, oracle, a, b):
    mp = sset.fill()
    Ho, Po, mpo = oracle.fill(a, b, order="wavefront")
    H, P = sset.gather()
    assert (H == Ho).all() and (P == Po).all() and mp == mpo
    for s in sset.strips[1:]:                                      # local column 0 = left neighbour's last column
        col0H = s.dH.view(s.n + 1, s.pitch)[:, 0].cpu().numpy()
        col0P = s.dP.view(s.n + 1, s.pitch)[:, 0].cpu().numpy()
        assert (col0H == Ho[:, s.col0]).all() and (col0P[1:] == 5).all() and col0P[0] == 0
    total, starts = sset.backtrack(mp)
    assert total == oracle.backtrack(Po, mpo)
    H2, P2 = sset.gather()
    assert (P2 == Po).all() and (H2 == Ho).all()
    return starts


@pytest.mark.parametrize("m,n,nstrips", [(300, 200, 2), (1027, 700, 3), (2000, 3000, 4), (130, 1500, 2), (4100, 260, 8), (40, 300, 8), (9, 70, 3)])
def test_strips_on_one_device(swb, strips, oracle, m, n, nstrips):
    a, b = make_pair(m + n, m, n)
    sset = strips.StripSet(a, b, nstrips)
    try:
        starts = check_against_oracle(sset, oracle, a, b)
        # a second call on the same buffers (epoch / double buffering)
        check_against_oracle(sset, oracle, a, b)
        if (m, n, nstrips) == (1027, 700, 3):
            assert len(starts) >= 2                                 # the planted path really crosses a boundary
    finally:
        sset.close()


def test_strip_mode_with_nul_bytes_and_other_scores(swb, strips, oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, 700, dtype=np.uint8); b = rng.integers(0, 256, 300, dtype=np.uint8)
    b[40:240] = a[350:550]
    sset = strips.StripSet(a.tobytes(), b.tobytes(), 3, scoring=(5, -3, -4))
    try:
        mp = sset.fill()
        Ho, Po, mpo = oracle.fill(a, b, scoring=(5, -3, -4))
        H, P = sset.gather()
        assert (H == Ho).all() and (P == Po).all() and mp == mpo
    finally:
        sset.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_strips_concurrent_on_two_devices(swb, strips, oracle):
    a, b = make_pair(77, 6000, 5000)
    sset = strips.StripSet(a, b, 2, devices=[0, 1])
    try:
        assert sset.concurrent
        check_against_oracle(sset, oracle, a, b)
        check_against_oracle(sset, oracle, a, b)
    finally:
        sset.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, m, n, out):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        importlib.import_module("smith-waterman_b200")
        st = importlib.import_module("smith-waterman_b200.strips")
        from oracle.swo import Oracle
        o = Oracle()
        a, b = make_pair(5, m, n)
        pipe = st.StripPipeline(a, b, rank)
        ok = True
        for rep in range(2):
            pipe.fill_async(stream=torch.cuda.current_stream())
            mp = pipe.maxpos()
            total = pipe.backtrack(mp)
            Ho, Po, mpo = o.fill(a, b, order="wavefront")
            want = o.backtrack(Po, mpo)
            H, P = pipe.strip.matrices()
            c0, w = pipe.parts[rank]
            ok &= mp == mpo and total == want
            ok &= bool((H.cpu().numpy() == Ho[:, c0 + 1:c0 + w + 1]).all()) and bool((P.cpu().numpy() == Po[:, c0 + 1:c0 + w + 1]).all())
        out[rank] = bool(ok)
        pipe.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_process_per_gpu_pipeline_world2(swb, strips):
    mp = torch.multiprocessing.get_context("spawn")
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        procs = [mp.Process(target=_nccl_worker, args=(r, 2, port, 5000, 4000, out)) for r in range(2)]
        for p in procs: p.start()
        for p in procs: p.join(300)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}
