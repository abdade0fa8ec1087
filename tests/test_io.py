"""Host-side input path (no GPU): FASTA / UCSC .2bit readers, batch manifests and the pair-wise shard
arithmetic behind include/swb200.h.  These replace generate() (omp_smithW.c:489-519)
as the source of the sequences."""
import struct

import pytest


def write_2bit(path, records, big_endian=False):
    """records: [(name, seq with N allowed)] -> UCSC .2bit"""
    e = ">" if big_endian else "<"
    enc = {"T": 0, "C": 1, "A": 2, "G": 3, "N": 0}
    bodies = []
    for _, seq in records:
        nb, i = [], 0
        while i < len(seq):
            if seq[i] == "N":
                j = i
                while j < len(seq) and seq[j] == "N":
                    j += 1
                nb.append((i, j - i)); i = j
            else:
                i += 1
        packed = bytearray()
        for i in range(0, len(seq), 4):
            byte = 0
            for k, ch in enumerate(seq[i:i + 4]):
                byte |= enc[ch] << (6 - 2 * k)
            packed.append(byte)
        body = struct.pack(e + "II", len(seq), len(nb))
        body += b"".join(struct.pack(e + "I", s) for s, _ in nb) + b"".join(struct.pack(e + "I", l) for _, l in nb)
        body += struct.pack(e + "II", 0, 0) + bytes(packed)
        bodies.append(body)
    hdr = struct.pack(e + "IIII", 0x1A412743, 0, len(records), 0)
    index_size = sum(1 + len(n) + 4 for n, _ in records)
    off = 16 + index_size
    index = b""
    for (name, _), body in zip(records, bodies):
        index += bytes([len(name)]) + name.encode() + struct.pack(e + "I", off)
        off += len(body)
    path.write_bytes(hdr + index + b"".join(bodies))


def test_fasta_records(swb, tmp_path):
    f = tmp_path / "x.fa"
    f.write_text(">seq1 some description\nACGTacgt\n  NNAC \n;comment\n>seq2\nGG\nGG\n\n")
    assert swb.read_sequences(f) == [("seq1", b"ACGTACGTNNAC"), ("seq2", b"GGGG")]
    assert swb.read_sequence(f, 1) == ("seq2", b"GGGG")
    plain = tmp_path / "plain.txt"
    plain.write_text("acgt\nTTTT\n")
    assert swb.read_sequences(plain) == [("", b"ACGTTTTT")]


@pytest.mark.parametrize("big", [False, True])
def test_2bit_records(swb, tmp_path, big):
    recs = [("chr1", "TCAGNNACGTTGCA"), ("chrM", "ACGTA"), ("empty_n", "NNNNACGT")]
    f = tmp_path / "y.2bit"
    write_2bit(f, recs, big_endian=big)
    assert swb.read_sequences(f) == [(n, s.encode()) for n, s in recs]


def test_manifest_and_errors(swb, tmp_path):
    (tmp_path / "a.fa").write_text(">a0\nACGT\n>a1\nTTTTTT\n")
    write_2bit(tmp_path / "b.2bit", [("b0", "GATTACA")])
    m = tmp_path / "pairs.txt"
    m.write_text("# batch\na.fa:1 b.2bit\n\na.fa b.2bit:0   # trailing comment\n")
    assert swb.load_manifest(m) == [(b"TTTTTT", b"GATTACA"), (b"ACGT", b"GATTACA")]
    with pytest.raises(swb.SwbError):
        swb.read_sequence(tmp_path / "missing.fa")
    with pytest.raises(swb.SwbError):
        swb.read_sequence(tmp_path / "a.fa", 7)
    bad = tmp_path / "bad.txt"
    bad.write_text("a.fa\n")
    with pytest.raises(swb.SwbError):
        swb.load_manifest(bad)


def test_shard_pairs_partition(swb):
    # contiguous blocks that cover every pair exactly once (65536 pairs -> 8192 per GPU)
    assert [swb.shard_pairs(65536, 8, g) for g in range(8)] == [(8192 * g, 8192) for g in range(8)]
    for npairs, ns in [(10, 3), (7, 8), (1, 1), (0, 4), (4097, 2)]:
        blocks = [swb.shard_pairs(npairs, ns, g) for g in range(ns)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == npairs
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(swb.SwbError):
        swb.shard_pairs(10, 0, 0)


def test_cigar_and_alignment_strings(swb, oracle):
    """Host helpers of the alignment emission against a walk of the oracle's P (omp_smithW.c:405-420)."""
    import numpy as np
    a, b = b"TGTTACGG", b"GGTTGACTA"                      # the built-in case: path 69 -> 59 -> 49 -> 40 -> 30 -> 20
    H, P, mp = oracle.fill(a, b, order="wavefront")
    m = len(a)
    moves, pos = bytearray(), mp
    while P.reshape(-1)[pos] != 0:
        code = int(P.reshape(-1)[pos]); moves.append(code)
        pos -= {3: m + 2, 1: m + 1, 2: 1}[code]
    assert len(moves) == 6
    cig = swb.cigar_from_moves(bytes(moves))
    ga, gb = swb.alignment_from_moves(bytes(moves), a, b, mp, m + 1)
    assert cig == "3M1I2M" and (ga, gb) == (b"GTT-AC", b"GTTGAC")
    # random pair: the gapped strings spell the two subsequences and their column scores add up to H[maxPos]
    rng = np.random.default_rng(3)
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), 300).tobytes(); b = rng.choice(np.frombuffer(b"ACGT", np.uint8), 260).tobytes()
    H, P, mp = oracle.fill(a, b, order="wavefront")
    m = len(a); moves, pos = bytearray(), mp
    while P.reshape(-1)[pos] != 0:
        code = int(P.reshape(-1)[pos]); moves.append(code)
        pos -= {3: m + 2, 1: m + 1, 2: 1}[code]
    ga, gb = swb.alignment_from_moves(bytes(moves), a, b, mp, m + 1)
    i1, j1 = divmod(mp, m + 1); i0, j0 = divmod(pos, m + 1)
    assert ga.replace(b"-", b"") == a[j0:j1] and gb.replace(b"-", b"") == b[i0:i1]
    score = sum(-2 if (x == 45 or y == 45) else (3 if x == y else -3) for x, y in zip(ga, gb))
    assert score == H.reshape(-1)[mp]
    cig = swb.cigar_from_moves(bytes(moves))
    import re
    ops = re.findall(r"(\d+)([MID])", cig)
    assert sum(int(n) for n, _ in ops) == len(moves)
    assert sum(int(n) for n, o in ops if o in "MD") == j1 - j0 and sum(int(n) for n, o in ops if o in "MI") == i1 - i0
