"""CPU-side checks of the product boundary: the C-ABI library builds/loads and exports
every symbol include/swb200.h declares; the schedule model of the kernel reproduces the
oracle; host-side helpers agree with the reference's generate().  No compute calls."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))


def declared_symbols():
    text = (ROOT / "include" / "swb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(swb_[A-Za-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(swb):
    names = declared_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(str(swb.LIB_PATH))
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/swb200.h but not exported"
    assert swb.lib.swb_version() >= 100
    assert swb.lib.swb_strerror(-2).decode().startswith("H/P")


def test_library_is_sm100a_only(swb):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(swb.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_generate_matches_reference_generator(swb, oracle, golden_dir):
    z = np.load(golden_dir / "ref_small.npz")
    for t in sorted({k.split("_")[0] for k in z.files}):
        cols, rows, seed = (int(x) for x in z[f"{t}_meta"][:3])
        if cols <= 0:
            continue
        a, b = swb.generate(seed, cols, rows)
        assert a == bytes(z[f"{t}_a"]) and b == bytes(z[f"{t}_b"])


def test_no_gpu_means_loud_failure(swb):
    if swb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(swb.SwbError):
        swb.align_host(b"ACGT", b"ACG", None, None)


def test_product_never_touches_the_oracle():
    for p in (ROOT / "smith-waterman_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"} or p.name == "Makefile":
            txt = p.read_text()
            assert "oracle" not in txt.replace("the oracle", "").lower() or "sw_oracle" not in txt, p
            assert "sw_oracle" not in txt and "libsworacle" not in txt and "swo_" not in txt, p


@pytest.mark.parametrize("m,n,wpc,pitch,policy", [
    (8, 9, 2, None, "lazy"), (130, 33, 1, None, "lazy"), (131, 64, 2, None, "eager"), (132, 65, 2, 140, "lazy"),
    (133, 100, 4, None, "random"), (303, 70, 3, None, "lazy"), (1, 40, 2, None, "lazy"), (37, 1, 1, None, "eager"),
    (700, 97, 2, None, "lazy"), (700, 97, 2, None, "eager"), (1100, 130, 2, 1128, "random"), (600, 200, 1, None, "lazy"),
])
def test_schedule_model_reproduces_oracle(oracle, m, n, wpc, pitch, policy):
    """Index arithmetic and flow-control rules of fill_kernel (staging ring + writer
    segmentation, tagged hand-off rings, band boundary + loader) on the CPU."""
    from schedule_model import fill_model
    rng = np.random.default_rng(m * 1000 + n)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    a, b = rng.choice(acgt, m), rng.choice(acgt, n)
    H, P, _ = oracle.fill(a, b)
    Hm, Pm, smax = fill_model(a, b, wpc=wpc, pitch=pitch, policy=policy, seed=m + n)
    assert (Hm == H).all() and (Pm == P).all()
    from schedule_model import STRIP_ROWS
    want = [H[1 + STRIP_ROWS * s: 1 + STRIP_ROWS * (s + 1)].max() for s in range((n + STRIP_ROWS - 1) // STRIP_ROWS)]
    assert list(smax) == want


def test_argument_checks_come_before_any_cuda_call(swb):
    """Bad arguments are rejected with a status code, never a crash or an exit ('all return int'); these
    checks run before the first CUDA call, so they work without a device."""
    import ctypes as C
    buf = (C.c_char * 64)()
    out = (C.c_int32 * 64)()          # stands in for device pointers: never dereferenced on these paths
    a = b = C.addressof(buf)
    H = P = (C.addressof(out) + 15) // 16 * 16

    def fill(m, n, scoring, dH=H, dP=P, pitch=None):
        sc = swb.Scoring(*scoring)
        return swb.lib.swb_fill_async(a, m, b, n, C.byref(sc), dH, dP, pitch or m + 1, None, None, 0, None, None)

    ARG, RANGE, ALIGN = -1, swb.lib.swb_fill_async(a, 1 << 30, b, 8, None, H, P, (1 << 30) + 1, None, None, 0, None, None), None
    assert fill(0, 8, (3, -3, -2)) == ARG and fill(8, -1, (3, -3, -2)) == ARG
    assert fill(8, 8, (3, -3, -2), dH=None) == ARG and fill(8, 8, (3, -3, -2), pitch=8) == ARG
    assert RANGE not in (0, ARG)
    assert fill(8, 8, (3, -3, 0)) == RANGE            # a gap must cost something
    assert fill(8, 8, (3, 1, -2)) == RANGE            # mismatch <= 0
    assert fill(8, 8, (1 << 21, -3, -2)) == RANGE
    assert fill(8, 8, (3, -3, -2), dH=H + 4) not in (0, ARG, RANGE)      # 16-byte alignment of H / P
    assert swb.lib.swb_strip_flag_count(0) == 0 and swb.lib.swb_strip_flag_count(33) == 2 and swb.lib.swb_strip_flag_count(64) == 2
    for code in (0, ARG, RANGE):
        assert len(swb.lib.swb_strerror(code)) > 0
