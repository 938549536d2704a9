"""CPU: the C-ABI library loads and exports every symbol include/llamarec_b200.h declares; the
ctypes table covers the header; size queries (no GPU work) answer sensibly."""
import ctypes
import re

from llamarec_b200 import _lib


def header_functions():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(lrb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported_and_bound():
    names = header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} missing from the ctypes signature table"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound but not declared in the header"


def test_size_queries_without_gpu():
    lib = _lib.load()
    assert lib.lrb_version() == 100
    assert lib.lrb_excl_stride(50) == 52 and lib.lrb_excl_stride(200) == 204
    assert lib.lrb_bias_blk_bytes(1683) == 7 * 8192
    assert lib.lrb_encoder_weight_floats(2) == 128 + 2 * 66816
    assert lib.lrb_encode_workspace_bytes(16, 200, 0) >= 16 * 200 * (64 * 2 + 256) * 4
    assert isinstance(lib.lrb_last_error(), (bytes, type(None)))
    # 9472 users per scoring launch on 148 SMs: the scratch grows with the number of chunk launches
    assert lib.lrb_score_scratch_bytes(32768) > lib.lrb_score_scratch_bytes(4096) > 0
    assert lib.lrb_ce_workspace_bytes(3200, 12087) >= 3200 * 3 * 4


def test_scoring_decomposition_without_gpu():
    """Host-side work decomposition of the scoring kernel (148 SMs assumed when no device is present): 74 CTA
    pairs; 4096 users = 16 pair tiles -> 4 full streams + a remainder stream = 12 partial lists per user;
    9472 users = 37 pair tiles -> 2 full streams, no remainder; batches beyond 74 pair tiles (18944 users) are
    chunked -- 32768 users = 18944 (one full stream, 2 lists) + 13824 (one full stream + a remainder stream, 6 lists)."""
    lib = _lib.load()
    s = ctypes.c_int(0)
    for B, rows, want in ((4096, 10_000_001, 12), (9472, 1_250_001, 4), (32768, 1_250_001, 6), (100, 1683, None)):
        assert lib.lrb_score_topk_slots(B, rows, 0, ctypes.byref(s)) == 0
        assert s.value >= 2 and s.value % 2 == 0
        if want is not None:
            assert s.value == want, (B, rows, s.value)
    assert lib.lrb_score_topk_slots(0, 10, 0, ctypes.byref(s)) != 0          # bad shape is an error, not a crash
    assert b"bad arguments" in lib.lrb_last_error()
