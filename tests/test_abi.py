"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/recnet_b200.h declares, and
the ctypes signature table agrees with the header's parameter counts.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import recnet_b200
from recnet_b200 import _lib as L

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "recnet_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|float\*|long long)\s+(recnet_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[name] = n
    return out


def test_library_is_built_and_loads():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    assert L.lib().recnet_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    decl = _declared()
    assert len(decl) >= 25
    handle = ctypes.CDLL(L.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
        assert len(L.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} params, ctypes table {len(L.SIGNATURES[name][1])}"
    for name in L.SIGNATURES:
        assert name in decl, f"{name} bound in ctypes but missing from the header"


def test_struct_layouts_match_header_field_order():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for cname, struct in (("recnet_decoder_tensors", L.decoder_tensors), ("recnet_local_tensors", L.local_tensors),
                          ("recnet_global_tensors", L.global_tensors)):
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname, src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = re.findall(r"\*(\w+)", body)
        assert tuple(fields) == struct.FIELDS + getattr(struct, "EXTRA", ())
    assert ctypes.sizeof(L.decoder_desc) == 10 * 4 + 3 * 4 + 4 + 8
    assert ctypes.sizeof(L.decoder_tensors) == (11 + 4 * 3) * 8
    assert ctypes.sizeof(L.local_desc) == 8 * 4 + 4 + 4 + 4
    assert ctypes.sizeof(L.global_desc) == 7 * 4 + 2 * 4 + 4 + 4


def test_workspace_size_queries_run_on_cpu():
    lib = L.lib()
    d = L.decoder_desc(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, L=31, precision=L.PREC_BF16, train=1,
                       embedding_scale=1.0, p_emb_drop=0.5, p_out_drop=0.5)
    n16 = lib.recnet_decoder_workspace_bytes(ctypes.byref(d))
    d.precision = L.PREC_FP32
    n32 = lib.recnet_decoder_workspace_bytes(ctypes.byref(d))
    assert 50e6 < n16 < n32 < 2e9
    ld = L.local_desc(B=100, S=28, R=1536, H=512, A=128, L=31, precision=L.PREC_BF16, train=1, p_drop=0.5)
    assert 50e6 < lib.recnet_local_workspace_bytes(ctypes.byref(ld)) < 2e9
    gd = L.global_desc(B=100, L=31, R=1536, H=512, T=28, precision=L.PREC_BF16, train=1, p_drop=0.5, caption_max_len=30.0)
    assert 50e6 < lib.recnet_global_workspace_bytes(ctypes.byref(gd)) < 2e9
    d.precision = 7
    assert lib.recnet_decoder_workspace_bytes(ctypes.byref(d)) == -4        # RECNET_ERR_UNSUPPORTED


def test_no_cpu_path():
    import torch
    dec = recnet_b200.Decoder("LSTM", 1, 16, 8, 1, 8, 8, 11, 0.5, 0.5, 0.5, precision="fp32")
    tok = torch.ones(3, 2, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dec.forward_sequence(tok, tok, torch.ones(3, 2), torch.randn(2, 4, 16))


def test_round2_entry_points_reject_bad_arguments_before_touching_a_device():
    """Argument checks of the phase-split / beam / background entry points are host code: they must answer on a box without a GPU."""
    lib = L.lib()
    d = L.decoder_desc(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, L=31, precision=L.PREC_BF16, train=1,
                       embedding_scale=1.0, p_emb_drop=0.5, p_out_drop=0.5)
    w = L.decoder_tensors()
    for bad in (0, 4, -1):
        assert lib.recnet_decoder_fwd_phase(ctypes.byref(d), ctypes.byref(w), None, None, None, None, None, None, 0, None, None, bad, None) < 0
    for bad in (0, 16):
        assert lib.recnet_decoder_bwd_phase(ctypes.byref(d), ctypes.byref(w), None, None, None, None, None, None, 0, None, None, None,
                                            ctypes.byref(w), bad, None) < 0
    ld = L.local_desc(B=100, S=28, R=1536, H=512, A=128, L=31, precision=L.PREC_BF16, train=1, p_drop=0.5, cell=L.CELL_LSTM, dec_layers=1)
    lw = L.local_tensors()
    for bad in (0, 4):
        assert lib.recnet_local_bwd_phase(ctypes.byref(ld), ctypes.byref(lw), None, None, None, None, 0, None, ctypes.byref(lw), None, bad, None) < 0
    assert lib.recnet_set_background_ctas(5000) < 0 and lib.recnet_set_background_ctas(0) == 0
    # the split applies to single-layer LSTM decoders on the projected-feature path only
    assert lib.recnet_decoder_bwd_is_split(ctypes.byref(d)) == 1
    d2 = L.decoder_desc(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, L=31, precision=L.PREC_BF16, train=1,
                        embedding_scale=1.0, p_emb_drop=0.5, p_out_drop=0.5, n_layers=2)
    assert lib.recnet_decoder_bwd_is_split(ctypes.byref(d2)) == 0
    dg = L.decoder_desc(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, L=31, precision=L.PREC_BF16, train=1,
                        embedding_scale=1.0, p_emb_drop=0.5, p_out_drop=0.5, cell=L.CELL_GRU)
    assert lib.recnet_decoder_bwd_is_split(ctypes.byref(dg)) == 0
    # beam search: width 1..8, at most 64 steps; workspace grows with the width
    db = L.decoder_desc(B=8 * 4, T=28, E=1536, H=512, A=128, EMB=468, V=4188, L=1, precision=L.PREC_BF16, train=0,
                        embedding_scale=1.0, p_emb_drop=0.0, p_out_drop=0.0)
    assert lib.recnet_beam_workspace_bytes(ctypes.byref(db), 9, 31) < 0 and lib.recnet_beam_workspace_bytes(ctypes.byref(db), 0, 31) < 0
    assert lib.recnet_beam_workspace_bytes(ctypes.byref(db), 4, 65) < 0
    assert lib.recnet_beam_workspace_bytes(ctypes.byref(db), 4, 31) > 0
