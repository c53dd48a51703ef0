"""The C-ABI library loads and exports every symbol include/lstur_b200.h declares (no GPU needed: only
host-side entry points are called)."""
import ctypes
import os
import re

import numpy as np

from mnexp_b200 import _lib


def test_every_declared_symbol_is_exported(lib):
    protos = _lib.parse_header()
    assert len(protos) >= 40
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), name
    # and nothing declared twice / missed by the parser
    src = open(_lib.HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    declared = set(re.findall(r'\b(lstur_[a-z0-9_]+)\s*\(', src))
    assert declared == set(protos), declared ^ set(protos)


def test_host_only_entry_points(lib):
    assert lib.lstur_version().decode().startswith('lstur_b200')
    assert lib.lstur_tc_padded_e(300) == 320 and lib.lstur_tc_padded_e(64) == 64
    assert lib.lstur_tc_supported(30, 300, 400, 3) == 1
    assert lib.lstur_tc_supported(50, 300, 400, 3) == 1          # 64-row title slots (BASELINE config C5)
    assert lib.lstur_tc_supported(64, 300, 400, 3) == 0
    assert lib.lstur_tc_slot(30) == 32 and lib.lstur_tc_slot(31) == 32 and lib.lstur_tc_slot(32) == 64 and lib.lstur_tc_slot(50) == 64
    assert lib.lstur_tc_dpre_img_bytes(10, 30, 400) == 10 * 2 * 4 * 32 * 128
    assert lib.lstur_tc_dpre_img_bytes(10, 50, 400) == 10 * 2 * 4 * 64 * 128
    assert lib.lstur_tc_wimg_dgrad_elems(300, 400) == 2 * 7 * 3 * 320 * 32
    assert lib.lstur_word_grad_workspace_bytes(1000, 50, 12) > 4 * 4 * 1000
    assert lib.lstur_tc_supported(30, 300, 400, 5) == 0
    assert lib.lstur_tc_wimg_elems(300, 400) == 3 * 320 * 400
    assert lib.lstur_attn_bwd_grid(10) == 10 and lib.lstur_attn_bwd_grid(10 ** 6) % 148 == 0


def _plan(lib, **kw):
    cfg = dict(B=8, W=5, C=3, L=7, E=12, F=16, KS=3, use_dense=1, Dd=8, dv=0, ds=0, G=8, Ue=8, U=8, arch=0,
               score_model=0, rec_act=0, precision=0, V=120, n_users=50, n_docs=81, dropout=0.0, save_for_backward=1)
    cfg.update(kw)
    c = _lib.lstur_config(**cfg)
    plan = ctypes.c_void_p()
    rc = lib.lstur_plan_create(ctypes.byref(c), ctypes.byref(plan))
    return rc, plan


def test_plan_layout_and_errors(lib):
    rc, plan = _plan(lib)
    assert rc == 0
    n = lib.lstur_plan_dense_count(plan)
    off, cnt = ctypes.c_longlong(), ctypes.c_longlong()
    seen = 0
    for name, want in [('conv_w', 3 * 12 * 16), ('conv_b', 16), ('att_w', 16), ('att_b', 1), ('dense_w', 16 * 8),
                       ('dense_b', 8), ('gru_wx', 8 * 24), ('gru_wh', 8 * 24), ('gru_b', 24)]:
        assert lib.lstur_plan_dense_offset(plan, name.encode(), ctypes.byref(off), ctypes.byref(cnt)) == 0
        assert cnt.value == want and off.value % 4 == 0 and off.value + cnt.value <= n
        seen += want
    assert n >= seen
    assert lib.lstur_plan_dense_offset(plan, b'con_w', ctypes.byref(off), ctypes.byref(cnt)) < 0     # not in LSTUR-ini
    assert lib.lstur_plan_workspace_bytes(plan) > 0
    lib.lstur_plan_destroy(plan)
    # the reference's error behaviour: unknown arch / scorer
    rc, _ = _plan(lib, arch=17)
    assert rc < 0 and b'Unsupport user model' in lib.lstur_last_error()
    rc, _ = _plan(lib, score_model=2)
    assert rc < 0
    rc, _ = _plan(lib, U=9)
    assert rc < 0
    # inference plan is smaller than a training plan
    rc, p1 = _plan(lib, save_for_backward=0)
    rc, p2 = _plan(lib, save_for_backward=1)
    assert lib.lstur_plan_workspace_bytes(p1) < lib.lstur_plan_workspace_bytes(p2)
    lib.lstur_plan_destroy(p1)
    lib.lstur_plan_destroy(p2)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    from mnexp_b200 import synth
    from mnexp_b200.engine import LsturEngine
    sh = synth.SHAPES['tiny']
    try:
        LsturEngine(synth.make_weights(sh), sh.B, sh.W, 1 + sh.K, sh.L)
    except _lib.LsturError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('engine must refuse to run without CUDA')
