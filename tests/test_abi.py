"""CPU: the C-ABI library loads, exports every symbol include/hrnb.h declares, and validates arguments
(no kernels are launched - there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from hrnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hrnb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hrnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)
    header = open(os.path.join(ROOT, "include", "hrnb.h")).read()
    assert int(re.search(r"#define HRNB_ABI_VERSION (\d+)", header).group(1)) == _lib.ABI_VERSION == lib.hrnb_abi_version()


def _params(**kw):
    p = _lib.ConvParams()
    base = dict(inp=8, in_ps=1024, wpk=8, bias=8, res=None, res_ps=0, out=8, out_ps=1024, N=2, H=16, W=16,
                in_phase_stride=0, out_phase_stride=0,
                in_H=16, in_W=16, cin=64, cout=64, taps=9, stride=1, KC=8, BN=64, MB=1, flags=0)
    base.update(kw)
    for k, v in base.items():
        setattr(p, k, v)
    return p


def test_conv_geometry_validation_and_smem_budget():
    lib = _lib.lib()
    ok = lib.hrnb_conv_smem_bytes(C.byref(_params()))
    assert 0 < ok <= 227 * 1024
    for bad in (dict(taps=4), dict(stride=2), dict(cin=24), dict(KC=3), dict(BN=24), dict(MB=3),
                dict(MB=4, BN=128), dict(in_H=32), dict(cout=48)):
        assert lib.hrnb_conv_smem_bytes(C.byref(_params(**bad))) < 0, bad
        assert lib.hrnb_last_error()
    # every conv shape of HRNet-W32/W48 at batch 64 must fit the shared-memory budget
    from hrnet_b200 import arch as A
    from hrnet_b200.config import make_cfg
    from hrnet_b200.ops import pick_tile
    for width in (32, 48):
        a = A.arch_from_cfg(make_cfg(width))
        for sp in A.layer_specs(a):
            if not isinstance(sp, A.Conv) or sp.key == "conv1":
                continue
            for batch in (1, 64):
                for hw in (64, 32, 16, 8):
                    P = batch * (hw + 1) * (hw + 1)
                    bn, mb, kc = pick_tile(P, hw, sp.cin, sp.cout, sp.k * sp.k, 2 if sp.stride == 2 else 1, True)
                    assert mb * bn <= 256
                    p = _params(cin=sp.cin, cout=sp.cout, taps=sp.k * sp.k, stride=sp.stride, KC=kc, BN=bn, MB=mb,
                                H=hw, W=hw, in_H=hw * sp.stride, in_W=hw * sp.stride, N=batch,
                                flags=(4 if sp.stride == 2 else 0) | (2 if sp.cout % 16 else 0))
                    assert lib.hrnb_conv_smem_bytes(C.byref(p)) > 0, (sp, hw, bn, mb, lib.hrnb_last_error())


def test_null_pointers_are_rejected_without_touching_the_gpu():
    lib = _lib.lib()
    assert lib.hrnb_decode_argmax(None, 1, 8, 8, 0, 1, None, None, None, None) == -1
    assert lib.hrnb_loss_heatmap(None, None, 1, 64, 0, None, None, None, None, None) == -1
    assert lib.hrnb_softmax_softargmax(None, None, 1, 8, 8, None, None, None) == -1
    with pytest.raises(_lib.HrnbError):
        _lib.check(lib.hrnb_fuse_sum(None, None))
