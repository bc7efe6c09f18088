"""CPU (-m "not gpu"): the C-ABI shared library builds/loads here and exports every symbol that
include/repmode_b200.h declares (no compute calls without a GPU); host-side API surface and the
no-CPU-fallback rule."""
import argparse
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "repmode_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mode_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from repmode_b200 import lib
    so = lib.load()
    names = _header_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(lib.SO_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    assert set(names) == set(lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert so.mode_version() == 1
    assert so.mode_packed_weight_elems(32, 32) == 125 * 32 * 32
    assert so.mode_packed_weight_elems(1, 8) == 125 * 8 * 32          # K padded to one 32-chunk
    assert so.mode_reparam_bwd_workspace_bytes(64, 32, 2) == 2 * 2 * 5 * 32 * 4


def test_sass_is_sm100a_tensor_core_code():
    """The shipped .so carries tcgen05 / TMA machine code for sm_100a (UTCHMMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    from repmode_b200 import lib
    lib.load()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    from repmode_b200.nn_modules import MoDEConv
    m = MoDEConv(5, 3, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 4, 4, 4), torch.tensor([0]))


def test_plugin_surface_matches_reference_keys():
    """`fnet.nn_modules.RepMode.Net(opts)` resolves, takes the reference ctor signature and produces the
    reference's 309 state_dict keys in the reference's order (checked against a golden state_dict)."""
    import importlib
    from tests.util import load_golden, params_of
    mod = importlib.import_module("fnet.nn_modules.RepMode")
    for name in ("Net", "MoDEConv", "MoDESubNet2Conv", "MoDEEncoderBlock", "MoDEDecoderBlock"):
        assert hasattr(mod, name)
    net = mod.Net(argparse.Namespace(adopted_datasets=list(range(3)), gpu_ids=-1), mult_chan=2)
    ref = params_of(load_golden("net_eval_small"))
    assert list(net.state_dict().keys()) == list(ref.keys())
    assert len(ref) == 309
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == ref[k].shape, k
    net.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()}, strict=True)
    with pytest.raises(AssertionError):
        mod.MoDEConv(5, 3, 4, 4, conv_type="other")


def test_pack_weights_helper_roundtrip():
    import numpy as np
    from tests.util import pack_weights
    w = np.random.RandomState(0).randn(1, 8, 40, 5, 5, 5).astype(np.float32)
    p = pack_weights(w, half=False).reshape(1, 125, 2, 8, 32)
    assert p[0, 62, 1, 3, 7] == w[0, 3, 39, 2, 2, 2] and p[0, 62, 1, 3, 8] == 0


def test_header_is_plain_c(tmp_path):
    """include/repmode_b200.h is the drop-in boundary: it must compile as C99 on its own (no C++-isms, every type it uses
    declared before use) and link against the shipped library."""
    import os
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "repmode_b200.h"\n'
                   "int main(void) {\n"
                   "    mode_rowmap_t m = {0, 1, 1, 1, 0};\n"
                   "    mode_reparam_item_t it;\n"
                   "    mode_planes_t p;\n"
                   "    (void)m; (void)it; (void)p;\n"
                   "    return mode_version() > 0 ? 0 : 1;\n"
                   "}\n")
    obj = tmp_path / "hdr.o"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src),
                        "-o", str(obj)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
