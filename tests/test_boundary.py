"""The drop-in boundary on a box without a GPU: the C-ABI library loads and exports every symbol
include/ai4xde_b200.h declares; the product never touches oracle/; compute fails loudly without
CUDA (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ai4xde_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ax_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ai4xde_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run `python -m ai4xde_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # the binding covers the header (nothing is bound that the header does not declare)
    assert set(_lib.exported_symbols()) <= set(names)
    assert _lib.load().ax_version() >= 100


def test_descriptor_queries_work_without_a_gpu():
    from ai4xde_b200 import _lib
    from ai4xde_b200.jet import JetSpec

    net = _lib.make_net_desc([2, 256, 256, 256, 256, 256, 256, 1], "tanh")
    assert _lib.param_count(net) == 329985
    assert _lib.param_count(_lib.make_net_desc([2, 20, 20, 20, 1], "tanh")) == 921
    assert _lib.param_count(_lib.make_net_desc([2, 128, 128, 128, 128, 128, 3], "tanh")) == 66819
    js = JetSpec.covering([0, 1], [(0, 0)])
    assert js.dirs == (0, 1) and js.pairs == ((0, 0),) and _lib.jet_supported(js.desc)
    js = JetSpec.covering([1], [(1, 1)])
    assert js.dirs == (1,) and js.pairs == ((0, 0),)
    js = JetSpec.covering([0, 1], [(1, 1)])
    assert js.dirs == (1, 0) and js.pairs == ((0, 0),)
    js = JetSpec.covering([], [(0, 0), (1, 1)])
    assert js.S == 5 and _lib.jet_supported(js.desc)
    js = JetSpec.covering([0], [(0, 1)])
    assert js.S == 6 and _lib.jet_supported(js.desc)
    assert _lib.jet_supported(JetSpec.full(4).desc) and JetSpec.full(4).S == 15
    with pytest.raises(NotImplementedError):
        JetSpec.covering([0, 1, 2, 3, 4], [])
    assert _lib.engine_supported(net, JetSpec.covering([], [(0, 0), (1, 1)]).desc, "tf32")
    assert not _lib.engine_supported(_lib.make_net_desc([2, 20, 20, 1], "tanh"), js.desc, "tf32")
    with pytest.raises(NotImplementedError):
        _lib.make_net_desc([2, 8, 1], "relu")


def test_tensor_core_switches_are_named_and_refused_when_unknown():
    """include/ai4xde_b200.h ax_tc_set_option: the documented switches of the L2-chained forward are accepted (they only
    set host-side state: no GPU needed), an unknown name fails loudly with the list of names."""
    from ai4xde_b200 import _lib

    for name, value in (("l2chain", 1), ("l2chain_group", 3), ("l2chain_hints", 1), ("l2chain_defer", 2), ("l2chain_wres", 1)):
        _lib.tc_set_option(name, value)
    with pytest.raises(_lib.NativeLibraryError, match="l2chain_defer"):
        _lib.tc_set_option("l2chain_nonsense", 1)


def test_no_cpu_fallback():
    from ai4xde_b200 import _lib
    from ai4xde_b200.jet import JetSpec

    net = _lib.make_net_desc([2, 8, 1], "tanh")
    with pytest.raises(_lib.NativeLibraryError, match="no CPU path"):
        _lib.jet_forward(net, JetSpec.value_only().desc, torch.zeros((1, 33)), torch.zeros((2, 32)), 4, True)
    if not torch.cuda.is_available():
        import ai4xde_b200.dde as dde

        fnn = dde.nn.FNN([2, 8, 1], "tanh", "Glorot uniform")
        with pytest.raises(_lib.NativeLibraryError):
            fnn(torch.zeros((4, 2)))


def test_missing_library_is_loud(tmp_path):
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from ai4xde_b200 import _lib\n"
        "_lib.LIB_PATH = %r\n"
        "try:\n    _lib.load()\nexcept _lib.NativeLibraryError as e:\n    print('LOUD', e)\n"
    ) % (ROOT, str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD" in out.stdout and "no CPU or PyTorch fallback" in out.stdout


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ai4xde_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "_taylor_double" not in text, f
    code = "import sys; sys.path.insert(0, %r); import ai4xde_b200, ai4xde_b200.algorithm, ai4xde_b200.cases; " \
           "print([m for m in sys.modules if m.split('.')[0] == 'oracle'])" % ROOT
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True)
    assert out.stdout.strip() == "[]", out.stdout + out.stderr


def test_install_as_deepxde_runs_reference_style_source():
    """`import deepxde as dde` resolves to the fused namespace after install_as_deepxde()."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import ai4xde_b200; ai4xde_b200.install_as_deepxde()\n"
        "import deepxde as dde, deepxde.backend as bkd\n"
        "from deepxde import gradients as grad, losses\n"
        "from deepxde.backend import backend_name, torch\n"
        "dde.optimizers.config.set_LBFGS_options(maxiter=1000)\n"
        "assert dde.optimizers.LBFGS_options['maxiter'] == 1000 and backend_name == 'pytorch'\n"
        "print('OK', dde.Model.__module__)\n"
    ) % ROOT
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True)
    assert "OK ai4xde_b200.dde.model" in out.stdout, out.stdout + out.stderr
