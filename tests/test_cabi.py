"""CPU: the C-ABI library builds, loads and exports every symbol include/tfswa_b200.h declares;
the host-side mirror refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    G.build()
    from tfswa_unet_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "tfswa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tfswa_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(built.SIGNATURES), (declared ^ set(built.SIGNATURES))
    lib = built.lib()
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_error_strings(built):
    lib = built.lib()
    assert b"sm_100a" in lib.tfswa_version()
    assert isinstance(lib.tfswa_last_error(), bytes)


def test_invalid_arguments_are_reported_not_crashed(built):
    import ctypes as C
    lib = built.lib()
    a = built.LinearArgs()      # all NULL
    rc = lib.tfswa_linear_fwd(C.byref(a), None)
    assert rc == -1 and b"null" in lib.tfswa_last_error()
    at = built.AttnArgs()
    assert lib.tfswa_attn_fwd(C.byref(at), None) == -1


def test_no_cpu_fallback():
    import tfswa_unet_b200 as T
    m = T.TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 16, 16))
    blk = T.TFSWABlock(32, 32, 8, 4, 8).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.zeros(1, 32, 16, 16))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tfswa-unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_reference_import_path_shim():
    import tfswa_unet_b200 as T
    T.install_as_reference("src_shim_test.models")
    from src_shim_test.models.tfswa_unet import TFSWAUNet
    from src_shim_test.models.attention import TemporalSequenceAttention, window_partition, window_reverse
    from src_shim_test.models.blocks import TFSWABlock
    assert TFSWAUNet is T.TFSWAUNet and "TFSWABlock" in TFSWABlock.__name__
    x = torch.arange(2 * 3 * 16 * 8, dtype=torch.float32).reshape(2, 3, 16, 8)
    assert torch.equal(window_reverse(window_partition(x, 8), 8, 16, 8), x)


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference tree not present (GPU box)")
def test_install_as_reference_keeps_the_reference_packages_importable():
    """with the real reference on sys.path the shim must not shadow `src` / `src.models` with empty stub
    packages - the unmodified Trainer / SourceSeparator have to import and must pick up the B200 model."""
    import subprocess
    import sys
    code = (
        "import sys, types\n"
        "for m in ('soundfile', 'musdb'): sys.modules[m] = types.ModuleType(m)\n"
        "import tfswa_unet_b200 as T\n"
        "T.install_as_reference()\n"
        "import src.training.trainer as tr, src.evaluation.inference as inf, src.data.stft_processor as sp\n"
        "from src.models.tfswa_unet import TFSWAUNet\n"
        "from src.models.blocks import TFSWABlock\n"
        "assert TFSWAUNet is T.TFSWAUNet and inf.TFSWAUNet is T.TFSWAUNet\n"
        "assert tr.Trainer.__module__ == 'src.training.trainer' and sp.STFTProcessor is not None\n"
        "assert 'reference' in sys.modules['src'].__path__[0]\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=REFERENCE + os.pathsep + ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
