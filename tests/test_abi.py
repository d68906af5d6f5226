"""CPU-only checks of the drop-in boundary: libsmb.so builds for sm_100a, loads without a GPU, exports every
symbol include/smb.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from scanner_colmap_b200 import matcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "smb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(built):
    L = matcher.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libsmb.so does not export {n}"
    assert L.smb_abi_version() == 2


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "smb.h")).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def test_default_options_match_colmap_proto(built):
    L = matcher.load_library()
    o = matcher.smb_options()
    L.smb_default_options(ctypes.byref(o))
    # /root/reference/integration/op_cpp/colmap.proto:14-24
    assert (o.max_ratio, o.max_distance, o.cross_check, o.max_num_matches) == (0.8, 0.7, 1, 32768)
    assert o.engine == matcher.ENGINE_TCGEN05


def test_sass_is_blackwell_native(built):
    """The shipped cubin must contain the tcgen05 / TMA / TMEM instructions, for sm_100a only."""
    out = subprocess.run(["cuobjdump", "-sass", matcher.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCIMMA", "UTMALDG", "LDTM", "VIMNMX3"):
        assert mnemonic in sass, mnemonic


def test_product_library_has_one_engine(built):
    """north_star: no multi-backend dispatch.  The CUDA-core cross-check engine exists only in the tests' build
    (libsmb_test.so, -DSMB_TEST_ENGINES); the product library must not contain it."""
    def kernels(path):
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True)
        if out.returncode != 0:
            pytest.skip("cuobjdump unavailable")
        return set(re.findall(r"Function : (\S+)", out.stdout))
    prod, test = kernels(matcher.LIB_PATH), kernels(matcher.TEST_LIB_PATH)
    assert not any("dp4a" in k for k in prod), prod
    assert any("dp4a" in k for k in test)
    assert any("score_tcgen05_kernel" in k for k in prod)
    assert prod < test          # same kernels otherwise


def test_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(matcher.SmbError) as e:
        matcher.SiftMatcher()
    assert e.value.code in (matcher.SMB_ENODEVICE, matcher.SMB_ECUDA)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "scanner_colmap_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn
