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


def test_tvg_options_mirror_and_defaults(built):
    """The ctypes mirror of smb_tvg_options has the C layout (checked by compiling a two-line probe against the header),
    its flag constants are the header's, and the defaults are colmap.proto's (24-44) with COLMAP's watermark test on."""
    import shutil
    import tempfile
    L = matcher.load_library()
    o = matcher.smb_tvg_options()
    L.smb_default_tvg_options(ctypes.byref(o))
    assert (o.min_num_inliers, o.min_num_trials, o.max_num_trials, o.flags) == (15, 30, 10000, 0)
    assert (o.max_error, o.confidence, o.min_inlier_ratio, o.max_h_inlier_ratio) == (4.0, 0.999, 0.25, 0.8)
    hdr = open(os.path.join(ROOT, "include", "smb.h")).read()
    for name, val in (("SMB_TVG_NO_WATERMARK", matcher.SMB_TVG_NO_WATERMARK), ("SMB_TVG_MULTIPLE_MODELS", matcher.SMB_TVG_MULTIPLE_MODELS)):
        assert re.search(rf"#define {name} {val}\b", hdr), name
    if shutil.which("gcc"):
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "probe.c")
            open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "smb.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                                 'sizeof(smb_tvg_options), offsetof(smb_tvg_options, flags), offsetof(smb_tvg_options, max_error), '
                                 'offsetof(smb_tvg_options, seed), sizeof(smb_tvg), offsetof(smb_tvg, F)); return 0; }\n')
            exe = os.path.join(d, "probe")
            subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
            got = [int(x) for x in subprocess.check_output([exe]).split()]
        t, g = matcher.smb_tvg_options, matcher.smb_tvg
        assert got == [ctypes.sizeof(t), t.flags.offset, t.max_error.offset, t.seed.offset, ctypes.sizeof(g), g.F.offset]


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
