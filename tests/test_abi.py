"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports exactly what
include/lrag.h declares; the ctypes binding covers every declared function; and nothing on the
product side imports the oracle or falls back to a CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "lrag.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(lrag_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from legal_rag_b200 import build
    return str(build.build_native())


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("lrag_init", "lrag_dense_topk_bf16", "lrag_bm25_topk", "lrag_maxsim_rerank_bf16", "lrag_fuse_topk",
                 "lrag_topk_merge", "lrag_topk_select_f32"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [n for n in _declared() if n not in exported]
    assert not missing, missing
    lib = ctypes.CDLL(lib_path)
    for n in _declared():
        assert getattr(lib, n) is not None


def test_ctypes_binding_covers_the_header(lib_path):
    from legal_rag_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared()
    lib = _native.load()
    assert lib.lrag_version() == 100


def test_library_is_sm100a_with_tcgen05_and_tma(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "no tcgen05.mma in the SASS"
    assert "LDTM" in sass, "no tcgen05.ld in the SASS"
    assert "UTMALDG" in sass, "no TMA tensor loads in the SASS"


def test_select_workspace_sizing_needs_no_gpu(lib_path):
    """The workspace queries are host arithmetic (the caller sizes its allocation before any launch): rows short enough
    for one CTA each need none; long rows need room for every streaming CTA's k keys, a handful of lists per row."""
    from legal_rag_b200 import _native
    lib = _native.load()
    assert lib.lrag_topk_select_workspace_bytes(4, 1000, 100) == 0
    assert lib.lrag_topk_select_workspace_bytes(4096, 20_000, 100) == 0          # many medium rows: one CTA per row
    for nq, N, k in [(64, 1_000_000, 100), (256, 1_000_000, 100), (8, 50_000_000, 100), (3, 700_001, 1024), (1, 4_200_000, 1)]:
        b = lib.lrag_topk_select_workspace_bytes(nq, N, k)
        assert b >= nq * k * 8 and b % 256 == 0
        assert b <= 8 * (nq * k * 8) * max(1, 2 * 148 // nq + 2), (nq, N, k, b)     # bounded by lists of ~one wave of CTAs + slack
    assert lib.lrag_topk_select_workspace_bytes(0, 10, 10) == 0


def test_init_without_a_gpu_fails_loudly(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from legal_rag_b200 import _native, engine
    with pytest.raises(_native.LragError, match="no CUDA device|CPU fallback"):
        _native.init()
    with pytest.raises(RuntimeError):
        engine.dense_topk(torch.zeros((4, 64), dtype=torch.bfloat16), torch.zeros((1, 64), dtype=torch.bfloat16), 2)


def test_product_code_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "legal_rag_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn), encoding="utf-8") as f:
                    src = f.read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or \
                        re.search(r"(__import__|import_module)\(\s*['\"]oracle", src):
                    bad.append(fn)
    assert not bad, bad


def _build_c_smoke(tmp_path, lib_path):
    """gcc-compiles tests/c_abi/smoke.c (a plain-C caller of the ABI) against the header and the library."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not installed")
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(os.path.abspath(str(lib_path)))
    cmd = ["gcc", "-O1", "-Wall", "-Werror", os.path.join(root, "tests", "c_abi", "smoke.c"), "-I" + os.path.join(root, "include"),
           "-I" + os.path.join(cuda, "include"), "-L" + libdir, "-llrag", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_a_plain_c_caller_compiles_and_links_against_the_header(tmp_path, lib_path):
    """The boundary is a C ABI: a C translation unit that includes lrag.h and calls the entry points with device pointers
    must compile with -Wall -Werror and link against liblrag.so (it is RUN by the GPU suite)."""
    _build_c_smoke(tmp_path, lib_path)


@pytest.mark.gpu
def test_a_plain_c_caller_gets_the_right_answers(tmp_path, lib_path):
    import subprocess
    exe = _build_c_smoke(tmp_path, lib_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c abi smoke ok" in r.stdout
