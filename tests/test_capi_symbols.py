"""CPU: libanimerec.so builds, loads and exports every symbol include/animerec.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from anime_recommendations_b200 import build
    return build.build_lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "animerec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ar_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared in animerec.h but not exported: %s" % missing


def test_binding_covers_header(libpath):
    from anime_recommendations_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_symbols()
    assert _capi.lib().ar_abi_version() == _capi.ABI_VERSION


def test_struct_layouts_match_header():
    """ctypes mirrors of ar_table / ar_plan / ar_train_ctx have the sizes the C compiler gives them."""
    import subprocess
    import tempfile
    from anime_recommendations_b200 import _capi
    prog = r'''
#include <stdio.h>
#include "animerec.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ar_table), sizeof(ar_plan), sizeof(ar_train_ctx), sizeof(ar_dist_ctx), sizeof(ar_shard_ctx), sizeof(ar_peer_ctx), sizeof(ar_sched)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_capi.ArTable), ctypes.sizeof(_capi.ArPlan), ctypes.sizeof(_capi.ArTrainCtx),
                     ctypes.sizeof(_capi.ArDistCtx), ctypes.sizeof(_capi.ArShardCtx), ctypes.sizeof(_capi.ArPeerCtx),
                     ctypes.sizeof(_capi.ArSched)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import anime_recommendations_b200 as ar
    with pytest.raises(ar.AnimerecError):
        ar.EmbeddingDotModel(10, 10, 8)
    with pytest.raises(ar.AnimerecError):
        ar.similarity.as_table([[1.0, 2.0, 3.0, 4.0]])


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline may touch oracle/."""
    pkg = os.path.join(ROOT, "anime_recommendations_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
