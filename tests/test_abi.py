"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/voxplat_b200.h declares,
and refuses to work without a GPU (no CPU fallback)."""
import os
import re

import pytest

import voxplat_b200
from voxplat_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "voxplat_b200.h")).read()
    return sorted(set(re.findall(r"VP_API[^;(]*?\b(vp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    names = header_functions()
    assert len(names) >= 25
    for must in ("vp_ctx_create", "vp_rebuild_batch", "vp_upload_chunks_rle", "vp_rle_compress", "vp_rle_decompress",
                 "vp_chunk_make_splatlists", "vp_chunk_make_mesh", "vp_halo_pack"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    for name in header_functions():
        assert hasattr(lib, name), name
    # and the Python binding covers them all
    assert set(header_functions()) == set(lib._vp_signatures)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(voxplat_b200.VoxplatError) as e:
        voxplat_b200.Context(5, (1, 1, 1))
    assert e.value.code == -2


def test_product_does_not_touch_the_oracle():
    """The shipped package must never import/link/execute anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "voxplat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"libvoxoracle|libvoxref|vox_oracle|ref_harness|oracle/_ref|from oracle|import oracle", txt):
                    bad.append(f)
    assert not bad, bad
