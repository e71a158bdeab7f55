"""The GPU tests of tests/test_zz_gpu_round2_unverified.py have not run on hardware yet.  What CAN be checked without a
GPU is everything around the device: the test code itself, the compiled-reference half of every comparison, and that the
expectations are the ones the oracle restatement meets.  This file runs those very test functions with
tests/hoststore.HostContext (the oracle behind the Context method surface) in place of voxplat_b200.Context."""
import pytest

import helpers
import hoststore
import voxplat_b200 as vpb

pytestmark = pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")


@pytest.fixture()
def zz(monkeypatch):
    monkeypatch.setattr(vpb, "Context", hoststore.HostContext)
    import test_zz_gpu_round2_unverified as mod
    return mod


def test_nodes_against_reference_gfx(zz):
    zz.test_lod_nodes_match_the_reference_gfx_update_svl(4, (2, 1, 3))


@pytest.mark.parametrize("rb,bits,kind", [(7, (1, 0, 0), "random"), (6, (1, 1, 1), "random")])
def test_large_chunks(zz, rb, bits, kind):
    zz.test_large_chunks_against_the_compiled_reference(rb, bits, kind)


def test_edit_chunk_128(zz):
    zz.test_edit_sphere_chunk_128_matches_compiled_reference()


def test_flat_codec(zz):
    zz.test_flat_codec_any_length_and_run_split()


def test_slab_edits(zz):
    zz.test_edit_sphere_on_slab_contexts()


def test_raycast(zz):
    zz.test_raycast_matches_compiled_reference(5, (2, 1, 2), "terrain")


def test_golden_rays_nodes_edits(zz):
    zz.test_device_rays_nodes_edits_match_golden()
