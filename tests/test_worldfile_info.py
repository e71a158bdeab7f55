"""vp_world_file_info needs no device: header parsing of the world file (deadcode.c:320-350 layout)."""
import pytest

import voxplat_b200 as vpb


def test_header_is_parsed(tmp_path):
    p = tmp_path / "w.bin"
    p.write_bytes(bytes([0x89]) + b"VOXPLAT" + bytes([6, 5, 2, 5]) + b"\0" * 64)
    assert vpb.world_file_info(str(p)) == (6, (5, 2, 5), 76)


def test_not_a_world_file(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"VOXPLAT\0\0\0\0\0\0")
    with pytest.raises(RuntimeError):
        vpb.world_file_info(str(p))
    with pytest.raises(RuntimeError):
        vpb.world_file_info(str(tmp_path / "missing.bin"))
