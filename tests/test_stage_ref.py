"""oracle/stage_ref.py: the staged copy of the reference (what travels to the GPU box) is byte-identical to the checkout and
a modified staged tree is refused.  CPU; skipped where neither the checkout nor a staged tree exists."""
import os
import shutil

import pytest

from oracle import ref_loader, stage_ref


def _source():
    for root in ("/root/reference", stage_ref.STAGED):
        if os.path.isfile(os.path.join(root, "codes", "models", "modules", "Sakuya_arch_test.py")):
            return root
    pytest.skip("no reference checkout and no staged copy")


def test_stage_is_byte_identical_and_tamper_evident(tmp_path):
    src = _source()
    dest = str(tmp_path / "_ref")
    manifest = stage_ref.stage(src, dest)
    assert "codes/models/modules/Sakuya_arch_test.py" in manifest and "codes/custom_video_test.py" in manifest
    for rel in manifest:
        with open(os.path.join(src, rel), "rb") as a, open(os.path.join(dest, rel), "rb") as b:
            assert a.read() == b.read(), rel
    assert stage_ref.verify(dest)
    with open(os.path.join(dest, "codes", "models", "modules", "SIREN.py"), "a") as f:
        f.write("\n# edited\n")
    assert not stage_ref.verify(dest)                       # a staged tree that was edited is not the reference any more
    shutil.rmtree(dest)
    assert not stage_ref.verify(dest)


def test_loader_resolves_a_reference_and_reports_its_kind():
    _source()
    assert ref_loader.reference_kind() in ("checkout", "staged")
    assert ref_loader.reference_available()
