"""The overlay (overlay/*.patch) applies to the reference tree: every patch is applied to a scratch copy of the file it
names and the result carries the calls INTEGRATION.md describes.  CPU suite; skipped where /root/reference is absent
(the GPU box)."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = sorted(glob.glob(os.path.join(ROOT, "overlay", "*.patch")))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None,
                                reason="needs the reference tree and patch(1)")

EXPECT = {
    "src/caffe/layers/conv_layer.cu": ["escort_sconv_forward(this->escort_plan_", "escort_sconv_backward_weight(",
                                       "escort_sconv_backward_data(", "escort_bias_backward("],
    "src/caffe/layers/conv_relu_layer.cu": ["/*fuse_relu=*/1"],
    "src/caffe/layers/base_conv_layer.cpp": ["escort_plan_create(&eg", "escort_plan_autotune(escort_plan_",
                                             "escort_plan_destroy(escort_plan_)"],
    "include/caffe/layers/base_conv_layer.hpp": ["escort_plan *escort_plan_;", "escort_plan_(NULL)"],
    "src/caffe/util/math_functions.cu": ["escort_pack_csr(", "escort_stretch(", "escort_copy_input(", "escort_sconv_padded("],
    "src/caffe/parallel.cpp": ["escort_allreduce_grads(comm_"],
    "include/caffe/util/device_alternate.hpp": ["#define ESCORT_CHECK(call)"],
    "Makefile": ["LIBRARIES += escort_b200"],
    # the dense layers
    "src/caffe/layers/inner_product_layer.cu": ["escort_inner_product_forward(M_, K_, N_"],
    "src/caffe/layers/esc_conv_layer.cu": ["escort_dense_conv_workspace_bytes(&eg", "escort_dense_conv_forward(&eg"],
    "src/caffe/layers/esc_conv_layer.cpp": ["cudaFree(escort_ws_)"],
    "include/caffe/layers/esc_conv_layer.hpp": ["void *escort_ws_;", "escort_ws_bytes_(0)"],
}


def _target(patch):
    with open(patch) as f:
        for line in f:
            if line.startswith("+++ b/"):
                return line[6:].strip()
    raise AssertionError("no target in " + patch)


def test_every_integration_file_has_a_patch():
    assert sorted(_target(p) for p in PATCHES) == sorted(EXPECT)


@pytest.mark.parametrize("patch", PATCHES, ids=lambda p: os.path.basename(p))
def test_patch_applies_to_the_reference(patch, tmp_path):
    rel = _target(patch)
    dst = tmp_path / rel
    dst.parent.mkdir(parents=True, exist_ok=True)
    shutil.copy(os.path.join(REF, rel), dst)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "fuzz" not in r.stdout, r.stdout + r.stderr
    txt = dst.read_text()
    for needle in EXPECT[rel]:
        assert needle in txt, "%s: missing %r after patching" % (rel, needle)


def test_committed_patches_are_what_the_generator_writes(tmp_path):
    out = tmp_path / "overlay"
    shutil.copytree(os.path.join(ROOT, "overlay"), out)
    for p in out.glob("*.patch"):
        p.unlink()
    subprocess.run([sys.executable, str(out / "make_overlay.py"), REF], check=True, capture_output=True)
    for p in PATCHES:
        assert (out / os.path.basename(p)).read_text() == open(p).read(), os.path.basename(p) + " is stale"
