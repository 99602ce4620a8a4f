"""The C++ host mirror of the reference's layer surface (caffe_escoin_b200/host/escort_conv_layer.hpp):
compiles everywhere (CPU check), runs its gtest-shaped program on the GPU box (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "host_mirror_test.cpp")
BIN = os.path.join(ROOT, "tests", "host", "host_mirror_test.bin")
CUDA = "/usr/local/cuda"


def _build():
    import __graft_entry__ as ge
    ge.build()
    pkg = os.path.join(ROOT, "caffe_escoin_b200")
    cmd = ["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + CUDA + "/include", SRC, "-o", BIN,
           "-L" + pkg, "-lescort_b200", "-L" + os.path.join(ROOT, "oracle"), "-lescort_oracle", "-L" + CUDA + "/lib64",
           "-lcudart", "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-rpath," + CUDA + "/lib64"]
    subprocess.run(cmd, check=True)


def test_host_mirror_compiles_and_links():
    """No GPU needed: the header, the C ABI it calls and the oracle symbols it checks against all resolve."""
    _build()
    assert os.path.exists(BIN)
    r = subprocess.run([BIN], capture_output=True, text=True)
    # without a device the program must refuse to run (exit 77), never fall back to the CPU
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_on_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "[  PASSED  ]" in r.stdout
