"""CPU-only checks of the drop-in boundary: libmrsb.so loads, exports exactly what include/mrsb.h
declares, its structs have the layout the Python mirror assumes, and — with no GPU — every compute
entry fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrsb.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mrs_multirotor_simulator_b200 import _lib

    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mrsb.h but not exported by libmrsb.so"
    assert sorted(_lib.SIGNATURES) == names, "the ctypes mirror and the header disagree on the symbol list"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\b(mrsb_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == names, "libmrsb.so exports symbols the header does not declare (or the reverse)"


def test_struct_layouts_match_the_c_header(tmp_path):
    from mrs_multirotor_simulator_b200._lib import ControllerParams, CreateInfo, DeviceView, ModelParams

    probe = tmp_path / "probe.c"
    probe.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "mrsb.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(mrsb_model_params), sizeof(mrsb_controller_params), sizeof(mrsb_create_info), sizeof(mrsb_device_view));
  printf("%zu %zu %zu %zu\\n", offsetof(mrsb_model_params, g), offsetof(mrsb_model_params, J), offsetof(mrsb_model_params, allocation_matrix),
         offsetof(mrsb_model_params, ground_z));
  printf("%zu %zu %zu\\n", offsetof(mrsb_controller_params, rate_kp), offsetof(mrsb_controller_params, pos_max_velocity), offsetof(mrsb_create_info, spawn_heading));
  printf("%d %d %d\\n", MRSB_MAX_MOTORS, MRSB_POSITION_CMD, MRSB_ERR_CAPACITY);
  return 0;
}''')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    sizes = list(map(int, lines[0].split()))
    assert sizes == [C.sizeof(ModelParams), C.sizeof(ControllerParams), C.sizeof(CreateInfo), C.sizeof(DeviceView)]
    offs = list(map(int, lines[1].split()))
    assert offs == [ModelParams.g.offset, ModelParams.J.offset, ModelParams.allocation_matrix.offset, ModelParams.ground_z.offset]
    offs = list(map(int, lines[2].split()))
    assert offs == [ControllerParams.rate_kp.offset, ControllerParams.pos_max_velocity.offset, CreateInfo.spawn_heading.offset]
    assert lines[3].split() == ["8", "10", "-4"]


def test_header_is_plain_c_and_cites_the_reference():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for banned in ("torch", "at::", "std::", "class ", "template", "&"):
        assert banned not in code, banned  # plain pointers and sizes only
    for cite in ("US:304-380", "US:175-248", "MM:24-88", "SIM:295-359", "US:254-272", "MM:424-433", "CTL/mixer.hpp:72-101"):
        assert cite in src, cite


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mrs_multirotor_simulator_b200 import UavBatch, airframe
    from mrs_multirotor_simulator_b200._lib import MrsbError

    with pytest.raises(MrsbError) as e:
        UavBatch([airframe("x500")], spawn_xyz=np.zeros((4, 3)), n=4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or name it."""
    pkg = os.path.join(ROOT, "mrs_multirotor_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#\s*include\s*[<\"][^>\"]*oracle", r"liboracle", r"libref_nanoflann", r"libref_uavsystem", r"oracle/_"):
                    assert not re.search(pat, text, flags=re.M), (os.path.join(dirpath, f), pat)
    from mrs_multirotor_simulator_b200 import _lib

    ldd = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "oracle" not in ldd


def test_graft_entry_build_is_importable():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    assert callable(g.build) and callable(g.smoke)
