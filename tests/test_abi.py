"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/lumb200.h
declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

from luminary_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lumb200.h")).read()
    declared = set(re.findall(r"\b(lumb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = api.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/lumb200.h but not exported"
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_struct_sizes_match_header():
    assert C.sizeof(api.Mesh) == 40
    assert C.sizeof(api.Instance) == 44
    assert C.sizeof(api.Material) == 68
    assert C.sizeof(api.Texture) == 48  # the mipmap word fills what was padding
    assert C.sizeof(api.OutputParams) == 72
    assert C.sizeof(api.AdaptiveSampling) == 44
    assert C.sizeof(api.Settings) == 16
    assert C.sizeof(api.Camera) == 52
    assert C.sizeof(api.Sky) == 124  # the full LuminarySky since the procedural atmosphere and its HDRI bake are on the path
    assert api.VERTEX_OUT.itemsize == 248  # 4 NEE slots
    assert C.sizeof(api.Stats) == 112


def test_device_count_without_gpu_reports_error_or_zero():
    import torch

    if torch.cuda.is_available():
        assert api.device_count() >= 1
    else:
        try:
            assert api.device_count() == 0
        except api.LuminaryError as e:
            assert e.code == 8


def test_struct_sizes_against_the_c_compiler(tmp_path):
    """sizeof() of every struct the Python mirror binds, as gcc lays it out from include/lumb200.h."""
    import subprocess

    names = {"Lumb200Mesh": C.sizeof(api.Mesh), "Lumb200Instance": C.sizeof(api.Instance), "Lumb200Material": C.sizeof(api.Material),
             "Lumb200Texture": C.sizeof(api.Texture), "Lumb200OutputParams": C.sizeof(api.OutputParams), "Lumb200Settings": C.sizeof(api.Settings),
             "Lumb200Camera": C.sizeof(api.Camera), "Lumb200Sky": C.sizeof(api.Sky), "Lumb200Stats": C.sizeof(api.Stats),
             "Lumb200AdaptiveSampling": C.sizeof(api.AdaptiveSampling), "Lumb200VertexIn": api.VERTEX_IN.itemsize,
             "Lumb200NeeSegment": api.NEE_SEGMENT.itemsize, "Lumb200VertexOut": api.VERTEX_OUT.itemsize}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "lumb200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, size in names.items():
        assert int(out[n]) == size, (n, out[n], size)
