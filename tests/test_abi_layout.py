"""The ctypes mirrors of the C ABI's structs (avatarclip_b200/{_lib,clip_vit,losses}.py) against include/avc_b200.h as a C
compiler lays them out: the header is compiled as plain C (gcc, no CUDA), a small program prints sizeof and every field's
offsetof, and each ctypes Structure must agree field for field.  Also proves the header is self-contained C."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "avc_b200.h")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def _pairs():
    from avatarclip_b200 import _lib, clip_vit, losses
    return [("avc_neus_cfg", _lib.NeusCfg), ("avc_neus_outputs", _lib.NeusOutputs), ("avc_neus_cotangents", _lib.NeusCotangents),
            ("avc_clip_cfg", clip_vit.ClipCfg), ("avc_clip_layer_weights", clip_vit.ClipLayerW),
            ("avc_clip_weights", clip_vit.ClipW), ("avc_loss_inputs", losses.LossInputs)]


def _c_fields(struct_name):
    """Field names of `typedef struct <name> { ... } <name>;` in declaration order (comments stripped)."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (struct_name, struct_name), src, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "const float* a", "float light_dir[3]", "float igr_weight, mask_weight, clip_weight", "int32_t R, S, H, W"
        first, *rest = decl.split(",")
        names.append(re.search(r"([A-Za-z_]\w*)\s*(\[\w+\])?\s*$", first).group(1))
        names += [re.search(r"([A-Za-z_]\w*)\s*(\[\w+\])?\s*$", r).group(1) for r in rest]
    return names


def test_ctypes_structures_match_the_c_layout(tmp_path):
    pairs = _pairs()
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "avc_b200.h"', "int main(void) {"]
    for cname, _ in pairs:
        prog.append(f'  printf("S {cname} %zu\\n", sizeof({cname}));')
        for f in _c_fields(cname):
            prog.append(f'  printf("F {cname} {f} %zu\\n", offsetof({cname}, {f}));')
    prog += ['  printf("V %d\\n", AVC_ABI_VERSION);', "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    sizes = {l.split()[1]: int(l.split()[2]) for l in out if l.startswith("S ")}
    offs = {(l.split()[1], l.split()[2]): int(l.split()[3]) for l in out if l.startswith("F ")}
    for cname, cls in pairs:
        assert C.sizeof(cls) == sizes[cname], cname
        c_names = _c_fields(cname)
        assert [n for n, *_ in cls._fields_] == c_names, cname                      # same fields, same order
        for n in c_names:
            assert getattr(cls, n).offset == offs[(cname, n)], (cname, n)
    version = int([l for l in out if l.startswith("V ")][0].split()[1])
    from avatarclip_b200 import _lib
    assert _lib.lib().avc_abi_version() == version == 3
