"""CPU tests of the drop-in boundary: the C-ABI libraries load, export every symbol their headers declare, the ctypes
views have the C layouts, and the product refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from voreen_b200 import abi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", src)))


def test_libb200rc_exports_every_declared_symbol():
    lib = abi.load_library()
    declared = _declared("b200rc.h", "b200rc_")
    assert set(declared) == set(abi.SYMBOLS), set(declared) ^ set(abi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.b200rc_abi_version() == abi.ABI_VERSION == 4


def test_libb200rc_host_exports_every_declared_symbol():
    lib = host.load_library()
    declared = _declared("b200rc_host.h", "b200host_")
    assert set(declared) == set(host.SYMBOLS), set(declared) ^ set(host.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_ctypes_layouts_match_the_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200rc_host.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200rc_volume_desc), sizeof(b200rc_light), sizeof(b200rc_camera), sizeof(b200rc_volume_params),
           sizeof(b200rc_params), sizeof(b200rc_eep_params), sizeof(b200rc_stats), sizeof(b200host_tf_key));
    printf("%zu %zu %zu %zu\n", offsetof(b200rc_params, light), offsetof(b200rc_params, sampler), offsetof(b200rc_params, image_band_remainder), offsetof(b200rc_eep_params, boxes));
    return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True).split("\n")
    sizes = [int(v) for v in out[0].split()]
    assert sizes == [C.sizeof(abi.VolumeDesc), C.sizeof(abi.Light), C.sizeof(abi.Camera), C.sizeof(abi.VolumeParams), C.sizeof(abi.Params),
                     C.sizeof(abi.EepParams), C.sizeof(abi.Stats), C.sizeof(host.TFKey)]
    offs = [int(v) for v in out[1].split()]
    assert offs == [abi.Params.light.offset, abi.Params.sampler.offset, abi.Params.image_band_remainder.offset, abi.EepParams.boxes.offset]


def test_no_cpu_fallback():
    """Without an sm_100 device the product refuses to create a context (B200RC_ERR_NO_DEVICE); nothing falls back to the CPU."""
    lib = abi.load_library()
    if lib.b200rc_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.b200rc_create(0, C.byref(h)) == abi.ERR_NO_DEVICE and not h.value
    import voreen_b200 as vb
    with pytest.raises(abi.B200RCError):
        vb.Context(0)
    with pytest.raises(abi.B200RCError):
        host.SingleVolumeRaycaster(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product package may reference it."""
    pkg = os.path.join(ROOT, "voreen_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "liborc" not in src and "vrc_oracle" not in src and "import orc" not in src, os.path.join(dp, f)
