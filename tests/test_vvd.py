"""Volume I/O next to the path: the .vvd reader of the host mirror against files written in the reference's
format (V/modules/core/io/vvdformat.cpp:59-120,269-330) and, when the reference tree is present, against the volumes it bundles."""
import os

import numpy as np
import pytest

from voreen_b200 import abi, host

REF_VOLUMES = "/root/reference/voreen/resource/voreenve/volumes"


def write_vvd(path, raw_name, fmt, dims, spacing=None, offset=None, transformation=None, rwm=None, old_style=False, extra_volume=False):
    """What VvdVolumeWriter emits (XmlSerializer with attributes): one <Volume> with RawData, MetaData and DerivedData."""
    meta = []
    if offset is not None:
        meta.append(f'<MetaItem name="Offset" type="Vec3MetaData"><value x="{offset[0]!r}" y="{offset[1]!r}" z="{offset[2]!r}" /></MetaItem>')
    if rwm is not None:
        meta.append(f'<MetaItem name="RealWorldMapping" type="RealWorldMappingMetaData"><value scale="{rwm[0]!r}" offset="{rwm[1]!r}" unit="HU &amp; more" /></MetaItem>')
    if spacing is not None:
        meta.append(f'<MetaItem name="Spacing" type="Vec3MetaData">\n <value x="{spacing[0]!r}" y="{spacing[1]!r}" z="{spacing[2]!r}" />\n</MetaItem>')
    if transformation is not None:
        rows = "".join(f'<value.row{r} x="{transformation[r][0]!r}" y="{transformation[r][1]!r}" z="{transformation[r][2]!r}" w="{transformation[r][3]!r}" />' for r in range(4))
        meta.append(f'<MetaItem name="Transformation" type="Mat4MetaData">{rows}</MetaItem>')
    meta.append('<MetaItem name="name" type="StringMetaData" value="a &lt;test&gt; volume" />')
    if old_style:
        raw = f'<RawData filename="{raw_name}" format="{fmt}" x="{dims[0]}" y="{dims[1]}" z="{dims[2]}" />'
    else:
        raw = (f'<RawData format="{fmt}" x="{dims[0]}" y="{dims[1]}" z="{dims[2]}"><Paths noPathSet="false"><paths><item value="{raw_name}" />'
               f'<item value="/somewhere/else/{raw_name}" /></paths></Paths></RawData>')
    vol = (f'<Volume>{raw}<MetaData>{"".join(meta)}</MetaData><DerivedData><DerivedItem type="VolumeMinMax" min="0" max="1" minNorm="0" maxNorm="1" />'
           f'</DerivedData></Volume>')
    other = '<Volume><RawData filename="other.raw" format="uint8" x="3" y="4" z="5" /><MetaData></MetaData></Volume>' if extra_volume else ""
    with open(path, "w") as f:
        f.write(f'<?xml version="1.0" ?>\n<!-- written by tests -->\n<VoreenData version="1">\n<Volumes>{other}{vol}</Volumes>\n</VoreenData>\n')


def test_vvd_new_style_with_all_metadata(tmp_path):
    t = [[0.0, -1.0, 0.0, 0.25], [1.0, 0.0, 0.0, -0.5], [0.0, 0.0, 2.0, 1.5], [0.0, 0.0, 0.0, 1.0]]
    p = tmp_path / "a.vvd"
    write_vvd(p, "a.raw", "uint16", (40, 33, 27), spacing=(0.5, 0.25, 0.125), offset=(-1.0, 2.0, 0.375), transformation=t, rwm=(4095.0, -1024.0), extra_volume=True)
    assert host.vvd_count(str(p)) == 2
    d, raw = host.parse_vvd(str(p), 1)
    assert raw == str(tmp_path / "a.raw")
    assert list(d.dims) == [40, 33, 27] and d.channels == 1 and d.dtype == abi.U16
    assert list(d.spacing) == [0.5, 0.25, 0.125] and list(d.offset) == [-1.0, 2.0, 0.375]
    assert np.allclose(np.array(d.phys_to_world[:]).reshape(4, 4), np.array(t, np.float32), rtol=0, atol=0)
    assert d.rwm_scale == 4095.0 and d.rwm_offset == -1024.0
    d0, raw0 = host.parse_vvd(str(p), 0)
    assert list(d0.dims) == [3, 4, 5] and d0.dtype == abi.U8 and raw0.endswith("other.raw")
    with pytest.raises(abi.B200RCError):
        host.parse_vvd(str(p), 2)


@pytest.mark.parametrize("fmt,dtype,ch,scale", [("uint8", abi.U8, 1, 255.0), ("uint16", abi.U16, 1, 65535.0), ("float", abi.F32, 1, 1.0),
                                                  ("Vector2(uint16)", abi.U16, 2, 65535.0), ("Vector3(float)", abi.F32, 3, 1.0), ("Vector4(uint8)", abi.U8, 4, 255.0)])
def test_vvd_old_style_formats_and_defaults(tmp_path, fmt, dtype, ch, scale):
    p = tmp_path / "b.vvd"
    write_vvd(p, "sub/b.raw", fmt, (7, 8, 9), old_style=True)
    d, raw = host.parse_vvd(str(p))
    assert raw == str(tmp_path / "sub/b.raw")
    assert (d.dtype, d.channels) == (dtype, ch)
    # no RealWorldMapping in the file -> the denormalising mapping (vvdvolumereader.cpp:103-118); default spacing 1, offset 0, identity
    assert d.rwm_scale == scale and d.rwm_offset == 0.0
    assert list(d.spacing) == [1.0, 1.0, 1.0] and list(d.offset) == [0.0, 0.0, 0.0]
    assert np.array_equal(np.array(d.phys_to_world[:]).reshape(4, 4), np.eye(4, dtype=np.float32))


@pytest.mark.parametrize("fmt", ["int16", "uint32", "double", "Vector5(uint8)", "Vector2(int8)"])
def test_vvd_unsupported_formats_fail_loudly(tmp_path, fmt):
    p = tmp_path / "c.vvd"
    write_vvd(p, "c.raw", fmt, (4, 4, 4))
    with pytest.raises(abi.B200RCError, match="unsupported format"):
        host.parse_vvd(str(p))
    with pytest.raises(abi.B200RCError, match="Failed to open"):
        host.parse_vvd(str(tmp_path / "missing.vvd"))


@pytest.mark.skipif(not os.path.isdir(REF_VOLUMES), reason="reference tree not present (GPU box)")
def test_vvd_bundled_reference_volumes():
    d, raw = host.parse_vvd(os.path.join(REF_VOLUMES, "nucleon.vvd"))
    assert list(d.dims) == [41, 41, 41] and d.dtype == abi.U8 and d.channels == 1 and raw == os.path.join(REF_VOLUMES, "nucleon.raw")
    assert np.allclose(d.spacing[:], [0.04878049] * 3, rtol=0, atol=1e-9) and np.allclose(d.offset[:], [-0.99999994] * 3, rtol=0, atol=1e-9)
    assert d.rwm_scale == 255.0 and os.path.getsize(raw) == 41 ** 3
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "nucleon.npz"))
    if "volume" in z.files:
        assert np.array_equal(np.fromfile(raw, np.uint8).reshape(41, 41, 41), z["volume"])
    d, raw = host.parse_vvd(os.path.join(REF_VOLUMES, "walnut.vvd"))
    assert d.dtype == abi.U16 and os.path.getsize(raw) == int(np.prod(d.dims[:])) * 2
    d, raw = host.parse_vvd(os.path.join(REF_VOLUMES, "aneurysm.vvd"))          # new-style <Paths>, Vector3(float), explicit RealWorldMapping
    assert list(d.dims) == [83, 83, 84] and d.dtype == abi.F32 and d.channels == 3 and raw == os.path.join(REF_VOLUMES, "aneurysm.raw")
    assert d.rwm_scale == 1.0 and d.rwm_offset == 0.0 and np.allclose(d.spacing[:], [0.600000024] * 3)
    assert os.path.getsize(raw) == 83 * 83 * 84 * 12
