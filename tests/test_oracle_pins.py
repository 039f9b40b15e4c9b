"""CPU tests that pin the oracle (the checker) against everything the reference holds in-tree for this path:

  - the reference's own tgt headers, compiled where they lie into oracle/_ref/tgt_probe (camera / matrix arithmetic);
  - known answers the reference computed itself and serialised into its bundled nucleon.vvd: VolumeMinMax, the
    intensity histogram (voxel normalisation VolumeRAM::getVoxelNormalized);
  - the numeric constants of the GLSL modules (0.95 ERT threshold + alpha snap, SAMPLING_BASE_INTERVAL_RCP = 200,
    ISO +-0.02, step-size formula), checked through closed-form scenes;
  - the committed rendering of the reference's standard workspace (regression fixture, tests/golden/).
"""
import math
import os
import subprocess

import numpy as np
import pytest

import nucleon_scene
import orc
import scenes
from voreen_b200 import abi, host
from voreen_b200 import raycaster as rc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- tgt headers -----------------------------------------------------------------------------------------------------
def _probe(lines):
    if not os.path.exists(orc.PROBE):
        pytest.skip("oracle/_ref/tgt_probe not built (needs /root/reference at build time)")
    out = subprocess.run([orc.PROBE], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.strip().split("\n")
    return [np.array([float.fromhex(t) for t in ln.split()], np.float32) if ln != "singular" else None for ln in out]


def _hex(a):
    # decimal with 9 significant digits round-trips every float32 (the probe parses with operator>>)
    return " ".join("%.9g" % float(v) for v in np.asarray(a, np.float32).reshape(-1))


def test_camera_matrices_match_tgt_headers():
    rng = np.random.default_rng(7)
    cases = [((1.2, 1.05, 3.1), (0, 0, 0), (0, 1, 0)), ((-1.409960508, 0.934277534, 1.960826397), (-0.024390221,) * 3, (0.182561591, -0.822090268, 0.539294899))]
    cases += [(tuple(rng.uniform(-4, 4, 3)), tuple(rng.uniform(-0.2, 0.2, 3)), tuple(rng.uniform(-1, 1, 3))) for _ in range(20)]
    ref = _probe([f"lookat {_hex(e)} {_hex(f)} {_hex(u)}" for e, f, u in cases])
    for (e, f, u), r in zip(cases, ref):
        assert np.array_equal(orc.look_at(e, f, u), r)
        cam = host.make_camera(e, f, u, W=64, H=64)
        assert np.array_equal(np.array(cam.view[:], np.float32), r)           # product host mirror, same bits
    persp = [(45.0, 1.0, 0.1, 50.0), (45.000003815, 1.0, 0.100000001, 502.0), (30.0, 1.7777, 0.5, 10.0), (175.0, 0.6, 0.01, 5.0)]
    deg2rad = _probe([f"deg2rad {_hex([p[0]])}" for p in persp])
    ref = _probe([f"persp {_hex(d)} {_hex([p[1]])} {_hex([p[2]])} {_hex([p[3]])}" for p, d in zip(persp, deg2rad)])
    for p, d, r in zip(persp, deg2rad, ref):
        got = np.zeros(16, np.float32)
        orc.lib().orc_perspective(float(d[0]), p[1], p[2], p[3], orc.fp(got))       # createPerspective on tgt::deg2rad's own output
        assert np.array_equal(got, r)
    # Camera::getProjectionMatrix for window ratio >= 1 is createPerspective(deg2rad(fovy), ratio, near, far)
    for p, r in zip(persp[:2], ref[:2]):
        assert np.array_equal(orc.camera_projection(p[0], 1.0, p[2], p[3], 1.0), r)
        cam = host.make_camera((0, 0, 3), fovy=p[0], near=p[2], far=p[3], W=128, H=128)
        assert np.array_equal(np.array(cam.proj[:], np.float32), r)


def test_matrix_inverse_and_product_match_tgt_headers():
    rng = np.random.default_rng(11)
    mats = [rng.uniform(-2, 2, 16).astype(np.float32) for _ in range(16)]
    mats.append(orc.look_at((1.2, 1.05, 3.1), (0, 0, 0), (0, 1, 0)))
    mats.append(orc.camera_projection(45.0, 1.0, 0.1, 50.0, 1.0))
    inv = _probe([f"invert {_hex(m)}" for m in mats])
    for m, r in zip(mats, inv):
        got = np.zeros(16, np.float32)
        ok = orc.lib().orc_mat4_invert(orc.fp(m), orc.fp(got))
        assert (r is None) == (not ok)
        if r is not None:
            assert np.array_equal(got, r)
    prod = _probe([f"mul {_hex(a)} {_hex(b)}" for a, b in zip(mats[:-1], mats[1:])])
    for a, b, r in zip(mats[:-1], mats[1:], prod):
        got = np.zeros(16, np.float32)
        orc.lib().orc_mat4_mul(orc.fp(a), orc.fp(b), orc.fp(got))
        assert np.array_equal(got, r)


# ---- known answers serialised by the reference into nucleon.vvd ----------------------------------------------------------
def _nucleon():
    return np.load(os.path.join(GOLDEN, "nucleon.npz"))


def test_nucleon_minmax_and_histogram_known_answers():
    z = _nucleon()
    vox = z["voxels"]
    norm = vox.astype(np.float32) / np.float32(255.0)            # VolumeRAM::getVoxelNormalized for uint8
    mn, mx, mnn, mxn = z["minmax"]
    assert abs(float(norm.min()) - mnn) < 1e-7 and abs(float(norm.max()) - mxn) < 5e-8
    # VolumeHistogramIntensity: 256 buckets over [min, max] of the normalised values (histogram.cpp)
    lo, hi = z["hist_range"]
    idx = np.floor((norm.astype(np.float64) - lo) / (hi - lo) * 256.0).astype(np.int64)
    idx = np.clip(idx, 0, 255)
    hist = np.bincount(idx.reshape(-1), minlength=256)
    assert hist.sum() == z["hist"].sum() == vox.size
    # bucket edges fall between two uint8 levels with 249 levels over 256 buckets: allow a sample to sit in the neighbouring bucket
    assert np.abs(np.cumsum(hist) - np.cumsum(z["hist"])).max() <= np.max(np.bincount(vox.reshape(-1)))
    assert np.array_equal(hist == 0, z["hist"] == 0) or (np.abs(hist - z["hist"]) > 0).sum() < 16


# (The 64x64 VolumePreview blob serialised in nucleon.vvd is NOT reproduced by the current VolumePreview::createFrom
#  (volumepreview.cpp:51-176) applied to nucleon.raw in any slice / orientation / gamma -- it predates that code -- so it is
#  not used as a known answer for the linear sampler.)


# ---- constants of the GLSL modules through closed-form scenes ---------------------------------------------------------------
def _uniform_scene(alpha8, n=16, rate=2.0, **kw):
    vol = np.full((n, n, n), 32768, np.uint16)
    keys = [(0.0, (255, 128, 64, alpha8)), (1.0, (255, 128, 64, alpha8))]
    return scenes.Scene(vol, keys=keys, W=32, H=32, eye=(0.0, 0.0, 3.5), sampling_rate=rate, shading=abi.SHADE["none"], **kw)


def test_dvr_opacity_correction_and_step_size():
    """compositeDVR: a' = 1 - (1 - a)^(step * 200), step = 1 / (max(dim) * samplingRate) (mod_compositing.frag:40-59,
    volumeraycaster.cpp:312-322): a ray through a uniform medium takes floor(tEnd/step)+1 samples."""
    s = _uniform_scene(1)
    outs, depth, st = s.oracle_render()
    assert abs(s.step - 1.0 / (16 * 2.0)) < 1e-9
    e, ed, x, xd = s.oracle_eep()
    c = (16, 16)
    tend = float(np.linalg.norm(x[c][:3] - e[c][:3]))
    nsamp = math.floor(tend / s.step) + 1
    a = 1.0 - (1.0 - 1 / 255.0) ** (s.step * 200.0)
    expect_alpha = 1.0 - (1.0 - a) ** nsamp
    assert expect_alpha < 0.95
    assert abs(outs[0][c][3] - expect_alpha) < 2e-5
    assert abs(outs[0][c][0] - expect_alpha * 1.0) < 2e-5          # premultiplied colour, r = 255/255
    assert st.samples_nominal == st.samples_classified             # no early termination below 0.95


def test_early_ray_termination_threshold_and_snap():
    """earlyRayTermination (mod_raysetup.frag:56-74): stop at alpha >= 0.95 and set alpha to exactly 1.0."""
    s = _uniform_scene(200)
    outs, depth, st = s.oracle_render()
    img = outs[0]
    hit = img[..., 3] > 0
    assert hit.any() and np.all(img[..., 3][hit] == 1.0)
    assert st.samples_classified < st.samples_nominal
    a = 1.0 - (1.0 - 200 / 255.0) ** (s.step * 200.0)
    k = math.ceil(math.log(1 - 0.95) / math.log(1 - a) - 1e-9)      # samples until A >= 0.95
    assert st.samples_classified == st.rays_cast * k
    # colour keeps the un-snapped accumulated value
    assert abs(img[16, 16, 0] - (1.0 - (1.0 - a) ** k)) < 2e-5
    # MIP / MOP / MIDA on a connected output disable ERT (volumeraycaster.cpp:225-233)
    s2 = _uniform_scene(200, compositing=[abi.COMP["dvr"], abi.COMP["mip"], abi.COMP["dvr"]], output_connected=[1, 1, 0])
    _, _, st2 = s2.oracle_render()
    assert st2.samples_classified == st2.samples_nominal


def test_iso_window():
    """compositeISO (mod_compositing.frag:156-166): hit iff TF alpha within isoValue +- 0.02."""
    for a8, iso, hit in ((128, 0.5, True), (128, 0.53, False), (128, 0.52, True), (100, 0.5, False)):
        s = _uniform_scene(a8, compositing=[abi.COMP["iso"]] * 3, iso_value=iso)
        outs, _, _ = s.oracle_render()
        assert bool((outs[0][..., 3] == 1.0).any()) == hit, (a8, iso)


def test_tf_texture_quirks():
    """createTextureData (transfunc1d.cpp:85-116): threshold compared with round(thr*width); key interpolation truncates."""
    tf = orc.make_tf([(0.0, (0, 0, 0, 0)), (1.0, (255, 255, 255, 255))], threshold=(0.25, 0.75))
    t = orc.tf1d_texture(tf, 256)
    assert t[63, 3] == 0 and t[64, 3] == int(64 / 255 * 255) and t[192, 3] == 192 and t[193, 3] == 0
    tf = orc.make_tf([(0.0, (10, 0, 0, 0)), (1.0, (20, 0, 0, 0))])
    t = orc.tf1d_texture(tf, 256)
    assert t[254, 0] == 19 and t[255, 0] == 20           # 10 + 10 * 254/255 = 19.96 -> truncated
    assert np.array_equal(t, host.tf1d_texture([(0.0, (10, 0, 0, 0)), (1.0, (20, 0, 0, 0))], 256))


def test_regression_comparator():
    """filecomparators.cpp:482-632: RGBA distance / 2 > tolerance, neighbour search radius 1."""
    a = np.zeros((8, 8, 4), np.float32)
    b = a.copy()
    b[3, 3] = (0.05, 0, 0, 0)                     # 0.025 > 0.02 and no matching neighbour in a? (neighbours are 0 -> |0-0.05|/2 > tol)
    assert orc.image_diff_count(a, b) == 1
    b[3, 3] = (0.03, 0, 0, 0)
    assert orc.image_diff_count(a, b) == 0
    a[3, 4] = (0.05, 0, 0, 0); b[3, 3] = (0.05, 0, 0, 0)
    assert orc.image_diff_count(a, b) <= 1         # the shifted pixel is found within radius 1


# ---- regression fixture: the reference's standard workspace --------------------------------------------------------------------
def test_nucleon_workspace_fixture_reproduces():
    z = np.load(os.path.join(GOLDEN, "nucleon_oracle.npz"))
    s = nucleon_scene.load()
    assert s.props["classification"] == "pre-integrated-gpu" and s.props["gradient"] == "sobel" and float(s.props["samplingRate"]) == 4.0
    outs, depth, st, eepi = nucleon_scene.oracle_render(s)
    assert st.rays_cast == int(z["rays"]) and st.samples_nominal == int(z["nominal"])
    q = z["rgba16"].astype(np.float32) / 65535.0
    frac, mx = scenes.parity_report(q, np.clip(np.nan_to_num(outs[0]), 0, 1))
    assert frac >= 0.9999 and mx < 1e-3, (frac, mx)        # libm differences between machines stay far below 1/255
    assert float((np.abs(depth - z["depth"]) <= 1e-5).mean()) >= 0.9999


def test_host_mirror_tf_texture_equals_the_oracle():
    """The product's TransFunc1DKeys::getTexture (host mirror, C++) against the oracle's restatement: split keys, thresholds,
    gamma, the three alpha modes, the widths the reference uses (256 for 8-bit data, 4096 for 16-bit, 1024 for the tables)."""
    from voreen_b200 import host
    rng = np.random.default_rng(11)
    key_sets = [
        [(0.0, (0, 0, 0, 0)), (1.0, (255, 255, 255, 255))],
        [(0.0, (0, 0, 0, 0)), (0.25, (0, 0, 0, 0)), (0.30, (255, 140, 60, 8)), (0.55, (255, 230, 200, 40)), (0.80, (255, 255, 255, 200)), (1.00, (255, 255, 255, 255))],
        [(0.1, (10, 20, 30, 40), (200, 100, 50, 25)), (0.5, (0, 255, 0, 128)), (0.5001, (255, 0, 0, 3), (1, 2, 3, 250)), (0.9, (9, 8, 7, 6))],     # split keys, keys not at 0/1
    ]
    for _ in range(6):
        n = int(rng.integers(2, 9))
        xs = np.sort(rng.random(n))
        key_sets.append([(float(x), tuple(int(v) for v in rng.integers(0, 256, 4))) for x in xs])
    checked = 0
    for keys in key_sets:
        for width in (256, 1024, 4096, 37):
            for thr in ((0.0, 1.0), (0.1, 0.9), (0.33, 0.34)):
                for gamma in (1.0, 0.5, 2.2):
                    for amode in (0, 1, 2):
                        want = orc.tf1d_texture(orc.make_tf(keys, threshold=thr, gamma=gamma, alpha_mode=amode), width)
                        got = host.tf1d_texture(keys, width, threshold=thr, gamma=gamma, alpha_mode=amode)
                        assert np.array_equal(got, want), (keys, width, thr, gamma, amode)
                        checked += 1
    assert checked == len(key_sets) * 4 * 3 * 3 * 3
