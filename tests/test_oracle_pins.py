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


# ---- the reference's own C++ (TransFunc1DKeys, TransFunc1D, PreIntegrationTable, VolumeRAM) compiled where it lies ----------------
# tests/golden/reference_pins.npz holds what oracle/_ref/voreen_probe printed (tests/golden/make_reference_pins.py); the probe itself
# is re-run live on further random inputs wherever the binary exists (this container, and the GPU box: it is a static binary).
VOREEN_PROBE = os.path.join(os.path.dirname(orc.PROBE), "voreen_probe")


def _pins():
    z = np.load(os.path.join(GOLDEN, "reference_pins.npz"))
    import json
    return z, json.loads(bytes(z["meta"]).decode())


def _keys_of(case):
    return [(k[0], tuple(k[1])) if k[2] is None else (k[0], tuple(k[1]), tuple(k[2])) for k in case["keys"]]


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64); b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a); b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def test_tf_texture_matches_the_reference_transfunc1dkeys():
    """TransFunc1D::createTextureData + TransFunc1DKeys::getMappingForValueUByte + TransFuncBase::applyAlpha, run from the reference's
    sources: the oracle's texels AND the product host mirror's texels are the same bytes (split keys, thresholds, gamma, alpha modes)."""
    z, meta = _pins()
    for case in meta["tf_cases"]:
        keys = _keys_of(case)
        want = z[f"tf_{case['name']}_texture"]
        tf = orc.make_tf(keys, domain=tuple(case["domain"]), threshold=tuple(case["threshold"]), gamma=case["gamma"], alpha_mode=case["alpha_mode"])
        got = orc.tf1d_texture(tf, case["width"])
        assert np.array_equal(got, want), case["name"]
        got_h = host.tf1d_texture(keys, case["width"], threshold=tuple(case["threshold"]), gamma=case["gamma"], alpha_mode=case["alpha_mode"])
        assert np.array_equal(got_h, want), ("host mirror", case["name"])
        # getMappingForValueUByte at arbitrary positions: a texture of the right width samples exactly x = i / (w - 1); check the
        # recorded mappings that fall on such positions of a 41-texel texture (gamma 1, full threshold)
        if case["gamma"] == 1.0 and case["threshold"] == [0.0, 1.0] and case["alpha_mode"] == 1:
            t41 = orc.tf1d_texture(tf, 41)
            xs, rgba = z[f"tf_{case['name']}_map_x"][:41], z[f"tf_{case['name']}_map_rgba"][:41]
            assert np.array_equal(xs, np.linspace(0, 1, 41, dtype=np.float32))
            same = np.float32(np.arange(41, dtype=np.float32) / np.float32(40)) == xs         # i / (w - 1) in float == the linspace value
            assert same.sum() >= 30 and np.array_equal(t41[same], rgba[same]), case["name"]


def test_preintegration_table_matches_the_reference_computeTable():
    """PreIntegrationTable::computeTable (both modes) and ::classify from the reference's sources vs orc_preint_cpu / orc_preint_classify."""
    z, meta = _pins()
    tfs = {c["name"]: c for c in meta["tf_cases"]}
    worst = 0
    for pc in meta["pi_cases"]:
        case = tfs[pc["tf"]]
        tf = orc.make_tf(_keys_of(case), domain=tuple(case["domain"]), threshold=tuple(case["threshold"]), gamma=case["gamma"], alpha_mode=case["alpha_mode"])
        want = z[f"pi_{pc['index']}_table"]
        got = orc.preint_cpu(tf, pc["res"], pc["step"], bool(pc["integral"]))
        assert np.isfinite(want).all() and np.isfinite(got).all()
        u = int(_ulp_diff(got, want).max())
        worst = max(worst, u)
        assert u <= 1, (pc, u, float(np.abs(got - want).max()))          # same libm, same operation order: bit-identical in practice
        fs = z[f"pi_{pc['index']}_classify_in"]; out = z[f"pi_{pc['index']}_classify_out"]
        for (a, b), w in zip(fs, out):
            c = np.zeros(4, np.float32)
            orc.lib().orc_preint_classify(orc.fp(np.ascontiguousarray(want)), pc["res"], float(a), float(b), orc.fp(c))
            assert np.array_equal(c, w)
    print(f"pre-integration tables vs the reference's computeTable: worst difference {worst} ulp over {len(meta['pi_cases'])} tables")


def test_voxel_access_matches_the_reference_volumeram():
    """VolumeAtomic<T>::getVoxelNormalized is the sampler's texel normalisation (bit-exact); VolumeRAM::getVoxelNormalizedLinear
    (volumeram.cpp:78-97) is a software trilinear filter in voxel coordinates written as a sum of eight weighted taps -- the GL-spec
    nested-lerp form the oracle uses differs from it by rounding only."""
    z, meta = _pins()
    for vc in meta["vol_cases"]:
        i = vc["index"]
        data = z[f"vol_{i}_data"]
        dims = tuple(vc["dims"])
        dt = {"uint8": abi.U8, "uint16": abi.U16, "float": abi.F32}[vc["format"]]
        desc = rc.make_volume_desc(dims, dtype=dt)
        ov = orc.make_volume(data, desc)
        d = np.array(dims, np.float32)
        for (x, y, zc), want in zip(z[f"vol_{i}_voxel_in"], z[f"vol_{i}_voxel_out"]):
            p = ((np.array([x, y, zc], np.float32) + np.float32(0.5)) / d).astype(np.float32)      # texel centre -> the texel itself
            out = np.zeros(4, np.float32)
            orc.lib().orc_sample3d(C_byref(ov), abi.FILTER_NEAREST, abi.CLAMP_TO_EDGE, 0.0, orc.fp(p), orc.fp(out))
            # VolumeAtomic<uint16_t>::getVoxelNormalized divides by 65535.f; the GL path the oracle restates multiplies by the float
            # constant 1/65535 (llvmpipe lp_build_unsigned_norm_to_float; GL 4.x 2.3.5 allows either): at most one ulp apart
            assert int(_ulp_diff(out[:1], np.array([want], np.float32)).max()) <= (0 if vc["format"] == "float" else 1), (vc, x, y, zc)
        scale = 1.0 if vc["format"] == "float" else 1.0
        errs = []
        for pos, want in zip(z[f"vol_{i}_linear_in"], z[f"vol_{i}_linear_out"]):
            p = ((pos + np.float32(0.5)) / d).astype(np.float32)
            out = np.zeros(4, np.float32)
            orc.lib().orc_sample3d(C_byref(ov), abi.FILTER_LINEAR, abi.CLAMP_TO_EDGE, 0.0, orc.fp(p), orc.fp(out))
            errs.append(abs(float(out[0]) - float(want)))
        # texture coordinate -> u = p * dim - 0.5 re-derives the voxel position with ~1e-6 relative error; the filter itself is linear
        assert max(errs) <= 4e-6 * scale, (vc, max(errs))


def C_byref(x):
    import ctypes
    return ctypes.byref(x)


def test_reference_probe_live_random_transfer_functions():
    """Where the probe binary exists: random key sets (incl. split keys, coinciding keys, thresholds, gamma) straight against the
    reference's TransFunc1DKeys and PreIntegrationTable."""
    if not os.path.exists(VOREEN_PROBE):
        pytest.skip("oracle/_ref/voreen_probe not built (needs /root/reference at build time)")
    import sys
    sys.path.insert(0, GOLDEN)
    import make_reference_pins as mk
    rng = np.random.default_rng(99)
    for trial in range(25):
        n = int(rng.integers(1, 9))
        xs = np.sort(rng.random(n).astype(np.float32))
        if n > 2 and trial % 3 == 0:
            xs[1] = xs[0]                                   # coinciding keys
        keys = []
        for x in xs:
            l = tuple(int(v) for v in rng.integers(0, 256, 4))
            r = tuple(int(v) for v in rng.integers(0, 256, 4)) if rng.random() < 0.3 else None
            keys.append((float(x), l, r))
        w = int(rng.choice([64, 256, 1024]))
        thr = (float(np.float32(rng.random() * 0.3)), float(np.float32(0.7 + rng.random() * 0.3))) if trial % 2 else (0.0, 1.0)
        gamma = float(np.float32(rng.choice([1.0, 0.5, 2.2])))
        am = int(rng.choice([0, 1, 2]))
        case = (f"r{trial}", w, keys, thr, gamma, am, (0.0, 1.0))
        res_ = int(rng.choice([32, 64]))
        step = float(np.float32(rng.choice([0.002, 0.01, 0.25, 1.0])))
        integral = int(trial % 4 == 3)
        res = mk.run(mk.tf_lines(case) + ["texture", f"preint {res_} {mk.f32(step)} {integral}"])
        want_tex = np.array([int(t) for t in res[1].split()], np.uint8).reshape(-1, 4)
        want_pi = mk.hexf(res[2]).reshape(res_, res_, 4)
        # TransFunc1DKeys::addKey (transfunc1dkeys.cpp:197-210) inserts a key BEFORE the first key whose intensity is not smaller:
        # coinciding keys end up in reverse order of insertion; the oracle takes the final key vector
        ordered = []
        for k in keys:
            j = 0
            while j < len(ordered) and k[0] > ordered[j][0]:
                j += 1
            ordered.insert(j, k)
        okeys = [(k[0], k[1]) if k[2] is None else (k[0], k[1], k[2]) for k in ordered]
        tf = orc.make_tf(okeys, domain=(0.0, 1.0), threshold=thr, gamma=gamma, alpha_mode=am)
        assert np.array_equal(orc.tf1d_texture(tf, w), want_tex), case
        assert np.array_equal(host.tf1d_texture(okeys, w, threshold=thr, gamma=gamma, alpha_mode=am), want_tex), ("host", case)
        got_pi = orc.preint_cpu(tf, res_, step, bool(integral))
        assert int(_ulp_diff(got_pi, want_pi).max()) <= 1, (case, res_, step, integral)


def _orc_gaussian_texture(curves, width, threshold, gamma, alpha_mode):
    import ctypes as C

    class Curve(C.Structure):
        _fields_ = [("intensity", C.c_float), ("width", C.c_float), ("base_value", C.c_float), ("color_l", C.c_uint8 * 4), ("color_r", C.c_uint8 * 4),
                    ("base_color_l", C.c_uint8 * 4), ("base_color_r", C.c_uint8 * 4), ("active", C.c_int32)]
    arr = (Curve * max(1, len(curves)))()
    for i, c in enumerate(curves):
        arr[i].intensity, arr[i].width, arr[i].base_value, arr[i].active = c["intensity"], c["width"], c["base_value"], 1 if c["active"] else 0
        arr[i].color_l[:] = c["color_l"]; arr[i].color_r[:] = c["color_r"]; arr[i].base_color_l[:] = c["base_color_l"]; arr[i].base_color_r[:] = c["base_color_r"]
    out = np.zeros((width, 4), np.uint8)
    thr = (C.c_float * 2)(*threshold)
    fn = orc.lib().orc_tf1d_gaussian_texture
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_int, C.c_int, C.c_void_p]
    fn.restype = None
    fn(C.cast(arr, C.c_void_p), len(curves), thr, gamma, alpha_mode, width, out.ctypes.data)
    return out


def test_gaussian_tf_texture_matches_the_reference_transfunc1dgaussian():
    """TransFunc1DGaussian::getMappingForValueUByte over TransFuncMappingCurve::getColorAt / getOpacityAt, run from the reference's sources:
    the oracle's and the product host mirror's texels are the same bytes; live random curve sets where the probe binary exists."""
    z, meta = _pins()
    cases = [(c["curves"], c["width"], tuple(c["threshold"]), c["gamma"], c["alpha_mode"], z[f"gauss_{c['name']}_texture"], c["name"]) for c in meta["gauss_cases"]]
    if os.path.exists(VOREEN_PROBE):
        import sys
        sys.path.insert(0, GOLDEN)
        import make_reference_pins as mk
        rng = np.random.default_rng(5)
        for trial in range(20):
            curves = [mk._curve(rng.random(), float(rng.choice([0.0005, 0.005, 0.02, 0.2])), float(rng.choice([0.0, 0.0, 0.1, 0.3])), tuple(int(v) for v in rng.integers(0, 256, 4)),
                                tuple(int(v) for v in rng.integers(0, 256, 4)) if rng.random() < 0.4 else None, tuple(int(v) for v in rng.integers(0, 256, 3)),
                                tuple(int(v) for v in rng.integers(0, 256, 3)), active=rng.random() < 0.85) for _ in range(int(rng.integers(1, 5)))]
            w = int(rng.choice([64, 256, 1024]))
            thr = (0.0, 1.0) if trial % 2 else (float(np.float32(0.1)), float(np.float32(0.95)))
            gamma = float(np.float32(rng.choice([1.0, 0.5, 2.0])))
            am = int(rng.choice([0, 1, 2]))
            res = mk.run(mk.gauss_lines((f"r{trial}", w, curves, thr, gamma, am)))
            cases.append((curves, w, thr, gamma, am, np.array([int(t) for t in res[0].split()], np.uint8).reshape(-1, 4), f"live{trial}"))
    for curves, w, thr, gamma, am, want, name in cases:
        got = _orc_gaussian_texture(curves, w, thr, gamma, am)
        assert np.array_equal(got, want), (name, np.argwhere(got != want)[:5], got[got != want][:5], want[got != want][:5])
        got_h = host.tf1d_gaussian_texture(curves, w, threshold=thr, gamma=gamma, alpha_mode=am)
        assert np.array_equal(got_h, want), ("host mirror", name)


def test_outer_faces_mesh_gives_the_entry_exit_of_the_brick_cubes():
    """OptimizedProxyGeometry "outerfaces" (optimizedproxygeometry.cpp:1392-1604) emits the cell faces between a non-empty and an empty
    cell (or the volume border).  MeshEntryExitPoints takes the nearest front face and the farthest back face per ray, so that mesh
    yields the same entry/exit points as the non-empty cells rendered as cubes ("visiblebricks", what B200RC_PROXY_OUTER_FACES emits):
    checked here on random occupancy grids with rays in float64 -- faces as rectangles against cubes as slabs."""
    rng = np.random.default_rng(12)
    for trial in range(6):
        n = (6, 5, 7)
        occ = rng.random(n) < (0.25 + 0.1 * trial)
        occ[2:4, 1:3, 2:5] = True
        faces = []          # (axis, plane coordinate, lo (2), hi (2), outward sign)
        for idx in np.argwhere(occ):
            for a in range(3):
                for sgn in (-1, 1):
                    nb = idx.copy(); nb[a] += sgn
                    if nb[a] < 0 or nb[a] >= n[a] or not occ[tuple(nb)]:
                        o = [b for b in range(3) if b != a]
                        faces.append((a, idx[a] + (1 if sgn > 0 else 0), (idx[o[0]], idx[o[1]]), (idx[o[0]] + 1, idx[o[1]] + 1), sgn))
        cubes = [(idx.astype(float), idx.astype(float) + 1.0) for idx in np.argwhere(occ)]
        for _ in range(300):
            org = rng.normal(size=3); org = org / np.linalg.norm(org) * 12.0 + np.array(n) / 2
            tgt_ = rng.random(3) * np.array(n)
            d = tgt_ - org; d /= np.linalg.norm(d)
            t_in, t_out = np.inf, -np.inf
            for lo, hi in cubes:
                with np.errstate(divide="ignore", invalid="ignore"):
                    t0 = (lo - org) / d; t1 = (hi - org) / d
                tn = np.max(np.minimum(t0, t1)); tf = np.min(np.maximum(t0, t1))
                if tn <= tf and tf > 0:
                    t_in = min(t_in, tn); t_out = max(t_out, tf)
            f_in, f_out = np.inf, -np.inf
            for a, plane, lo, hi, sgn in faces:
                if d[a] == 0:
                    continue
                t = (plane - org[a]) / d[a]
                p = org + t * d
                o = [b for b in range(3) if b != a]
                if t > 0 and lo[0] <= p[o[0]] <= hi[0] and lo[1] <= p[o[1]] <= hi[1]:
                    if d[a] * sgn < 0:
                        f_in = min(f_in, t)         # front face: outward normal against the ray
                    else:
                        f_out = max(f_out, t)
            assert np.isinf(t_in) == np.isinf(f_in)
            if np.isfinite(t_in):
                assert abs(t_in - f_in) < 1e-9 and abs(t_out - f_out) < 1e-9


# ---- TransFunc2DPrimitives: host mirror against the oracle's restatement, plus what the reference's paint code
#      implies for specific texels.  The rasteriser itself is OpenGL in the reference: PARITY UNPINNED (no reference bit pattern exists) ---
TF2D_PRIMS = [
    {"type": "quad", "points": [(0.30, 0.10), (0.70, 0.10), (0.70, 0.60), (0.30, 0.60)], "colors": (255, 128, 0, 200), "fuzziness": 0.5},
    {"type": "triangle", "points": [(0.05, 0.55), (0.45, 0.95), (0.10, 0.98)], "colors": [(0, 255, 0, 255), (0, 0, 255, 128), (255, 0, 255, 64)], "fuzziness": 0.7},
    {"type": "banana", "points": [(0.55, 0.65), (0.75, 0.97), (0.75, 0.80), (0.97, 0.66)], "colors": (20, 200, 250, 180), "fuzziness": 0.6},
    {"type": "quad", "points": [(0.0, 0.0), (0.2, 0.0), (0.2, 0.2), (0.0, 0.2)], "colors": (255, 255, 255, 255), "fuzziness": 1.0},
]


@pytest.mark.parametrize("size", [(256, 256), (128, 96), (33, 57)])
@pytest.mark.parametrize("variant", ["plain", "thresholds+gamma", "one-alpha", "zero-alpha"])
def test_tf2d_primitives_texture(size, variant):
    W, H = size
    kw = {"plain": {}, "thresholds+gamma": {"threshold_x": (0.1, 0.8), "threshold_y": (0.05, 0.9), "gamma": (1.5, 0.7)}, "one-alpha": {"alpha_mode": 2},
          "zero-alpha": {"alpha_mode": 0}}[variant]
    f, u = host.tf2d_primitives_texture(TF2D_PRIMS, W, H, **kw)
    want = orc.tf2d_primitives_texture(TF2D_PRIMS, W, H, **kw)
    assert np.array_equal(f, want)
    assert np.array_equal(u, np.clip(np.floor(want * np.float32(255.0) + np.float32(0.5)), 0, 255).astype(np.uint8))
    if variant == "zero-alpha":
        assert not want[..., 3].any() and want[..., :3].any()
        return
    if variant == "one-alpha":
        assert set(np.unique(want[..., 3])) <= {0.0, 1.0}
        return
    if variant == "plain":
        def texel(x, y):
            return want[min(int(y * H), H - 1), min(int(x * W), W - 1)]
        # the core of the first quad (inside its inner quad: fuzziness 0.5 -> [0.4, 0.6] x [0.225, 0.475]) carries the full colour
        assert np.allclose(texel(0.5, 0.35), np.array([255, 128, 0, 200]) / 255.0, atol=1e-6)
        # its faded border: full rgb, alpha strictly between 0 and the full value, growing towards the core
        a_out, a_in = texel(0.33, 0.35)[3], texel(0.38, 0.35)[3]
        assert 0.0 < a_out < a_in < 200 / 255.0 and np.allclose(texel(0.33, 0.35)[:3], np.array([255, 128, 0]) / 255.0, atol=1e-6)
        # nothing painted between the primitives
        assert not texel(0.85, 0.3).any()
        # fuzziness 1: no border, the whole quad is opaque white
        assert np.allclose(texel(0.1, 0.1), 1.0, atol=1e-6) and np.allclose(texel(0.01, 0.19), 1.0, atol=1e-6)
        # the triangle interpolates its vertex colours: green dominates near the first vertex, blue near the second
        g, b = texel(0.12, 0.66), texel(0.36, 0.9)
        assert g[1] > g[2] and b[2] > b[1]
        # the banana: opaque along its centre line between the two middle control points, transparent above the upper curve
        assert texel(0.75, 0.86)[3] > 0.5 and texel(0.75, 0.99)[3] == 0.0
    if variant == "thresholds+gamma":
        xf, xb = int(math.floor(0.1 * W + 0.5)), int(math.floor(0.8 * W + 0.5))
        yf, yb = int(math.floor(0.05 * H + 0.5)), int(math.floor(0.9 * H + 0.5))
        assert not want[:, :xf].any() and not want[:, xb + 1:].any() and not want[:yf].any() and not want[yb + 1:].any()
        assert want[yf:yb + 1, xf:xb + 1, 3].any()


def test_tf2d_primitives_invalid_input():
    with pytest.raises(abi.B200RCError):
        host.tf2d_primitives_texture(TF2D_PRIMS, 1, 64)
    lib = host.load_library()
    arr = host.make_primitives(TF2D_PRIMS[:1])
    arr[0].type = 7
    import ctypes as C
    thr, g = (C.c_float * 2)(0.0, 1.0), (C.c_float * 2)(1.0, 1.0)
    out = np.zeros((8, 8, 4), np.float32)
    assert lib.b200host_tf2d_primitives_texture(arr, 1, 8, 8, thr, thr, g, 1, out.ctypes.data_as(C.POINTER(C.c_float)), None) != 0
    f, u = host.tf2d_primitives_texture([], 16, 16)       # no primitives: a transparent texture
    assert not f.any() and not u.any()
