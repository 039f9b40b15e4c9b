#!/usr/bin/env python
"""Generates tests/golden/reference_pins.npz: outputs of the REFERENCE'S OWN C++ code for the transfer-function, pre-integration and
voxel-access arithmetic of the ray-casting path, produced by oracle/_ref/voreen_probe (the reference's sources compiled where they
lie, see oracle/voreen_probe.cpp / oracle/Makefile `ref`).  Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_reference_pins.py

The committed fixture lets the pins run where /root/reference does not exist (the GPU box); tests/test_oracle_pins.py re-runs the
probe live, on more inputs, whenever the binary is present.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PROBE = os.path.join(ROOT, "oracle", "_ref", "voreen_probe")

# (name, width, keys [(intensity, left rgba, right rgba or None)], threshold, gamma, alpha_mode, domain)
CT_KEYS = [(0.00, (0, 0, 0, 0), None), (0.25, (0, 0, 0, 0), None), (0.30, (255, 140, 60, 8), None), (0.55, (255, 230, 200, 40), None),
           (0.80, (255, 255, 255, 200), None), (1.00, (255, 255, 255, 255), None)]
TF_CASES = [
    ("ct4096", 4096, CT_KEYS, (0.0, 1.0), 1.0, 1, (0.0, 65535.0)),
    ("ct1024", 1024, CT_KEYS, (0.0, 1.0), 1.0, 1, (0.0, 65535.0)),
    ("ct256_thr", 256, CT_KEYS, (0.2, 0.9), 1.0, 1, (0.0, 255.0)),
    ("gamma", 1024, CT_KEYS, (0.05, 0.97), 2.2, 1, (0.0, 65535.0)),
    ("split", 512, [(0.0, (10, 20, 30, 0), None), (0.2, (200, 10, 90, 0), (0, 255, 0, 120)), (0.5, (255, 255, 0, 255), (3, 2, 1, 4)),
                    (0.5001, (9, 8, 7, 6), None), (0.9, (255, 0, 255, 77), None)], (0.0, 1.0), 1.0, 1, (-100.0, 300.0)),
    ("one_alpha", 256, CT_KEYS, (0.0, 1.0), 1.0, 2, (0.0, 1.0)),
    ("zero_alpha", 256, CT_KEYS, (0.0, 1.0), 0.5, 0, (0.0, 1.0)),
    ("nucleon", 256, [(0.0, (0, 0, 0, 0), None), (0.0784, (0, 0, 0, 0), None), (0.1255, (26, 77, 128, 6), None), (0.2, (204, 128, 51, 24), None),
                      (0.4, (255, 255, 128, 110), None), (0.7, (255, 255, 255, 255), None), (1.0, (255, 255, 255, 255), None)], (0.0, 1.0), 1.0, 1, (0.0, 249.0)),
    ("single_key", 64, [(0.4, (1, 2, 3, 4), None)], (0.0, 1.0), 1.0, 1, (0.0, 1.0)),
    ("opaque", 128, [(0.0, (0, 0, 0, 0), None), (0.3, (0, 0, 0, 0), None), (0.32, (255, 200, 100, 255), None), (1.0, (255, 255, 255, 255), None)], (0.0, 1.0), 1.0, 1, (0.0, 1.0)),
]
# pre-integration tables: (tf case, resolution, sampling step, use_integral)
PI_CASES = [("ct1024", 256, 1.0 / 4096.0, 0), ("ct1024", 128, 1.0 / 1024.0, 1), ("ct256_thr", 96, 1.0 / 96.0, 0), ("gamma", 64, 0.004, 0), ("gamma", 64, 0.004, 1),
            ("split", 80, 0.01, 0), ("one_alpha", 48, 0.02, 0), ("zero_alpha", 48, 0.02, 0), ("nucleon", 256, 1.0 / 164.0, 0), ("opaque", 128, 1.0 / 96.0, 0),
            ("opaque", 128, 1.0 / 96.0, 1), ("nucleon", 256, 1.0, 0)]


# Gaussian transfer functions: (name, width, curves, threshold, gamma, alpha_mode)
def _curve(x, w, base, cl, cr=None, bl=None, br=None, active=True):
    return {"intensity": float(np.float32(x)), "width": float(np.float32(w)), "base_value": float(np.float32(base)), "color_l": list(cl), "color_r": list(cr or cl),
            "base_color_l": list(bl or cl)[:3] + [0], "base_color_r": list(br or cr or cl)[:3] + [0], "active": bool(active)}


GAUSS_CASES = [
    ("g_single", 256, [_curve(0.5, 0.01, 0.0, (255, 128, 0, 200))], (0.0, 1.0), 1.0, 1),
    ("g_two", 1024, [_curve(0.3, 0.004, 0.0, (255, 0, 0, 255), (200, 40, 0, 120), (20, 20, 90), (0, 0, 0)), _curve(0.7, 0.02, 0.1, (0, 255, 30, 100), None, (9, 9, 9))], (0.0, 1.0), 1.0, 1),
    ("g_inactive", 128, [_curve(0.2, 0.01, 0.05, (10, 200, 250, 180)), _curve(0.6, 0.05, 0.0, (250, 250, 250, 250), active=False), _curve(0.8, 0.001, 0.0, (90, 60, 30, 255))], (0.1, 0.9), 2.0, 1),
    ("g_one_alpha", 64, [_curve(0.5, 0.03, 0.0, (255, 255, 255, 128))], (0.0, 1.0), 1.0, 2),
    ("g_empty", 32, [], (0.0, 1.0), 1.0, 1),
]


def gauss_lines(case):
    name, w, curves, thr, gamma, am = case
    out = [f"gauss {w} {len(curves)} {thr[0]!r} {thr[1]!r} {gamma!r} {am}"]
    for c in curves:
        v = c["color_l"] + c["color_r"] + c["base_color_l"][:3] + c["base_color_r"][:3]
        out.append(f"{c['intensity']!r} {c['width']!r} {c['base_value']!r} " + " ".join(str(int(t)) for t in v) + f" {1 if c['active'] else 0}")
    return out


def tf_lines(case):
    name, w, keys, thr, gamma, am, dom = case
    out = [f"tf {w} {len(keys)} {thr[0]!r} {thr[1]!r} {gamma!r} {am} {dom[0]!r} {dom[1]!r}"]
    for x, l, r in keys:
        rr = r if r is not None else (0, 0, 0, 0)
        out.append(f"{float(np.float32(x))!r} {l[0]} {l[1]} {l[2]} {l[3]} {rr[0]} {rr[1]} {rr[2]} {rr[3]} {1 if r is not None else 0}")
    return out


def run(lines):
    r = subprocess.run([PROBE], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    return r.stdout.strip().split("\n")


def hexf(line):
    return np.array([float.fromhex(t) for t in line.split()], np.float32)


def f32(v):
    return "%.9g" % float(np.float32(v))


def main():
    if not os.path.exists(PROBE):
        sys.exit("oracle/_ref/voreen_probe missing: make -C oracle ref (needs /root/reference)")
    rng = np.random.default_rng(20261018)
    out = {}
    meta = {"tf_cases": [], "pi_cases": [], "vol_cases": []}
    for case in TF_CASES:
        name = case[0]
        xs = np.concatenate([np.linspace(0, 1, 41, dtype=np.float32), rng.random(60, dtype=np.float32), np.array([k[0] for k in case[2]], np.float32)])
        lines = tf_lines(case) + ["texture"] + [f"mapping {f32(x)}" for x in xs]
        res = run(lines)
        assert res[0].startswith("ok")
        out[f"tf_{name}_texture"] = np.array([int(t) for t in res[1].split()], np.uint8).reshape(-1, 4)
        out[f"tf_{name}_map_x"] = xs
        out[f"tf_{name}_map_rgba"] = np.array([[int(t) for t in ln.split()] for ln in res[2:]], np.uint8)
        meta["tf_cases"].append({"name": name, "width": case[1], "keys": [[float(np.float32(k[0])), list(k[1]), list(k[2]) if k[2] else None] for k in case[2]],
                                 "threshold": list(case[3]), "gamma": case[4], "alpha_mode": case[5], "domain": list(case[6])})
    by_name = {c[0]: c for c in TF_CASES}
    for i, (name, res_, step, integral) in enumerate(PI_CASES):
        fs = rng.random((40, 2), dtype=np.float32)
        lines = tf_lines(by_name[name]) + [f"preint {res_} {f32(step)} {integral}"] + [f"classify {f32(a)} {f32(b)}" for a, b in fs]
        res = run(lines)
        out[f"pi_{i}_table"] = hexf(res[1]).reshape(res_, res_, 4)
        out[f"pi_{i}_classify_in"] = fs
        out[f"pi_{i}_classify_out"] = np.stack([hexf(ln) for ln in res[2:]])
        meta["pi_cases"].append({"index": i, "tf": name, "res": res_, "step": float(np.float32(step)), "integral": integral})
    for i, (fmt, dims) in enumerate([("uint16", (5, 4, 3)), ("uint8", (3, 6, 4)), ("float", (4, 4, 4)), ("uint16", (2, 2, 2))]):
        n = dims[0] * dims[1] * dims[2]
        if fmt == "uint16":
            data = rng.integers(0, 65536, n).astype(np.uint16); data[:3] = [0, 65535, 32768]
            txt = " ".join(str(int(v)) for v in data)
        elif fmt == "uint8":
            data = rng.integers(0, 256, n).astype(np.uint8); data[:2] = [0, 255]
            txt = " ".join(str(int(v)) for v in data)
        else:
            data = rng.random(n, dtype=np.float32)
            txt = " ".join(f32(v) for v in data)
        d = np.array(dims, np.float32)
        pos = np.concatenate([rng.random((60, 3), dtype=np.float32) * (d - 1), rng.integers(0, 2, (12, 3)).astype(np.float32) * (d - 1),
                              (rng.integers(0, 4, (12, 3)) % np.array(dims)).astype(np.float32)]).astype(np.float32)
        vox = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])][:40])
        lines = [f"vol {fmt} {dims[0]} {dims[1]} {dims[2]}", txt] + [f"voxel {x} {y} {z}" for x, y, z in vox] + [f"linear {f32(p[0])} {f32(p[1])} {f32(p[2])}" for p in pos]
        res = run(lines)
        out[f"vol_{i}_data"] = data.reshape(dims[::-1])
        out[f"vol_{i}_voxel_in"] = vox
        out[f"vol_{i}_voxel_out"] = np.array([float.fromhex(t) for t in res[1:1 + len(vox)]], np.float32)
        out[f"vol_{i}_linear_in"] = pos
        out[f"vol_{i}_linear_out"] = np.array([float.fromhex(t) for t in res[1 + len(vox):]], np.float32)
        meta["vol_cases"].append({"index": i, "format": fmt, "dims": list(dims)})
    meta["gauss_cases"] = []
    for case in GAUSS_CASES:
        res = run(gauss_lines(case))
        out[f"gauss_{case[0]}_texture"] = np.array([int(t) for t in res[0].split()], np.uint8).reshape(-1, 4)
        meta["gauss_cases"].append({"name": case[0], "width": case[1], "curves": case[2], "threshold": list(case[3]), "gamma": case[4], "alpha_mode": case[5]})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(HERE, "reference_pins.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(TF_CASES)} TFs, {len(PI_CASES)} tables, 4 volumes")


if __name__ == "__main__":
    main()
