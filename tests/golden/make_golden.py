#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, so tests read the fixtures, never the reference).

Inputs (read where they lie, nothing is copied as source):
  V/resource/voreenve/volumes/nucleon.{vvd,raw}            the reference's bundled sample volume (41^3 uint8)
  V/resource/voreenve/workspaces/rendering-raycasting.vws  the reference's standard ray-casting workspace (config C1)
Outputs:
  nucleon.npz          voxels + meta data + the DERIVED DATA the reference itself computed and serialised into the .vvd:
                       VolumeMinMax, the 256-bucket intensity histogram and the 64x64 VolumePreview -- known answers of
                       reference code (V/src/core/datastructures/volume/{volumeminmax,histogram,volumepreview}.cpp,
                       VolumeRAM::getVoxelNormalizedLinear volumeram.cpp:78-97) that pin the oracle's voxel
                       normalisation and linear sampler on real reference data.
  nucleon_scene.json   camera, light, transfer-function keys and raycaster properties parsed from the workspace
  nucleon_oracle.npz   the oracle's rendering of that scene (RGBA16 + depth) -- regression fixture for the oracle and
                       for the CUDA path on the GPU box ("parity unpinned" at image level: the reference's own PNGs
                       live in the un-vendored voreen-testdata submodule).
"""
from __future__ import annotations

import base64
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
V = "/root/reference/voreen"


def parse_vvd(path):
    vol = ET.parse(path).getroot().find("Volumes/Volume")
    raw = vol.find("RawData")
    dims = tuple(int(raw.get(a)) for a in "xyz")
    meta = {m.get("name"): m for m in vol.findall("MetaData/MetaItem")}
    vec = lambda m: [float(m.find("value").get(a)) for a in "xyz"]
    out = {"dims": dims, "format": raw.get("format"), "raw": os.path.join(os.path.dirname(path), raw.get("filename")),
           "offset": vec(meta["Offset"]), "spacing": vec(meta["Spacing"])}
    for d in vol.findall("DerivedData/DerivedItem"):
        if d.get("type") == "VolumeMinMax":
            out["minmax"] = [float(d.get(k)) for k in ("min", "max", "minNorm", "maxNorm")]
        elif d.get("type") == "VolumePreview":
            out["preview_height"] = int(d.get("height"))
            out["preview"] = np.frombuffer(base64.b64decode(d.get("prevData")), np.uint8)
        elif d.find("histogram") is not None:
            h = d.find("histogram")
            out["hist_min"] = float(h.find("minValues/item").get("value"))
            out["hist_max"] = float(h.find("maxValues/item").get("value"))
            out["hist"] = np.array([int(i.get("value")) for i in h.findall("buckets/item")], np.int64)
    return out


def parse_workspace(path):
    root = ET.parse(path).getroot()
    procs = {p.get("name"): p for p in root.iter("Processor") if p.get("name") and p.find("Properties") is not None}
    rc = procs["SingleVolumeRaycaster"]
    props = {}
    scene = {"properties": props}
    for p in rc.find("Properties"):
        name = p.get("name")
        if name is None:
            continue
        if p.get("value") is not None and name not in ("renderingQuality", "interQuality"):
            props[name] = p.get("value")
        elif p.find("value") is not None and p.find("TransferFunction") is None:
            v = p.find("value")
            props[name] = [float(v.get(a)) for a in "xyzw" if v.get(a) is not None]
        if name == "camera":
            g = lambda tag: [float(p.find(tag).get(a)) for a in "xyz"]
            scene["camera"] = {"position": g("position"), "focus": g("focus"), "up": g("upVector"), "fovy": float(p.get("fovy")),
                               "near": float(p.get("frustNear")), "far": float(p.get("frustFar")),
                               "frustum_ratio": (float(p.get("frustRight")) - float(p.get("frustLeft"))) / (float(p.get("frustTop")) - float(p.get("frustBottom")))}
        if name == "transferFunction":
            tf = p.find("TransferFunction")
            keys = []
            for k in tf.find("Keys"):
                cl = k.find("colorL"); cr = k.find("colorR")
                col = lambda c: [int(c.get(a)) for a in "rgba"]
                keys.append({"intensity": float(k.get("intensity")), "split": k.get("split") == "true", "colorL": col(cl),
                             "colorR": col(cr) if cr is not None else col(cl)})
            scene["transferFunction"] = {"keys": keys, "domain": [float(tf.find("domain").get(a)) for a in "xy"],
                                         "threshold": [float(tf.find("threshold").get(a)) for a in "xy"],
                                         "gamma": float(tf.get("gammaValue")), "alphaMode": int(tf.get("alphaMode")),
                                         "width": int(tf.find("texDimensions").get("x"))}
    eep = procs["MeshEntryExitPoints"].find("Properties")
    scene["entryexit"] = {p.get("name"): p.get("value") for p in eep if p.get("value") is not None}
    pg = procs["OptimizedProxyGeometry"].find("Properties")
    scene["proxygeometry"] = {p.get("name"): p.get("value") for p in pg if p.get("value") is not None}
    cs = procs["Canvas"].find("Properties")
    for p in cs:
        if p.get("name") == "canvasSize":
            scene["canvasSize"] = [int(p.find("value").get(a)) for a in "xy"]
    return scene


def main():
    vvd = parse_vvd(f"{V}/resource/voreenve/volumes/nucleon.vvd")
    assert vvd["format"] == "uint8"
    dx, dy, dz = vvd["dims"]
    vox = np.fromfile(vvd["raw"], np.uint8).reshape(dz, dy, dx)
    np.savez_compressed(os.path.join(HERE, "nucleon.npz"), voxels=vox, offset=np.array(vvd["offset"], np.float32),
                        spacing=np.array(vvd["spacing"], np.float32), minmax=np.array(vvd["minmax"], np.float64),
                        hist=vvd["hist"], hist_range=np.array([vvd["hist_min"], vvd["hist_max"]], np.float64),
                        preview=vvd["preview"].reshape(vvd["preview_height"], -1))
    scene = parse_workspace(f"{V}/resource/voreenve/workspaces/rendering-raycasting.vws")
    scene["source"] = "V/resource/voreenve/workspaces/rendering-raycasting.vws + V/resource/voreenve/volumes/nucleon.vvd"
    with open(os.path.join(HERE, "nucleon_scene.json"), "w") as f:
        json.dump(scene, f, indent=1)
    # the oracle's rendering of that scene (fixture)
    import nucleon_scene
    outs, depth, st, eepi = nucleon_scene.oracle_render(nucleon_scene.load())
    q = np.clip(np.floor(np.nan_to_num(outs[0]) * 65535.0 + 0.5), 0, 65535).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "nucleon_oracle.npz"), rgba16=q, depth=depth.astype(np.float32),
                        rays=np.int64(st.rays_cast), nominal=np.int64(st.samples_nominal), classified=np.int64(st.samples_classified),
                        shaded=np.int64(st.samples_shaded), entry_rays=np.int64(np.any(eepi[0][..., :3] != eepi[2][..., :3], axis=2).sum()))
    print("wrote nucleon.npz, nucleon_scene.json, nucleon_oracle.npz; rays", st.rays_cast, "classified", st.samples_classified)


if __name__ == "__main__":
    main()
