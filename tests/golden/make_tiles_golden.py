"""Generates the tile-reader fixtures by RUNNING THE REFERENCE'S OWN
`sbmc/datasets.py` (imported unmodified from /root/reference) in this container:

    python tests/golden/make_tiles_golden.py     # writes tests/golden/tiles/

What is stubbed because it is absent from the image: `lz4.frame.decompress` ->
LZ4F_decompress of the system liblz4 (the C library the python `lz4` package
wraps, tests/tile_io.py) and `ttools.get_logger`.  The tiles themselves are
written with LZ4F_compressFrame(default preferences) of the same library, i.e.
byte-for-byte what the reference renderer's writer emits
(pbrt_patches/sbmc_pbrt.diff:6140-6158), in the layout its reader expects.

Fixture contents (tests/golden/tiles/):
  data/<scene>/*.bin, data/list.txt      two scenes of 2x2 tiles, 8x8 px, 3 spp
  expected.npz                           per configuration: sha256 of every output
                                         array that is pure data movement / exact
                                         arithmetic, the full arrays where numpy's
                                         log / reductions are involved.
"""
import hashlib
import importlib.util
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
OUT = os.path.join(HERE, "tiles")

from tests import tile_io  # noqa: E402


def import_reference_datasets():
    from sbmc_b200 import _compat
    ttools = types.ModuleType("ttools")
    ttools.get_logger = _compat.get_logger
    lz4 = types.ModuleType("lz4")
    frame = types.ModuleType("lz4.frame")
    frame.decompress = tile_io.decompress_frame
    lz4.frame = frame
    sys.modules.update({"ttools": ttools, "lz4": lz4, "lz4.frame": frame})
    spec = importlib.util.spec_from_file_location(
        "reference_datasets", os.path.join(REFERENCE, "sbmc", "datasets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


# configurations: name -> (class, path kind, kwargs)
CONFIGS = {
    "tiles_sbmc_all": ("TilesDataset", "folder", dict()),
    "tiles_sbmc_subset": ("TilesDataset", "list", dict(spp=2, load_coords=False, load_p=False)),
    "tiles_sbmc_nogbuf": ("TilesDataset", "folder", dict(spp=1, load_gbuffer=False, load_ld=False,
                                                         load_bt=False)),
    "tiles_raw": ("TilesDataset", "folder", dict(mode="raw")),
    "tiles_kpcn": ("TilesDataset", "folder", dict(mode="kpcn")),
    "full_sbmc_all": ("FullImagesDataset", "folder", dict()),
    "full_sbmc_spp2": ("FullImagesDataset", "folder", dict(spp=2, load_bt=False)),
    "full_kpcn": ("FullImagesDataset", "folder", dict(mode="kpcn", spp=2)),
}
LOG_KEYS = ("features",)      # contain np.log values in sbmc mode: stored, not hashed


def record(store, prefix, item, sbmc_mode, kpcn):
    for k, v in item.items():
        if k.startswith("_"):
            continue
        key = "%s/%s" % (prefix, k)
        if isinstance(v, np.ndarray):
            if kpcn and k.startswith("kpcn"):
                store[key] = v
            elif k in LOG_KEYS and sbmc_mode:
                store[key + "#sha_nolog"] = np.array(sha(drop_log(v, item)))
                store[key + "#log"] = log_channels(v, item)
            else:
                store[key + "#sha"] = np.array(sha(v))
        else:
            store[key] = np.array(v)


def drop_log(feats, item):
    i = item["_i_diffuse"]
    return np.concatenate([feats[:, :i], feats[:, i + 6:]], 1)


def log_channels(feats, item):
    i = item["_i_diffuse"]
    return feats[:, i:i + 6].copy()


def main():
    ref = import_reference_datasets()
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    data = os.path.join(OUT, "data")
    rng = np.random.default_rng(20190401)
    names = []
    for scene in ("scene_a", "scene_b"):
        tile_io.write_scene(data, scene, rng, ts=8, tiles_x=2, tiles_y=2, sample_count=3,
                            quantize=1.0 / 16, scene_radius=4.0 if scene == "scene_a" else 2.5,
                            aperture_radius=0.25 if scene == "scene_a" else 0.0,
                            focus_distance=1.5 if scene == "scene_a" else float("nan"))
        names += ["%s/%s" % (scene, f) for f in sorted(os.listdir(os.path.join(data, scene)))]
    with open(os.path.join(data, "list.txt"), "w") as fid:
        fid.write("\n".join(names[::-1]) + "\n")      # reversed: order comes from the list

    store = {}
    for name, (cls, kind, kw) in CONFIGS.items():
        path = data if kind == "folder" else os.path.join(data, "list.txt")
        dset = getattr(ref, cls)(path, **kw)
        mode = kw.get("mode", "sbmc")
        store["%s/len" % name] = np.array(len(dset))
        store["%s/num_features" % name] = np.array(dset.num_features)
        store["%s/num_global_features" % name] = np.array(dset.num_global_features)
        store["%s/repr" % name] = np.array(repr(dset))
        tiles = dset.tiles_dset if cls == "FullImagesDataset" else dset
        store["%s/labels" % name] = np.array("|".join(tiles.labels))
        for idx in range(len(dset)):
            item = dict(dset[idx])
            item.pop("path", None)
            item["_i_diffuse"] = tiles.labels.index("diffuse_r")
            record(store, "%s/%d" % (name, idx), item, mode == "sbmc", mode == "kpcn")
    multi = ref.MultiSampleCountDataset(data, spp=3)
    store["multi/len"] = np.array(len(multi))
    store["multi/spp_of_items"] = np.array([int(multi[i]["spp"].ravel()[0])
                                            for i in range(len(multi))])
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **store)
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(OUT) for f in fs)
    print("wrote %s: %d arrays, %d bytes in all" % (OUT, len(store), total))


if __name__ == "__main__":
    main()
