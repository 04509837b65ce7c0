"""numpy restatement of the reference's tile reader -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py for who may import this).

Follows sbmc/datasets.py: header :504-520 (+ value checks :522-538), chunk
reader :570-579, `_read_data` :581-739 (pixel statistics :592-606, per-sample
planes :625-698, channel selection :700-711, radiance / low_spp :716-729),
`_preprocess_standard` :744-778 and the tile pasting of
`FullImagesDataset.__getitem__` :920-957.  LZ4 frames are inflated by
oracle/lz4_oracle.c.

PARITY PIN: tests/golden/tiles/ holds tiles written with the real liblz4 plus
the outputs of the reference's own, unmodified `sbmc/datasets.py` on them
(tests/golden/make_tiles_golden.py); tests/test_tiles.py checks this
restatement against those outputs bit for bit.
"""
import struct

import numpy as np

from . import lz4_frame_decompress

DEPTH = 6
GLABELS = ("aperture_radius", "focus_distance", "fov")


def read_header(buf):
    vals = struct.unpack_from("<9i4f", buf, 0)
    meta = dict(zip(("version", "tile_size", "image_width", "image_height", "sample_count",
                     "gt_sample_count", "sample_features", "pixel_features", "path_depth"),
                    vals[:9]))
    g = dict(zip(("focus_distance", "aperture_radius", "fov", "scene_radius"), vals[9:]))
    if g["aperture_radius"] == 0:
        g["focus_distance"] = 0.0
    return meta, g


def read_tile(buf, spp=None, load_coords=True, load_gbuffer=True, load_p=True, load_ld=True,
              load_bt=True, log_radiance=True):
    """bytes of one .bin tile -> dict of numpy arrays, as `TilesDataset[...]` in
    sbmc mode (log_radiance=True) or raw mode (False, with the raw-mode flag
    overrides left to the caller)."""
    meta, g = read_header(buf)
    ts, nsf, depth = meta["tile_size"], meta["sample_features"], meta["path_depth"]
    spp = meta["sample_count"] if spp is None else spp
    pos = 52
    block_x, block_y = struct.unpack_from("<2i", buf, pos)
    pos += 8

    def chunk():
        nonlocal pos
        (n,) = struct.unpack_from("<i", buf, pos)
        raw = lz4_frame_decompress(buf[pos + 4:pos + 4 + n])
        pos += 4 + n
        return raw

    image = np.frombuffer(chunk(), np.float32).reshape(meta["pixel_features"], ts, ts)
    nch = image.shape[0] // 2
    out = {"block_x": block_x, "block_y": block_y,
           "global_features": np.array([g[k] for k in GLABELS], np.float32).reshape(3, 1, 1),
           "image_data": image[:nch], "image_data_var": image[nch:2 * nch],
           "target_image": image[:3] + image[3:6],
           "spp": spp * np.ones((1, 1, 1), np.int32), "scene_radius": g["scene_radius"]}
    if spp <= 0:
        out["low_spp"] = np.zeros(out["target_image"].shape)
        return out

    nfloat = nsf + 6 * depth
    planes = []
    for _ in range(spp):
        raw = chunk()
        fl = np.frombuffer(raw, np.float32, count=nfloat * ts * ts).reshape(nfloat, ts, ts)
        bits = np.frombuffer(raw, np.int16, count=depth * ts * ts,
                             offset=nfloat * ts * ts * 4).reshape(depth, ts, ts)
        keep = []
        if load_coords:
            keep.append(fl[0:5])
        keep.append(fl[5:11])
        if load_gbuffer:
            keep.append(fl[11:27])
        if load_p:
            keep.append(fl[nsf:nsf + 4 * depth])
        if load_ld:
            keep.append(fl[nsf + 4 * depth:nsf + 6 * depth])
        if load_bt:
            for bit in range(5):          # reflection, transmission, diffuse, glossy, specular
                keep.append(((bits & (1 << bit)) != 0).astype(np.float32))
        planes.append(np.concatenate(keep, 0))
    feats = np.stack(planes, 0)
    i = 5 if load_coords else 0
    diffuse, specular = feats[:, i:i + 3], feats[:, i + 3:i + 6]
    out["radiance"] = diffuse + specular
    out["low_spp"] = out["radiance"].mean(0)
    if log_radiance:
        d = np.maximum(diffuse, 0)
        s = np.maximum(specular, 0)
        feats = feats.copy()
        feats[:, i:i + 3] = np.log(1 + (d + s)) / 10.0
        feats[:, i + 3:i + 6] = np.log(1 + s) / 10.0
    out["features"] = feats
    return out


def read_image(tile_bufs, **kw):
    """All tiles of one scene pasted at (block_y, block_x), as
    `FullImagesDataset[...]` (sbmc / raw mode)."""
    tiles = [read_tile(b, **kw) for b in tile_bufs]
    meta, _ = read_header(tile_bufs[0])
    h, w, ts = meta["image_height"], meta["image_width"], meta["tile_size"]
    first = tiles[0]
    out = {"global_features": first["global_features"], "scene_radius": first["scene_radius"]}
    keys = [k for k, v in first.items() if isinstance(v, np.ndarray) and k != "global_features"]
    for k in keys:
        out[k] = np.zeros(first[k].shape[:-2] + (h, w), first[k].dtype)
    for t in tiles:
        for k in keys:
            out[k][..., t["block_y"]:t["block_y"] + ts, t["block_x"]:t["block_x"] + ts] = t[k]
    return out
