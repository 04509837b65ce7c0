"""Test helpers for the sample-buffer (.bin tile) format: a writer that follows
the reference renderer's layout (pbrt_patches/sbmc_pbrt.diff:5796-5806 field
counts, :6140-6158 `write_compressed` = int32 size + LZ4F_compressFrame with
default preferences; header order as read by sbmc/datasets.py:504-520,592-593)
and a ctypes face of the real liblz4 (the C library the reference links), used
to produce frames and as an independent decompressor.
"""
import ctypes
import ctypes.util
import os
import struct

import numpy as np

PATH_DEPTH = 6
SAMPLE_FEATURES = 27
PIXEL_FEATURES = 30
VERSION = 20190401


class _Prefs(ctypes.Structure):
    """LZ4F_preferences_t (lz4frame.h, v1.9): frameInfo {blockSizeID, blockMode,
    contentChecksumFlag, frameType, contentSize, dictID, blockChecksumFlag},
    compressionLevel, autoFlush, favorDecSpeed, reserved[3]."""
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int),
                ("contentChecksumFlag", ctypes.c_int), ("frameType", ctypes.c_int),
                ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint),
                ("blockChecksumFlag", ctypes.c_int), ("compressionLevel", ctypes.c_int),
                ("autoFlush", ctypes.c_uint), ("favorDecSpeed", ctypes.c_uint),
                ("reserved", ctypes.c_uint * 3)]


_liblz4 = None


def liblz4():
    """The system liblz4 (None if absent)."""
    global _liblz4
    if _liblz4 is None:
        name = ctypes.util.find_library("lz4") or "liblz4.so.1"
        try:
            L = ctypes.CDLL(name)
        except OSError:
            _liblz4 = False
            return None
        sz, vp = ctypes.c_size_t, ctypes.c_void_p
        L.LZ4F_compressFrameBound.restype = sz
        L.LZ4F_compressFrameBound.argtypes = [sz, vp]
        L.LZ4F_compressFrame.restype = sz
        L.LZ4F_compressFrame.argtypes = [vp, sz, vp, sz, vp]
        L.LZ4F_isError.restype = ctypes.c_uint
        L.LZ4F_isError.argtypes = [sz]
        L.LZ4F_createDecompressionContext.restype = sz
        L.LZ4F_createDecompressionContext.argtypes = [ctypes.POINTER(vp), ctypes.c_uint]
        L.LZ4F_freeDecompressionContext.restype = sz
        L.LZ4F_freeDecompressionContext.argtypes = [vp]
        L.LZ4F_decompress.restype = sz
        L.LZ4F_decompress.argtypes = [vp, vp, ctypes.POINTER(sz), vp, ctypes.POINTER(sz), vp]
        _liblz4 = L
    return _liblz4 or None


def compress_frame(raw, block_size_id=0, independent=False, content_checksum=False,
                   block_checksum=False, content_size=False, level=0):
    """LZ4F_compressFrame; all-default arguments = the reference writer's NULL
    preferences (64 KiB linked blocks, no checksums, no content size)."""
    L = liblz4()
    if L is None:
        raise RuntimeError("liblz4 not available")
    raw = bytes(raw)
    default = not (block_size_id or independent or content_checksum or block_checksum
                   or content_size or level)
    prefs = None
    if not default:
        prefs = _Prefs()
        prefs.blockSizeID = block_size_id
        prefs.blockMode = 1 if independent else 0
        prefs.contentChecksumFlag = 1 if content_checksum else 0
        prefs.blockChecksumFlag = 1 if block_checksum else 0
        prefs.contentSize = len(raw) if content_size else 0
        prefs.compressionLevel = level
        prefs = ctypes.byref(prefs)
    bound = L.LZ4F_compressFrameBound(len(raw), prefs)
    dst = ctypes.create_string_buffer(bound)
    n = L.LZ4F_compressFrame(dst, bound, raw, len(raw), prefs)
    if L.LZ4F_isError(n):
        raise RuntimeError("LZ4F_compressFrame failed")
    return dst.raw[:n]


def decompress_frame(buf):
    """LZ4F_decompress of the real library: what `lz4.frame.decompress` wraps."""
    L = liblz4()
    if L is None:
        raise RuntimeError("liblz4 not available")
    buf = bytes(buf)
    ctx = ctypes.c_void_p()
    if L.LZ4F_isError(L.LZ4F_createDecompressionContext(ctypes.byref(ctx), 100)):
        raise RuntimeError("LZ4F_createDecompressionContext failed")
    try:
        out = []
        pos = 0
        chunk = ctypes.create_string_buffer(1 << 18)
        src = ctypes.create_string_buffer(buf, len(buf))
        base = ctypes.addressof(src)
        hint = 1
        while pos < len(buf):
            dst_n = ctypes.c_size_t(len(chunk))
            src_n = ctypes.c_size_t(len(buf) - pos)
            hint = L.LZ4F_decompress(ctx, chunk, ctypes.byref(dst_n), base + pos,
                                     ctypes.byref(src_n), None)
            if L.LZ4F_isError(hint):
                raise RuntimeError("LZ4F_decompress failed")
            out.append(chunk.raw[:dst_n.value])
            pos += src_n.value
            if src_n.value == 0 and dst_n.value == 0:
                break
        if hint != 0:
            raise RuntimeError("LZ4F_decompress: frame incomplete")
        return b"".join(out)
    finally:
        L.LZ4F_freeDecompressionContext(ctx)


def stored_frame(raw, block=1 << 16):
    """A valid LZ4 frame made of uncompressed blocks only (no library needed)."""
    from oracle import xxh32
    desc = bytes([0x60, 0x40])                     # v1, independent blocks, 64 KiB
    out = [struct.pack("<I", 0x184D2204), desc, bytes([(xxh32(desc) >> 8) & 0xFF])]
    raw = bytes(raw)
    for i in range(0, len(raw), block):
        part = raw[i:i + block]
        out.append(struct.pack("<I", 0x80000000 | len(part)))
        out.append(part)
    out.append(struct.pack("<I", 0))
    return b"".join(out)


def synth_tile(rng, ts, sample_count, scale=1.0, quantize=None):
    """Random tile content in the renderer's planar layout: dict with
    image [30, ts, ts] f32, floats [spp, 27 + 36, ts, ts] f32, bt [spp, 6, ts, ts] i16."""
    image = rng.standard_normal((PIXEL_FEATURES, ts, ts)).astype(np.float32) * scale
    floats = rng.standard_normal((sample_count, SAMPLE_FEATURES + 6 * PATH_DEPTH, ts, ts))
    floats = floats.astype(np.float32) * scale
    # radiance: mostly positive with some negatives / exact zeros (clamped by the reader)
    rad = np.abs(floats[:, 5:11]) * (rng.random(floats[:, 5:11].shape) > 0.1)
    rad = np.where(rng.random(rad.shape) < 0.05, -rad, rad)
    floats[:, 5:11] = rad.astype(np.float32)
    floats[:, 20:21] = (rng.random(floats[:, 20:21].shape) > 0.3)      # hasHit-like flags
    if quantize:      # fewer distinct bit patterns: the frames get real matches
        image = (np.round(image / quantize) * quantize).astype(np.float32)
        floats = (np.round(floats / quantize) * quantize).astype(np.float32)
    bt = rng.integers(0, 32, size=(sample_count, PATH_DEPTH, ts, ts)).astype(np.int16)
    bt[rng.random(bt.shape) < 0.5] = 0                                  # long zero runs
    return {"image": image, "floats": floats, "bt": bt}


def tile_bytes(content, ts, image_width, image_height, block_x, block_y, gt_sample_count=512,
               focus_distance=1.5, aperture_radius=0.25, fov=35.0, scene_radius=4.0,
               version=VERSION, path_depth=PATH_DEPTH, compress=None,
               sample_features=SAMPLE_FEATURES, pixel_features=PIXEL_FEATURES):
    """Serialises one tile (header, globals, block position, 1 + spp chunks)."""
    compress = compress or compress_frame
    spp = content["floats"].shape[0]
    out = [struct.pack("<9i", version, ts, image_width, image_height, spp, gt_sample_count,
                       sample_features, pixel_features, path_depth),
           struct.pack("<4f", focus_distance, aperture_radius, fov, scene_radius),
           struct.pack("<2i", block_x, block_y)]

    def chunk(raw):
        frame = compress(raw)
        out.append(struct.pack("<i", len(frame)))
        out.append(frame)

    chunk(np.ascontiguousarray(content["image"], np.float32).tobytes())
    for s in range(spp):
        chunk(np.ascontiguousarray(content["floats"][s], np.float32).tobytes()
              + np.ascontiguousarray(content["bt"][s], np.int16).tobytes())
    return b"".join(out)


def write_scene(root, scene, rng, ts, tiles_x, tiles_y, sample_count, quantize=None, **kw):
    """Writes tiles_x * tiles_y tiles of one scene; returns their contents keyed
    by (block_x, block_y)."""
    folder = os.path.join(root, scene)
    os.makedirs(folder, exist_ok=True)
    contents = {}
    idx = 0
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            content = synth_tile(rng, ts, sample_count, quantize=quantize)
            contents[(tx * ts, ty * ts)] = content
            with open(os.path.join(folder, "%s_tile%03d.bin" % (scene, idx)), "wb") as fid:
                fid.write(tile_bytes(content, ts, tiles_x * ts, tiles_y * ts, tx * ts, ty * ts, **kw))
            idx += 1
    return contents
