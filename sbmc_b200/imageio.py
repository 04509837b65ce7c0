"""Output images of the denoise script: OpenEXR (float32 RGB, uncompressed
scanlines) and 8-bit PNG.  The reference writes both through external packages
(`pyexr.write` and `skimage.io.imsave`, scripts/denoise.py:163-166) that are not in
this image; these are small writers for exactly those two calls, and a reader for
the EXR subset the writer emits (tests read the files back).
"""
import struct
import zlib

import numpy as np

__all__ = ["write_exr", "read_exr", "write_png"]

_EXR_MAGIC = 20000630


def _attr(name, kind, payload):
    return name.encode() + b"\0" + kind.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path, image):
    """image: [h, w, 3] (or [h, w]) float array -> scanline OpenEXR, FLOAT channels
    B, G, R (or Y), no compression."""
    image = np.asarray(image, np.float32)
    if image.ndim == 2:
        image = image[..., None]
    h, w, c = image.shape
    if c not in (1, 3):
        raise ValueError("write_exr wants 1 or 3 channels, got %d" % c)
    names = ["Y"] if c == 1 else ["B", "G", "R"]          # stored alphabetically
    order = [0] if c == 1 else [2, 1, 0]
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1)
                      for n in names) + b"\0"
    window = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = b"".join([
        struct.pack("<ii", _EXR_MAGIC, 2),
        _attr("channels", "chlist", chlist),
        _attr("compression", "compression", b"\0"),
        _attr("dataWindow", "box2i", window),
        _attr("displayWindow", "box2i", window),
        _attr("lineOrder", "lineOrder", b"\0"),
        _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0)),
        _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)),
        b"\0"])
    row_bytes = c * w * 4
    first = len(header) + 8 * h
    offsets = struct.pack("<%dQ" % h, *[first + y * (8 + row_bytes) for y in range(h)])
    planar = np.ascontiguousarray(image[:, :, order].transpose(0, 2, 1))     # [h, c, w]
    with open(path, "wb") as fid:
        fid.write(header)
        fid.write(offsets)
        for y in range(h):
            fid.write(struct.pack("<ii", y, row_bytes))
            fid.write(planar[y].tobytes())


def read_exr(path):
    """Reads back what write_exr wrote: -> [h, w, c] float32 (RGB order)."""
    buf = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", buf, 0)
    if magic != _EXR_MAGIC or version & 0xFF != 2:
        raise ValueError("not an OpenEXR v2 file")
    pos = 8
    attrs = {}
    while buf[pos] != 0:
        end = buf.index(b"\0", pos)
        name = buf[pos:end].decode()
        pos = end + 1
        end = buf.index(b"\0", pos)
        pos = end + 1
        (size,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        attrs[name] = buf[pos:pos + size]
        pos += size
    pos += 1
    if attrs["compression"] != b"\0":
        raise ValueError("only uncompressed files are supported")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names = []
    ch = attrs["channels"]
    p = 0
    while ch[p] != 0:
        end = ch.index(b"\0", p)
        names.append(ch[p:end].decode())
        p = end + 1 + 16
    c = len(names)
    offsets = struct.unpack_from("<%dQ" % h, buf, pos)
    out = np.zeros((h, c, w), np.float32)
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", buf, off)
        out[y - y0] = np.frombuffer(buf, np.float32, c * w, off + 8).reshape(c, w)
    order = [names.index(n) for n in (("R", "G", "B") if c == 3 else names)]
    return out[:, order].transpose(0, 2, 1)


def write_png(path, image):
    """image: [h, w, 3] or [h, w] uint8 -> PNG (filter 0, zlib)."""
    image = np.ascontiguousarray(image, np.uint8)
    if image.ndim == 2:
        image = image[..., None]
    h, w, c = image.shape
    color = {1: 0, 3: 2, 4: 6}[c]

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    rows = np.concatenate([np.zeros((h, 1), np.uint8), image.reshape(h, w * c)], 1)
    with open(path, "wb") as fid:
        fid.write(b"\x89PNG\r\n\x1a\n")
        fid.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0)))
        fid.write(chunk(b"IDAT", zlib.compress(rows.tobytes(), 6)))
        fid.write(chunk(b"IEND", b""))
