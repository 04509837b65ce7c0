"""Sample-buffer readers: the input side of the denoiser's callers.

Same classes, constructor arguments, properties, dictionary keys, tensor shapes
and error behaviour as the reference's ``sbmc/datasets.py`` (``TilesDataset``
:35-880, ``FullImagesDataset`` :883-1012, ``MultiSampleCountDataset`` :1015-1043),
with the data path rebuilt for the GPU:

  * the host only parses the 60-byte tile header and the chunk sizes;
  * the COMPRESSED chunks go to the device in one copy per item (pinned staging),
    so PCIe carries the LZ4 frames, not the inflated planes;
  * ``sbmc_lz4_frames_inflate`` inflates every frame of every tile of the item
    in one launch (one warp per frame) -- replaces ``lz4.frame.decompress``
    (datasets.py:570-579);
  * ``sbmc_tile_assemble_f32`` builds ``features`` / ``radiance`` / ``low_spp`` /
    ``image_data`` / ``image_data_var`` / ``target_image`` in one launch, pasting
    each tile at its (block_y, block_x) -- replaces the numpy code of
    ``_read_data`` (:581-739), ``_preprocess_standard`` (:744-778) and the tile
    loop of ``FullImagesDataset.__getitem__`` (:920-957).

Items are dictionaries of torch tensors already resident on ``device`` (the
reference returns numpy arrays that the DataLoader collates and the caller then
moves to the GPU); use ``num_workers=0`` -- there is no CPU work to parallelise.
There is no CPU data path: reading an item without the CUDA library raises.
"""
import os
import struct
import threading

import numpy as np
import torch as th
from torch.utils.data import ConcatDataset, Dataset

from . import _lib
from ._compat import get_logger

LOG = get_logger(__name__)

__all__ = ["TilesDataset", "FullImagesDataset", "MultiSampleCountDataset", "PrefetchLoader"]

_HEADER = struct.Struct("<9i4f")       # metadata (9 x int32) + global features (4 x float32)
_META_FIELDS = ("version", "tile_size", "image_width", "image_height", "sample_count",
                "gt_sample_count", "sample_features", "pixel_features", "path_depth")
_LZ4_ERRORS = {1: "not an LZ4 frame", 2: "unsupported LZ4 frame header", 3: "truncated frame",
               4: "frame larger than the tile header implies", 5: "invalid match offset",
               6: "inflated size differs from what the tile header implies",
               7: "block exceeds the frame's maximum block size", 8: "checksum mismatch"}
# flags of sbmc_tile_assemble_f32 (include/sbmc_b200.h)
_F_COORDS, _F_GBUFFER, _F_P, _F_LD, _F_BT, _F_LOG, _F_ALIGNED = 1, 2, 4, 8, 16, 32, 64


def _align(n, a=256):
    return (n + a - 1) // a * a


class _Staging(object):
    """One growing pinned host buffer per process: files are read straight into
    it and leave in a single asynchronous copy."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = th.empty(_align(max(nbytes, 1 << 20), 1 << 20), dtype=th.uint8,
                                pin_memory=th.cuda.is_available())
        return self.buf


_STAGING = _Staging()      # items read directly (dataset[i], DataLoader) stage here

# Threads that read tile files into the staging buffer (the only host work of an
# item that scales with its size; `readinto` releases the GIL).
_IO_THREADS = max(1, min(16, (os.cpu_count() or 1)))
_POOL = None
_POOL_LOCK = threading.Lock()


def _io_pool():
    global _POOL
    with _POOL_LOCK:        # the main thread and a loader's planner thread both get here
        if _POOL is None:
            from concurrent.futures import ThreadPoolExecutor
            _POOL = ThreadPoolExecutor(max_workers=_IO_THREADS, thread_name_prefix="sbmc-tiles")
        return _POOL


class _CudaBackend(object):
    """Where the two launches of an item run: libsbmc_b200's kernels on a CUDA
    device.  (The only backend of the product; the test-suite swaps in the device
    sources compiled for the host, tests/native/tiles_emul.cpp, to exercise the
    Python side of this module on machines without a GPU.)"""

    def __init__(self, device):
        if device is not None:
            self.device = th.device(device)
        elif th.cuda.is_available():
            self.device = th.device("cuda", th.cuda.current_device())
        else:
            self.device = None
        if self.device is None or self.device.type != "cuda" or not th.cuda.is_available():
            raise _lib.SbmcB200Error(
                "sbmc_b200.datasets inflates and assembles tiles on the GPU; no CUDA device is "
                "available and there is no CPU data path")
        self.lib = _lib.load()

    def scope(self):
        return th.cuda.device(self.device)

    def _stream(self):
        return th.cuda.current_stream(self.device).cuda_stream

    def inflate(self, comp, table, nframes, raw, status):
        _lib.check(self.lib.sbmc_lz4_frames_inflate(
            comp.data_ptr(), table.data_ptr(), nframes, raw.data_ptr(), status.data_ptr(),
            self._stream()), "lz4_frames_inflate")

    def assemble(self, *args):
        _lib.check(self.lib.sbmc_tile_assemble_f32(*(args + (self._stream(),))), "tile_assemble")


def _backend(device):
    return _CudaBackend(device)


class TilesDataset(Dataset):
    """Tiles stored one per .bin file (format: reference docstring,
    datasets.py:36-155).  `path` is a .txt list of .bin files or a root folder of
    scene folders.  Arguments as in the reference (datasets.py:190-192); `device`
    (extra) is where the tensors are produced, default the current CUDA device."""

    FILELIST_MODE = 0
    FOLDERS_MODE = 1
    PATH_DEPTH = 6
    N_BT_FEATURES = 5
    SBMC_MODE = "sbmc"
    RAW_MODE = "raw"
    KPCN_MODE = "kpcn"

    def __init__(self, path, spp=None, load_coords=True, load_gbuffer=True,
                 load_p=True, load_ld=True, load_bt=True, mode="sbmc", device=None):
        if mode not in (TilesDataset.SBMC_MODE, TilesDataset.RAW_MODE, TilesDataset.KPCN_MODE):
            LOG.error("Unknown dataset loading mode %s", mode)
            raise RuntimeError("Unknown dataset loading mode %s" % mode)
        self.mode = mode
        self.device = device
        sbmc = mode == TilesDataset.SBMC_MODE
        # `raw` and `kpcn` only keep radiance + g-buffer (datasets.py:210-216)
        self.load_coords = load_coords if sbmc else False
        self.load_gbuffer = load_gbuffer if sbmc else True
        self.load_p = load_p if sbmc else False
        self.load_ld = load_ld if sbmc else False
        self.load_bt = load_bt if sbmc else False
        self.count = 0
        self.spp = None
        for field in _META_FIELDS:
            setattr(self, field, None)
        self.root = self.files = self.scenes = self.tiles = self.indices = None
        self._init_filelist(path)
        self._init_labels()
        self._init_metadata(spp)

    # -- listing (datasets.py:245-297) -------------------------------------------------
    def _init_filelist(self, path):
        if os.path.splitext(path)[-1] == ".txt":
            self.io_mode = TilesDataset.FILELIST_MODE
            self.root = os.path.dirname(path)
            with open(path) as fid:
                self.files = [os.path.join(self.root, line.strip()) for line in fid.readlines()]
            self.count = len(self.files)
        elif os.path.isdir(path):
            self.io_mode = TilesDataset.FOLDERS_MODE
            self.root = path
            scenes = [os.path.join(path, d) for d in sorted(os.listdir(path))]
            self.scenes = [s for s in scenes if os.path.isdir(s)]
            self.tiles, self.indices = {}, {}
            idx = 0
            for scene in self.scenes:
                names = [f for f in sorted(os.listdir(scene)) if os.path.splitext(f)[-1] == ".bin"]
                self.tiles[scene] = [os.path.join(scene, f) for f in names]
                self.indices[scene] = (idx, idx + len(names))
                idx += len(names)
            self.count = idx
        else:
            LOG.error("Unknown dataset format at path %s, maybe the folder is empty?", path)
            raise RuntimeError("Incorrect data path.")
        if self.count == 0:
            LOG.error("Dataset is empty, please check the file format / folder structure.")
            raise RuntimeError("Empty dataset")

    # -- channel names (datasets.py:299-354) -------------------------------------------
    def _init_labels(self):
        self.image_channels = ["diffuse_r", "diffuse_g", "diffuse_b", "specular_r", "specular_g",
                               "specular_b", "albedo_r", "albedo_g", "albedo_b", "normal_x",
                               "normal_y", "normal_z", "depth", "visibility", "hasHit"]
        self.valid_versions = [20181212, 20190401]
        self.glabels = ["aperture_radius", "focus_distance", "fov"]
        depth = TilesDataset.PATH_DEPTH
        labels = []
        if self.load_coords:
            labels += ["dx", "dy", "lens_u", "lens_v", "t"]
        labels += ["diffuse_r", "diffuse_g", "diffuse_b", "specular_r", "specular_g", "specular_b"]
        if self.load_gbuffer:
            labels += ["normal_first_x", "normal_first_y", "normal_first_z", "normal_x", "normal_y",
                       "normal_z", "depth_first", "depth", "visibility", "hasHit", "albedo_first_r",
                       "albedo_first_g", "albedo_first_b", "albedo_r", "albedo_g", "albedo_b"]
        if self.load_p:
            labels += ["p"] * (depth * 4)
        if self.load_ld:
            for i in range(depth):
                labels += ["ld_theta_%d" % i, "ld_phi_%d" % i]
        if self.load_bt:
            for txt in ("reflection", "transmisson", "diffuse", "glossy", "specular"):
                labels += ["bt_%s_%d" % (txt, i) for i in range(depth)]
        self.labels = labels

    def _init_metadata(self, spp):
        fname = self._filename(0)
        with open(fname, "rb") as fid:
            try:
                self._parse_header(fid.read(_HEADER.size))
            except Exception:
                LOG.error("Could not read %s", fname)
                raise
        if spp is None:
            self.spp = self.sample_count
        else:
            if spp > self.sample_count:
                LOG.error("Requested %d samples, which is higher that what the data has %d",
                          spp, self.sample_count)
                raise RuntimeError("Requested too many samples.")
            self.spp = spp

    def __len__(self):
        return self.count

    def _filename(self, idx):
        if self.io_mode == TilesDataset.FOLDERS_MODE:
            scene = next(k for k in self.scenes if self.indices[k][1] > idx)
            return self.tiles[scene][idx - self.indices[scene][0]]
        return self.files[idx]

    @property
    def num_features(self):
        return 27 if self.mode == TilesDataset.KPCN_MODE else len(self.labels)

    @property
    def num_global_features(self):
        return 0 if self.mode == TilesDataset.KPCN_MODE else len(self.glabels)

    def __repr__(self):
        s = "Dataset v%d\n" % self.version
        s += "  .image size: %dx%d\n" % (self.image_width, self.image_height)
        s += "  .block size: %d\n" % self.tile_size
        s += "  .sample count: %d (of %d)\n" % (self.spp, self.sample_count)
        s += "  .suffix length: %d\n" % self.pixel_features
        s += "  .sample feature size: %d\n" % self.sample_features
        s += "  .path depth: %d\n" % self.path_depth
        for flag, txt in ((self.load_p, "probabilities"), (self.load_ld, "light direction"),
                          (self.load_bt, "bounce types")):
            if not flag:
                s += "  .NO %s\n" % txt
        s += "  .total feature count: %d (and %d global)\n" % (len(self.labels), len(self.glabels))
        return s

    # -- header (datasets.py:458-520) ----------------------------------------------------
    def _rcheck(self, field, value):
        current = getattr(self, field)
        if current is not None:
            if current != value:
                LOG.error("metadata do not match, got %s for field %s, should be %s",
                          value, field, current)
                raise ValueError("Metadata do not match.")
        else:
            if field == "version" and value not in self.valid_versions:
                raise ValueError("Version unsupported: got %s, valid are %s"
                                 % (value, self.valid_versions))
            setattr(self, field, value)
        if field == "path_depth" and value != TilesDataset.PATH_DEPTH:
            LOG.error("The path depth of the rendered data shoud be %d", TilesDataset.PATH_DEPTH)
            raise RuntimeError("Incorrect path depth in the data")

    def _parse_header(self, head):
        """Checks the nine metadata fields against the dataset's and returns the
        global features; raises struct.error on a short header like the reference's
        field-by-field `struct.unpack` (datasets.py:504-520)."""
        values = _HEADER.unpack(head)
        for field, value in zip(_META_FIELDS, values[:9]):
            self._rcheck(field, value)
        g = dict(zip(("focus_distance", "aperture_radius", "fov", "scene_radius"), values[9:]))
        if g["aperture_radius"] == 0:      # no depth of field: the focus distance is NaN on disk
            g["focus_distance"] = 0.0
        for key, what, err in (
                ("focus_distance", "Focus distance", "Incorrect focus distance feature."),
                ("aperture_radius", "Aperture radius", "Incorrect aperture radius feature."),
                ("fov", "Field of view", "Incorrect field of view feature."),
                ("scene_radius", "Scene radius", "Incorrect scene radius.")):
            if g[key] < 0:
                LOG.error("%s is negative: data is corrupt.", what)
                raise RuntimeError(err)
        return g

    # -- the GPU read path ----------------------------------------------------------------
    def _sample_frame_bytes(self):
        ts, depth = self.tile_size, self.path_depth
        return (self.sample_features + 6 * depth) * ts * ts * 4 + depth * ts * ts * 2

    def _flags(self):
        flags = 0
        for on, bit in ((self.load_coords, _F_COORDS), (self.load_gbuffer, _F_GBUFFER),
                        (self.load_p, _F_P), (self.load_ld, _F_LD), (self.load_bt, _F_BT),
                        (self.mode == TilesDataset.SBMC_MODE, _F_LOG)):
            if on:
                flags |= bit
        return flags

    def _needed_bytes(self, fname):
        """Length of the file's prefix that holds the header and the first `spp`
        sample chunks (the whole file when every sample is wanted)."""
        size = os.path.getsize(fname)
        if self.spp == self.sample_count:
            return size
        pos = _HEADER.size + 8
        with open(fname, "rb") as fid:
            for _ in range(1 + self.spp):
                fid.seek(pos)
                head = fid.read(4)
                if len(head) < 4:
                    return size             # truncated: the parser reports it
                (nbytes,) = struct.unpack("<i", head)
                if nbytes < 0:
                    return size
                pos += 4 + nbytes
        return min(pos, size)

    def _plan(self, fnames, staging=None):
        """Reads the files (only the chunks of the first `spp` samples) into a
        pinned staging buffer -- in parallel, file reads release the GIL -- and
        walks their chunk headers.  Pure host work (no CUDA call): PrefetchLoader
        runs it on a background thread, into its own buffers, while the GPU is busy
        with the previous item.
        Returns (staging view, frame table rows, per-tile records, inflated bytes)."""
        workers = min(_IO_THREADS, len(fnames))
        pool = _io_pool() if workers > 1 else None
        needed = list(pool.map(self._needed_bytes, fnames)) if pool else \
            [self._needed_bytes(f) for f in fnames]
        offsets, total = [], 0
        for need in needed:
            offsets.append(total)
            total += _align(need, 16)
        stage = (staging or _STAGING).get(total)
        host = stage.numpy()

        def read(job):
            fname, off, need = job
            with open(fname, "rb") as fid:
                return fid.readinto(memoryview(host[off:off + need]))

        jobs = list(zip(fnames, offsets, needed))
        gots = list(pool.map(read, jobs)) if pool else [read(j) for j in jobs]

        ts = self.tile_size
        ts_bytes = (self.pixel_features * ts * ts * 4, self._sample_frame_bytes())
        frames, tiles = [], []
        dst = 0      # offset in the inflated buffer
        for fname, src, got in zip(fnames, offsets, gots):
            try:
                view = host[src:src + got]
                try:
                    gfeatures = self._parse_header(view[:_HEADER.size].tobytes())
                except struct.error:
                    LOG.error("reading meta for file %s failed", fname)
                    raise
                pos = _HEADER.size
                try:
                    block_x, block_y = struct.unpack_from("<2i", view, pos)
                    pos += 8
                    record = {"path": fname, "gfeatures": gfeatures, "block_x": block_x,
                              "block_y": block_y, "image_off": dst}
                    for i in range(1 + self.spp):
                        (nbytes,) = struct.unpack_from("<i", view, pos)
                        pos += 4
                        if nbytes < 0 or pos + nbytes > got:
                            raise RuntimeError("chunk %d of %s runs past the end of the file"
                                               % (i, fname))
                        want = ts_bytes[0] if i == 0 else ts_bytes[1]
                        if i == 1:
                            record["samples_off"] = dst
                        frames.append((src + pos, nbytes, dst, want))
                        dst += _align(want)
                        pos += nbytes
                    record.setdefault("samples_off", dst)
                except (struct.error, RuntimeError):
                    LOG.error("reading data from file %s failed", fname)
                    raise
            except Exception:
                LOG.error("could not read %s", fname)
                raise
            tiles.append(record)
        return stage[:max(total, 1)], frames, tiles, dst

    def _read_tiles(self, fnames, height, width, positions=None, row0=0, clip_rows=False,
                    planned=None):
        """Inflates and assembles `fnames` into one set of [.., height, width]
        tensors holding image rows row0 .. row0 + height - 1; `positions` overrides
        the tiles' (block_x, block_y).  With `clip_rows` tiles may stick out of the
        row range (their rows outside are skipped: a rank's band), otherwise every
        tile must fit.  `planned` is the result of an earlier `_plan(fnames)`."""
        backend = _backend(self.device)
        dev = backend.device
        stage, frames, tiles, raw_bytes = planned if planned is not None else self._plan(fnames)
        ts, spp = self.tile_size, self.spp
        nchans = self.pixel_features // 2
        nf = len(self.labels)
        rows = []
        aligned = True
        for i, t in enumerate(tiles):
            bx, by = positions[i] if positions is not None else (t["block_x"], t["block_y"])
            fits = by >= row0 and by + ts <= row0 + height
            if bx < 0 or by < 0 or bx + ts > width or not (fits or clip_rows):
                raise ValueError("tile %s at (%d, %d) does not fit a %dx%d image"
                                 % (t["path"], bx, by, width, height))
            aligned &= bx % 4 == 0
            rows.append((t["image_off"], t["samples_off"], bx, by))
        with backend.scope():
            comp = stage.to(dev, non_blocking=True)
            table = th.tensor(frames, dtype=th.int64).reshape(-1, 4).to(dev, non_blocking=True)
            raw = th.empty(max(raw_bytes, 16), dtype=th.uint8, device=dev)
            status = th.empty(len(frames), dtype=th.int32, device=dev)
            backend.inflate(comp, table, len(frames), raw, status)
            tile_table = th.tensor(rows, dtype=th.int64).reshape(-1, 4).to(dev, non_blocking=True)
            whole = len(tiles) == 1 and height == ts and width == ts
            alloc = th.empty if whole else th.zeros      # uncovered pixels stay 0 (datasets.py:944)
            out = {"image_data": alloc(nchans, height, width, device=dev),
                   "image_data_var": alloc(nchans, height, width, device=dev),
                   "target_image": alloc(3, height, width, device=dev)}
            if spp > 0:
                out["features"] = alloc(spp, nf, height, width, device=dev)
                out["radiance"] = alloc(spp, 3, height, width, device=dev)
                out["low_spp"] = alloc(3, height, width, device=dev)
            ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
            backend.assemble(
                raw.data_ptr(), tile_table.data_ptr(), len(tiles), _align(self._sample_frame_bytes()),
                ts, spp, self.sample_features, self.pixel_features, self.path_depth,
                self._flags() | (_F_ALIGNED if aligned else 0), ptr("features"), ptr("radiance"),
                ptr("low_spp"), out["image_data"].data_ptr(), out["image_data_var"].data_ptr(),
                out["target_image"].data_ptr(), height, width, row0)
            bad = status.cpu()      # also orders the staging buffer's reuse after the copy
        if bool(bad.any()):
            f = int(bad.nonzero()[0])
            code = int(bad[f])
            per_tile = 1 + spp
            LOG.error("reading data from file %s failed", tiles[f // per_tile]["path"])
            raise RuntimeError("LZ4 frame %d of %s: %s" % (
                f % per_tile, tiles[f // per_tile]["path"], _LZ4_ERRORS.get(code, "code %d" % code)))
        if spp <= 0:
            LOG.warning("No sample requested, setting low_spp to 0")
            out["low_spp"] = th.zeros(3, height, width, dtype=th.float64, device=dev)
        return out, tiles

    def _global_features(self, gfeatures, dev):
        return th.tensor([gfeatures[k] for k in self.glabels], dtype=th.float32,
                         device=dev).reshape(len(self.glabels), 1, 1)

    def _get_raw_data(self, idx, planned=None):
        """One tile as the reference's raw sample dict (datasets.py:401-456)."""
        fname = self._filename(idx)
        ts = self.tile_size
        out, tiles = self._read_tiles([fname], ts, ts, positions=[(0, 0)], planned=planned)
        t = tiles[0]
        dev = out["target_image"].device
        sample = {"block_x": t["block_x"], "block_y": t["block_y"],
                  "global_features": self._global_features(t["gfeatures"], dev)}
        sample.update(out)
        sample["spp"] = th.full((1, 1, 1), self.spp, dtype=th.int32, device=dev)
        sample["scene_radius"] = t["gfeatures"]["scene_radius"]
        sample["path"] = fname
        return sample

    def __getitem__(self, idx):
        sample = self._get_raw_data(idx)
        if self.mode == TilesDataset.KPCN_MODE:
            sample = self._preprocess_kpcn(sample)
        return sample       # the sbmc-mode log compression ran inside the assembly kernel

    def __getitems__(self, indices, planned=None):
        """Batched fetch (the DataLoader calls this once per batch when it exists):
        all tiles of the batch are inflated and assembled by one pair of launches,
        stacked along rows of one scratch image; every sample is its slice."""
        indices = list(indices)
        if self.mode == TilesDataset.KPCN_MODE:
            return [self[i] for i in indices]
        if len(indices) == 1:
            return [self._get_raw_data(indices[0], planned=planned)]
        ts = self.tile_size
        fnames = [self._filename(i) for i in indices]
        out, tiles = self._read_tiles(fnames, ts * len(fnames), ts, planned=planned,
                                      positions=[(0, i * ts) for i in range(len(fnames))])
        dev = out["target_image"].device
        samples = []
        for i, t in enumerate(tiles):
            sample = {"block_x": t["block_x"], "block_y": t["block_y"],
                      "global_features": self._global_features(t["gfeatures"], dev)}
            for k, v in out.items():
                sample[k] = v[..., i * ts:(i + 1) * ts, :]
            sample["spp"] = th.full((1, 1, 1), self.spp, dtype=th.int32, device=dev)
            sample["scene_radius"] = t["gfeatures"]["scene_radius"]
            sample["path"] = t["path"]
            samples.append(sample)
        return samples

    # -- [Bako2017] inputs (datasets.py:780-859); second caller, composed from torch ops --
    def _preprocess_kpcn(self, sample):
        src_f, tgt = sample["features"], sample["image_data"]
        spp = src_f.shape[0]
        eps = 0.00316

        def mean_var(label):
            i = self.labels.index(label)
            block = src_f[:, i:i + 3]
            return block.mean(0), block.var(0, unbiased=False).mean(0, keepdim=True) / spp

        i = self.labels.index("depth")
        depth = src_f[:, i:i + 1].mean(0)
        depth_v = src_f[:, i:i + 1].var(0, unbiased=False)
        max_depth = depth.max()
        if max_depth > 0:
            depth = depth / max_depth
            depth_v = depth_v / (max_depth * max_depth * spp)
        depth = depth.clamp(0, 1)

        albedo, albedo_v = mean_var("albedo_r")
        albedo = albedo + eps
        j = self.image_channels.index("albedo_r")
        albedo_r = tgt[j:j + 3] + eps
        albedo_sqr = (albedo * albedo).mean(0, keepdim=True)

        diffuse, diffuse_v = mean_var("diffuse_r")
        diffuse = diffuse.clamp_min(0)
        j = self.image_channels.index("diffuse_r")
        diffuse_r = tgt[j:j + 3].clamp_min(0)
        specular, specular_v = mean_var("specular_r")
        specular = specular.clamp_min(0)
        j = self.image_channels.index("specular_r")
        specular_r = tgt[j:j + 3].clamp_min(0)

        diffuse = diffuse / albedo
        diffuse_v = diffuse_v / albedo_sqr
        specular = th.log(1 + specular)
        specular_v = specular_v / (((1 + specular) * (1 + specular)).mean(0, keepdim=True) + 1e-5)
        normals, normals_v = mean_var("normal_x")

        grads = self._gradients
        normals_g, depth_g, albedo_g = grads(normals), grads(depth), grads(albedo)
        specular_g, diffuse_g = grads(specular), grads(diffuse)
        specular_r = th.log(1 + specular_r.clamp_min(0))      # transformed targets: computed,
        diffuse_r = diffuse_r / albedo_r                      # not returned (datasets.py:836-837)
        del specular_r, diffuse_r

        shared = [normals_g, normals_v, depth_g, depth_v, albedo_g, albedo_v]
        out = {"kpcn_diffuse_in": th.cat([diffuse] + shared + [diffuse_g, diffuse_v], 0),
               "kpcn_specular_in": th.cat([specular] + shared + [specular_g, specular_v], 0),
               "kpcn_diffuse_buffer": diffuse, "kpcn_specular_buffer": specular,
               "kpcn_albedo": albedo}
        for k in ("target_image", "low_spp", "spp", "block_x", "block_y"):
            out[k] = sample[k]
        return out

    @staticmethod
    def _gradients(buf):
        """[c, h, w] -> [2c, h, w]: backward differences in x then y, first column /
        row zero (datasets.py:861-877)."""
        dx = th.zeros_like(buf)
        dy = th.zeros_like(buf)
        dx[:, :, 1:] = buf[:, :, 1:] - buf[:, :, :-1]
        dy[:, 1:] = buf[:, 1:] - buf[:, :-1]
        return th.cat([dx, dy], 0)


class FullImagesDataset(Dataset):
    """Whole images assembled from the tiles of one scene folder
    (datasets.py:883-1012).  In sbmc / raw mode every tile of the scene is inflated
    and pasted by ONE pair of launches; kpcn mode preprocesses tile by tile like
    the reference does (its depth normalisation and gradients are per tile)."""

    def __init__(self, *args, **kwargs):
        self.tiles_dset = TilesDataset(*args, **kwargs)
        if self.tiles_dset.io_mode != TilesDataset.FOLDERS_MODE:
            LOG.error("Full image dataset needs to point to a folder containing scenes, got '%s'.",
                      args[0] if args else kwargs.get("path"))
            raise RuntimeError("TilesDataset should be in folder mode.")
        self.scenes = self.tiles_dset.scenes

    def __len__(self):
        return len(self.scenes)

    def __repr__(self):
        return self.tiles_dset.__repr__()

    def get_scene_name(self, idx):
        return self.scenes[idx]

    def _scene_files(self, idx):
        d = self.tiles_dset
        start, end = d.indices[self.scenes[idx]]
        return [d._filename(i) for i in range(start, end)]

    def __getitem__(self, idx, planned=None):
        d = self.tiles_dset
        scene = self.scenes[idx]
        start, end = d.indices[scene]
        height, width, ts = d.image_height, d.image_width, d.tile_size
        if d.mode == TilesDataset.KPCN_MODE:
            return self._paste_tiles(start, end)
        out, tiles = d._read_tiles(self._scene_files(idx), height, width, planned=planned)
        dev = out["target_image"].device
        first = tiles[0]["gfeatures"]
        sample = {"global_features": d._global_features(first, dev),
                  "scene_radius": first["scene_radius"]}
        sample.update(out)
        # the reference pastes the [1, 1, 1] sample count like an image (datasets.py:936-949)
        spp_img = th.zeros(1, height, width, dtype=th.int32, device=dev)
        for t in tiles:
            spp_img[:, t["block_y"]:t["block_y"] + ts, t["block_x"]:t["block_x"] + ts] = d.spp
        sample["spp"] = spp_img
        return sample

    def tile_positions(self, idx):
        """[(file, block_x, block_y)] of scene `idx` from the 60-byte tile headers
        (cached): what a rank needs to pick the tiles of its row band."""
        cache = self.__dict__.setdefault("_positions", {})
        if idx not in cache:
            d = self.tiles_dset
            start, end = d.indices[self.scenes[idx]]
            found = []
            for i in range(start, end):
                fname = d._filename(i)
                with open(fname, "rb") as fid:
                    head = fid.read(_HEADER.size + 8)
                d._parse_header(head[:_HEADER.size])
                found.append((fname,) + struct.unpack_from("<2i", head, _HEADER.size))
            cache[idx] = found
        return cache[idx]

    def read_rows(self, idx, y_lo, y_hi):
        """Rows [y_lo, y_hi) of scene `idx`: same keys as `self[idx]` with tensors
        [.., y_hi - y_lo, width].  Only the tiles that intersect the rows are read,
        shipped and inflated -- the unit of multi-GPU sharding of the reader: tiles
        are independent, so ranks split an image by rows without any exchange
        (sbmc_b200.sharding.BandPlan gives the rows)."""
        d = self.tiles_dset
        if d.mode == TilesDataset.KPCN_MODE:
            raise ValueError("read_rows serves the sample-based modes (sbmc / raw)")
        height, width, ts = d.image_height, d.image_width, d.tile_size
        if not 0 <= y_lo < y_hi <= height:
            raise ValueError("rows [%d, %d) are not inside the image (%d rows)"
                             % (y_lo, y_hi, height))
        mine = [(f, bx, by) for f, bx, by in self.tile_positions(idx)
                if by < y_hi and by + ts > y_lo]
        if not mine:
            raise ValueError("no tile of %s covers rows [%d, %d)" % (self.scenes[idx], y_lo, y_hi))
        out, tiles = d._read_tiles([m[0] for m in mine], y_hi - y_lo, width, row0=y_lo,
                                   clip_rows=True)
        dev = out["target_image"].device
        first = tiles[0]["gfeatures"]
        sample = {"global_features": d._global_features(first, dev),
                  "scene_radius": first["scene_radius"]}
        sample.update(out)
        spp_img = th.zeros(1, y_hi - y_lo, width, dtype=th.int32, device=dev)
        for t in tiles:
            lo, hi = max(t["block_y"], y_lo) - y_lo, min(t["block_y"] + ts, y_hi) - y_lo
            spp_img[:, lo:hi, t["block_x"]:t["block_x"] + ts] = d.spp
        sample["spp"] = spp_img
        return sample

    def _paste_tiles(self, start, end):
        d = self.tiles_dset
        height, width, ts = d.image_height, d.image_width, d.tile_size
        first = d[start]
        sample = {}
        tensor_keys = []
        for k, v in first.items():
            if k in ("global_features", "scene_radius"):
                sample[k] = v
            elif isinstance(v, th.Tensor):
                tensor_keys.append(k)
                sample[k] = th.zeros(tuple(v.shape[:-2]) + (height, width), dtype=v.dtype,
                                     device=v.device)
        for tidx in range(start, end):
            tile = first if tidx == start else d[tidx]
            bx, by = tile["block_x"], tile["block_y"]
            for k in tensor_keys:
                sample[k][..., by:by + ts, bx:bx + ts] = tile[k]
        return sample

    num_features = property(lambda self: self.tiles_dset.num_features)
    num_global_features = property(lambda self: self.tiles_dset.num_global_features)
    spp = property(lambda self: self.tiles_dset.spp)
    sample_count = property(lambda self: self.tiles_dset.sample_count)
    gt_sample_count = property(lambda self: self.tiles_dset.gt_sample_count)
    load_p = property(lambda self: self.tiles_dset.load_p)
    load_ld = property(lambda self: self.tiles_dset.load_ld)
    load_bt = property(lambda self: self.tiles_dset.load_bt)
    labels = property(lambda self: self.tiles_dset.labels)
    glabels = property(lambda self: self.tiles_dset.glabels)
    version = property(lambda self: self.tiles_dset.version)
    image_channels = property(lambda self: self.tiles_dset.image_channels)


class MultiSampleCountDataset(ConcatDataset):
    """Tiles at every sample count in [2, spp] (datasets.py:1015-1043); the sample
    dimension varies between items, so batch with batch_size = 1."""

    def __init__(self, *args, **kwargs):
        spp = kwargs.get("spp", None)
        if spp is None:
            LOG.error("MultiSampleCountDataset requires a number of spps")
            raise RuntimeError("spp not provided.")
        if spp < 2:
            LOG.error("MultiSampleCountDataset needs at least 2spp")
            raise RuntimeError("spp too low to randomize sample count, should be at least 2.")
        datasets = []
        for count in range(2, spp + 1):
            kwargs["spp"] = count
            datasets.append(TilesDataset(*args, **kwargs))
        super(MultiSampleCountDataset, self).__init__(datasets)
        first = datasets[0]
        self.labels, self.glabels, self.version = first.labels, first.glabels, first.version
        self.num_features = first.num_features
        self.num_global_features = first.num_global_features


class PrefetchLoader(object):
    """Iterates a dataset of this module in batches like
    `DataLoader(dataset, batch_size, shuffle, num_workers=0)` (same collation),
    with the host half of batch i+1 -- reading its tile files into the second
    staging buffer and walking their chunk tables -- running on a background
    thread while batch i is inflated, assembled and consumed on the GPU.  (An
    extra on top of the reference's interface: its DataLoader workers overlapped
    the CPU decoding the same way; here only file reads are left on the host.)

    TilesDataset: any batch size.  FullImagesDataset and MultiSampleCountDataset:
    batch_size = 1 (the sample dimension varies / images are whole).  kpcn mode has
    no split host half and is iterated plainly."""

    def __init__(self, dataset, batch_size=1, shuffle=False, drop_last=False, generator=None,
                 device_prefetch=0):
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        if not isinstance(dataset, TilesDataset) and batch_size != 1:
            raise ValueError("batch_size must be 1 for %s" % type(dataset).__name__)
        self.dataset, self.batch_size = dataset, batch_size
        self.shuffle, self.drop_last, self.generator = shuffle, drop_last, generator
        # device_prefetch = D > 0 (TilesDataset in sbmc mode on a GPU): D batches at a time
        # are read, inflated and assembled as ONE group on a side stream by a background
        # thread while the consumer trains on the previous group.  The inflater decodes one
        # LZ4 frame per warp and a frame is a serial chain (~0.1 s for a 128 x 128 sample
        # plane of 27 floats): a single batch of 8 tiles keeps 72 warps busy for that long,
        # D batches keep 72 D warps busy for the same time -- the per-batch cost drops D-fold
        # and the decode overlaps the training step (benchmarks/train_e2e_bench.py).
        self.device_prefetch = int(device_prefetch)
        # while the copy / kernels of one batch read the first buffer, the files of
        # the next batch are read into the second; both belong to this loader only
        self._stagings = (_Staging(), _Staging())

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _resolve(self, idx):
        """-> (tiles dataset or full-images dataset, local index)."""
        d = self.dataset
        if isinstance(d, ConcatDataset):
            import bisect
            which = bisect.bisect_right(d.cumulative_sizes, idx)
            return d.datasets[which], idx - (d.cumulative_sizes[which - 1] if which else 0)
        return d, idx

    def _host_half(self, batch, slot, cuda_device=None):
        if cuda_device is not None:     # pinned allocations belong to the caller's device
            th.cuda.set_device(cuda_device)
        d, first = self._resolve(batch[0])
        tiles = d.tiles_dset if isinstance(d, FullImagesDataset) else d
        if tiles.mode == TilesDataset.KPCN_MODE:
            return None
        if isinstance(d, FullImagesDataset):
            return tiles._plan(d._scene_files(first), self._stagings[slot])
        return tiles._plan([d._filename(self._resolve(i)[1]) for i in batch], self._stagings[slot])

    def _device_half(self, batch, planned):
        d, first = self._resolve(batch[0])
        if isinstance(d, FullImagesDataset):
            return [d.__getitem__(first, planned=planned)]
        if planned is None:
            return [d[self._resolve(i)[1]] for i in batch]
        return d.__getitems__([self._resolve(i)[1] for i in batch], planned=planned)

    def _can_prefetch_on_device(self):
        d = self.dataset
        parts = d.datasets if isinstance(d, ConcatDataset) else [d]
        return (len(parts) == 1 and isinstance(parts[0], TilesDataset)
                and parts[0].mode == TilesDataset.SBMC_MODE)

    def _iter_device_prefetch(self, batches):
        """Groups of `device_prefetch` batches, each decoded by one pair of launches on a side
        stream from a background thread; the consumer's stream waits on the group's event."""
        import contextlib
        import queue
        from concurrent.futures import ThreadPoolExecutor
        from torch.utils.data import default_collate
        depth = self.device_prefetch
        groups = [batches[g:g + depth] for g in range(0, len(batches), depth)]
        d = self.dataset.datasets[0] if isinstance(self.dataset, ConcatDataset) else self.dataset
        on_gpu = _backend(d.device).device.type == "cuda"     # (host emulation in the tests: no streams)
        cuda_device = th.cuda.current_device() if on_gpu else None
        side = th.cuda.Stream(device=cuda_device) if on_gpu else None
        ready = queue.Queue(maxsize=2)
        stop = threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    ready.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def worker():
            try:
                if on_gpu:
                    th.cuda.set_device(cuda_device)
                with ThreadPoolExecutor(max_workers=1, thread_name_prefix="sbmc-plan") as planner, \
                        (th.cuda.stream(side) if on_gpu else contextlib.nullcontext()):
                    flat = lambda grp: [i for b in grp for i in b]          # noqa: E731
                    pending = planner.submit(self._host_half, flat(groups[0]), 0, cuda_device)
                    for g, grp in enumerate(groups):
                        if stop.is_set():
                            return
                        planned = pending.result()
                        if g + 1 < len(groups):
                            pending = planner.submit(self._host_half, flat(groups[g + 1]),
                                                     (g + 1) % 2, cuda_device)
                        items = self._device_half(flat(grp), planned)
                        out, k = [], 0
                        for b in grp:
                            out.append(default_collate(items[k:k + len(b)]))
                            k += len(b)
                        del items
                        ev = None
                        if on_gpu:
                            ev = th.cuda.Event()
                            ev.record(side)
                        if not put((out, ev)):
                            return
                put(None)
            except BaseException as exc:            # surfaces in the consumer
                put(exc)

        thread = threading.Thread(target=worker, name="sbmc-device-prefetch", daemon=True)
        thread.start()
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                out, ev = item
                if on_gpu:
                    cur = th.cuda.current_stream(cuda_device)
                    cur.wait_event(ev)
                    for batch in out:
                        for v in batch.values():        # allocated on the side stream, used here
                            if isinstance(v, th.Tensor) and v.is_cuda:
                                v.record_stream(cur)
                for batch in out:
                    yield batch
        finally:
            stop.set()
            while thread.is_alive():
                try:
                    ready.get_nowait()
                except queue.Empty:
                    thread.join(timeout=0.05)

    def __iter__(self):
        from concurrent.futures import ThreadPoolExecutor
        from torch.utils.data import default_collate
        n = len(self.dataset)
        order = th.randperm(n, generator=self.generator).tolist() if self.shuffle else list(range(n))
        batches = [order[i:i + self.batch_size] for i in range(0, n, self.batch_size)]
        if self.drop_last and batches and len(batches[-1]) < self.batch_size:
            batches.pop()
        if not batches:
            return
        if self.device_prefetch > 0 and self._can_prefetch_on_device():
            yield from self._iter_device_prefetch(batches)     # (forwards close() to the worker's owner)
            return
        # one planner thread, distinct from the file-read pool it fans out to
        cuda_device = th.cuda.current_device() if th.cuda.is_available() else None
        with ThreadPoolExecutor(max_workers=1, thread_name_prefix="sbmc-plan") as planner:
            pending = planner.submit(self._host_half, batches[0], 0, cuda_device)
            for i, batch in enumerate(batches):
                planned = pending.result()
                if i + 1 < len(batches):
                    pending = planner.submit(self._host_half, batches[i + 1], (i + 1) % 2,
                                             cuda_device)
                yield default_collate(self._device_half(batch, planned))
