"""The sample-buffer reader (`.bin` + lz4 tiles,
sbmc/datasets.py).

CPU suite (`-m "not gpu"`):
  * oracle/lz4_oracle.c against the real liblz4 and pyarrow's lz4 codec;
  * the DEVICE inflater and assembly body compiled for the host
    (tests/native/tiles_emul.cpp) against the same data -- cursor / index
    arithmetic of the kernels without a GPU;
  * oracle/tiles_ref.py and the emulated kernels against fixtures produced by
    the reference's own sbmc/datasets.py (tests/golden/make_tiles_golden.py);
  * host logic of sbmc_b200.datasets: listing modes, labels, header checks and
    the reference's error behaviour; reading an item without CUDA raises.
GPU suite (`-m gpu`): the product path (C ABI kernels behind
sbmc_b200.datasets) against the same fixtures and against the oracle on larger
synthetic tiles.

Tolerances: everything that is data movement or a single fp32 add / divide is
compared BIT FOR BIT (sha256 of the reference's arrays); the six log-compressed
radiance channels go through `log` (numpy's SIMD log on the reference side,
libm logf / CUDA logf here): 2 ulp-ish, |a-b| <= 1e-6 * max(|b|, 1e-3).
"""
import ctypes
import hashlib
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch as th

import oracle
from oracle import tiles_ref
from sbmc_b200 import _lib, datasets
from tests import tile_io

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "tiles")
DATA = os.path.join(GOLD, "data")
LOG_RTOL = 1e-6

needs_liblz4 = pytest.mark.skipif(tile_io.liblz4() is None, reason="system liblz4 not found")


# ------------------------------------------------------------------ helpers
def emul(reverse_lanes=False, wide=False):
    """tests/native/tiles_emul.cpp built with g++ (device sources on the host);
    `reverse_lanes` runs the 32 lanes of the inflater from 31 down to 0, `wide`
    builds the experimental 16-byte copy path (-DSBMC_LZ4_WIDE_COPY)."""
    src = os.path.join(HERE, "native", "tiles_emul.cpp")
    out = os.path.join(HERE, "native", "libtiles_emul%s%s.so" % (
        "_rev" if reverse_lanes else "", "_wide" if wide else ""))
    deps = [src] + [os.path.join(HERE, "..", "sbmc_b200", "csrc", f)
                    for f in ("lz4_warp.cuh", "tiles_body.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = "/usr/local/cuda/include"
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I", cuda_inc, "-o", out, src]
                              + (["-DSBMC_LZ4_REVERSE_LANES"] if reverse_lanes else [])
                              + (["-DSBMC_LZ4_WIDE_COPY"] if wide else []))
    lib = ctypes.CDLL(out)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.emul_lz4_frames_inflate.argtypes = [vp, vp, i64, vp, vp]
    lib.emul_tile_assemble_f32.argtypes = [vp, vp, i64, i64, i32, i32, i32, i32, i32, i32,
                                           vp, vp, vp, vp, vp, vp, i64, i64, i64]
    lib.emul_last_error.restype = ctypes.c_char_p
    return lib


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def frame_corpus(rng):
    raws = []
    for n in (0, 1, 13, 100, 4096, 65536, 70001, 200000):
        raws.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes())          # incompressible
        raws.append(rng.integers(0, 2, n, dtype=np.uint8).tobytes())            # short matches
        raws.append(np.repeat(rng.standard_normal(n // 16 + 1).astype(np.float32), 4).tobytes()[:n])
        raws.append(bytes(n))                                                   # offset-1 runs
        raws.append((b"abc" * (n // 3 + 1))[:n])                                # offset-3 overlap
    return raws


FRAME_OPTIONS = [dict(), dict(independent=True),
                 dict(content_checksum=True, block_checksum=True, content_size=True),
                 dict(level=9), dict(block_size_id=5)]


def layout(frames, sizes):
    """Frame table the way sbmc_b200.datasets lays frames out."""
    table, so, do = [], 0, 0
    for f, n in zip(frames, sizes):
        table.append((so, len(f), do, n))
        so += len(f)
        do += (n + 255) // 256 * 256
    return np.array(table, np.int64).reshape(-1, 4), max(do, 16)


def emul_inflate(frames, sizes, reverse_lanes=False, wide=False):
    table, total = layout(frames, sizes)
    src = np.frombuffer(b"".join(frames) + b"\0", np.uint8)
    dst = np.zeros(total, np.uint8)
    status = np.zeros(len(frames), np.int32)
    vp = ctypes.c_void_p
    emul(reverse_lanes, wide).emul_lz4_frames_inflate(
        src.ctypes.data_as(vp), table.ctypes.data_as(vp), len(frames), dst.ctypes.data_as(vp),
        status.ctypes.data_as(vp))
    return [dst[t[2]:t[2] + t[3]].tobytes() for t in table], status


def expected():
    return np.load(os.path.join(GOLD, "expected.npz"), allow_pickle=False)


def check_item(exp, prefix, item, i_diffuse, exact=True):
    """Compares one dataset item with what the reference produced for it."""
    keys = [k for k in exp.files if k.startswith(prefix + "/")]
    assert keys, prefix
    seen = set()
    for key in keys:
        name = key[len(prefix) + 1:]
        base = name.split("#")[0]
        seen.add(base)
        got = item[base]
        if isinstance(got, th.Tensor):
            got = got.cpu().numpy()
        if name.endswith("#sha"):
            assert sha(got) == str(exp[key]), "%s differs from the reference" % key
        elif name.endswith("#sha_nolog"):
            rest = np.concatenate([got[:, :i_diffuse], got[:, i_diffuse + 6:]], 1)
            assert sha(rest) == str(exp[key]), "%s differs from the reference" % key
        elif name.endswith("#log"):
            ref = exp[key]
            mine = got[:, i_diffuse:i_diffuse + 6]
            assert mine.dtype == ref.dtype and mine.shape == ref.shape
            tol = LOG_RTOL * np.maximum(np.abs(ref), 1e-3)
            assert (np.abs(mine.astype(np.float64) - ref) <= tol).all(), key
        elif isinstance(got, np.ndarray):      # kpcn arrays: reductions in another order
            ref = exp[key]
            assert got.shape == ref.shape and got.dtype == ref.dtype, key
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6, err_msg=key)
        else:
            ref = exp[key]
            if np.issubdtype(ref.dtype, np.floating):
                assert float(got) == pytest.approx(float(ref), rel=1e-7), key
            else:
                assert got == ref.item(), key
    return seen


SBMC_CONFIGS = {      # name -> (full image?, list mode?, reader kwargs) as in make_tiles_golden.py
    "tiles_sbmc_all": (False, False, dict()),
    "tiles_sbmc_subset": (False, True, dict(spp=2, load_coords=False, load_p=False)),
    "tiles_sbmc_nogbuf": (False, False, dict(spp=1, load_gbuffer=False, load_ld=False,
                                             load_bt=False)),
    "tiles_raw": (False, False, dict(mode="raw")),
    "full_sbmc_all": (True, False, dict()),
    "full_sbmc_spp2": (True, False, dict(spp=2, load_bt=False)),
}


def fixture_files(listed):
    if listed:
        with open(os.path.join(DATA, "list.txt")) as fid:
            return [os.path.join(DATA, l.strip()) for l in fid]
    out = []
    for scene in sorted(os.listdir(DATA)):
        folder = os.path.join(DATA, scene)
        if os.path.isdir(folder):
            out += [os.path.join(folder, f) for f in sorted(os.listdir(folder))]
    return out


def oracle_kwargs(kw):
    kw = dict(kw)
    mode = kw.pop("mode", "sbmc")
    if mode != "sbmc":
        kw.update(load_coords=False, load_gbuffer=True, load_p=False, load_ld=False, load_bt=False)
    kw["log_radiance"] = mode == "sbmc"
    return kw


def i_diffuse_of(kw):
    return 5 if (kw.get("load_coords", True) and kw.get("mode", "sbmc") == "sbmc") else 0


# ------------------------------------------------------------------ LZ4: oracle + device code on the host
@needs_liblz4
@pytest.mark.parametrize("opts", FRAME_OPTIONS, ids=lambda o: "-".join(o) or "default")
def test_lz4_oracle_and_emulated_warp_inflater_match_liblz4(opts):
    rng = np.random.default_rng(7)
    raws = frame_corpus(rng)
    frames = [tile_io.compress_frame(r, **opts) for r in raws]
    for f, r in zip(frames, raws):
        assert tile_io.decompress_frame(f) == r          # the real library round-trips
        assert oracle.lz4_frame_decompress(f) == r
    outs, status = emul_inflate(frames, [len(r) for r in raws])
    assert not status.any(), status
    assert outs == raws
    # the copies of one sequence are independent of the order the lanes run in
    # (on the GPU they run concurrently): same bytes with the lanes reversed
    outs, status = emul_inflate(frames, [len(r) for r in raws], reverse_lanes=True)
    assert not status.any() and outs == raws
    # the experimental 16-byte copy path (off in the product build) gives the same bytes
    for rev in (False, True):
        outs, status = emul_inflate(frames, [len(r) for r in raws], reverse_lanes=rev, wide=True)
        assert not status.any() and outs == raws


def test_lz4_default_preferences_are_the_reference_writers():
    if tile_io.liblz4() is None:
        pytest.skip("system liblz4 not found")
    small = tile_io.compress_frame(b"x" * 100)
    large = tile_io.compress_frame(bytes(200000))
    assert small[:4] == large[:4] == struct.pack("<I", 0x184D2204)
    assert large[4] == 0x40 and large[5] == 0x40      # v1, linked blocks, no checksums; 64 KiB


def test_lz4_against_pyarrow_codec_and_stored_frames():
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(11)
    raws = frame_corpus(rng)
    codec = pa.Codec("lz4")
    frames = [codec.compress(r, asbytes=True) for r in raws]
    frames += [tile_io.stored_frame(r) for r in raws]
    for f, r in zip(frames, raws + raws):
        assert oracle.lz4_frame_decompress(f) == r
    outs, status = emul_inflate(frames, [len(r) for r in raws + raws])
    assert not status.any()
    assert outs == raws + raws


def test_lz4_error_statuses():
    raw = (b"sample-based monte carlo denoising " * 400)
    frame = tile_io.stored_frame(raw)
    if tile_io.liblz4() is not None:
        frame = tile_io.compress_frame(raw)
    n = len(raw)
    cases = [(frame[:len(frame) // 2], n, 3),              # truncated
             (b"\0\0\0\0junk", 10, 1),                    # bad magic
             (frame, n - 1, 4),                            # overflows the expected size
             (frame, n + 1, 6),                            # shorter than expected
             (frame[:4] + bytes([0x00]) + frame[5:], n, 2),  # version bits cleared
             (frame + frame, 2 * n, 0),                    # concatenated frames
             (struct.pack("<II", 0x184D2A50, 3) + b"abc" + frame, n, 0)]  # skippable frame first
    outs, status = emul_inflate([c[0] for c in cases], [c[1] for c in cases])
    assert list(status) == [c[2] for c in cases]
    assert outs[5] == raw + raw and outs[6] == raw
    # checksums are verified like lz4.frame.decompress does: header byte always,
    # block / content xxHash32 when the frame carries them
    hdr = bytearray(tile_io.stored_frame(raw))
    hdr[6] ^= 0xFF
    sums = []
    if tile_io.liblz4() is not None:
        good = tile_io.compress_frame(raw, content_checksum=True, block_checksum=True)
        flipped = bytearray(good)
        flipped[len(good) // 2] ^= 0x01            # inside a block: malformed or checksum error
        tail = bytearray(good)
        tail[-1] ^= 0x01                           # the content checksum itself
        only_content = bytearray(tile_io.compress_frame(raw, content_checksum=True))
        only_content[20] ^= 0x01                   # a literal byte: inflates, hash differs
        sums = [bytes(flipped), bytes(tail), bytes(only_content), good]
    outs2, status2 = emul_inflate([bytes(hdr)] + sums, [n] * (1 + len(sums)))
    assert status2[0] == 8
    if sums:
        assert status2[1] != 0 and list(status2[2:]) == [8, 8, 0]
    for blob in sums[:3]:
        with pytest.raises(RuntimeError):
            tile_io.decompress_frame(blob)         # the real library rejects them too
    with pytest.raises(oracle.Lz4Error):
        oracle.lz4_frame_decompress(frame[:len(frame) // 2])
    bad = bytearray(tile_io.stored_frame(raw))
    bad[6] ^= 0xFF                                          # header checksum byte
    with pytest.raises(oracle.Lz4Error) as err:
        oracle.lz4_frame_decompress(bytes(bad))
    assert err.value.code == 8


def test_xxh32_known_answers():
    # published xxHash32 test vectors (seed 0): empty input and "abc"
    assert oracle.xxh32(b"") == 0x02CC5D05
    assert oracle.xxh32(b"abc") == 0x32D153FF


# ------------------------------------------------------------------ reader: oracle vs the reference's outputs
@pytest.mark.parametrize("name", sorted(SBMC_CONFIGS))
def test_oracle_reader_matches_reference_fixtures(name):
    full, listed, kw = SBMC_CONFIGS[name]
    exp = expected()
    files = fixture_files(listed)
    okw = oracle_kwargs(kw)
    if full:
        per_scene = 4
        assert int(exp[name + "/len"]) == len(files) // per_scene
        for s in range(len(files) // per_scene):
            bufs = [open(f, "rb").read() for f in files[s * per_scene:(s + 1) * per_scene]]
            seen = check_item(exp, "%s/%d" % (name, s), tiles_ref.read_image(bufs, **okw),
                              i_diffuse_of(kw))
            assert {"features", "radiance", "low_spp", "target_image", "spp"} <= seen
    else:
        assert int(exp[name + "/len"]) == len(files)
        for i, f in enumerate(files):
            item = tiles_ref.read_tile(open(f, "rb").read(), **okw)
            check_item(exp, "%s/%d" % (name, i), item, i_diffuse_of(kw))


# ------------------------------------------------------------------ GPU: the inflater through the C ABI
def _gpu_inflate(frames, sizes):
    table, total = layout(frames, sizes)
    dev = th.device("cuda")
    src = th.from_numpy(np.frombuffer(b"".join(frames) + b"\0", np.uint8).copy()).to(dev)
    t = th.from_numpy(table).to(dev)
    dst = th.zeros(total, dtype=th.uint8, device=dev)
    status = th.full((len(frames),), -7, dtype=th.int32, device=dev)
    before = _lib.launch_count()
    _lib.check(_lib.load().sbmc_lz4_frames_inflate(
        src.data_ptr(), t.data_ptr(), len(frames), dst.data_ptr(), status.data_ptr(),
        th.cuda.current_stream().cuda_stream), "inflate")
    th.cuda.synchronize()
    assert _lib.launch_count() == before + 1
    host = dst.cpu().numpy()
    return [host[r[2]:r[2] + r[3]].tobytes() for r in table], status.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("opts", FRAME_OPTIONS, ids=lambda o: "-".join(o) or "default")
def test_gpu_inflater_matches_oracle(opts):
    rng = np.random.default_rng(7)
    raws = frame_corpus(rng)
    if tile_io.liblz4() is not None:
        frames = [tile_io.compress_frame(r, **opts) for r in raws]
    else:
        pa = pytest.importorskip("pyarrow")
        frames = [pa.Codec("lz4").compress(r, asbytes=True) for r in raws]
    for f, r in zip(frames, raws):
        assert oracle.lz4_frame_decompress(f) == r
    outs, status = _gpu_inflate(frames, [len(r) for r in raws])
    assert not status.any(), status
    assert outs == raws


@pytest.mark.gpu
def test_gpu_inflater_error_statuses_match_the_emulation():
    raw = (b"sample-based monte carlo denoising " * 400)
    frame = tile_io.compress_frame(raw) if tile_io.liblz4() else tile_io.stored_frame(raw)
    n = len(raw)
    hdr = bytearray(tile_io.stored_frame(raw))
    hdr[6] ^= 0xFF                                   # header checksum byte
    cases = [(frame[:len(frame) // 2], n), (b"\0\0\0\0junk", 10), (frame, n - 1), (frame, n + 1),
             (frame + frame, 2 * n), (frame, n), (bytes(hdr), n)]
    expect = [3, 1, 4, 6, 0, 0, 8]
    if tile_io.liblz4() is not None:
        good = tile_io.compress_frame(raw, content_checksum=True, block_checksum=True)
        bad = bytearray(good)
        bad[-1] ^= 0x01                              # content checksum
        cases += [(good, n), (bytes(bad), n)]
        expect += [0, 8]
    _, want = emul_inflate([c[0] for c in cases], [c[1] for c in cases])
    outs, got = _gpu_inflate([c[0] for c in cases], [c[1] for c in cases])
    assert list(got) == list(want) == expect
    assert outs[4] == raw + raw and outs[5] == raw


# ------------------------------------------------------------------ the dataset classes, two backends
class EmulBackend(object):
    """Stands in for sbmc_b200.datasets._CudaBackend: the two launches of an item
    run as the device sources compiled for the host, everything else (file reads,
    chunk tables, tile tables, allocation, error handling) is the product's own
    Python code working on CPU tensors."""
    device = th.device("cpu")

    def __init__(self, device=None):
        self.lib = emul()

    def scope(self):
        import contextlib
        return contextlib.nullcontext()

    def inflate(self, comp, table, nframes, raw, status):
        self.lib.emul_lz4_frames_inflate(comp.data_ptr(), table.data_ptr(), nframes, raw.data_ptr(),
                                         status.data_ptr())

    def assemble(self, *args):
        rc = self.lib.emul_tile_assemble_f32(*args)
        if rc < 0:
            raise RuntimeError(self.lib.emul_last_error().decode())
        self.last_vec = rc


BACKENDS = ["host-emulation", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request, monkeypatch):
    """"host-emulation": product Python + device code built for the host (runs
    everywhere); "gpu": the product as shipped."""
    if request.param == "host-emulation":
        monkeypatch.setattr(datasets, "_backend", EmulBackend)
    return request.param


def launches():
    return _lib.launch_count()


ALL_CONFIGS = sorted(SBMC_CONFIGS) + ["tiles_kpcn", "full_kpcn"]


def config_of(name):
    if name in SBMC_CONFIGS:
        return SBMC_CONFIGS[name]
    full = name.startswith("full")
    return full, False, (dict(mode="kpcn", spp=2) if full else dict(mode="kpcn"))


@pytest.mark.parametrize("scalar", [False, True], ids=["vec4", "scalar"])
@pytest.mark.parametrize("name", ALL_CONFIGS)
def test_datasets_match_reference_fixtures(name, scalar, backend, monkeypatch):
    """Every item of every configuration against what the reference's own
    sbmc/datasets.py produced for the same files."""
    if scalar:
        if name.endswith("kpcn"):
            pytest.skip("same kernels as the sbmc configurations")
        monkeypatch.setattr(datasets, "_F_ALIGNED", 0)        # forces the 1-pixel kernel
    exp = expected()
    full, listed, kw = config_of(name)
    path = os.path.join(DATA, "list.txt") if listed else DATA
    d = (datasets.FullImagesDataset if full else datasets.TilesDataset)(path, **kw)
    assert len(d) == int(exp[name + "/len"])
    before = launches()
    for i in range(len(d)):
        item = d[i]
        if backend == "gpu":
            assert all(v.is_cuda for v in item.values() if isinstance(v, th.Tensor))
            assert _lib.last_path() == (2 if scalar else 1)
        item = {k: v for k, v in item.items() if k != "path"}
        check_item(exp, "%s/%d" % (name, i), item, i_diffuse_of(kw))
    if backend == "gpu":
        assert launches() >= before + 2 * len(d)          # our kernels did the work


@pytest.mark.parametrize("ts,tiles_x,tiles_y,spp", [(32, 3, 2, 4), (6, 2, 3, 2), (128, 1, 1, 8)])
def test_full_image_matches_oracle_on_larger_tiles(tmp_path, ts, tiles_x, tiles_y, spp, backend):
    """Bigger tiles (multi-block linked frames at ts = 128), a tile size that is not
    a multiple of 4 (scalar kernel), several tiles pasted by one launch."""
    if backend == "host-emulation" and ts == 128:
        spp = 2                                            # keep the CPU suite quick
    rng = np.random.default_rng(ts)
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(str(tmp_path), "scene", rng, ts, tiles_x, tiles_y, spp, quantize=1.0 / 64,
                        compress=compress)
    d = datasets.FullImagesDataset(str(tmp_path))
    item = d[0]
    files = sorted(os.listdir(tmp_path / "scene"))
    ref = tiles_ref.read_image([open(tmp_path / "scene" / f, "rb").read() for f in files])
    if backend == "gpu":
        assert _lib.last_path() == (1 if ts % 4 == 0 else 2)
    for k, v in ref.items():
        if not isinstance(v, np.ndarray):
            continue
        got = item[k].cpu().numpy()
        assert got.shape == v.shape and got.dtype == v.dtype, k
        if k == "features":
            i = 5
            rest = lambda a: np.concatenate([a[:, :i], a[:, i + 6:]], 1)  # noqa: E731
            assert np.array_equal(rest(got).view(np.int32), rest(v).view(np.int32))
            tol = LOG_RTOL * np.maximum(np.abs(v[:, i:i + 6]), 1e-3)
            assert (np.abs(got[:, i:i + 6].astype(np.float64) - v[:, i:i + 6]) <= tol).all()
        else:
            assert np.array_equal(got.view(np.int32), v.view(np.int32)), k
    # a tiles-dataset item of the same scene agrees with its region of the full image
    t = d.tiles_dset[len(d.tiles_dset) - 1]
    by, bx = t["block_y"], t["block_x"]
    assert th.equal(t["radiance"], item["radiance"][..., by:by + ts, bx:bx + ts])
    assert th.equal(t["features"], item["features"][..., by:by + ts, bx:bx + ts])


def test_row_bands_tile_the_full_image(backend):
    """Row-sharded reading: what each rank of a multi-GPU job reads (BandPlan rows +
    halo) equals those rows of the whole image; only intersecting tiles are read."""
    from sbmc_b200.sharding import BandPlan
    d = datasets.FullImagesDataset(DATA)
    whole = d[1]
    plan = BandPlan(16, 2, 5, align=4)
    bands = [(plan.y0[r] - plan.halo_top(r), plan.y1[r] + plan.halo_bot(r)) for r in range(2)]
    for lo, hi in bands + [(0, 5), (3, 11), (8, 16), (12, 13)]:
        mine = [f for f, bx, by in d.tile_positions(1) if by < hi and by + 8 > lo]
        assert (len(mine) == 2) == (hi <= 8 or lo >= 8)
        band = d.read_rows(1, lo, hi)
        assert set(band) == set(whole)
        for k, v in whole.items():
            if isinstance(v, th.Tensor) and v.dim() >= 3 and v.shape[-1] == 16:
                assert th.equal(band[k], v[..., lo:hi, :]), (k, lo, hi)
    with pytest.raises(ValueError):
        d.read_rows(0, 4, 4)
    with pytest.raises(ValueError):
        datasets.FullImagesDataset(DATA, mode="kpcn").read_rows(0, 0, 8)


def test_batched_fetch_equals_single_items(backend):
    """DataLoader batches go through TilesDataset.__getitems__ (one pair of
    launches per batch) and must equal the items fetched one by one."""
    from torch.utils.data import DataLoader
    d = datasets.TilesDataset(DATA, spp=2)
    before = launches()
    batch = next(iter(DataLoader(d, batch_size=5, shuffle=False, num_workers=0)))
    if backend == "gpu":
        assert launches() == before + 2
    assert batch["features"].shape == (5, 2, 93, 8, 8) and batch["spp"].shape == (5, 1, 1, 1)
    for i in range(5):
        one = d[i]
        for k, v in one.items():
            if isinstance(v, th.Tensor):
                assert th.equal(batch[k][i], v), (k, i)
            elif isinstance(v, str):
                assert batch[k][i] == v
            else:
                assert batch[k][i].item() == pytest.approx(v)


def _same_batch(a, b):
    assert set(a) == set(b)
    for k, v in a.items():
        if isinstance(v, th.Tensor):
            assert th.equal(v, b[k]), k
        else:
            assert v == b[k], k


def test_prefetch_loader_equals_the_dataloader(backend):
    """PrefetchLoader overlaps the next batch's file reads with the current batch:
    same batches, same order, same collation as DataLoader(num_workers=0)."""
    from torch.utils.data import DataLoader
    tiles = datasets.TilesDataset(DATA, spp=2)
    got = list(datasets.PrefetchLoader(tiles, batch_size=3))
    want = list(DataLoader(tiles, batch_size=3, shuffle=False, num_workers=0))
    assert len(got) == len(want) == len(datasets.PrefetchLoader(tiles, batch_size=3)) == 3
    for a, b in zip(got, want):
        _same_batch(a, b)
    assert len(list(datasets.PrefetchLoader(tiles, batch_size=3, drop_last=True))) == 2
    g = th.Generator().manual_seed(3)
    order = th.randperm(len(tiles), generator=th.Generator().manual_seed(3)).tolist()
    shuffled = list(datasets.PrefetchLoader(tiles, batch_size=1, shuffle=True, generator=g))
    assert [b["path"][0] for b in shuffled] == [tiles._filename(i) for i in order]
    full = datasets.FullImagesDataset(DATA)
    for a, b in zip(datasets.PrefetchLoader(full), DataLoader(full, batch_size=1, num_workers=0)):
        _same_batch(a, b)
    multi = datasets.MultiSampleCountDataset(DATA, spp=3)
    spps = [int(b["spp"].reshape(-1)[0]) for b in datasets.PrefetchLoader(multi)]
    assert spps == expected()["multi/spp_of_items"].tolist()
    kp = datasets.TilesDataset(DATA, mode="kpcn", spp=2)
    a = next(iter(datasets.PrefetchLoader(kp, batch_size=2)))
    b = next(iter(DataLoader(kp, batch_size=2, num_workers=0)))
    _same_batch(a, b)
    with pytest.raises(ValueError):
        datasets.PrefetchLoader(full, batch_size=2)
    # direct reads interleaved with a live prefetching iteration (e.g. validation
    # inside a training epoch) do not share its staging buffers
    it = iter(datasets.PrefetchLoader(tiles, batch_size=3))
    first = next(it)
    probe = tiles[7]                   # main-thread read while batch 2 is being planned
    _same_batch(first, want[0])
    _same_batch(next(it), want[1])
    assert th.equal(probe["features"], want[2]["features"][1])


def test_device_prefetch_yields_the_same_batches(backend):
    """device_prefetch = D: groups of D batches decoded by one pair of launches on a side
    stream from a background thread (GPU backend; elsewhere the plain path serves it) --
    same batches, same order; leaving the loop early stops the worker."""
    tiles = datasets.TilesDataset(DATA, spp=2)
    want = list(datasets.PrefetchLoader(tiles, batch_size=2, shuffle=True,
                                        generator=th.Generator().manual_seed(5)))
    for depth in (1, 2, 3, 16):
        got = list(datasets.PrefetchLoader(tiles, batch_size=2, shuffle=True, device_prefetch=depth,
                                           generator=th.Generator().manual_seed(5)))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            _same_batch(a, b)
    it = iter(datasets.PrefetchLoader(tiles, batch_size=1, device_prefetch=2))
    first = next(it)
    _same_batch(first, next(iter(datasets.PrefetchLoader(tiles, batch_size=1))))
    it.close()                          # generator exit: the worker thread is told to stop
    import threading
    assert not [t for t in threading.enumerate() if t.name == "sbmc-device-prefetch" and t.is_alive()]


def test_corrupt_tile_raises_like_the_reference(tmp_path, backend):
    src = fixture_files(False)[0]
    folder = tmp_path / "scene"
    folder.mkdir()
    blob = bytearray(open(src, "rb").read())
    (n0,) = struct.unpack_from("<i", blob, 60)
    blob[64 + n0 + 4:64 + n0 + 8] = b"\0\0\0\0"          # destroy the first sample frame's magic
    (folder / "bad.bin").write_bytes(bytes(blob))
    d = datasets.TilesDataset(str(tmp_path))
    with pytest.raises(RuntimeError, match="LZ4 frame 1"):
        d[0]
    with pytest.raises(ValueError, match="does not fit"):
        datasets.TilesDataset(DATA)._read_tiles([fixture_files(False)[3]], 8, 8)   # tile at (8, 8)


def test_zero_samples_and_multi_sample_count(backend):
    d0 = datasets.TilesDataset(DATA, spp=0)
    item = d0[0]
    assert "features" not in item and "radiance" not in item
    assert item["low_spp"].dtype == th.float64 and not item["low_spp"].any()
    assert item["target_image"].shape == (3, 8, 8)
    multi = datasets.MultiSampleCountDataset(DATA, spp=3)
    exp = expected()
    got = [int(multi[i]["spp"].reshape(-1)[0]) for i in range(len(multi))]
    assert got == exp["multi/spp_of_items"].tolist()
    assert multi[0]["features"].shape[0] == 2 and multi[len(multi) - 1]["features"].shape[0] == 3


# ------------------------------------------------------------------ host logic of sbmc_b200.datasets
def test_dataset_metadata_matches_the_reference():
    exp = expected()
    for name, (full, listed, kw) in SBMC_CONFIGS.items():
        path = os.path.join(DATA, "list.txt") if listed else DATA
        cls = datasets.FullImagesDataset if full else datasets.TilesDataset
        d = cls(path, **kw)
        assert len(d) == int(exp[name + "/len"])
        assert d.num_features == int(exp[name + "/num_features"])
        assert d.num_global_features == int(exp[name + "/num_global_features"])
        assert repr(d) == str(exp[name + "/repr"])
        assert "|".join(d.labels) == str(exp[name + "/labels"])
    k = datasets.TilesDataset(DATA, mode="kpcn")
    assert k.num_features == 27 and k.num_global_features == 0
    assert "|".join(k.labels) == str(exp["tiles_kpcn/labels"])
    multi = datasets.MultiSampleCountDataset(DATA, spp=3)
    assert len(multi) == int(exp["multi/len"])
    assert [ds.spp for ds in multi.datasets] == [2, 3]
    assert multi.num_features == 93 and multi.num_global_features == 3


def test_dataset_listing_modes():
    d = datasets.TilesDataset(DATA)
    assert d.io_mode == datasets.TilesDataset.FOLDERS_MODE
    assert [os.path.basename(s) for s in d.scenes] == ["scene_a", "scene_b"]
    assert d.indices[d.scenes[1]] == (4, 8)
    assert d._filename(5).endswith("scene_b_tile001.bin")
    lst = datasets.TilesDataset(os.path.join(DATA, "list.txt"))
    assert lst.io_mode == datasets.TilesDataset.FILELIST_MODE
    assert lst._filename(0).endswith("scene_b_tile003.bin")      # the list is reversed
    assert (d.tile_size, d.image_width, d.image_height, d.sample_count) == (8, 16, 16, 3)
    assert d.spp == 3 and datasets.TilesDataset(DATA, spp=2).spp == 2


def _one_tile_root(tmp_path, **kw):
    rng = np.random.default_rng(3)
    content = tile_io.synth_tile(rng, 4, 2)
    folder = tmp_path / "root" / "scene"
    folder.mkdir(parents=True)
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    (folder / "t0.bin").write_bytes(tile_io.tile_bytes(content, 4, 4, 4, 0, 0, compress=compress, **kw))
    return str(tmp_path / "root"), content


def test_dataset_error_behaviour_follows_the_reference(tmp_path):
    T = datasets.TilesDataset
    with pytest.raises(RuntimeError, match="Unknown dataset loading mode"):
        T(DATA, mode="nope")
    with pytest.raises(RuntimeError, match="Incorrect data path"):
        T(str(tmp_path / "missing"))
    (tmp_path / "empty").mkdir()
    with pytest.raises(RuntimeError, match="Empty dataset"):
        T(str(tmp_path / "empty"))
    with pytest.raises(RuntimeError, match="Requested too many samples"):
        T(DATA, spp=4)
    with pytest.raises(RuntimeError, match="spp not provided"):
        datasets.MultiSampleCountDataset(DATA)
    with pytest.raises(RuntimeError, match="spp too low"):
        datasets.MultiSampleCountDataset(DATA, spp=1)
    with pytest.raises(RuntimeError, match="folder mode"):
        datasets.FullImagesDataset(os.path.join(DATA, "list.txt"))

    def root(sub, **kw):
        (tmp_path / sub).mkdir()
        return _one_tile_root(tmp_path / sub, **kw)[0]

    with pytest.raises(ValueError, match="Version unsupported"):
        T(root("v", version=20170101))
    with pytest.raises(RuntimeError, match="Incorrect path depth"):
        T(root("d", path_depth=5))
    with pytest.raises(RuntimeError, match="focus distance"):
        T(root("f", focus_distance=-1.0))
    with pytest.raises(RuntimeError, match="aperture radius"):
        T(root("a", aperture_radius=-0.5))
    with pytest.raises(RuntimeError, match="field of view"):
        T(root("o", fov=-3.0))
    with pytest.raises(RuntimeError, match="scene radius"):
        T(root("r", scene_radius=-1.0))
    # aperture 0 => the (NaN) focus distance is replaced by 0, not rejected
    ok = T(root("n", aperture_radius=0.0, focus_distance=float("nan")))
    with open(ok._filename(0), "rb") as fid:
        assert ok._parse_header(fid.read(52))["focus_distance"] == 0.0
    short = tmp_path / "short" / "scene"
    short.mkdir(parents=True)
    (short / "t.bin").write_bytes(b"\0" * 20)
    with pytest.raises(struct.error):
        T(str(tmp_path / "short"))


def test_metadata_mismatch_between_tiles_is_rejected(tmp_path):
    root, _ = _one_tile_root(tmp_path)
    rng = np.random.default_rng(4)
    other = tile_io.synth_tile(rng, 4, 3)       # 3 samples instead of 2
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    with open(os.path.join(root, "scene", "t1.bin"), "wb") as fid:
        fid.write(tile_io.tile_bytes(other, 4, 4, 4, 0, 0, compress=compress))
    d = datasets.TilesDataset(root, device="cpu")
    with pytest.raises(ValueError, match="Metadata do not match"):
        d._plan([d._filename(1)])


def test_plan_ships_only_the_requested_chunks(tmp_path):
    d = datasets.TilesDataset(DATA, spp=1, device="cpu")
    files = fixture_files(False)[:2]
    stage, frames, tiles, raw_bytes = d._plan(files)
    assert len(frames) == 2 * (1 + 1) and len(tiles) == 2
    shipped = stage.numel()
    assert shipped < sum(os.path.getsize(f) for f in files) * 0.6      # 1 of 3 sample chunks
    for (so, n, do, want), nxt in zip(frames, frames[1:]):
        assert do % 256 == 0 and nxt[2] >= do + want
    buf = stage.numpy()
    so, n, _, want = frames[1]
    assert len(oracle.lz4_frame_decompress(buf[so:so + n].tobytes())) == want
    with open(files[0], "r+b") as fid:
        pass
    trunc = tmp_path / "s"
    trunc.mkdir()
    shutil.copy(files[0], trunc / "t.bin")
    with open(trunc / "t.bin", "r+b") as fid:
        fid.truncate(os.path.getsize(files[0]) // 3)
    with pytest.raises((RuntimeError, struct.error)):
        d._plan([str(trunc / "t.bin")])


@pytest.mark.skipif(th.cuda.is_available(), reason="checks the behaviour without a GPU")
def test_reading_an_item_without_cuda_raises():
    d = datasets.TilesDataset(DATA)
    with pytest.raises(_lib.SbmcB200Error, match="no CPU data path"):
        d[0]


def test_tile_assemble_argument_validation_needs_no_gpu():
    lib = _lib.load()
    args = [None, None, 1, 0, 8, 3, 27, 30, 6, 31, None, None, None, None, None, None, 8, 8, 0,
            None]
    assert lib.sbmc_tile_assemble_f32(*args) == -1 and b"null" in lib.sbmc_b200_last_error()
    args[6] = 26
    assert lib.sbmc_tile_assemble_f32(*args) == -1 and b"27" in lib.sbmc_b200_last_error()
    args[6], args[2] = 27, 0
    assert lib.sbmc_tile_assemble_f32(*args) == 0          # no tiles: nothing to do
    assert lib.sbmc_lz4_frames_inflate(None, None, 0, None, None, None) == 0
    assert lib.sbmc_lz4_frames_inflate(None, None, 2, None, None, None) == -1


# ------------------------------------------------------------------ GPU: reader -> model
@pytest.mark.gpu
def test_gpu_dataset_feeds_the_denoiser():
    """FullImagesDataset -> DataLoader -> Multisteps, the denoise.py call chain
    (scripts/denoise.py:113-114,150-157 of the reference)."""
    from torch.utils.data import DataLoader
    from sbmc_b200 import models
    d = datasets.FullImagesDataset(DATA, spp=2)
    loader = DataLoader(d, batch_size=1, shuffle=False, num_workers=0)
    batch = next(iter(loader))
    assert batch["features"].shape == (1, 2, 93, 16, 16) and batch["features"].is_cuda
    assert batch["global_features"].shape == (1, 3, 1, 1)
    th.manual_seed(0)
    model = models.Multisteps(d.num_features, d.num_global_features, ksize=3, nsteps=1,
                              width=16, embedding_width=16).cuda().eval()
    with th.no_grad():
        out = model(batch)["radiance"]
    assert out.shape == (1, 3, 14, 14) and bool(th.isfinite(out).all())
