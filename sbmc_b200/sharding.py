"""H-sharding of the kernel-weighting path across GPUs (one process per GPU).

The reference is single-device; its only mechanism for
large images is overlap-tiling with recompute (scripts/denoise.py:54-93).  Every
output pixel of KernelWeighting depends on a K x K neighbourhood only, so the
image is cut into contiguous row bands, one per rank, and instead of recomputing
the overlap the ranks exchange it:

* forward / d_weights read `data` rows up to c0 = (K-1)//2 above and K-1-c0
  below the band  ->  each rank receives those rows from its two neighbours
  (`exchange_halo`: one NCCL all-gather of every rank's edge rows per call,
  [world, 2, B, C, pad, W] floats -- KBs to a few MB);
* d_data is a scatter: a band's samples contribute to rows up to K-1-c0 above
  and c0 below the band  ->  each rank computes its partial d_data on band+halo
  rows and the halo rows are sent to the neighbours, which add them
  (`reduce_halo`).  The K*K-channel weight tensor never crosses the link.

The exchange is written against torch.distributed collectives only, so it
runs over NCCL/NVLink on GPUs and over gloo on CPU tensors (tests/test_sharding.py
drives it with world_size 2 on CPU).  The compute goes through the row-band
entry points of the C ABI (sbmc_kernel_weighting_{fwd,bwd}_band_f32).
"""
import torch as th
import torch.distributed as dist

from . import _lib

__all__ = ["BandPlan", "exchange_halo", "reduce_halo", "kernel_weighting_fwd_sharded",
           "kernel_weighting_bwd_sharded", "ShardedKernelWeighting", "gather_bands",
           "HaloPipeline", "kernel_weighting_fwd_band", "kernel_weighting_bwd_band",
           "multisteps_forward_sharded", "multisteps_forward_halo", "halo_mode_rows"]


class BandPlan:
    """Contiguous row bands of an image of `height` rows over `world` ranks.

    Rows are split as evenly as possible (the first `height % world` bands get
    one extra row).  `pad` = K-1-c0 >= c0 is the halo kept on both sides of a
    band (clamped at the image border), which covers the rows read by
    forward / d_weights and the rows reached by d_data for odd and even K.
    """

    def __init__(self, height, world, kh, kw=None, pad=None, align=1):
        if height < world * align:
            raise ValueError("cannot split %d rows over %d ranks" % (height, world))
        self.height, self.world, self.kh, self.kw = height, world, kh, kw or kh
        c0 = (kh - 1) // 2
        # `pad` overrides the halo width (e.g. the receptive field of a U-net);
        # `align` makes every band start on a multiple of `align` rows (the U-net's
        # 2x2 poolings need band origins that are multiples of 4)
        self.pad = max(c0, kh - 1 - c0) if pad is None else pad
        units = (height + align - 1) // align
        base, extra = divmod(units, world)
        self.y0, self.y1 = [], []
        y = 0
        for r in range(world):
            rows = (base + (1 if r < extra else 0)) * align
            self.y0.append(min(y, height))
            y += rows
            self.y1.append(min(y, height))
        if world > 1 and min(b - a for a, b in zip(self.y0, self.y1)) < self.pad:
            raise ValueError("bands (%d rows) are shorter than the halo (%d rows)"
                             % (base, self.pad))

    def rows(self, rank):
        return self.y1[rank] - self.y0[rank]

    def halo_top(self, rank):
        return min(self.pad, self.y0[rank])

    def halo_bot(self, rank):
        return min(self.pad, self.height - self.y1[rank])

    def band(self, rank, full, dim):
        """The slice of a full-image tensor that `rank` owns along `dim`."""
        return full.narrow(dim, self.y0[rank], self.rows(rank))


def _all_gather(x, group=None):
    """[...] -> [world, ...]: one NCCL / gloo all-gather, stream-ordered (it never
    blocks the host, unlike grouped send/recv whose wait() was measured to stall
    the launch queue: profiles/r1j_shard_timing.txt)."""
    world = dist.get_world_size(group)
    x = x.contiguous()
    out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))   # concatenated form
    dist.all_gather_into_tensor(out, x, group=group)
    return out.view((world,) + tuple(x.shape))


def exchange_halo(plan, rank, band, group=None):
    """band [..., rows, W] -> ext [..., top + rows + bot, W] with the neighbours'
    rows in the halos (image borders have no halo).  Every rank publishes its
    first and last `pad` rows in ONE all-gather and picks its neighbours' edges
    (2 * pad rows per rank: KBs to a few MB, the K*K weight volume stays put)."""
    top, bot, pad = plan.halo_top(rank), plan.halo_bot(rank), plan.pad
    rows = band.shape[-2]
    ext = band.new_empty(band.shape[:-2] + (top + rows + bot, band.shape[-1]))
    ext.narrow(-2, top, rows).copy_(band)
    if plan.world == 1:
        return ext
    edges = th.stack([band.narrow(-2, 0, pad), band.narrow(-2, rows - pad, pad)])
    allv = _all_gather(edges, group)               # [world, 2, ..., pad, W]
    if top:   # the upper neighbour's last rows
        ext.narrow(-2, 0, top).copy_(allv[rank - 1, 1].narrow(-2, pad - top, top))
    if bot:   # the lower neighbour's first rows
        ext.narrow(-2, top + rows, bot).copy_(allv[rank + 1, 0].narrow(-2, 0, bot))
    return ext


def reduce_halo(plan, rank, ext, out=None, group=None):
    """Adjoint of exchange_halo: ext [..., top + rows + bot, W] holds this rank's
    partial sums for its band and for the neighbours' edge rows; the halo rows
    are published in one all-gather and every rank adds its neighbours' halos to
    its own edge rows.  Returns the band [..., rows, W] (written into `out` if
    given)."""
    top, bot, pad = plan.halo_top(rank), plan.halo_bot(rank), plan.pad
    rows = ext.shape[-2] - top - bot
    band = ext.narrow(-2, top, rows)
    if out is None:
        out = band.clone()
    else:
        out.copy_(band)
    if plan.world == 1:
        return out
    halos = ext.new_zeros((2,) + tuple(ext.shape[:-2]) + (pad, ext.shape[-1]))
    if top:   # my contribution to the upper neighbour's last rows
        halos[0].narrow(-2, pad - top, top).copy_(ext.narrow(-2, 0, top))
    if bot:   # ... and to the lower neighbour's first rows
        halos[1].narrow(-2, 0, bot).copy_(ext.narrow(-2, top + rows, bot))
    allv = _all_gather(halos, group)               # [world, 2, ..., pad, W]
    if top:   # the upper neighbour's bottom halo lands on my first rows
        cnt = plan.halo_bot(rank - 1)
        out.narrow(-2, 0, cnt).add_(allv[rank - 1, 1].narrow(-2, 0, cnt))
    if bot:   # the lower neighbour's top halo lands on my last rows
        cnt = plan.halo_top(rank + 1)
        out.narrow(-2, rows - cnt, cnt).add_(allv[rank + 1, 0].narrow(-2, pad - cnt, cnt))
    return out


class HaloPipeline:
    """The halo exchange of a STREAM of KernelWeighting calls on one band, taken off
    the critical path ("on a side stream overlapped with interior
    tiles"; round-1 VERDICT: 0.21 ms per call of eager glue on the compute stream).

    * every buffer is allocated once: the caller keeps `data` and `d_data` in their
      halo-extended layout (`new_ext()`; the band is the view `band(ext)`), so there is
      no 44 MB band copy per call, and the edge staging / all-gather buffers are reused;
    * `exchange_async` / `reduce_async` run on a side stream: stage 2 x pad edge rows,
      ONE all-gather, write (or add) the neighbours' rows -- the compute stream only
      waits on the returned event, so the exchange of call s + 1 and the reduction of
      call s overlap the kernels of the neighbouring calls.

    On CPU tensors (gloo, tests) everything runs synchronously in program order."""

    def __init__(self, plan, rank, band_shape, device, dtype=th.float32, group=None):
        self.plan, self.rank, self.group = plan, rank, group
        self.top, self.bot, self.pad = plan.halo_top(rank), plan.halo_bot(rank), plan.pad
        self.rows, self.width = band_shape[-2], band_shape[-1]
        self.lead = tuple(band_shape[:-2])
        self.device, self.dtype = th.device(device), dtype
        self.cuda = self.device.type == "cuda"
        self.side = th.cuda.Stream(self.device) if self.cuda else None
        edge = (2,) + self.lead + (self.pad, self.width)
        kw = dict(device=self.device, dtype=dtype)
        self._edges = [th.empty(edge, **kw), th.empty(edge, **kw)]            # fwd, bwd staging
        self._all = [th.empty((plan.world,) + edge, **kw), th.empty((plan.world,) + edge, **kw)]

    def new_ext(self):
        return th.empty(self.lead + (self.top + self.rows + self.bot, self.width),
                        device=self.device, dtype=self.dtype)

    def band(self, ext):
        """The rows of an extended tensor this rank owns (a view)."""
        return ext.narrow(-2, self.top, self.rows)

    def _gather(self, which):
        out = self._all[which]
        dist.all_gather_into_tensor(out.view((-1,) + tuple(out.shape[2:])), self._edges[which],
                                    group=self.group)
        return out

    def _on_side(self, fn):
        if not self.cuda:
            fn()
            return None
        ready = th.cuda.Event()
        ready.record(th.cuda.current_stream(self.device))
        with th.cuda.stream(self.side):
            self.side.wait_event(ready)
            fn()
            done = th.cuda.Event()
            done.record(self.side)
        return done

    def wait(self, done):
        """Makes the current stream wait for an exchange / reduction."""
        if done is not None:
            th.cuda.current_stream(self.device).wait_event(done)

    def exchange_async(self, ext):
        """Fills the halo rows of `ext` with the neighbours' rows (its band rows must
        have been written on the current stream).  Returns an event to `wait` on."""
        if self.plan.world == 1:
            return None
        top, bot, pad, rows, rank = self.top, self.bot, self.pad, self.rows, self.rank

        def run():
            band = self.band(ext)
            self._edges[0][0].copy_(band.narrow(-2, 0, pad))
            self._edges[0][1].copy_(band.narrow(-2, rows - pad, pad))
            allv = self._gather(0)
            if top:
                ext.narrow(-2, 0, top).copy_(allv[rank - 1, 1].narrow(-2, pad - top, top))
            if bot:
                ext.narrow(-2, top + rows, bot).copy_(allv[rank + 1, 0].narrow(-2, 0, bot))
        return self._on_side(run)

    def reduce_async(self, ext):
        """Adjoint: `ext` holds partial sums for the band and for the neighbours' edge
        rows; adds the neighbours' halo rows into the band rows of `ext` in place
        (`band(ext)` is the result).  Returns an event to `wait` on."""
        if self.plan.world == 1:
            return None
        top, bot, pad, rows, rank, plan = self.top, self.bot, self.pad, self.rows, self.rank, self.plan

        def run():
            if top:
                self._edges[1][0].narrow(-2, pad - top, top).copy_(ext.narrow(-2, 0, top))
            if bot:
                self._edges[1][1].narrow(-2, 0, bot).copy_(ext.narrow(-2, top + rows, bot))
            allv = self._gather(1)
            band = self.band(ext)
            if top:
                cnt = plan.halo_bot(rank - 1)
                band.narrow(-2, 0, cnt).add_(allv[rank - 1, 1].narrow(-2, 0, cnt))
            if bot:
                cnt = plan.halo_top(rank + 1)
                band.narrow(-2, rows - cnt, cnt).add_(allv[rank + 1, 0].narrow(-2, pad - cnt, cnt))
        return self._on_side(run)


def kernel_weighting_fwd_band(plan, rank, data_ext, weights, output, sum_w):
    """KernelWeighting forward on this rank's band, halo-extended `data_ext`
    [B,C,top+rows+bot,W] already exchanged (HaloPipeline.exchange_async)."""
    n, kh, kw, rows, w = weights.shape
    c = data_ext.shape[1]
    lib = _lib.load()
    with th.cuda.device(weights.device):
        _lib.check(lib.sbmc_kernel_weighting_fwd_band_f32(
            data_ext.data_ptr(), weights.data_ptr(), output.data_ptr(), sum_w.data_ptr(),
            n, c, rows, w, kh, kw, plan.halo_top(rank), plan.halo_bot(rank),
            _stream(weights)), "kernel_weighting (band)")


def kernel_weighting_bwd_band(plan, rank, data_ext, weights, d_output, d_sum_w, d_data_ext,
                              d_weights):
    """KernelWeighting backward on this rank's band: d_weights complete, d_data_ext
    [B,C,top+rows+bot,W] = this rank's partial sums (HaloPipeline.reduce_async adds
    the neighbours')."""
    n, kh, kw, rows, w = weights.shape
    c = data_ext.shape[1]
    lib = _lib.load()
    with th.cuda.device(weights.device):
        _lib.check(lib.sbmc_kernel_weighting_bwd_band_f32(
            data_ext.data_ptr(), weights.data_ptr(), d_output.data_ptr(),
            d_sum_w.data_ptr(), d_data_ext.data_ptr(), d_weights.data_ptr(),
            n, c, rows, w, kh, kw, plan.halo_top(rank), plan.halo_bot(rank),
            _stream(weights)), "kernel_weighting_grad (band)")


def _stream(t):
    return th.cuda.current_stream(t.device).cuda_stream


def kernel_weighting_fwd_sharded(plan, rank, data, weights, output, sum_w, group=None,
                                 data_ext=None):
    """KernelWeighting forward on this rank's band.  data [B,C,rows,W] and weights
    [B,K,K,rows,W] are the band's slices; output / sum_w are caller-allocated
    (reference ownership rule, sbmc/functions.py:91-94).  Returns data_ext so the
    backward pass can reuse it."""
    if data_ext is None:
        data_ext = exchange_halo(plan, rank, data, group)
    n, c, rows, w = data.shape
    _, kh, kw, _, _ = weights.shape
    lib = _lib.load()
    with th.cuda.device(data.device):
        _lib.check(lib.sbmc_kernel_weighting_fwd_band_f32(
            data_ext.data_ptr(), weights.data_ptr(), output.data_ptr(), sum_w.data_ptr(),
            n, c, rows, w, kh, kw, plan.halo_top(rank), plan.halo_bot(rank),
            _stream(data)), "kernel_weighting (band)")
    return data_ext


def kernel_weighting_bwd_sharded(plan, rank, data, weights, d_output, d_sum_w, d_data,
                                 d_weights, group=None, data_ext=None):
    """KernelWeighting backward on this rank's band (see module docstring)."""
    if data_ext is None:
        data_ext = exchange_halo(plan, rank, data, group)
    n, c, rows, w = data.shape
    _, kh, kw, _, _ = weights.shape
    top, bot = plan.halo_top(rank), plan.halo_bot(rank)
    d_data_ext = th.empty_like(data_ext)
    lib = _lib.load()
    with th.cuda.device(data.device):
        _lib.check(lib.sbmc_kernel_weighting_bwd_band_f32(
            data_ext.data_ptr(), weights.data_ptr(), d_output.data_ptr(),
            d_sum_w.data_ptr(), d_data_ext.data_ptr(), d_weights.data_ptr(),
            n, c, rows, w, kh, kw, top, bot, _stream(data)), "kernel_weighting_grad (band)")
    reduce_halo(plan, rank, d_data_ext, out=d_data, group=group)
    return d_data


class ShardedKernelWeighting(th.autograd.Function):
    """`KernelWeighting` (sbmc/functions.py:74-115) on one row band per rank:
    same (output, sum_w) / (d_data, d_weights) contract for the band's slices."""

    @staticmethod
    def forward(ctx, data, weights, plan, rank, group=None):
        data = data.contiguous()
        weights = weights.contiguous()
        n, c, rows, w = data.shape
        output = th.empty_like(data)
        sum_w = data.new_empty((n, rows, w))
        data_ext = kernel_weighting_fwd_sharded(plan, rank, data, weights, output, sum_w, group)
        ctx.save_for_backward(data, weights, data_ext)
        ctx.plan, ctx.rank, ctx.group = plan, rank, group
        return output, sum_w

    @staticmethod
    def backward(ctx, d_output, d_sum_w):
        data, weights, data_ext = ctx.saved_tensors
        d_data = th.empty_like(data)
        d_weights = th.empty_like(weights)
        kernel_weighting_bwd_sharded(ctx.plan, ctx.rank, data, weights,
                                     d_output.contiguous(), d_sum_w.contiguous(), d_data,
                                     d_weights, ctx.group, data_ext)
        return d_data, d_weights, None, None, None


def gather_bands(plan, rank, band, dim, group=None):
    """Final gather: every rank ends with the full tensor (bands concatenated
    along `dim`).  Bands may differ by one row, so this is an all_gather of
    padded slices followed by a trim."""
    world = plan.world
    rows_max = max(plan.rows(r) for r in range(world))
    moved = band.movedim(dim, 0).contiguous()
    padded = moved.new_zeros((rows_max,) + moved.shape[1:])
    padded[:moved.shape[0]].copy_(moved)
    parts = [th.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    full = th.cat([p[:plan.rows(r)] for r, p in enumerate(parts)], dim=0)
    return full.movedim(0, dim).contiguous()


# -- tiled multi-GPU inference of the whole model (BASELINE.json config 5) ---------
def model_band_rows(plan, rank, overlap):
    """Rows [a, b) of the full image that `rank` feeds to the model: its band plus
    `overlap` rows of context on each side (clamped to the image), and the rows
    [top, top + rows) of the model's *input tile* that the band occupies."""
    a = max(0, plan.y0[rank] - overlap)
    b = min(plan.height, plan.y1[rank] + overlap)
    return a, b, plan.y0[rank] - a


def multisteps_forward_sharded(model, samples, rank, world, overlap=144, group=None):
    """Denoise one (batch of) full image(s) with `world` ranks, one row band each.

    The reference handles large images by overlap-tiling with recompute
    (scripts/denoise.py:54-93, 256-pixel pad); here the tiles are the ranks' row
    bands.  `samples` holds the FULL-image tensors ("radiance" [bs, spp, 3, H, W],
    "features" [bs, spp, nf, H, W], "global_features") on every rank -- only the
    band plus `overlap` rows of context are moved to the GPU and pushed through
    the model; the context covers the receptive field of the three U-nets and the
    K x K splat (~127 px), so band rows see exactly what the
    unsharded forward sees.  `overlap` must be a multiple of 4 (pooling
    alignment) as must the band origins.  The denoised bands are all-gathered:
    every rank returns the full {"radiance": [bs, 3, H - K + 1, W - K + 1]}.
    """
    radiance = samples["radiance"]
    height = radiance.shape[-2]
    crop = (model.ksize - 1) // 2
    if overlap % 4 or overlap < crop:
        raise ValueError("overlap must be a multiple of 4 and >= (K-1)/2")
    if height % 4:
        # the U-net pools twice: with H % 4 != 0 MaxPool2d floors and the bilinear
        # upsampling rescales by the BAND's height, so band rows would differ from the
        # unsharded forward (pad the image to a multiple of 4 first)
        raise ValueError("tiled inference needs an image height that is a multiple of 4, "
                         "got %d" % height)
    plan = BandPlan(height, world, model.ksize, align=4)
    a, b, top = model_band_rows(plan, rank, overlap)
    dev = th.device("cuda", th.cuda.current_device()) if th.cuda.is_available() \
        else radiance.device
    tile = {"radiance": radiance[..., a:b, :].to(dev),
            "features": samples["features"][..., a:b, :].to(dev),
            "global_features": samples["global_features"].to(dev)}
    out = model(tile)["radiance"]                  # [bs, 3, (b - a) - 2 crop, W - 2 crop]
    # rows of the cropped output that belong to this band (full-image output row
    # y_out = y - crop): the band's rows, minus the image-border ring
    y_lo = max(plan.y0[rank], crop)
    y_hi = min(plan.y1[rank], height - crop)
    band = out[..., y_lo - a - crop:y_hi - a - crop, :].contiguous()
    if world == 1:
        return {"radiance": band}
    counts = [min(plan.y1[r], height - crop) - max(plan.y0[r], crop) for r in range(world)]
    rows_max = max(counts)
    moved = band.movedim(-2, 0).contiguous()
    padded = moved.new_zeros((rows_max,) + tuple(moved.shape[1:]))
    padded[:moved.shape[0]].copy_(moved)
    parts = moved.new_empty((world * rows_max,) + tuple(moved.shape[1:]))
    dist.all_gather_into_tensor(parts, padded, group=group)
    parts = parts.view((world, rows_max) + tuple(moved.shape[1:]))
    full = th.cat([parts[r, :counts[r]] for r in range(world)], dim=0)
    return {"radiance": full.movedim(0, -2).contiguous()}


def halo_mode_rows(height, world, ksize, rank):
    """Image rows [a, b) that `rank` needs as input in halo mode: its band plus
    the K x K splat halo.  A rank can read just these from disk
    (`FullImagesDataset.read_rows`) and pass `image_height=height, row0=a`."""
    plan = BandPlan(height, world, ksize, align=4)
    return plan.y0[rank] - plan.halo_top(rank), plan.y1[rank] + plan.halo_bot(rank)


def multisteps_forward_halo(model, samples, rank, world, unet_pad=64, group=None,
                            image_height=None, row0=0):
    """Tiled inference with halo EXCHANGE instead of full overlap recompute
    (inference pipeline of `Multisteps._forward_nhwc`, needs `model.bf16_chains`).

    Per rank: the per-sample 1x1 chains, the kernel regressor and the fused splat
    run on the band plus (K-1)/2 rows (everything a band pixel's K x K splat
    neighbourhood needs); before each of the `nsteps` U-nets the ranks exchange
    `unet_pad` rows of the pixel-domain tensor `reduced` with their neighbours
    (one NCCL all-gather of the band edges per step, [bs, 128 ch, unet_pad, W]),
    so the U-net sees real neighbour rows instead of recomputing them through the
    whole network; `unet_pad` only has to cover ONE U-net's receptive field
    (~39 px) because every step exchanges again.  Final all-gather of the bands.
    Same result as the unsharded forward up to the U-net's zero padding at
    distance >= unet_pad (exact when unet_pad covers the receptive field).

    `samples` holds the full image by default; with `image_height` / `row0` it
    holds only rows [row0, row0 + samples_rows) of an image of `image_height`
    rows, which must cover `halo_mode_rows(...)` of this rank.
    """
    from . import conv1x1 as _c
    from . import unet_fast as _u
    from ._compat import crop_like
    radiance, features = samples["radiance"], samples["features"]
    bs, spp, nf, local_rows, w = features.shape
    height = local_rows if image_height is None else image_height
    k = model.ksize
    crop = (k - 1) // 2
    if unet_pad % 4 or not model._nhwc_pipeline_ok(nf):
        raise ValueError("halo mode needs unet_pad % 4 == 0 and the 128-wide 1x1 chains")
    if height % 4:
        raise ValueError("tiled inference needs an image height that is a multiple of 4, "
                         "got %d (see multisteps_forward_sharded)" % height)
    plan = BandPlan(height, world, k, align=4)                 # K x K halo (splat)
    uplan = BandPlan(height, world, k, pad=unet_pad, align=4)  # U-net halo
    dev = th.device("cuda", th.cuda.current_device())
    y0, y1 = plan.y0[rank], plan.y1[rank]
    top, bot = plan.halo_top(rank), plan.halo_bot(rank)
    a, b = y0 - top, y1 + bot                          # rows of the chain / splat domain
    rows_ext, rows = b - a, y1 - y0
    if a < row0 or b > row0 + local_rows:
        raise ValueError("samples cover rows [%d, %d), rank %d needs [%d, %d)"
                         % (row0, row0 + local_rows, rank, a, b))
    rad = radiance[..., a - row0:b - row0, :].to(dev)
    feats = _c.to_nhwc_bf16(features[..., a - row0:b - row0, :].to(dev))  # [bs, spp, rows_ext*w, 128]
    gf = samples["global_features"].to(dev).reshape(bs, -1).float()
    hw = rows_ext * w
    prop, ca = None, nf
    utop, ubot = uplan.halo_top(rank), uplan.halo_bot(rank)
    own_convs = getattr(model, "own_convs", True)
    for step in range(model.nsteps):
        embed = getattr(model, "embedding_{:02d}".format(step))
        # all samples in one launch; the sample mean comes out of the same kernel
        new, reduced = _c.chain_samples_nhwc(embed, feats, ca, prop=prop,
                                             gf=gf if step == 0 else None, want_mean=True)
        feats, ca = new, 128
        band = reduced.view(bs, rows_ext, w * 128)[:, top:top + rows]
        ext = exchange_halo(uplan, rank, band, group)              # [bs, utop+rows+ubot, w*128]
        x = ext.view(bs, utop + rows + ubot, w, 128).permute(0, 3, 1, 2)
        y = _u.autoencoder_forward(getattr(model, "propagation_{:02d}".format(step)), x,
                                   own_convs=own_convs)
        y = y.permute(0, 2, 3, 1)[:, utop - top:utop + rows + bot]  # back to the ext rows
        prop = y.to(th.bfloat16).contiguous().view(bs, hw, 128)
    sum_r = sum_w = max_w = None
    k2 = k * k
    group_n = max(2, min(spp, int((4 << 30) // max(1, bs * k2 * hw * 4)) // 2 * 2))
    for s0 in range(0, spp, group_n):
        ns = min(group_n, spp - s0)
        logits = _c.chain_samples_nhwc(model.kernel_regressor, feats, 128, prop=prop,
                                       regress=True, sample0=s0, nsamples=ns)
        for i in range(ns):
            kernels = logits[:, i].view(bs, k2, rows_ext, w)
            sum_r, sum_w, max_w = model.kernel_update(
                crop_like(rad[:, s0 + i], kernels), kernels, sum_r, sum_w, max_w)
        del logits
    out = sum_r / (sum_w + model.eps)                  # [bs, 3, rows_ext, w]
    y_lo, y_hi = max(y0, crop), min(y1, height - crop)
    band_out = out[..., y_lo - a:y_hi - a, crop:w - crop].contiguous()
    if world == 1:
        return {"radiance": band_out}
    counts = [min(plan.y1[r], height - crop) - max(plan.y0[r], crop) for r in range(world)]
    rows_max = max(counts)
    moved = band_out.movedim(-2, 0).contiguous()
    padded = moved.new_zeros((rows_max,) + tuple(moved.shape[1:]))
    padded[:moved.shape[0]].copy_(moved)
    parts = moved.new_empty((world * rows_max,) + tuple(moved.shape[1:]))
    dist.all_gather_into_tensor(parts, padded, group=group)
    parts = parts.view((world, rows_max) + tuple(moved.shape[1:]))
    full = th.cat([parts[r, :counts[r]] for r in range(world)], dim=0)
    return {"radiance": full.movedim(0, -2).contiguous()}
