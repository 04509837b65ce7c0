"""Training-time display of denoising results: the reference's
`DenoisingDisplayCallback` (sbmc/callbacks.py:28-60) builds a gallery (low-spp
input, output, target, |difference|, stacked vertically, tonemapped) and hands it
to a Visdom server through `ttools.ImageDisplayCallback`.  Visdom and ttools are
external and absent: the gallery is the reference's, the sink is a PNG file per
`frequency` steps (sbmc_b200.imageio).
"""
import os

import torch as th

from . import imageio
from ._compat import crop_like

__all__ = ["DenoisingDisplayCallback"]


class DenoisingDisplayCallback(object):
    """Writes `<out_dir>/<win>_<step>.png` every `frequency` training steps.
    `env` / `port` are accepted for signature parity with the Visdom version."""

    def __init__(self, frequency=100, out_dir=None, env=None, port=None, win="images"):
        self.frequency, self.out_dir, self.win = frequency, out_dir, win
        self.step = 0

    def caption(self, batch, fwd_result):
        spp = batch["spp"][0].item()
        return "vertically: %dspp, ours, target, difference" % spp

    def visualized_image(self, batch, fwd_result):
        output = fwd_result["radiance"].detach()
        lowspp = crop_like(batch["low_spp"].detach().to(output.device), output)
        target = crop_like(batch["target_image"].detach().to(output.device), output)
        diff = (output - target).abs()
        data = th.cat([lowspp, output, target, diff], -2)
        data = th.clamp(data, 0)                 # clip and tonemap (callbacks.py:53-57)
        data = data / (1 + data)
        data = th.pow(data, 1.0 / 2.2)
        return th.clamp(data, 0, 1)

    def batch_end(self, batch, fwd_result, bwd_result):
        self.step += 1
        if self.out_dir is None or self.frequency <= 0 or self.step % self.frequency:
            return None
        gallery = self.visualized_image(batch, fwd_result)       # [bs, 3, 4h, w]
        image = th.cat(list(gallery), -1).permute(1, 2, 0)        # batch side by side
        os.makedirs(self.out_dir, exist_ok=True)
        path = os.path.join(self.out_dir, "%s_%06d.png" % (self.win, self.step))
        imageio.write_png(path, (image * 255).to(th.uint8).cpu().numpy())
        return path
