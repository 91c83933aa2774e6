"""Oracle (test infrastructure): the slice of diffusers' `VideoProcessor` / `VaeImageProcessor` that the
reference pipeline calls.  Used ONLY by tests/golden/_reference_shim.py as the stand-in for the absent
`diffusers.video_processor.VideoProcessor` when the reference's own `__call__` is executed to produce goldens,
so that the product's `aether_b200.pipeline.VideoProcessor` is compared with an independent restatement and
not with itself.

Reference call sites (/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py):
  :459  self.video_processor.preprocess(image, height, width)                    numpy / list-of-numpy input
  :474-496  self.video_processor.preprocess(x, height, width, resize_mode="crop")  PIL input
  :932  self.video_processor.postprocess_video(video=rgb_video, output_type="np")

PARITY UNPINNED (third-party diffusers >= 0.32.2, image_processor.py / video_processor.py): restated from the
published algorithm --
  numpy input  [N, H, W, C] in [0, 1] -> torch [N, C, H, W] (`numpy_to_pt`: transpose(0, 3, 1, 2)), resized with
               `F.interpolate(size)` (nearest) when (H, W) differs from the multiple-of-`vae_scale_factor` target,
               then `normalize`: 2x - 1;
  PIL input    `resize(..., resize_mode="crop")` = scale to cover, centre paste on a black canvas, LANCZOS;
  postprocess  per batch item: [C, F, H, W] -> [F, C, H, W] -> denormalize (x / 2 + 0.5).clamp(0, 1)
               -> `pt_to_numpy` (cpu, permute(0, 2, 3, 1), float().numpy()) -> stacked [B, F, H, W, C].
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class OracleVideoProcessor:
    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = int(vae_scale_factor)

    # ------------------------------------------------------------------ helpers (diffusers VaeImageProcessor)
    @staticmethod
    def _numpy_to_pt(images: np.ndarray) -> torch.Tensor:
        if images.ndim == 3:
            images = images[..., None]
        return torch.from_numpy(images.transpose(0, 3, 1, 2))

    @staticmethod
    def _cover_resize_pil(image, width: int, height: int):
        import PIL.Image
        want = width / height
        have = image.width / image.height
        w_scaled = width if want > have else image.width * height // image.height
        h_scaled = height if want <= have else image.height * width // image.width
        scaled = image.resize((w_scaled, h_scaled), resample=PIL.Image.LANCZOS)
        canvas = PIL.Image.new("RGB", (width, height))
        canvas.paste(scaled, box=(width // 2 - w_scaled // 2, height // 2 - h_scaled // 2))
        return canvas

    def _target_size(self, t: torch.Tensor, height, width):
        h = t.shape[-2] if height is None else height
        w = t.shape[-1] if width is None else width
        return h - h % self.vae_scale_factor, w - w % self.vae_scale_factor

    # ------------------------------------------------------------------ public surface used by the reference
    def preprocess(self, image, height=None, width=None, resize_mode: str = "default") -> torch.Tensor:
        try:
            import PIL.Image
            pil_type = PIL.Image.Image
        except Exception:  # pragma: no cover
            pil_type = ()
        if isinstance(image, pil_type):
            image = [image]
        if isinstance(image, (list, tuple)) and len(image) and isinstance(image[0], pil_type):
            frames = []
            for im in image:
                im = im.convert("RGB")
                if resize_mode == "crop":
                    im = self._cover_resize_pil(im, width, height)
                elif resize_mode == "default":
                    import PIL.Image
                    im = im.resize((width, height), resample=PIL.Image.LANCZOS)
                else:
                    raise ValueError(f"resize_mode {resize_mode} is not supported")
                frames.append(np.asarray(im, dtype=np.float32) / 255.0)
            t = self._numpy_to_pt(np.stack(frames, axis=0))
        else:
            if isinstance(image, (list, tuple)):
                image = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            elif image.ndim == 3:
                image = image[None]
            t = self._numpy_to_pt(np.asarray(image))
        th, tw = self._target_size(t, height, width)
        if tuple(t.shape[-2:]) != (th, tw):
            t = F.interpolate(t, size=(th, tw))
        return 2.0 * t - 1.0

    def postprocess_video(self, video: torch.Tensor, output_type: str = "np") -> np.ndarray:
        if output_type != "np":
            raise ValueError("the reference only asks for output_type='np' (:932)")
        per_item = []
        for item in video:                                     # [C, F, H, W]
            frames = item.permute(1, 0, 2, 3)                  # [F, C, H, W]
            frames = (frames / 2 + 0.5).clamp(0, 1)
            per_item.append(frames.cpu().permute(0, 2, 3, 1).float().numpy())
        return np.stack(per_item)
