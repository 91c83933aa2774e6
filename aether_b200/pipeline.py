"""`AetherV1PipelineCogVideoX` -- host-side mirror of the reference pipeline class, same public surface.

Reference: /root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py (class at :255-965; cited per
method below).  The reference subclasses diffusers' CogVideoXImageToVideoPipeline; diffusers is not
available in this image, so the few inherited members the reference relies on (SURVEY.md A.4:
`vae_scale_factor_*`, `decode_latents`, `prepare_extra_step_kwargs`, `video_processor.preprocess /
postprocess_video`, `_execution_device`, `progress_bar`, `maybe_free_model_hooks`, `guidance_scale`,
`interrupt`) are restated here.  Names, argument meaning, defaults and error behaviour follow the
reference so its call sites (scripts/demo.py:565-630, evaluation/video_depth/launch_aether.py:151-158) work
unchanged against this class.

The three module objects (`transformer`, `vae`, `scheduler`) are duck-typed exactly like the reference
does; with the aether_b200 CUDA modules the denoise-loop body (:832-916) collapses to two C-ABI calls per
step (`aether_dit_forward`, `aether_cfg_dpm_step`) and no host synchronisation (the reference's per-step
`t.item()` at :886 is replaced by a host copy of the timesteps taken once).
"""
from __future__ import annotations

import inspect
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

try:  # PIL is optional at import time; only PIL inputs need it
    import PIL.Image
except Exception:  # pragma: no cover
    PIL = None

from .rope import prepare_rotary_positional_embeddings


# ---------------------------------------------------------------------------------------------------------
# reference aether/utils/preprocess_utils.py:4-39 (imcrop_center / crop): aspect-preserving centre crop
# ---------------------------------------------------------------------------------------------------------
def _crop(img, start_h, start_w, crop_h, crop_w):
    out = np.zeros((crop_h, crop_w, *img.shape[2:]), dtype=img.dtype)
    hsize, wsize = crop_h, crop_w
    dh, dw, sh, sw = start_h, start_w, 0, 0
    if dh < 0:
        sh = -dh
        hsize += dh
        dh = 0
    if dh + hsize > img.shape[0]:
        hsize = img.shape[0] - dh
    if dw < 0:
        sw = -dw
        wsize += dw
        dw = 0
    if dw + wsize > img.shape[1]:
        wsize = img.shape[1] - dw
    out[sh:sh + hsize, sw:sw + wsize] = img[dh:dh + hsize, dw:dw + wsize]
    return out


def imcrop_center(img_list, crop_p_h, crop_p_w):
    new = []
    for im in img_list:
        if crop_p_h / crop_p_w > im.shape[0] / im.shape[1]:
            start_h = 0
            start_w = int((im.shape[1] - im.shape[0] / crop_p_h * crop_p_w) / 2)
            size = (im.shape[0], int(im.shape[0] / crop_p_h * crop_p_w))
        else:
            start_h = int((im.shape[0] - im.shape[1] / crop_p_w * crop_p_h) / 2)
            start_w = 0
            size = (int(im.shape[1] / crop_p_w * crop_p_h), im.shape[1])
        new.append(_crop(im, start_h, start_w, size[0], size[1]))
    return new


class VideoProcessor:
    """The slice of diffusers' VideoProcessor / VaeImageProcessor the reference uses (:459, :474-496, :932)."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    def _pil_resize_crop(self, image, width, height):
        ratio = width / height
        src_ratio = image.width / image.height
        src_w = width if ratio > src_ratio else image.width * height // image.height
        src_h = height if ratio <= src_ratio else image.height * width // image.width
        resized = image.resize((src_w, src_h), resample=PIL.Image.LANCZOS)
        res = PIL.Image.new("RGB", (width, height))
        res.paste(resized, box=(width // 2 - src_w // 2, height // 2 - src_h // 2))
        return res

    def preprocess(self, image, height=None, width=None, resize_mode: str = "default", device=None) -> torch.Tensor:
        """-> float32 [N, 3, H, W] in [-1, 1].  With a CUDA `device` the frames are uploaded first and the layout
        change and the 2x-1 map run there (same IEEE operations, same values; ~0.2 s less host work per 41-frame clip)."""
        if PIL is not None and isinstance(image, PIL.Image.Image):
            image = [image]
        if isinstance(image, list) and PIL is not None and isinstance(image[0], PIL.Image.Image):
            out = []
            for im in image:
                im = im.convert("RGB")
                if resize_mode == "crop":
                    im = self._pil_resize_crop(im, width, height)
                else:
                    im = im.resize((width, height), resample=PIL.Image.LANCZOS)
                out.append(np.array(im).astype(np.float32) / 255.0)
            arr = np.stack(out, axis=0)
        elif isinstance(image, list):
            arr = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
        else:
            arr = image if image.ndim == 4 else image[None]
        if arr.ndim == 3:
            arr = arr[..., None]
        if device is not None and torch.device(device).type == "cuda" and arr.dtype in (np.float32, np.float64):
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(device).permute(0, 3, 1, 2).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 3, 1, 2)))
        h = height or t.shape[2]
        w = width or t.shape[3]
        h, w = h - h % self.vae_scale_factor, w - w % self.vae_scale_factor
        if (t.shape[2], t.shape[3]) != (h, w):
            t = torch.nn.functional.interpolate(t, size=(h, w))
        return 2.0 * t - 1.0

    @staticmethod
    def postprocess_video(video: torch.Tensor, output_type: str = "np"):
        """video [B, C, F, H, W] in [-1, 1] -> float32 [B, F, H, W, C] in [0, 1]: numpy ("np", like the reference :932)
        or a tensor on the video's device ("pt")."""
        outs = []
        for b in range(video.shape[0]):
            v = video[b].permute(1, 0, 2, 3)
            v = (v / 2 + 0.5).clamp(0, 1)
            # layout change and widening on the tensor's own device (exact), then one contiguous copy to the host
            v = v.permute(0, 2, 3, 1).float().contiguous()
            outs.append(v if output_type == "pt" else v.cpu().numpy())
        return torch.stack(outs) if output_type == "pt" else np.stack(outs)


@dataclass
class AetherV1PipelineOutput:   # reference :248-252
    rgb: np.ndarray
    disparity: np.ndarray
    raymap: np.ndarray


def retrieve_latents(encoder_output, generator=None, sample_mode: str = "sample"):   # reference :233-245
    if hasattr(encoder_output, "latent_dist") and sample_mode == "sample":
        return encoder_output.latent_dist.sample(generator)
    if hasattr(encoder_output, "latent_dist") and sample_mode == "argmax":
        return encoder_output.latent_dist.mode()
    if hasattr(encoder_output, "latents"):
        return encoder_output.latents
    raise AttributeError("Could not access latents of provided encoder_output")


class _Progress:
    def __init__(self, total, enabled):
        self.enabled = enabled
        self.bar = None
        if enabled:
            try:
                from tqdm.auto import tqdm
                self.bar = tqdm(total=total)
            except Exception:
                self.bar = None

    def __enter__(self):
        return self

    def update(self):
        if self.bar is not None:
            self.bar.update()

    def __exit__(self, *a):
        if self.bar is not None:
            self.bar.close()


class AetherV1PipelineCogVideoX:
    _supported_tasks = ["reconstruction", "prediction", "planning"]
    _default_num_inference_steps = {"reconstruction": 4, "prediction": 50, "planning": 50}     # :257-261
    _default_guidance_scale = {"reconstruction": 1.0, "prediction": 3.0, "planning": 3.0}      # :262-266
    _default_use_dynamic_cfg = {"reconstruction": False, "prediction": True, "planning": True}  # :267-271
    _base_fps = 12

    def __init__(self, tokenizer=None, text_encoder=None, vae=None, scheduler=None, transformer=None,
                 empty_prompt_embeds: Optional[torch.Tensor] = None, progress: bool = False):
        """Same five modules as the reference (:274-288).  The reference runs T5 on the empty prompt once here
        (:290-297); pass `empty_prompt_embeds` ([1, 226, 4096]) instead of tokenizer/text_encoder to skip T5."""
        self.tokenizer, self.text_encoder = tokenizer, text_encoder
        self.vae, self.scheduler, self.transformer = vae, scheduler, transformer
        self.vae_scale_factor_spatial = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.vae_scale_factor_temporal = vae.config.temporal_compression_ratio if vae is not None else 4
        self.vae_scaling_factor_image = vae.config.scaling_factor if vae is not None else 0.7
        self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)
        self._progress = progress
        self._device = None
        self._guidance_scale = None
        self._interrupt = False
        self._current_timestep = None
        self._attention_kwargs = None
        self._num_timesteps = 0
        if empty_prompt_embeds is None:
            if tokenizer is None or text_encoder is None:
                raise ValueError("either (tokenizer, text_encoder) or empty_prompt_embeds has to be provided")
            empty_prompt_embeds = self.encode_prompt("")
        self.empty_prompt_embeds = empty_prompt_embeds.to(dtype=torch.bfloat16)

    # ------------------------------------------------------------------ inherited-surface restatements
    @torch.no_grad()
    def encode_prompt(self, prompt: str, max_sequence_length: int = 226):
        ti = self.tokenizer(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                            add_special_tokens=True, return_tensors="pt")
        dev = next(self.text_encoder.parameters()).device
        return self.text_encoder(ti.input_ids.to(dev))[0]

    def to(self, device):
        self._device = torch.device(device)
        for m in (self.vae, self.transformer, self.text_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(self._device)
        return self

    @property
    def _execution_device(self):
        if self._device is not None:
            return self._device
        dev = getattr(self.transformer, "device", None)
        return torch.device(dev) if dev is not None else torch.device("cpu")

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def num_timesteps(self):
        return self._num_timesteps

    def progress_bar(self, total):
        return _Progress(total, self._progress)

    def maybe_free_model_hooks(self):
        pass

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        latents = latents.permute(0, 2, 1, 3, 4)          # [B, C, F, H, W]
        latents = 1 / self.vae_scaling_factor_image * latents
        return self.vae.decode(latents).sample

    # ------------------------------------------------------------------ reference :299-348
    def _prepare_rotary_positional_embeddings(self, height, width, num_frames, device, fps=None):
        c = self.transformer.config
        if c.patch_size_t is not None:
            raise NotImplementedError("CogVideoX-1.5 rotary branch (:333-346) is not used by AetherV1")
        return prepare_rotary_positional_embeddings(
            height, width, num_frames, patch_size=c.patch_size, vae_scale_factor_spatial=self.vae_scale_factor_spatial,
            sample_height=c.sample_height, sample_width=c.sample_width, attention_head_dim=c.attention_head_dim,
            base_fps=self._base_fps, fps=fps, device=device)

    # ------------------------------------------------------------------ reference :350-449 (same messages)
    def check_inputs(self, task, image, video, goal, raymap, height, width, num_frames, fps):
        is_pil = lambda x: PIL is not None and isinstance(x, PIL.Image.Image)
        if task not in self._supported_tasks:
            raise ValueError(f"`task` has to be one of {self._supported_tasks}.")
        if image is None and video is None:
            raise ValueError("`image` or `video` has to be provided.")
        if image is not None and video is not None:
            raise ValueError("`image` and `video` cannot both be provided.")
        if image is not None:
            if task == "reconstruction":
                raise ValueError("`image` is not supported for `reconstruction` task.")
            if not isinstance(image, (torch.Tensor, np.ndarray)) and not is_pil(image):
                raise ValueError("`image` has to be of type `torch.Tensor` or `np.ndarray` or `PIL.Image.Image` but is"
                                 f" {type(image)}")
        if goal is not None:
            if task != "planning":
                raise ValueError("`goal` is only supported for `planning` task.")
            if not isinstance(goal, (torch.Tensor, np.ndarray)) and not is_pil(goal):
                raise ValueError("`goal` has to be of type `torch.Tensor` or `np.ndarray` or `PIL.Image.Image` but is"
                                 f" {type(goal)}")
        if video is not None:
            if task != "reconstruction":
                raise ValueError("`video` is only supported for `reconstruction` task.")
            if not isinstance(video, (torch.Tensor, np.ndarray)) and not (
                    isinstance(video, list) and all(is_pil(v) for v in video)):
                raise ValueError("`video` has to be of type `torch.Tensor` or `np.ndarray` or `List[PIL.Image.Image]` "
                                 f"but is {type(video)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if num_frames is None:
            raise ValueError("`num_frames` is required.")
        if num_frames not in [17, 25, 33, 41]:
            raise ValueError("`num_frames` has to be one of [17, 25, 33, 41].")
        if fps not in [8, 10, 12, 15, 24]:
            raise ValueError("`fps` has to be one of [8, 10, 12, 15, 24].")
        if raymap is not None and not isinstance(raymap, (torch.Tensor, np.ndarray)):
            raise ValueError("`raymap` has to be of type `torch.Tensor` or `np.ndarray`.")
        if raymap is not None:
            want = (num_frames, 6, height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial)
            if tuple(raymap.shape[-4:]) != want:
                raise ValueError(f"`raymap` shape is not correct. Expected {want}, got {raymap.shape}.")

    # ------------------------------------------------------------------ reference :451-512
    def _preprocess_image(self, image, height, width, device=None):
        if (isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 4 and device is not None
                and torch.device(device).type == "cuda" and tuple(image.shape[1:]) == (height, width, 3)):
            image = torch.from_numpy(np.ascontiguousarray(image)).to(device)      # 3 bytes per pixel cross PCIe
        if (isinstance(image, torch.Tensor) and image.is_cuda and image.dtype == torch.uint8 and image.ndim == 4
                and tuple(image.shape[1:]) == (height, width, 3) and height % self.vae_scale_factor_spatial == 0
                and width % self.vae_scale_factor_spatial == 0 and image.stride(3) == 1 and image.stride(2) == 3):
            # uint8 frames resident on the device at the target size: `/ 255` (:454-455), the identity crop / resize,
            # the layout change, 2x - 1 and the bf16 cast in ONE kernel (bit-identical to the host path for every
            # pixel value, tests/test_input_gpu.py)
            from . import ops
            return ops.u8_frames_to_model_input(image)
        if (isinstance(image, torch.Tensor) and image.is_cuda and image.is_floating_point() and image.ndim == 4
                and tuple(image.shape[1:]) == (height, width, 3) and height % self.vae_scale_factor_spatial == 0
                and width % self.vae_scale_factor_spatial == 0):
            # frames already resident on the device at the target size [F, H, W, 3]: the centre crop (:451-458) and
            # the resize are identities, what remains is the layout change and 2x - 1 in the frames' own dtype --
            # the same IEEE operations the host path performs, without the reference's .cpu().numpy() round trip
            return 2.0 * image.permute(0, 3, 1, 2) - 1.0
        if isinstance(image, torch.Tensor):
            image = image.cpu().numpy()
        if image.dtype == np.uint8:
            image = image.astype(np.float32) / 255.0
        if image.ndim == 3:
            image = [image]
        image = imcrop_center(image, height, width)
        return self.video_processor.preprocess(image, height, width, device=device)

    def preprocess_inputs(self, image, goal, video, raymap, height, width, num_frames):
        dev = self._execution_device
        is_pil = lambda x: PIL is not None and isinstance(x, PIL.Image.Image)

        def one(x):
            if x is None:
                return None
            if is_pil(x) or (isinstance(x, list) and all(is_pil(v) for v in x)):
                t = self.video_processor.preprocess(x, height, width, resize_mode="crop", device=dev)
            else:
                t = self._preprocess_image(x, height, width, device=dev)
            return t.to(device=dev, dtype=torch.bfloat16)

        image, goal, video = one(image), one(goal), one(video)
        if raymap is not None:
            if isinstance(raymap, np.ndarray):
                raymap = torch.from_numpy(raymap).to(dev, dtype=torch.bfloat16)
            if raymap.ndim == 4:
                raymap = raymap.unsqueeze(0).to(dev, dtype=torch.bfloat16)
        return image, goal, video, raymap

    # ------------------------------------------------------------------ reference :514-688
    @torch.no_grad()
    def prepare_latents(self, image=None, goal=None, video=None, raymap=None, batch_size: int = 1,
                        num_frames: int = 13, height: int = 60, width: int = 90, dtype=None, device=None,
                        generator=None):
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}. Make sure the batch size matches the length of "
                             "the generators.")
        sf_t, sf_s = self.vae_scale_factor_temporal, self.vae_scale_factor_spatial
        num_frames = (num_frames - 1) // sf_t + 1
        shape = (batch_size, num_frames, 56, height // sf_s, width // sf_s)
        scale = self.vae_scaling_factor_image
        invert = getattr(self.vae.config, "invert_scale_latents", False)

        def encode(x5):   # x5: [B, C, F, H, W]
            if isinstance(generator, list):
                lat = [retrieve_latents(self.vae.encode(x5[i].unsqueeze(0)), generator[i]) for i in range(batch_size)]
            else:
                lat = [retrieve_latents(self.vae.encode(x.unsqueeze(0)), generator) for x in x5]
            lat = torch.cat(lat, dim=0).to(dtype).permute(0, 2, 1, 3, 4)      # [B, F, C, H, W]
            return scale * lat if not invert else 1 / scale * lat

        image_latents = goal_latents = video_latents = None
        if image is not None:
            image_latents = encode(image.unsqueeze(2))
        if goal is not None:
            goal_latents = encode(goal.unsqueeze(2))
        if video is not None:
            if video.ndim == 4:
                video = video.unsqueeze(0)
            video_latents = encode(video.permute(0, 2, 1, 3, 4))

        if image is not None and goal is None:
            pad = torch.zeros((batch_size, num_frames - image_latents.shape[1], *image_latents.shape[2:]),
                              device=device, dtype=dtype)
            condition_latents = torch.cat([image_latents, pad], dim=1)
        elif goal is not None:
            pad = torch.zeros((batch_size, num_frames - goal_latents.shape[1] - image_latents.shape[1],
                               *image_latents.shape[2:]), device=device, dtype=dtype)
            condition_latents = torch.cat([image_latents, pad, goal_latents], dim=1)
        elif video is not None:
            condition_latents = video_latents

        if raymap is not None:
            if raymap.shape[1] % sf_t != 0:
                raymap = torch.cat([raymap[:, : sf_t - raymap.shape[1] % sf_t], raymap], dim=1)
            # einops "b (n t) c h w -> b t (n c) h w", n = sf_t   (:666-670)
            b, nt, c, h, w = raymap.shape
            t = nt // sf_t
            camera_conditions = raymap.reshape(b, sf_t, t, c, h, w).permute(0, 2, 1, 3, 4, 5).reshape(b, t, sf_t * c, h, w)
        else:
            camera_conditions = torch.zeros(batch_size, num_frames, 24, height // sf_s, width // sf_s, device=device,
                                            dtype=dtype)
        condition_latents = torch.cat([condition_latents, camera_conditions], dim=2)
        # diffusers randn_tensor(shape, generator, device, dtype)
        if isinstance(generator, list):
            latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=device, dtype=dtype)
                                 for g in generator], dim=0)
        else:
            gdev = generator.device if generator is not None else device
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        latents = latents * self.scheduler.init_noise_sigma
        return latents, condition_latents

    # ------------------------------------------------------------------ reference :690-965
    @torch.no_grad()
    def __call__(self, task: Optional[str] = None, image=None, video=None, goal=None,
                 raymap: Optional[Union[torch.Tensor, np.ndarray]] = None, height: Optional[int] = None,
                 width: Optional[int] = None, num_frames: Optional[int] = None,
                 num_inference_steps: Optional[int] = None, timesteps: Optional[List[int]] = None,
                 guidance_scale: Optional[float] = None, use_dynamic_cfg: bool = False,
                 num_videos_per_prompt: int = 1, eta: float = 0.0, generator=None, return_dict: bool = True,
                 attention_kwargs: Optional[Dict] = None, fps: Optional[int] = None, output_latents: bool = False,
                 output_type: str = "np", decode_rgb: bool = True):
        """Reference signature (:691-712) plus three extensions that default to the reference behaviour:
        `output_latents` (return the denoised latents, tests), `output_type="pt"` (rgb / disparity / raymap stay
        float32 CUDA tensors -- the sliding-window path keeps per-tile disparities on the device), `decode_rgb=False`
        (skip the rgb VAE decode; rgb is then None -- the sliding-window evaluation only uses the rgb of tile 0)."""
        if output_type not in ("np", "pt"):
            raise ValueError("`output_type` has to be 'np' or 'pt'.")
        if task is None:
            task = "reconstruction" if video is not None else ("planning" if goal is not None else "prediction")
        tc = self.transformer.config
        height = height or tc.sample_height * self.vae_scale_factor_spatial
        width = width or tc.sample_width * self.vae_scale_factor_spatial
        num_frames = num_frames or tc.sample_frames
        fps = fps or self._base_fps
        num_videos_per_prompt = 1

        self.check_inputs(task=task, image=image, video=video, goal=goal, raymap=raymap, height=height, width=width,
                          num_frames=num_frames, fps=fps)
        image, goal, video, raymap = self.preprocess_inputs(image=image, goal=goal, video=video, raymap=raymap,
                                                            height=height, width=width, num_frames=num_frames)
        self._guidance_scale = guidance_scale
        self._current_timestep = None
        self._attention_kwargs = attention_kwargs
        self._interrupt = False
        batch_size = 1
        device = self._execution_device
        prompt_embeds = self.empty_prompt_embeds.to(device)

        num_inference_steps = num_inference_steps or self._default_num_inference_steps[task]
        guidance_scale = guidance_scale or self._default_guidance_scale[task]
        use_dynamic_cfg = use_dynamic_cfg or self._default_use_dynamic_cfg[task]
        do_cfg = guidance_scale > 1.0

        # retrieve_timesteps (:167-229): custom timesteps only if the scheduler supports them
        if timesteps is not None:
            if "timesteps" not in set(inspect.signature(self.scheduler.set_timesteps).parameters.keys()):
                raise ValueError(f"The current scheduler class {self.scheduler.__class__}'s `set_timesteps` does not "
                                 "support custom timestep schedules. Please check whether you are using the correct "
                                 "scheduler.")
            self.scheduler.set_timesteps(timesteps=timesteps, device=device)
            num_inference_steps = len(self.scheduler.timesteps)
        else:
            self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        self._num_timesteps = len(timesteps)
        t_host = [int(v) for v in timesteps.tolist()]       # one D2H here instead of t.item() every step (:886)

        latents, condition_latents = self.prepare_latents(image, goal, video, raymap,
                                                          batch_size * num_videos_per_prompt, num_frames, height,
                                                          width, prompt_embeds.dtype, device, generator)
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        image_rotary_emb = (self._prepare_rotary_positional_embeddings(height, width, latents.size(1), device, fps=fps)
                            if tc.use_rotary_positional_embeddings else None)
        ofs_emb = None if tc.ofs_embed_dim is None else latents.new_full((1,), fill_value=2.0)

        # loop-invariant part of :839-855: the (un)conditional condition latents
        if do_cfg:
            lc = self.vae.config.latent_channels if self.vae is not None else 16
            uncond = condition_latents.clone()
            if task == "planning":
                assert goal is not None
                uncond[:, :, :lc] = 0
            elif task == "prediction":
                uncond[:, :1, :lc] = 0
            else:
                raise ValueError(f"Task {task} not supported for classifier-free guidance.")
            latent_condition = torch.cat([uncond, condition_latents])
        else:
            latent_condition = condition_latents
        n_cfg = 2 if do_cfg else 1
        text = prompt_embeds.repeat(n_cfg, 1, 1)
        fused = hasattr(self.scheduler, "step_fused")
        # the reference dispatches on isinstance(scheduler, CogVideoXDPMScheduler) (:902); without diffusers the same
        # decision is taken on the capability that class adds: the `old_pred_original_sample` / `timestep_back` arguments
        dpm_signature = "old_pred_original_sample" in inspect.signature(self.scheduler.step).parameters
        # with the aether_b200 transformer the concat / repeat / expand of :832-869 happen inside the patch-gather
        # kernel (aether_dit_forward_split): the loop body is then exactly two C-ABI calls
        split = fused and hasattr(self.transformer, "forward_split") and ofs_emb is None
        num_warmup_steps = max(len(timesteps) - num_inference_steps * self.scheduler.order, 0)

        with self.progress_bar(total=num_inference_steps) as progress_bar:
            old_pred_original_sample = None
            for i, t in enumerate(timesteps):
                if self.interrupt:
                    continue
                self._current_timestep = t
                if split:
                    noise_pred = self.transformer.forward_split(latents, latent_condition, prompt_embeds, t.reshape(1),
                                                                image_rotary_emb)
                else:
                    lmi = torch.cat([latents] * 2) if do_cfg else latents
                    lmi = self.scheduler.scale_model_input(lmi, t)
                    lmi = torch.cat([lmi, latent_condition], dim=2)
                    timestep = t.expand(lmi.shape[0])
                    noise_pred = self.transformer(hidden_states=lmi, encoder_hidden_states=text, timestep=timestep,
                                                  ofs=ofs_emb, image_rotary_emb=image_rotary_emb,
                                                  attention_kwargs=attention_kwargs, return_dict=False)[0]
                if use_dynamic_cfg:   # :879-893, python-float arithmetic on the raw timestep value
                    self._guidance_scale = 1 + guidance_scale * (
                        (1 - math.cos(math.pi * ((num_inference_steps - t_host[i]) / num_inference_steps) ** 5.0)) / 2)
                else:
                    self._guidance_scale = guidance_scale
                t_back = timesteps[i - 1] if i > 0 else None
                if fused:
                    latents, old_pred_original_sample = self.scheduler.step_fused(
                        noise_pred, self.guidance_scale if do_cfg else 1.0, old_pred_original_sample, t_host[i],
                        t_host[i - 1] if i > 0 else None, latents, generator=extra_step_kwargs.get("generator"))
                else:
                    noise_pred = noise_pred.float()
                    if do_cfg:
                        nu, ntxt = noise_pred.chunk(2)
                        noise_pred = nu + self.guidance_scale * (ntxt - nu)
                    if dpm_signature:     # reference :902-915: CogVideoXDPMScheduler vs. DDIM-style signature
                        latents, old_pred_original_sample = self.scheduler.step(
                            noise_pred, old_pred_original_sample, t, t_back, latents, **extra_step_kwargs,
                            return_dict=False)
                    else:
                        latents = self.scheduler.step(noise_pred, t, latents, **extra_step_kwargs,
                                                      return_dict=False)[0]
                    latents = latents.to(prompt_embeds.dtype)
                if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % self.scheduler.order == 0):
                    progress_bar.update()
        self._current_timestep = None
        if output_latents:
            return latents

        lc = self.vae.config.latent_channels
        rgb_latents = latents[:, :, :lc]
        disparity_latents = latents[:, :, lc:lc * 2]
        camera_latents = latents[:, :, lc * 2:]

        to_host = output_type == "np"
        rgb_video = None
        if decode_rgb:
            rgb_video = self.decode_latents(rgb_latents)
            rgb_video = self.video_processor.postprocess_video(video=rgb_video, output_type=output_type)
        disparity_video = self.decode_latents(disparity_latents)
        n_out_frames = disparity_video.shape[2]
        disparity_video = disparity_video.mean(dim=1, keepdim=False)
        disparity_video = disparity_video * 0.5 + 0.5
        disparity_video = torch.square(disparity_video)
        disparity_video = disparity_video.float()
        if to_host:
            disparity_video = disparity_video.cpu().numpy()
        # einops "b t (n c) h w -> b (n t) c h w", n = 4, keep the last F frames   (:942-949)
        b, tl, nc, h, w = camera_latents.shape
        rm = camera_latents.reshape(b, tl, 4, nc // 4, h, w).permute(0, 2, 1, 3, 4, 5).reshape(b, 4 * tl, nc // 4, h, w)
        raymap_out = rm[:, -n_out_frames:, :, :].float()
        if to_host:
            raymap_out = raymap_out.cpu().numpy()
        self.maybe_free_model_hooks()
        if not return_dict:
            return (rgb_video, disparity_video, raymap_out)
        sq = lambda a: None if a is None else a.squeeze(0)
        return AetherV1PipelineOutput(rgb=sq(rgb_video), disparity=sq(disparity_video), raymap=sq(raymap_out))
