"""Inference entry with the reference's API (face_replace/inference/test.py:38-163):

    predictor = Predictor(checkpoint_path)
    pred_image, visualization, attn_probs = predictor.predict(input_img, cond_imgs, target_img=None)

`checkpoint_path` is a reference training checkpoint: `torch.load` dict {'state_dict', 'cfg'[, 'optimizer']}
(written by face_replace/training/coach.py:712-718) with the whole FaceReplaceModel under the `net.` prefix
(`net.module.` under DDP, stripped like test.py:49): net.unet.*, net.original_unet.*, net.vae.*, net.original_vae.*,
net.text_encoder.*. LoRA leaves (peft key layout) are merged at load.

The caption encoding is a constant of the model (pix2pix_turbo.py:100-106) but is NOT stored in the state_dict; it is
taken from, in order: the `caption_enc` argument, a 'caption_enc' entry of the checkpoint dict, or recomputed with
transformers' CLIPTextModel from net.text_encoder.* when the sd-turbo tokenizer files are available locally.
"""
from __future__ import annotations

import logging
from pathlib import Path
from types import SimpleNamespace
from typing import Any, List, Optional

import numpy as np
import torch
from PIL import Image

from . import _lib as L
from .pipeline import ModelFlags, RestorePipeline

PROMPT = "A high-quality photo of a person; professional, 8k"       # pix2pix_turbo.py:100
MODEL_NAME = "stabilityai/sd-turbo"                                    # pix2pix_turbo.py:17


def decode_cfg(cfg: Any):
    """`pyrallis.decode(TrainConfig, ckpt['cfg'])` of the reference (test.py:43) without pyrallis: nested mapping ->
    face_replace.configs.train_config.TrainConfig (typed `model` / `data` sections with the reference defaults,
    configs/train_config.py:94-147; training-only sections kept loosely)."""
    from face_replace.configs.train_config import TrainConfig, decode
    return decode(TrainConfig, cfg)


def split_state_dict(sd: dict) -> dict:
    """'net.<part>.<key>' (optionally with DDP's '.module.') -> {part: {key: tensor}}."""
    parts: dict = {}
    for k, v in sd.items():
        k = k.replace(".module.", ".")
        if k.startswith("module."):
            k = k[len("module."):]
        if k.startswith("net."):
            k = k[len("net."):]
        head, _, rest = k.partition(".")
        parts.setdefault(head, {})[rest] = v
    return parts


def image_to_tensor(img: Image.Image, size: int = 512) -> torch.Tensor:
    """test.py:54-59: Resize(size, LANCZOS) (shorter side), CenterCrop(size), ToTensor, Normalize(0.5, 0.5) -> [-1, 1]."""
    img = img.convert("RGB")
    w, h = img.size
    if (w, h) != (size, size):
        if w <= h:
            nw, nh = size, max(size, int(size * h / w))
        else:
            nw, nh = max(size, int(size * w / h)), size
        img = img.resize((nw, nh), Image.LANCZOS)
        left, top = int(round((nw - size) / 2.0)), int(round((nh - size) / 2.0))
        img = img.crop((left, top, left + size, top + size))
    x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)
    return x.sub_(0.5).div_(0.5)


def tensor2im(var: torch.Tensor, unnorm: bool = False) -> Image.Image:
    """face_replace/training/utils/vis_utils.py:14-23, in the tensor's own dtype like the reference's in-place ops (an
    fp16 prediction is scaled, shifted and multiplied by 255 with fp16 rounding at every step), truncation to uint8."""
    var = var.detach().cpu().clone()
    if unnorm:
        var *= 0.5
        var += 0.5
    arr = var.permute(1, 2, 0).contiguous().numpy()
    arr[arr < 0] = 0
    arr[arr > 1] = 1
    arr *= 255
    return Image.fromarray(arr.astype("uint8"))


def _caption_from_text_encoder(text_encoder_sd: dict, device) -> Optional[torch.Tensor]:
    try:
        from transformers import AutoTokenizer, CLIPTextConfig, CLIPTextModel
        tok = AutoTokenizer.from_pretrained(MODEL_NAME, subfolder="tokenizer", local_files_only=True)
        cfg = CLIPTextConfig.from_pretrained(MODEL_NAME, subfolder="text_encoder", local_files_only=True)
        enc = CLIPTextModel(cfg)
        enc.load_state_dict({k: v.float() for k, v in text_encoder_sd.items()}, strict=False)
        ids = tok(PROMPT, max_length=tok.model_max_length, padding="max_length", truncation=True, return_tensors="pt").input_ids
        with torch.no_grad():
            return enc.eval()(ids)[0]
    except Exception as e:  # noqa: BLE001
        logging.info("caption encoding could not be recomputed (%s)", e)
        return None


class FaceReplaceModel:
    """`face_replace.models.face_replace_model.FaceReplaceModel` of the reference (:8-45) for inference: holds the
    configuration and, once `load_state_dict` has seen the checkpoint, `.net` — the B200 RestorePipeline standing in
    for `Pix2Pix_Turbo` (same forward signature, `noise_timesteps`, `unet.attn_processors`). Geometry (UNet / VAE
    widths) is read from the checkpoint, so the released SD-Turbo checkpoints and reduced test models load alike."""

    def __init__(self, cfg, full_cfg=None, evaluating: bool = False, device="cuda:0", use_cuda_graph: bool = True,
                 caption_enc: Optional[torch.Tensor] = None):
        self.cfg, self.full_cfg = cfg, full_cfg
        if cfg.net_type != "pix2pix_turbo":
            raise ValueError(f"Invalid encoder type: {cfg.net_type}")
        self.device = torch.device(device)
        self.use_cuda_graph = use_cuda_graph
        self.caption_enc = caption_enc
        self.net: Optional[RestorePipeline] = None

    def load_state_dict(self, state_dict: dict, strict: bool = True, caption_enc: Optional[torch.Tensor] = None):
        from .unet_engine import UNetSpec
        from .weights import infer_unet_geometry, infer_vae_channels
        parts = split_state_dict(state_dict)
        need = ("unet", "original_unet", "vae", "original_vae") if self.cfg.use_shared_attention else ("unet", "vae")
        missing = [f"net.{n}.*" for n in need if n not in parts]
        if missing and strict:
            raise KeyError(f"checkpoint has no {', '.join(missing)} weights")
        cap = caption_enc if caption_enc is not None else self.caption_enc
        if cap is None and "text_encoder" in parts:
            cap = _caption_from_text_encoder(parts["text_encoder"], self.device)
        if cap is None:
            raise RuntimeError("caption_enc is not in the checkpoint and the sd-turbo tokenizer/text-encoder config are "
                               "not available locally: pass caption_enc=<(1,77,1024) tensor>")
        m = self.cfg
        flags = ModelFlags(use_shared_attention=m.use_shared_attention, use_adain=m.use_adain, train_input=m.train_input,
                           condition_on_face_embeds=m.condition_on_face_embeds, lora_rank_unet=m.lora_rank_unet)
        spec = UNetSpec(**infer_unet_geometry(parts["unet"]))
        self.net = RestorePipeline(parts["unet"], parts.get("original_unet"), parts["vae"], parts.get("original_vae"), cap,
                                   flags, spec=spec, vae_block_out_channels=infer_vae_channels(parts["vae"]),
                                   use_shortcuts=m.use_shortcuts, device=self.device, noise_timestep=m.noise_timestep,
                                   use_cuda_graph=self.use_cuda_graph)
        return SimpleNamespace(missing_keys=[], unexpected_keys=[])

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self

    def forward(self, *args, **kwargs):
        return self.net.forward(*args, **kwargs)

    __call__ = forward


class Predictor:
    logging.basicConfig(level=logging.INFO)

    def __init__(self, checkpoint_path: Path, caption_enc: Optional[torch.Tensor] = None, device="cuda:0",
                 use_cuda_graph: bool = True):
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        self.cfg = decode_cfg(ckpt.get("cfg"))
        if caption_enc is None:
            caption_enc = ckpt.get("caption_enc")
        self.device = torch.device(device)
        self.face_replace_model = FaceReplaceModel(cfg=self.cfg.model, full_cfg=self.cfg, device=device,
                                                   use_cuda_graph=use_cuda_graph, caption_enc=caption_enc)
        logging.info("Moving model to GPU")
        self.face_replace_model.load_state_dict(ckpt["state_dict"], strict=True)   # '.module.' stripped like test.py:49
        self.face_replace_model.eval()
        self.max_conditioning_images = self.cfg.data.max_conditioning_images
        self.face_replace_model.net.noise_timesteps = [249]             # test.py:62
        self.dtype = torch.float16
        # pre / post-processing on the GPU (bit-identical to the PIL / torchvision path, tests/test_preprocess.py)
        from .preprocess import GpuPreprocessor
        self.pre = GpuPreprocessor(self.device)
        self._pinned: dict = {}

    @classmethod
    def from_pipeline(cls, net: RestorePipeline, max_conditioning_images: int = 4, cfg: Any = None) -> "Predictor":
        """A Predictor around an already built RestorePipeline (no checkpoint file): what a serving process does after it
        has loaded / received the weights once. Same `predict` / `predict_many` behaviour as the checkpoint constructor."""
        from .preprocess import GpuPreprocessor
        self = cls.__new__(cls)
        self.cfg = cfg
        self.device = net.dev
        self.face_replace_model = SimpleNamespace(net=net, cfg=None if cfg is None else cfg.model)
        self.max_conditioning_images = max_conditioning_images
        net.noise_timesteps = [249]
        self.dtype = torch.float16
        self.pre = GpuPreprocessor(self.device)
        self._pinned = {}
        return self

    @property
    def net(self) -> RestorePipeline:
        return self.face_replace_model.net

    def _apply_transforms_on_image_list(self, images: List[Image.Image]) -> List[torch.Tensor]:
        return [image_to_tensor(im) for im in images]

    def _forward_batch(self, input_images: torch.Tensor, conditioning_images: torch.Tensor, face_embeds=None,
                       calc_attn_probs: bool = False):
        valid_indices = torch.ones(input_images.size(0), dtype=torch.int64) * self.max_conditioning_images
        x_pred, _, maps = self.net.forward(input_images.to(self.device, self.dtype), face_embeds=face_embeds,
                                           conditioning_images=conditioning_images.to(self.device, self.dtype),
                                           valid_indices=valid_indices, return_self_attention_maps=calc_attn_probs)
        if calc_attn_probs:                                         # test.py:107-110
            return x_pred, [p.float().cpu().detach() for p in maps]
        return x_pred, None

    def prepare_conditioning_images(self, cond_imgs: List[Image.Image]):
        if self.cfg.model.condition_on_face_embeds:
            raise NotImplementedError("condition_on_face_embeds is False in the released configs")
        return torch.stack(self._apply_transforms_on_image_list(cond_imgs), dim=0), None, None

    def parse_results(self, outputs: torch.Tensor, input_img: Image.Image, target_img: Optional[Image.Image] = None):
        pred_image = [tensor2im(out, unnorm=True) for out in outputs][0]
        to_join = [np.asarray(input_img.convert("RGB").resize(pred_image.size)), np.asarray(pred_image)]
        if target_img is not None:
            to_join.append(np.asarray(target_img.convert("RGB").resize(pred_image.size)))
        return pred_image, Image.fromarray(np.concatenate(to_join, axis=0))

    def predict(self, input_img: Image.Image, cond_imgs: Optional[List[Image.Image]] = None,
                target_img: Optional[Image.Image] = None, calc_attn_probs: bool = False):
        """test.py:149-163. The transform (:54-59) and tensor2im run on the GPU; results are bit-identical to the host path."""
        if calc_attn_probs:
            input_t = image_to_tensor(input_img).unsqueeze(0)
            conds_t, _, _ = self.prepare_conditioning_images(cond_imgs)
            outputs, attn_probs = self._forward_batch(input_t, conditioning_images=conds_t.unsqueeze(0), calc_attn_probs=True)
            pred_image, visualization = self.parse_results(outputs, input_img=input_img, target_img=target_img)
            return pred_image, visualization, attn_probs
        pred_image = next(self.predict_many([(input_img, cond_imgs)], in_flight=1))
        return pred_image, self._visualization(pred_image, input_img, target_img), None

    @staticmethod
    def _visualization(pred_image, input_img, target_img):
        to_join = [np.asarray(input_img.convert("RGB").resize(pred_image.size)), np.asarray(pred_image)]
        if target_img is not None:
            to_join.append(np.asarray(target_img.convert("RGB").resize(pred_image.size)))
        return Image.fromarray(np.concatenate(to_join, axis=0))

    # ------------------------------------------------------------------------------------------ batched / pipelined entry
    def _stage(self, slot: int, images: List[Image.Image]) -> List[torch.Tensor]:
        """PIL images -> uint8 HWC tensors on the device through this slot's pinned staging buffer (one async H2D each)."""
        arrs = [np.array(im.convert("RGB"), dtype=np.uint8) for im in images]
        total = sum(a.size for a in arrs)
        buf = self._pinned.get(("in", slot))
        if buf is None or buf.numel() < total:
            buf = torch.empty(max(total, 1 << 22), dtype=torch.uint8).pin_memory()
            self._pinned[("in", slot)] = buf
        dev_buf = torch.empty(total, dtype=torch.uint8, device=self.device)
        off = 0
        views = []
        for a in arrs:
            n = a.size
            buf[off:off + n].copy_(torch.from_numpy(a.reshape(-1)))
            views.append((off, a.shape))
            off += n
        dev_buf.copy_(buf[:total], non_blocking=True)
        return [dev_buf[o:o + int(np.prod(sh))].view(*sh) for o, sh in views]

    def predict_many(self, requests, in_flight: int = 3):
        """Generator over (input_img, cond_imgs) pairs -> restored PIL images, in order. `in_flight` requests are kept on
        the GPU at once (one CUDA stream and one graph instance each): while request i runs, the host stages request
        i+1 (uint8 through pinned memory; Lanczos resize, crop, normalisation and the uint8 packing of the result run on
        the device) and converts result i-1. This is the serving loop `bench.py` times as `e2e`, behind the
        reference's Predictor."""
        net, dev = self.net, self.device
        n_fly = max(1, in_flight)
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_fly)]
        pending = []                                   # (event, pinned result, slot)
        size = self.pre.size

        def finish(item):
            ev, host = item
            ev.synchronize()
            return Image.fromarray(host.numpy().copy())

        cur = torch.cuda.current_stream(dev)
        for i, (input_img, cond_imgs) in enumerate(requests):
            k = i % n_fly
            if len(pending) >= n_fly:
                yield finish(pending.pop(0))
            cond_imgs = list(cond_imgs or [])
            st = streams[k]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                dev_imgs = self._stage(k, [input_img] + cond_imgs)
                x = torch.empty((1 + len(cond_imgs), 3, size, size), dtype=torch.float16, device=dev)
                for j, im in enumerate(dev_imgs):
                    self.pre(im, x[j])
                valid = torch.ones(1, dtype=torch.int64) * self.max_conditioning_images
                cond = x[1:].unsqueeze(0) if cond_imgs else None
                out, _, _ = net.forward(x[:1], conditioning_images=cond, valid_indices=valid, slot=k)
                u8 = L.image_out_u8(out.contiguous())
                host = self._pinned.get(("out", k))
                if host is None:
                    host = torch.empty((size, size, 3), dtype=torch.uint8).pin_memory()
                    self._pinned[("out", k)] = host
                host.copy_(u8[0], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            pending.append((ev, host))
        while pending:
            yield finish(pending.pop(0))
