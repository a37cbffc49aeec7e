"""ORACLE (test infrastructure): CPU restatement of the reference single-step pipeline at the latent boundary.

Follows reference face_replace/models/pix2pix_turbo.py: get_conditioning_keys_values :242-279 (noise the reference
latents to t=1, run the frozen UNet, gather the 9 captured key/value pairs as (B, N, S, C), ZERO the padded slots)
and forward :281-343 (noise the degraded latent to t=noise_timestep, run the LoRA-tuned UNet with
cross_attention_kwargs={'ref_keys','ref_values'}, take the scheduler's pred_original_sample). The four stochastic
draws of the reference (two VAE posterior samples, two randn_like) are injected as arguments so the path is
deterministic; the VAE on either side of this boundary is a later row of SURVEY.md 8f.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import attn_processors as _oracle_processors
from .diffusers024 import DDPMScheduler1Step
from .synth import ModelFlags
from .unet import UNet2DConditionModel


class LatentRestorePipeline:
    def __init__(self, unet: UNet2DConditionModel, original_unet: UNet2DConditionModel, caption_enc: torch.Tensor,
                 flags: ModelFlags, save_self_attentions: bool = False, noise_timestep: int = 249,
                 processors=_oracle_processors):
        """`processors`: module providing AttnProcessor / register_attention_processor(_kv_unet); the oracle's by
        default, the reference's own face_replace.models.attn_processors when generating golden vectors."""
        self.unet, self.original_unet = unet, original_unet
        self.caption_enc = caption_enc
        self.flags = flags
        self.sched = DDPMScheduler1Step()
        self.noise_timestep = noise_timestep          # reference test.py:62 sets noise_timesteps = [249]
        self._kv_proc_type = processors.AttnProcessor
        processors.register_attention_processor_kv_unet(original_unet)                              # pix2pix_turbo.py:97
        processors.register_attention_processor(unet, cfg=flags, save_self_attentions=save_self_attentions)  # :98

    @torch.no_grad()
    def conditioning_keys_values(self, ref_latents: torch.Tensor, noise: torch.Tensor, valid_indices) -> Tuple[List, List]:
        b, n = ref_latents.shape[:2]
        enc = ref_latents.reshape(b * n, *ref_latents.shape[2:])
        t = torch.tensor([1], device=enc.device)
        noisy = self.sched.add_noise(enc, noise, t.long().repeat(enc.shape[0]))
        cap = self.caption_enc.to(enc.device).repeat(enc.shape[0], 1, 1)
        self.original_unet(noisy, t, encoder_hidden_states=cap)
        procs = [p for p in self.original_unet.attn_processors.values() if type(p) is self._kv_proc_type]
        keys = [p.keys.reshape(-1, n, p.keys.shape[1], p.keys.shape[2]) for p in procs]
        values = [p.values.reshape(-1, n, p.values.shape[1], p.values.shape[2]) for p in procs]
        for k, v in zip(keys, values):
            for i in range(k.shape[0]):
                idx = int(valid_indices[i])
                k[i, idx:] = 0
                v[i, idx:] = 0
        for p in procs:
            p.reset()
        return keys, values

    @torch.no_grad()
    def forward_latents(self, enc_control: torch.Tensor, ref_latents: Optional[torch.Tensor], noise_main: torch.Tensor,
                        noise_ref: Optional[torch.Tensor], valid_indices=None, face_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        keys = values = None
        if ref_latents is not None and self.flags.use_shared_attention:
            if valid_indices is None:
                valid_indices = [ref_latents.shape[1]] * ref_latents.shape[0]
            keys, values = self.conditioning_keys_values(ref_latents, noise_ref, valid_indices)
        t = torch.tensor([self.noise_timestep], device=enc_control.device)
        noisy = self.sched.add_noise(enc_control, noise_main, t.long().repeat(enc_control.shape[0]))
        if self.flags.condition_on_face_embeds and face_embeds is not None:      # pix2pix_turbo.py:316-320
            cap = face_embeds.to(noisy.device)
        else:
            cap = self.caption_enc.to(noisy.device).repeat(noisy.shape[0], 1, 1)
        pred = self.unet(noisy, t, encoder_hidden_states=cap,
                         cross_attention_kwargs={"ref_keys": keys, "ref_values": values})
        pred = getattr(pred, "sample", pred)
        return self.sched.pred_original_sample(pred, self.noise_timestep, noisy)


class ImageRestorePipeline:
    """ORACLE: the whole reference forward on images (face_replace/models/pix2pix_turbo.py:281-343 with
    get_conditioning_keys_values :242-279): VAE-encode the degraded image (:291) and the reference images (:245),
    run the latent pipeline above, decode with the skip activations of the degraded-image encode (:332-333) and clamp
    to [-1, 1]. The two posterior draws are injected (`eps_main`, `eps_ref`). The reference also decodes the reference
    latents (:277-278); that output is unused at inference (test.py:100-105) and is not computed here."""

    def __init__(self, latent_pipeline: LatentRestorePipeline, vae, original_vae):
        self.latent = latent_pipeline
        self.vae, self.original_vae = vae, original_vae

    @torch.no_grad()
    def forward(self, c_t, conditioning_images, eps_main, eps_ref, noise_main, noise_ref, valid_indices=None):
        sf = self.vae.config.scaling_factor
        enc = self.vae.encode_sample(c_t, eps_main) * sf
        ref_lat = None
        if conditioning_images is not None and self.latent.flags.use_shared_attention:
            b, n = conditioning_images.shape[:2]
            cond = conditioning_images.reshape(b * n, *conditioning_images.shape[2:])
            ref_lat = (self.original_vae.encode_sample(cond, eps_ref) * self.original_vae.config.scaling_factor)
            ref_lat = ref_lat.reshape(b, n, *ref_lat.shape[1:])
        x0 = self.latent.forward_latents(enc, ref_lat, noise_main, noise_ref, valid_indices)
        self.vae.decoder.incoming_skip_acts = self.vae.encoder.current_down_blocks
        return self.vae.decode(x0 / sf).clamp(-1, 1)
