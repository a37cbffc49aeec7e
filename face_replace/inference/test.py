"""Reference import path `face_replace.inference.test` (reference inference/test.py:38-187): `Predictor(checkpoint_path)`
with `.cfg`, `.face_replace_model.net`, `.predict(input_img, cond_imgs, target_img=None, calc_attn_probs=False)`,
and the `__main__` loop over a folder of identities (`<identity>/degraded.png`, `<identity>/conditioning/*.png`,
optional `<identity>/gt.png`). Runs on the B200 engine; there is no CPU path.

    python -m face_replace.inference.test --checkpoint ckpt.pt --data path/to/data --results path/to/results
"""
from __future__ import annotations

import re
from pathlib import Path

from instantrestore_b200.inference import FaceReplaceModel, Predictor, image_to_tensor, tensor2im  # noqa: F401
from face_replace.models.attn_processors import SharedAttnProcessor  # noqa: F401  (imported by the reference module too)

__all__ = ["Predictor", "run_folder"]


def _natural_key(path: Path):
    """natsort-style ordering: digit runs compare as numbers (1.png < 2.png < 10.png)."""
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split(r"(\d+)", path.name)]


def run_folder(predictor: Predictor, data_root: Path, results_dir: Path, max_refs: int = 4, calc_attn_probs: bool = False):
    """The reference's __main__ loop (:165-187): one restored image per identity directory."""
    from PIL import Image
    results_dir.mkdir(parents=True, exist_ok=True)
    written = []
    for identity in sorted(p for p in Path(data_root).glob("*") if p.is_dir()):
        input_img = Image.open(identity / "degraded.png").convert("RGB")
        cond_paths = sorted((identity / "conditioning").glob("*.png"), key=_natural_key)[:max_refs]
        cond_imgs = [Image.open(p).convert("RGB") for p in cond_paths]
        gt = identity / "gt.png"
        target_img = Image.open(gt).convert("RGB") if gt.exists() else None
        pred, _vis, _probs = predictor.predict(input_img, cond_imgs=cond_imgs, target_img=target_img,
                                               calc_attn_probs=calc_attn_probs)
        out = results_dir / f"{identity.name}.png"
        pred.save(out)
        written.append(out)
    return written


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--data", required=True)
    ap.add_argument("--results", required=True)
    ap.add_argument("--calc-attn-probs", action="store_true")
    a = ap.parse_args()
    run_folder(Predictor(checkpoint_path=Path(a.checkpoint)), Path(a.data), Path(a.results), calc_attn_probs=a.calc_attn_probs)
