"""Integer / byte path either side of the network: bit-exact parity.

CPU (`-m "not gpu"`): the oracle restatement of Pillow's 8-bit Lanczos resampling (oracle/pil_resample.py) is pinned
against Pillow itself, the product's host-side coefficient tables (instantrestore_b200/preprocess.py) against the
oracle's, and the oracle's normalise / tensor2im against the host functions of instantrestore_b200/inference.py.
GPU (`-m gpu`): the CUDA passes through the C ABI against the PIL path of the reference transform, bit for bit."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pil_resample as O

SIZES = [(700, 525), (512, 640), (1024, 768), (300, 400), (513, 512), (1500, 1000), (512, 512), (640, 512), (2048, 1365), (517, 519)]


def _photo(rng, w, h):
    """Half smooth content, half white noise: exercises both the clipping and the negative Lanczos lobes."""
    low = rng.random((h // 16 + 2, w // 16 + 2, 3))
    img = np.kron(low, np.ones((16, 16, 1)))[:h, :w] * 255
    img[:, w // 2:] = rng.integers(0, 256, (h, w - w // 2, 3))
    return img.astype(np.uint8)


def _pil_transform(img_u8):
    """torchvision Resize(512, LANCZOS) + CenterCrop(512) of test.py:54-57 with PIL."""
    h, w = img_u8.shape[:2]
    nw, nh, left, top = O.resize_geometry(w, h)
    im = Image.fromarray(img_u8)
    if (nw, nh) != (w, h):
        im = im.resize((nw, nh), Image.LANCZOS)
    return np.asarray(im.crop((left, top, left + 512, top + 512)))


@pytest.mark.parametrize("w,h", SIZES)
def test_oracle_resampling_is_bit_exact_with_pillow(w, h):
    img = _photo(np.random.default_rng(w * 10007 + h), w, h)
    assert np.array_equal(O.transform_u8(img), _pil_transform(img))


@pytest.mark.parametrize("n_in,n_out", [(700, 682), (525, 512), (300, 512), (2048, 768), (1365, 512), (519, 514), (513, 512)])
def test_host_coefficient_tables_match_the_oracle(n_in, n_out):
    from instantrestore_b200.preprocess import lanczos_tables
    b, k = lanczos_tables(n_in, n_out)
    bo, ko = O.coeffs_8bpc(n_in, n_out)
    assert np.array_equal(b, bo) and np.array_equal(k, ko)
    assert (k.sum(1) - (1 << 22)).__abs__().max() <= k.shape[1]          # fixed-point weights sum to ~1


def test_resize_geometry_matches_image_to_tensor():
    from instantrestore_b200.inference import image_to_tensor
    from instantrestore_b200.preprocess import resize_geometry
    rng = np.random.default_rng(5)
    for (w, h) in SIZES[:6]:
        img = _photo(rng, w, h)
        assert resize_geometry(w, h) == O.resize_geometry(w, h)
        ref = image_to_tensor(Image.fromarray(img)).half().numpy()
        assert np.array_equal(O.normalize_f16(O.transform_u8(img)), ref)


def test_tensor2im_matches_the_oracle_on_fp16_predictions():
    from instantrestore_b200.inference import tensor2im
    g = torch.Generator().manual_seed(3)
    pred = (torch.randn(3, 64, 96, generator=g) * 0.8).clamp(-1, 1).half()
    pred[0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 0.99951171875, -0.99951171875, 0.5, -0.5, 0.00390625]).half()
    assert np.array_equal(np.asarray(tensor2im(pred, unnorm=True)), O.tensor2im_u8(pred.numpy()))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("w,h", SIZES)
def test_gpu_preprocess_is_bit_exact_with_the_reference_transform(w, h):
    from instantrestore_b200.inference import image_to_tensor
    from instantrestore_b200.preprocess import GpuPreprocessor
    img = _photo(np.random.default_rng(w * 31 + h), w, h)
    pre = GpuPreprocessor("cuda:0")
    out = torch.full((3, 512, 512), float("nan"), dtype=torch.float16, device="cuda")
    pre(torch.from_numpy(img).cuda(), out)
    ref = image_to_tensor(Image.fromarray(img)).half()                    # PIL resize + crop, ToTensor, Normalize, fp16 cast
    assert torch.equal(out.cpu(), ref)
    out2 = torch.empty_like(out)
    pre(torch.from_numpy(img).cuda(), out2)                               # cached tables, same result
    assert torch.equal(out2, out)


@pytest.mark.gpu
def test_gpu_tensor2im_is_bit_exact():
    from instantrestore_b200 import _lib as L
    from instantrestore_b200.inference import tensor2im
    g = torch.Generator().manual_seed(4)
    pred = (torch.randn(2, 3, 128, 160, generator=g) * 0.9).clamp(-1, 1).half()
    pred[0, :, 0, :4] = torch.tensor([[-1.0, 1.0, 0.0, 0.5]] * 3).half()
    u8 = L.image_out_u8(pred.cuda()).cpu().numpy()
    for b in range(2):
        assert np.array_equal(u8[b], np.asarray(tensor2im(pred[b], unnorm=True)))
        assert np.array_equal(u8[b], O.tensor2im_u8(pred[b].numpy()))


@pytest.mark.gpu
def test_gpu_preprocess_rejects_bad_inputs():
    from instantrestore_b200.preprocess import GpuPreprocessor
    pre = GpuPreprocessor("cuda:0")
    out = torch.empty((3, 512, 512), dtype=torch.float16, device="cuda")
    with pytest.raises(TypeError):
        pre(torch.zeros(600, 600, 3), out)                                # host tensor
    with pytest.raises(TypeError):
        pre(torch.zeros(600, 600, 4, dtype=torch.uint8, device="cuda"), out)
