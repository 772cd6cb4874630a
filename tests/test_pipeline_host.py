"""Host-side pieces of the drop-in pipeline surface that need no GPU: the result containers of
Kandinsky5T2VPipeline.__call__ (kandinsky/t2v_pipeline.py:165-189: PIL images in image mode, png / mp4 files for
`save_path`) and the "sta_mask" entry of get_sparse_params (generation_utils.py:10-36) against the oracle's restatement
of fast_sta_nabla."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))

from oracle import dit_oracle as O  # noqa: E402


def test_image_mode_returns_pil_images_like_to_pil_image():
    from kandinsky.t2v_pipeline import to_pil_images

    g = torch.Generator().manual_seed(0)
    out = torch.randint(0, 256, (2, 3, 1, 8, 12), generator=g, dtype=torch.uint8)
    imgs = to_pil_images(out)
    assert len(imgs) == 2 and imgs[0].size == (12, 8) and imgs[0].mode == "RGB"
    for i, img in enumerate(imgs):
        assert np.array_equal(np.asarray(img), out[i, :, 0].permute(1, 2, 0).numpy())
    tv = pytest.importorskip("torchvision.transforms")
    ref = tv.ToPILImage()(out[0, :, 0])                                       # what the reference calls (:168)
    assert np.array_equal(np.asarray(ref), np.asarray(imgs[0]))


def test_write_video_writes_a_readable_mp4(tmp_path):
    import cv2

    from kandinsky.t2v_pipeline import write_video

    T, H, W = 9, 64, 96
    t = torch.arange(T).view(T, 1, 1, 1)
    frames = ((torch.arange(H).view(1, H, 1, 1) * 2 + torch.arange(W).view(1, 1, W, 1) + 10 * t) % 256).expand(T, H, W, 3)
    path = os.path.join(tmp_path, "v.mp4")
    backend = write_video(path, frames.to(torch.uint8), fps=24)
    assert backend in ("pyav", "opencv") and os.path.getsize(path) > 0
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == T
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (W, H)
    assert abs(cap.get(cv2.CAP_PROP_FPS) - 24.0) < 0.5
    ok, first = cap.read()
    assert ok and abs(float(first.mean()) - float(frames[0].float().mean())) < 6.0      # lossy codec
    cap.release()


@pytest.mark.parametrize("T,H,W,w", [(3, 2, 2, (3, 3, 3)), (5, 4, 6, (11, 3, 3)), (2, 8, 12, (1, 5, 3))])
def test_sparse_params_carry_the_reference_sta_mask(T, H, W, w):
    from kandinsky.generation_utils import get_sparse_params

    conf = {"model": {"dit_params": {"patch_size": [1, 2, 2]},
                      "attention": {"type": "nabla", "P": 0.9, "wT": w[0], "wH": w[1], "wW": w[2], "add_sta": True}}}
    sp = get_sparse_params(conf, {"visual": torch.zeros(T, 16 * H, 16 * W, 16)}, "cpu")
    assert set(sp) == {"sta_mask", "attention_type", "to_fractal", "P", "wT", "wW", "wH", "add_sta", "visual_shape", "method"}
    assert sp["visual_shape"] == (T, 8 * H, 8 * W) and sp["method"] == "topcdf" and sp["to_fractal"] is True
    m = sp["sta_mask"]
    assert m.dtype == torch.bool and tuple(m.shape) == (1, 1, T * H * W, T * H * W)
    assert torch.equal(m[0, 0], O.sta_mask(T, H, W, *w).bool())
    conf["model"]["attention"]["type"] = "flash"
    assert get_sparse_params(conf, {"visual": torch.zeros(T, 16 * H, 16 * W, 16)}, "cpu") is None
