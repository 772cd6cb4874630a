"""Drop-in for the reference's `kandinsky/t2v_pipeline.py` (Kandinsky5T2VPipeline.__call__, :90-189)."""
import torch

from .generation_utils import generate_sample


def to_pil_images(images):
    """t2v_pipeline.py:166-169: uint8 [B, 3, 1, H, W] -> list of PIL RGB images (torchvision's ToPILImage on a CHW
    uint8 tensor = the HWC byte array handed to PIL unchanged)."""
    from PIL import Image

    return [Image.fromarray(img.permute(1, 2, 0).contiguous().numpy(), mode="RGB") for img in images.squeeze(2).cpu()]


def write_video(path, frames, fps=24, crf=5):
    """t2v_pipeline.py:181-186: torchvision.io.write_video(path, [T, H, W, 3] frames, fps=24, options={"crf": "5"}).
    torchvision's writer needs PyAV; where PyAV is not installed the same frames go through OpenCV's mp4 writer
    (codec mp4v, no CRF control).  Returns the backend used."""
    frames = torch.as_tensor(frames).to(torch.uint8).cpu()
    assert frames.dim() == 4 and frames.shape[-1] == 3, "frames must be [T, H, W, 3]"
    try:
        import av  # noqa: F401
        import torchvision

        torchvision.io.write_video(path, frames.numpy(), fps=fps, options={"crf": str(crf)})
        return "pyav"
    except ImportError:
        pass
    import cv2

    T, H, W, _ = frames.shape
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (W, H))
    if not wr.isOpened():
        raise RuntimeError(f"cannot open a video writer for {path}")
    for f in frames.numpy():
        wr.write(f[:, :, ::-1].copy())              # OpenCV takes BGR
    wr.release()
    return "opencv"


class Kandinsky5T2VPipeline:
    def __init__(self, device_map, dit, text_embedder, vae, resolution=512, local_dit_rank=0, world_size=1, conf=None,
                 offload=False):
        if resolution not in [512]:
            raise ValueError("Resolution can be only 512")
        self.dit, self.text_embedder, self.vae = dit, text_embedder, vae
        self.resolution = resolution
        self.device_map = device_map
        self.local_dit_rank = local_dit_rank
        self.world_size = world_size
        self.conf = conf
        self.num_steps = conf.model.num_steps
        self.guidance_weight = conf.model.guidance_weight
        self.offload = offload
        self.RESOLUTIONS = {512: [(512, 512), (512, 768), (768, 512)]}

    def __call__(self, text, time_length=5, width=768, height=512, seed=None, num_steps=None, guidance_weight=None,
                 scheduler_scale=10.0, negative_caption="Static, 2D cartoon, cartoon, 2d animation, paintings, images, "
                 "worst quality, low quality, ugly, deformed, walking backwards", expand_prompts=True, save_path=None,
                 progress=True):
        """t2v_pipeline.py:90-189: uint8 video [1, 3, F, H, W], or a list of PIL images when time_length == 0; `save_path`
        (a path or a list with one path per result) writes png / mp4 on rank 0 exactly where the reference does.
        Prompt expansion (an LLM generate call, :127-147) is outside the hot path: `expand_prompts` is accepted and ignored."""
        num_steps = self.num_steps if num_steps is None else num_steps
        guidance_weight = self.guidance_weight if guidance_weight is None else guidance_weight
        if seed is None:
            if self.local_dit_rank == 0:
                seed = torch.randint(2 ** 63 - 1, (1,)).to(self.local_dit_rank)
            else:
                seed = torch.empty((1,), dtype=torch.int64).to(self.local_dit_rank)
            if self.world_size > 1:
                torch.distributed.broadcast(seed, 0)
            seed = seed.item()
        if self.resolution != 512:
            raise NotImplementedError("Only 512 resolution is available for now")
        if (height, width) not in self.RESOLUTIONS[self.resolution]:
            raise ValueError(f"Wrong height, width pair. Available (height, width) are: {self.RESOLUTIONS[self.resolution]}")
        num_frames = 1 if time_length == 0 else time_length * 24 // 4 + 1
        shape = (1, num_frames, height // 8, width // 8, 16)
        out = generate_sample(shape, text, self.dit, self.vae, self.conf, text_embedder=self.text_embedder,
                              num_steps=num_steps, guidance_weight=guidance_weight, scheduler_scale=scheduler_scale,
                              negative_caption=negative_caption, seed=seed, device=self.device_map["dit"],
                              vae_device=self.device_map["vae"], progress=progress, offload=self.offload)
        if self.vae is None or self.local_dit_rank != 0:
            return out                              # no decoder: the fp32 latent (generate_sample); ranks > 0: t2v_pipeline.py:165
        if time_length == 0:
            images = to_pil_images(out)
            if save_path is not None:
                paths = [save_path] if isinstance(save_path, str) else save_path
                if len(paths) == len(images):
                    for path, image in zip(paths, images):
                        image.save(path)
            return images
        if save_path is not None:
            paths = [save_path] if isinstance(save_path, str) else save_path
            if len(paths) == len(out):
                for path, video in zip(paths, out):
                    write_video(path, video.permute(1, 2, 3, 0), fps=24, crf=5)
        return out
