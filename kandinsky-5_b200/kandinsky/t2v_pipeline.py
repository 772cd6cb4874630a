"""Drop-in for the reference's `kandinsky/t2v_pipeline.py` (Kandinsky5T2VPipeline.__call__, :90-189)."""
import torch

from .generation_utils import generate_sample


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
        """t2v_pipeline.py:90-189.  Prompt expansion (an LLM generate call) and mp4 writing are outside the hot path;
        `expand_prompts` is accepted and ignored, `save_path` is honoured only for tensors (torch.save)."""
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
        if save_path is not None and self.local_dit_rank == 0:
            torch.save(out.cpu(), save_path)
        return out
