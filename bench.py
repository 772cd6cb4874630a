#!/usr/bin/env python
"""Headline benchmark: DiT denoising-step latency and latent-tokens/s of the Kandinsky-5 2B Lite DiT at
768x512x121 (latent 31x64x96x16 -> S = 47 616 tokens, L = 256 text tokens), random-init weights, cached synthetic
text embeddings (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl k5|reference] [--workload 5s_nocfg|5s_sft|10s_sft_nabla]

A "step" is one iteration of the flow-matching sampler (generation_utils.py:105-128): one DiT forward (two with CFG)
plus the Euler update.  `value` = visual tokens through the DiT per second with inputs resident in HBM; `e2e` = the
same through the reference-facing API (DiffusionTransformer3D.forward -> C ABI) with pinned HOST buffers, H2D / D2H
copies inside the timed region.  N > 1: one process per GPU (torchrun), ONE video sharded along the latent's time
axis: every rank owns a slab of frames, K | V are all-gathered per visual block inside the QKV
projection kernel over NVLink peer memory (no NCCL on the data path; NCCL only returns the finished latent slabs once
per sample), value = the video's tokens / max-over-ranks time, scaling "strong".
`--impl reference` times the reference algorithm's CPU path (the oracle port, oracle/dit_oracle.py) on the host
cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "kandinsky-5_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

LITE = dict(in_visual_dim=16, out_visual_dim=16, time_dim=512, patch_size=(1, 2, 2), model_dim=1792, ff_dim=7168,
            num_text_blocks=2, num_visual_blocks=32, axes_dims=(16, 24, 24), visual_cond=True, in_text_dim=3584,
            in_text_dim2=768)
WORKLOADS = {
    # name: (latent T, H, W, text L, null L, guidance, scheduler_scale, nabla dict or None, NFE of the full config)
    "5s_nocfg": dict(T=31, H=64, W=96, L=256, Ln=64, w=1.0, sched=5.0, nabla=None, nfe=50,
                     name="config_5s_nocfg: 768x512x121, 50 NFE, no CFG"),
    "5s_sft": dict(T=31, H=64, W=96, L=256, Ln=64, w=5.0, sched=5.0, nabla=None, nfe=100,
                   name="config_5s_sft: 768x512x121, 50 steps x CFG = 100 NFE"),
    "10s_sft_nabla": dict(T=61, H=64, W=96, L=256, Ln=64, w=5.0, sched=10.0,
                          nabla=dict(P=0.9, wT=11, wH=3, wW=3, add_sta=True), nfe=100,
                          name="config_10s_sft: 768x512x241, NABLA P=0.9 (11,3,3), 100 NFE"),
    # the same with the adaptive part switched off (P -> 0: only the sliding-tile window survives, density 4.8 %):
    # random weights give near-uniform attention maps (density ~0.9), trained ones sit between the two
    "10s_sft_sta": dict(T=61, H=64, W=96, L=256, Ln=64, w=5.0, sched=10.0,
                        nabla=dict(P=0.0, wT=11, wH=3, wW=3, add_sta=True), nfe=100,
                        name="config_10s_sft with P=0: 768x512x241, STA window (11,3,3) only, 100 NFE"),
}


def dit_flops(S, L, rho=1.0, D=1792, F=7168, nvis=32, ntext=2):
    """Algorithmic FLOPs of one forward."""
    per_vis = (2 * S * D * 3 * D + 2 * S * D * D + 2 * S * D * D * 2 + 2 * L * D * 2 * D + 2 * S * D * F * 2
               + rho * 4 * S * S * D + 4 * S * L * D + 2 * 512 * 9 * D)
    per_text = 8 * L * D * D + 4 * L * D * F + 4 * L * L * D + 2 * 512 * 6 * D
    return nvis * per_vis + ntext * per_text + 2 * S * 132 * D + 2 * L * 3584 * D + 2 * S * D * 64


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1386.8), d.get("bf16_tflops", 1642.7), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = f"/tmp/k5_clocks_{os.getpid()}.csv"
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(wl, budget_s):
    """Times the oracle port (the reference algorithm on the CPU, eager torch, all host threads) on a bounded sample
    of the workload: ONE visual TransformerDecoderBlock (1/32 of the stack, >99.9 % of a forward's FLOPs are in the 32
    blocks) for the first `Sq` query tokens against the full-length K/V, then extrapolates linearly in the query
    count (every op of the block except the K/V projection is per query token) and x32 blocks."""
    import torch

    from oracle import dit_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(O.LITE_CFG, num_visual_blocks=1, num_text_blocks=0)
    sd = {k: v for k, v in O.synthetic_state_dict(cfg, seed=0).items() if k.startswith("visual_transformer_blocks.0.")}
    S = wl["T"] * (wl["H"] // 2) * (wl["W"] // 2)
    L, D = wl["L"], cfg["model_dim"]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(S, D, generator=g).to(torch.bfloat16)
    text = torch.randn(L, D, generator=g).to(torch.bfloat16)
    tm = torch.randn(1, cfg["time_dim"], generator=g)
    ang = torch.randn(S, 32, generator=g)
    cos, sin = torch.cos(ang), torch.sin(ang)
    pfx = "visual_transformer_blocks.0."

    def kv_part():
        m = O.modulation(sd, pfx + "visual_modulation.", tm)
        shift, scale, _ = torch.chunk(torch.chunk(m, 3, dim=-1)[0], 3, dim=-1)
        xn = O.scale_shift_norm(x, scale, shift, "cuda")
        p = pfx + "self_attention."
        k = O._lin(xn, sd[p + "to_key.weight"], sd[p + "to_key.bias"], "cuda").reshape(S, -1, 64)
        v = O._lin(xn, sd[p + "to_value.weight"], sd[p + "to_value.bias"], "cuda").reshape(S, -1, 64)
        k = O.apply_rotary(O.rms_norm_heads(k, sd[p + "key_norm.weight"], "cuda"), cos, sin, "cuda")
        return xn, k, v, m

    def q_part(Sq, xn, k, v, m):
        sa, ca, ff = torch.chunk(m, 3, dim=-1)
        p = pfx + "self_attention."
        xq = x[:Sq]
        q = O._lin(xn[:Sq], sd[p + "to_query.weight"], sd[p + "to_query.bias"], "cuda").reshape(Sq, -1, 64)
        q = O.apply_rotary(O.rms_norm_heads(q, sd[p + "query_norm.weight"], "cuda"), cos[:Sq], sin[:Sq], "cuda")
        o = O.attention(q, k, v, "cuda")
        o = O._lin(o, sd[p + "out_layer.weight"], sd[p + "out_layer.bias"], "cuda")
        xq = O.gate_sum(xq, o, torch.chunk(sa, 3, dim=-1)[2], "cuda")
        shift, scale, gate = torch.chunk(ca, 3, dim=-1)
        o = O._cross_attention(sd, pfx + "cross_attention.", O.scale_shift_norm(xq, scale, shift, "cuda"), text, "cuda")
        xq = O.gate_sum(xq, o, gate, "cuda")
        shift, scale, gate = torch.chunk(ff, 3, dim=-1)
        o = O.feed_forward(sd, pfx + "feed_forward.", O.scale_shift_norm(xq, scale, shift, "cuda"), "cuda")
        return O.gate_sum(xq, o, gate, "cuda")

    with torch.no_grad():
        t0 = time.perf_counter()
        xn, k, v, m = kv_part()
        t_kv = time.perf_counter() - t0
        # calibrate the per-query cost on a small slice, then size the sample to the budget
        Sq0 = min(S, 512)
        t0 = time.perf_counter()
        q_part(Sq0, xn, k, v, m)
        per_q = (time.perf_counter() - t0) / Sq0
        Sq = int(max(512, min(S, (budget_s - t_kv) / max(per_q, 1e-9))))
        Sq = min(S, (Sq // 64) * 64)
        t0 = time.perf_counter()
        q_part(Sq, xn, k, v, m)
        t_q = time.perf_counter() - t0
    t_block = t_kv + t_q * (S / Sq)
    fwd_per_step = 2 if abs(wl["w"] - 1.0) > 1e-6 else 1
    t_step = 32 * t_block * fwd_per_step
    sample = (f"1 visual decoder block of 32 (oracle port, eager torch CPU, bf16 operands): K/V path for all S={S} tokens "
              f"({t_kv:.2f} s) + query path for the first {Sq} of {S} tokens ({t_q:.2f} s), extrapolated linearly in queries, "
              f"x32 blocks x{fwd_per_step} forward(s) per step")
    return S * fwd_per_step / t_step, t_step * 1e3, cores, sample


def cpu_reference_block(wl, budget_s):
    """The reference's OWN TransformerDecoderBlock (kandinsky/models/dit.py:47-79, imported unmodified from
    baseline/_ref through baseline/ref_loader.py) on the host cores: one visual block = 1/32 of the stack that holds
    > 99.9 % of a forward's FLOPs, at the workload's full token count, eager, FlashAttention replaced by torch SDPA (no
    CPU flash_attn exists), under torch.autocast('cpu', bf16) as SURVEY.md section 8c prescribes.  Returns None when
    baseline/_ref is absent (then the oracle port is timed instead)."""
    import torch

    from baseline import ref_loader

    if not ref_loader.available():
        return None
    import torch._dynamo

    from oracle import dit_oracle as O

    torch._dynamo.config.disable = True
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(O.LITE_CFG, num_visual_blocks=1, num_text_blocks=0)
    sd = O.synthetic_state_dict(cfg, seed=0)
    model = ref_loader.build_model(cfg, sd, "cpu")
    block = model.visual_transformer_blocks[0]
    T, Hp, Wp = wl["T"], wl["H"] // 2, wl["W"] // 2
    S, L, D = T * Hp * Wp, wl["L"], cfg["model_dim"]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(S, D, generator=g).to(torch.bfloat16)
    text = torch.randn(L, D, generator=g).to(torch.bfloat16)
    tm = torch.randn(1, cfg["time_dim"], generator=g)
    with torch.no_grad():
        rope = model.visual_rope_embeddings((T, Hp, Wp), [torch.arange(T), torch.arange(Hp), torch.arange(Wp)], (1.0, 2.0, 2.0))
        rope = rope.flatten(0, 2)

    def call():
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            return block(x, text, tm, rope, None)

    t0 = time.perf_counter()
    call()
    t_block = time.perf_counter() - t0
    if t_block < budget_s / 2:                   # a second, warm call when the budget allows
        t0 = time.perf_counter()
        call()
        t_block = time.perf_counter() - t0
    fwd_per_step = 2 if abs(wl["w"] - 1.0) > 1e-6 else 1
    t_step = 32 * t_block * fwd_per_step
    sample = (f"the reference's own TransformerDecoderBlock (baseline/_ref, eager, FA -> torch SDPA, autocast cpu bf16) on all "
              f"S={S} tokens: 1 visual block of 32 in {t_block:.2f} s, x32 blocks x{fwd_per_step} forward(s) per step")
    return S * fwd_per_step / t_step, t_step * 1e3, cores, sample, "reference"


def cpu_arm(wl, budget_s):
    """(tokens/s, ms/step, cores, sample, kind): the reference's own block when baseline/_ref travelled, else the port."""
    try:
        r = cpu_reference_block(wl, budget_s)
    except Exception as e:                       # noqa: BLE001 - a broken reference copy must not take the bench down
        print(f"reference block failed ({e!r}); timing the oracle port instead", file=sys.stderr)
        r = None
    if r is not None:
        return r
    v, m, cores, sample = cpu_reference_sample(wl, budget_s)
    return v, m, cores, sample, "port"


def default_dist_mode(world):
    """The all-gather form engine_dist_init (csrc/engine.cu) picks for `world` ranks in separate processes."""
    if world <= 1:
        return 0
    ov = os.environ.get("K5_DIST_OVERLAP", "")
    return (2 if ov != "0" else 1) if ov != "" else (2 if world >= 8 else 1)


def bench_config(wl, world, nf, density, dist_mode=1):
    """The `config` object of the JSON line - the same for the GPU arm and the `--impl reference` arm."""
    T = wl["T"]
    S = T * (wl["H"] // 2) * (wl["W"] // 2)
    fwd_per_step = 2 if abs(wl["w"] - 1.0) > 1e-6 else 1
    return {"workload": wl["name"], "tokens": S, "text_tokens": wl["L"], "forwards_per_step": fwd_per_step,
            "model": "Kandinsky-5 T2V Lite DiT 2.0B (random init, modulation re-randomised)",
            "parallelism": (f"temporal shard x{world}: {nf} of {T} latent frames on rank 0, " + (
                "K|V all-gather overlapped with attention (copy-engine pushes over NVLink, per-slab arrival flags, local slab "
                "attended first)" if dist_mode == 2 else
                "K|V all-gather fused into the QKV GEMM epilogue over NVLink peer memory")) if world > 1 else "single GPU",
            "l2_policy": "per-step working set (>2 GB activations + 4 GB weights) exceeds the 126 MB L2",
            "nabla_density": density if wl["nabla"] else None}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kandinsky.models.parallelize import frame_partition      # host-side partition arithmetic only (no CUDA)

    total = args.steps + args.warmup
    per_step_budget = max(4.0, min(30.0, 150.0 / max(total, 1)))
    vals, ms = [], []
    cores, sample, kind = 1, "", "port"
    t_start = time.perf_counter()
    executed = 0
    for i in range(total):
        # every step is one execution of the sample; when the host is so slow that K + W executions would not end within
        # a few minutes, the remaining steps re-use the mean of the executed ones (said in `sample`)
        if executed >= 1 + args.warmup and time.perf_counter() - t_start > 210.0 and vals:
            vals.append(sum(vals) / len(vals))
            ms.append(sum(ms) / len(ms))
            continue
        v, m, cores, sample, kind = cpu_arm(wl, per_step_budget)
        executed += 1
        if i >= args.warmup:
            vals.append(v)
            ms.append(m)
    if executed < total:
        sample += f"; {executed} of {total} steps executed (210 s cap), the others carry their mean"
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "dit_latent_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / len(ms), "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": bench_config(wl, args.gpus, frame_partition(wl["T"], args.gpus)[0][1] if args.gpus > 1 else wl["T"], None,
                               default_dist_mode(args.gpus)),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def synthetic_state_dict_gpu(cfg, device, seed=0):
    """Random-init checkpoint generated on the device (same distributions as oracle.synthetic_state_dict: nn.Linear
    style uniform weights, modulation re-randomised with N(0, 0.02) because the reference zero-inits it)."""
    import math

    import torch

    from kandinsky.models.dit import state_dict_shapes

    g = torch.Generator(device=device).manual_seed(seed)
    shapes = state_dict_shapes(cfg)
    sd = {}
    for key, shape in shapes.items():
        fp32 = ("modulation" in key) or key.startswith("time_embeddings.") or ("norm" in key)
        if "modulation" in key:
            t = torch.randn(shape, device=device, generator=g) * 0.02
        elif key.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, device=device, generator=g)
        elif key.endswith("norm.bias"):
            t = 0.05 * torch.randn(shape, device=device, generator=g)
        else:
            fan_in = shape[1] if key.endswith(".weight") else shapes[key[:-4] + "weight"][1]
            t = (torch.rand(shape, device=device, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[key] = t if fp32 else t.to(torch.bfloat16)
    return sd


def build_bench_vae(dev, H, W):
    """Engine-backed AutoencoderKLHunyuanVideo mirror with random-init fp16 weights (the published VAE file is fp16)."""
    import torch

    from kandinsky.models.vae import AutoencoderKLHunyuanVideo, decoder_state_dict_shapes

    g = torch.Generator(device=dev).manual_seed(3)
    shapes = decoder_state_dict_shapes()
    sd = {}
    for k, shp in shapes.items():
        if "norm" in k:
            t = (1.0 + 0.1 * torch.randn(shp, device=dev, generator=g)) if k.endswith("weight") else 0.05 * torch.randn(shp, device=dev, generator=g)
        else:
            wshape = shp if k.endswith("weight") else shapes[k[:-4] + "weight"]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            t = (torch.rand(shp, device=dev, generator=g) * 2 - 1) / fan_in ** 0.5
        sd[k] = t.half()
    vae = AutoencoderKLHunyuanVideo(max_latent=(5, H, W))
    vae.load_state_dict(sd)
    vae.to(dev)
    return vae


def time_vae_decode(vae, latent, dev):
    """vae.decode of the latent the sampler produced, as generate_sample does it (generation_utils.py:210-222), through
    the engine-backed AutoencoderKLHunyuanVideo mirror.  One warm-up, one timed decode."""
    import torch

    T, H, W, _ = latent.shape
    z = (latent.reshape(1, T, H, W, -1) / vae.config.scaling_factor).permute(0, 4, 1, 2, 3).contiguous()
    z = z / z.std().clamp_min(1e-6)              # random-init DiT latents are not unit-scale; keep activations finite
    video = vae.decode(z).sample
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    video = vae.decode(z).sample
    u8 = ((video.clamp(-1.0, 1.0) + 1.0) * 127.5).to(torch.uint8)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert bool(torch.isfinite(video.float()).all()), "non-finite video"
    tiles = len(range(0, T - 4 + 1, 2)) if T > 5 else 1
    tflop = 118.84 * tiles * (H * W) / (64 * 96)
    sustained, _, src = measured_peaks()
    return {"ms": ms, "tiles": tiles, "frames": int(u8.shape[2]), "height": int(u8.shape[3]), "width": int(u8.shape[4]),
            "algorithmic_tflop": tflop, "tflops_achieved": tflop / ms * 1e3, "frac_of_sustained_peak": tflop / ms * 1e3 / sustained,
            "peak_source": src, "api": "kandinsky.models.vae.AutoencoderKLHunyuanVideo.decode -> k5_vae_decode (+ uint8 conversion)"}


class CachedTextEmbedder:
    """Stand-in for Kandinsky5TextEmbedder.encode (text_embedders.py) returning cached synthetic embeddings held in
    PINNED HOST memory, as BASELINE.json's configs prescribe (Qwen2.5-VL / CLIP are outside the hot path)."""

    def __init__(self, L, Ln):
        import torch

        g = torch.Generator().manual_seed(1)
        self.pos = {"text_embeds": torch.randn(L, 3584, generator=g).to(torch.bfloat16).pin_memory(),
                    "pooled_embed": torch.randn(1, 768, generator=g).to(torch.bfloat16).pin_memory()}
        self.neg = {"text_embeds": torch.randn(Ln, 3584, generator=g).to(torch.bfloat16).pin_memory(),
                    "pooled_embed": torch.randn(1, 768, generator=g).to(torch.bfloat16).pin_memory()}

    def encode(self, texts, type_of_content="video"):
        import torch

        e = self.pos if texts[0] else self.neg
        return dict(e), torch.tensor([0, e["text_embeds"].shape[0]], dtype=torch.int32)


def other_configs(model, vae, dev, args):
    """BASELINE.json configs 3, 4 and 5 as short samples on the same engine (N = 1 only; the headline line stays
    config 2).  Config 5 is ONE call of the pipeline's generate_sample - 16 NFE + VAE decode + uint8, text embeddings
    from pinned host memory, the video read back to the host - timed by the wall clock around the call."""
    import torch

    from kandinsky.generation_utils import generate, generate_sample

    out = {}
    sustained, _, _ = measured_peaks()

    def conf_for(wl):
        att = {"type": "nabla", **wl["nabla"]} if wl["nabla"] else {"type": "flash"}
        return {"metrics": {"scale_factor": (1.0, 2.0, 2.0)}, "model": {"dit_params": dict(LITE), "attention": att}}

    def sampler_steps(wl, steps, warm):
        T, H, W, L, Ln = wl["T"], wl["H"], wl["W"], wl["L"], wl["Ln"]
        S = T * (H // 2) * (W // 2)
        emb = CachedTextEmbedder(L, Ln)
        te = {k: v.to(dev) for k, v in emb.pos.items()}
        nte = {k: v.to(dev) for k, v in emb.neg.items()}
        noise = torch.randn(T, H, W, 16, device=dev, generator=torch.Generator(device=dev).manual_seed(6554))
        pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]

        def run(n):
            return generate(model, dev, (T, H, W, 16), n, te, nte, pos, torch.arange(L), torch.arange(Ln), wl["w"],
                            wl["sched"], conf_for(wl), noise=noise)

        if warm:
            run(warm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = run(steps)
        e1.record()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(lat).all()), "non-finite latent"
        ms = e0.elapsed_time(e1) / steps
        fwd = 2 if abs(wl["w"] - 1.0) > 1e-6 else 1
        rho = model.last_sparse_density() if wl["nabla"] else 1.0
        fl = dit_flops(S, L, rho) * fwd
        return {"workload": wl["name"], "tokens": S, "forwards_per_step": fwd, "steps_timed": steps, "ms_per_step": ms,
                "tokens_per_s": S * fwd / (ms * 1e-3), "nabla_density": rho if wl["nabla"] else None,
                "full_config_seconds_at_this_rate": ms * 1e-3 * wl["nfe"] / fwd,
                "model_frac_of_sustained_peak": fl / (ms * 1e-3) / 1e12 / sustained}

    out["config_3_5s_sft"] = sampler_steps(WORKLOADS["5s_sft"], 2, 0)
    if model.max_tokens >= 93696:
        out["config_4_10s_sft_sta_only"] = sampler_steps(WORKLOADS["10s_sft_sta"], 2, 1)
        out["config_4_10s_sft_nabla_P0.9"] = sampler_steps(WORKLOADS["10s_sft_nabla"], 1, 0)
    if vae is not None:
        # config 5: config_5s_distil = 16 NFE (w = 1) + VAE decode, one pipeline call
        emb = CachedTextEmbedder(256, 64)
        conf = conf_for(WORKLOADS["5s_nocfg"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        video = generate_sample((1, 31, 64, 96, 16), "a synthetic prompt", model, vae, conf, text_embedder=emb, num_steps=16,
                                guidance_weight=1.0, scheduler_scale=5.0, negative_caption="", seed=6554, device=dev,
                                vae_device=dev, progress=False)
        host = video.cpu()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        assert host.dtype == torch.uint8 and tuple(host.shape) == (1, 3, 121, 512, 768)
        h2d = sum(v.numel() * 2 for v in emb.pos.values()) + sum(v.numel() * 2 for v in emb.neg.values())
        out["config_5_5s_distil_e2e"] = {
            "workload": "config_5s_distil: 768x512x121, 16 NFE (no CFG) + VAE decode (14 temporal tiles) -> uint8 video on the host",
            "seconds": sec, "nfe": 16, "video_shape": list(host.shape), "h2d_bytes": h2d, "d2h_bytes": host.numel(),
            "api": "kandinsky.generation_utils.generate_sample (k5_sample + k5_vae_decode), one call, wall clock",
            "note": "random-init weights: the latent is not unit-scale, pixel values are not meaningful, the work is the same"}
    return out


def run_k5(args, wl):
    import torch
    import torch.distributed as dist

    from kandinsky import _lib
    from kandinsky.generation_utils import generate
    from kandinsky.models.dit import DiffusionTransformer3D

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the DiT hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    T, H, W, L, Ln = wl["T"], wl["H"], wl["W"], wl["L"], wl["Ln"]
    S = T * (H // 2) * (W // 2)
    cfg_on = abs(wl["w"] - 1.0) > 1e-6
    fwd_per_step = 2 if cfg_on else 1
    with_configs = world == 1 and not args.no_configs and args.workload == "5s_nocfg"
    # (the 10 s workloads of the `configs` block need the larger token workspace; buffer sizes do not change timings)
    model = DiffusionTransformer3D(**LITE, max_tokens=max(S, 93696 if with_configs else 0), max_text_tokens=max(L, Ln))
    sd = synthetic_state_dict_gpu(LITE, dev, seed=0)
    model.load_state_dict(sd, assign=True)
    model.to(dev)
    del sd
    torch.cuda.empty_cache()

    if world > 1:
        from kandinsky.models.parallelize import parallelize_dit

        parallelize_dit(model)
    g = torch.Generator(device=dev).manual_seed(1)          # identical inputs on every rank of the shard
    text = torch.randn(L, 3584, device=dev, generator=g).to(torch.bfloat16)
    pooled = torch.randn(1, 768, device=dev, generator=g).to(torch.bfloat16)
    ntext = torch.randn(Ln, 3584, device=dev, generator=g).to(torch.bfloat16)
    npooled = torch.randn(1, 768, device=dev, generator=g).to(torch.bfloat16)
    noise = torch.randn(T, H, W, 16, device=dev, generator=torch.Generator(device=dev).manual_seed(6554))
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    att = {"type": "nabla", **wl["nabla"]} if wl["nabla"] else {"type": "flash"}
    conf = {"metrics": {"scale_factor": (1.0, 2.0, 2.0)}, "model": {"dit_params": dict(LITE), "attention": att}}
    te = {"text_embeds": text, "pooled_embed": pooled}
    nte = {"text_embeds": ntext, "pooled_embed": npooled}

    def sample(nsteps, start):
        # the schedule of a K-step run; cost per step does not depend on t
        return generate(model, dev, (T, H, W, 16), nsteps, te, nte, pos, torch.arange(L), torch.arange(Ln), wl["w"],
                        wl["sched"], conf, noise=start)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -------------------------------------------------------------------
    sample(max(args.warmup, 3), noise)
    barrier()
    # how long the HOST needs to enqueue one sampler step (the call returns before the device has run it): what a CUDA
    # graph could save at most, reported beside the device time
    t_h = time.perf_counter()
    sample(1, noise)
    host_enqueue_ms = (time.perf_counter() - t_h) * 1e3
    barrier()
    import ctypes

    lib.k5_engine_attention_timing(model._engine, 1, None, None)
    lib.k5_launch_count(1)
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    latent = sample(args.steps, noise)
    e1.record()
    barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.k5_launch_count(1))
    att_ms, att_n = ctypes.c_double(0.0), ctypes.c_int64(0)
    lib.k5_engine_attention_timing(model._engine, 0, ctypes.byref(att_ms), ctypes.byref(att_n))
    density = model.last_sparse_density() if wl["nabla"] else 1.0
    assert bool(torch.isfinite(latent).all()), "non-finite latent"
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = S * fwd_per_step / (ms_step * 1e-3)
    model.set_grid((T, H, W), pos, (1.0, 2.0, 2.0), wl["nabla"] is not None)
    f0, nf = model.local_frames()
    S_loc = nf * (H // 2) * (W // 2)

    # ---- end to end through the reference-facing API with host buffers (e2e) -----------------------------
    h_img = torch.randn(T, H, W, 16).pin_memory()
    h_text, h_pooled = text.cpu().pin_memory(), pooled.cpu().pin_memory()
    h_ntext, h_npooled = ntext.cpu().pin_memory(), npooled.cpu().pin_memory()
    h_out = torch.empty(T, H, W, 16, dtype=torch.bfloat16).pin_memory()
    sparse = None
    if wl["nabla"]:
        sparse = {"to_fractal": True, **wl["nabla"]}
    t1000 = torch.tensor([500.0])

    def e2e_step():
        x = h_img.to(dev, non_blocking=True)
        v = model(x, h_text.to(dev, non_blocking=True), h_pooled.to(dev, non_blocking=True), t1000, pos, torch.arange(L),
                  scale_factor=(1.0, 2.0, 2.0), sparse_params=sparse)
        if cfg_on:
            vu = model(x, h_ntext.to(dev, non_blocking=True), h_npooled.to(dev, non_blocking=True), t1000, pos,
                       torch.arange(Ln), scale_factor=(1.0, 2.0, 2.0), sparse_params=sparse)
            v = vu + wl["w"] * (v - vu)
        h_out[f0:f0 + nf].copy_(v[f0:f0 + nf], non_blocking=True)      # a rank produces (and returns) its own frames
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = h_img.numel() * 4 + h_text.numel() * 2 + h_pooled.numel() * 2
    if cfg_on:
        h2d += h_ntext.numel() * 2 + h_npooled.numel() * 2
    d2h = h_out[f0:f0 + nf].numel() * 2

    # ---- VAE decode of the sampled latent (BASELINE.json configs[4], decode leg): reported beside the DiT metric ----
    vae_info, vae, configs = None, None, None
    if world == 1 and not args.no_vae and not wl["nabla"]:
        vae = build_bench_vae(dev, H, W)
        vae_info = time_vae_decode(vae, latent, dev)
    if with_configs:
        configs = other_configs(model, vae, dev, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sustained, burst, src = measured_peaks()
    flops_fwd = dit_flops(S, L, density)
    attn_flops = density * 4.0 * S_loc * S * 1792          # per launch on this rank: own query rows x all keys
    # dram bytes per launch of the dominant kernel: from the committed ncu --set full capture of the same kernel at the same
    # size (profiles/attention_traffic.json names the .ncu-rep); null when that report is not in the tree
    traffic, traffic_src = None, "not measured in this run"
    tpath = os.path.join(ROOT, "profiles", "attention_traffic.json")
    if world == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if os.path.exists(os.path.join(ROOT, tj.get("_report", "missing"))):
            traffic = tj.get(args.workload)
            traffic_src = "ncu --set full capture " + tj["_report"] if traffic is not None else traffic_src
    att_avg_ms = att_ms.value / max(att_n.value, 1)
    achieved = attn_flops / (att_avg_ms * 1e-3) / 1e12 if att_n.value else None
    line = {
        "metric": "dit_latent_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": bench_config(wl, world, nf, density, model.dist_mode()),
        "ms_per_forward": ms_step / fwd_per_step,
        "model_tflops_per_forward": flops_fwd / 1e12,
        "model_tflops_achieved": flops_fwd * fwd_per_step / (ms_step * 1e-3) / 1e12,
        "model_frac_of_sustained_peak": flops_fwd * fwd_per_step / (ms_step * 1e-3) / 1e12 / sustained,
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clk,
        "e2e": {"value": S * fwd_per_step / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "kandinsky.models.dit.DiffusionTransformer3D.forward -> k5_dit_forward (pinned host buffers)"},
        "roofline": {"kernel": "attention_fwd_kernel (visual self-attention, tcgen05)", "bound": "tensor",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                     "frac": (achieved / sustained) if achieved else None, "peak_source": f"{src} (sustained bf16 GEMM)",
                     "flops_per_launch": attn_flops, "avg_launch_ms": att_avg_ms, "launches_timed": int(att_n.value),
                     "share_of_step": att_ms.value / ms_total if ms_total else None, "traffic": traffic,
                     "traffic_source": traffic_src},
    }
    if vae_info is not None:
        line["vae_decode"] = vae_info
    if configs is not None:
        line["configs"] = configs
    if not args.no_cpu_baseline:
        v, m, cores, smp, kind = cpu_arm(wl, 20.0)
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": smp}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line of this run, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    # Libraries (NCCL's version banner, for one) write to fd 1; keep the real stdout for the JSON line only.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="k5", choices=["k5", "reference"])
    ap.add_argument("--workload", default="5s_nocfg", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="skip the VAE-decode measurement that follows the DiT timing")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the short samples of BASELINE.json configs 3 / 4 / 5 that follow the headline measurement (N = 1)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_k5(args, wl)


if __name__ == "__main__":
    main()
