#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref, see baseline/ref_loader.py) on this B200 beside the engine, on the
same synthetic weights and inputs (not a pytest file):  one DiT forward at the 5 s size (BASELINE.json configs[1]:
S = 47 616, L = 256) for 2 and 32 visual blocks, eager (dynamo disabled) and - when K5_REF_COMPILED=1 - with the
reference's own @torch.compile decorators live.  Writes one JSON line per measurement to stdout / the file given as
argv[1]; the numbers land in profiles/r2_reference_b200.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))

from baseline import ref_loader  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402


def synth(nblocks, T, H, W, L):
    cfg = dict(O.LITE_CFG, num_visual_blocks=nblocks)
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = O.dit_state_dict_shapes(cfg)
    sd = {}
    for k, shp in shapes.items():
        if "modulation" in k:
            t = torch.randn(shp, device="cuda", generator=g) * 0.02
        elif k.endswith("norm.weight"):
            t = torch.ones(shp, device="cuda")
        elif k.endswith("norm.bias"):
            t = torch.zeros(shp, device="cuda")
        else:
            fan_in = shp[1] if k.endswith(".weight") else shapes[k[:-4] + "weight"][1]
            t = (torch.rand(shp, device="cuda", generator=g) * 2 - 1) / fan_in ** 0.5
        sd[k] = t if O.is_fp32_key(k) else t.to(torch.bfloat16)
    img = torch.randn(T, H, W, 16, device="cuda", generator=g)
    text = torch.randn(L, 3584, device="cuda", generator=g).to(torch.bfloat16)
    pooled = torch.randn(1, 768, device="cuda", generator=g).to(torch.bfloat16)
    return cfg, sd, img, text, pooled


def events(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    out = open(sys.argv[1], "a") if len(sys.argv) > 1 else None
    # "1": the reference exactly as shipped (the whole forward under @torch.compile(mode="max-autotune-no-cudagraphs"),
    # dit.py:155, with the 20 inner @torch.compile regions nested in it) - under torch 2.11 this raises
    # InternalTorchDynamoError ('CatchErrorsWrapper' object has no attribute '__closure__', gpurun_out/r2_reference_compiled.log);
    # "inner": the outermost wrapper is bypassed (its _torchdynamo_orig_callable is called), every inner region still compiles
    mode_env = os.environ.get("K5_REF_COMPILED", "0")
    compiled = mode_env in ("1", "inner")
    import torch._dynamo

    torch._dynamo.config.disable = not compiled
    ref_loader.import_reference(keep_registered=compiled)
    T, H, W, L = 31, 64, 96, 256
    S = T * (H // 2) * (W // 2)
    for nblocks in [int(x) for x in os.environ.get("K5_REF_BLOCKS", "2,32").split(",")]:
        cfg, sd, img, text, pooled = synth(nblocks, T, H, W, L)
        x = O.model_input(img, True)
        pos = [torch.arange(T, device="cuda"), torch.arange(H // 2, device="cuda"), torch.arange(W // 2, device="cuda")]
        t1000 = torch.tensor([700.0], device="cuda")
        tpos = torch.arange(L, device="cuda")
        model = ref_loader.build_model(cfg, sd, "cuda")

        bypassed = []
        if mode_env == "inner":
            # regions whose torch.compile wrapper NESTS other compiled regions trip the same dynamo error under torch 2.11;
            # they run through their original Python body, everything they call stays compiled
            for name in os.environ.get("K5_REF_BYPASS", "forward,after_blocks").split(","):
                cur = getattr(type(model), name)
                orig = getattr(cur, "_torchdynamo_orig_callable", None)
                if orig is not None:                     # (already replaced on the class by an earlier model of this process)
                    setattr(type(model), name, orig)
                else:
                    assert not hasattr(cur, "_torchdynamo_orig_callable") and callable(cur)
                bypassed.append(f"DiffusionTransformer3D.{name}")
        fwd = model

        def ref_fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return fwd(x, text, pooled, t1000, pos, tpos, scale_factor=(1.0, 2.0, 2.0))

        t0 = time.time()
        ref_out = ref_fwd()
        torch.cuda.synchronize()
        first = time.time() - t0
        ms_ref = events(ref_fwd, 3 if nblocks > 8 else 5, 1)
        rec = {"what": "reference forward on B200", "mode": {"1": "compiled", "inner": "compiled (inner regions only)"}.get(mode_env, "eager"), "visual_blocks": nblocks,
               "tokens": S, "text_tokens": L, "ms_per_forward": ms_ref, "first_call_s": first,
               "attention": getattr(ref_loader.import_reference()["nn"].FA, "__module__", "?"), "torch": torch.__version__}
        if bypassed:
            rec["compile_wrappers_bypassed"] = bypassed
        del model
        torch.cuda.empty_cache()
        if not compiled:
            from kandinsky.models.dit import DiffusionTransformer3D

            eng = DiffusionTransformer3D(**cfg, max_tokens=S, max_text_tokens=L)
            eng.load_state_dict(sd, assign=True)
            eng.to("cuda:0")
            cpos = [p.cpu() for p in pos]

            def eng_fwd():
                return eng(img, text, pooled, t1000, cpos, torch.arange(L), scale_factor=(1.0, 2.0, 2.0))

            eng_out = eng_fwd()
            rec["engine_ms_per_forward"] = events(eng_fwd, 5, 2)
            rec["engine_vs_reference_rel_l2"] = float((eng_out.float() - ref_out.float()).norm() / ref_out.float().norm())
            del eng
            torch.cuda.empty_cache()
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()


if __name__ == "__main__":
    main()
