"""Loads the UNMODIFIED reference (ai-forever/Kandinsky-5) from the git-ignored copy under ``baseline/_ref/``.

Test / benchmark infrastructure only: nothing under ``kandinsky-5_b200/`` imports this.  ``baseline/_ref/`` is
filled by ``populate()`` (called from ``__graft_entry__.build()`` in the build container, where ``/root/reference``
exists) with verbatim copies of the reference's own files; it is git-ignored but travels to the GPU box with the
``gpurun`` snapshot.  The reference cannot be pip-installed (it ships no setup.py / pyproject) and its package
``__init__`` pulls omegaconf / diffusers (absent), so the modules on the DiT path are imported directly under empty
package objects - the same way ``tests/golden/make_golden.py`` does from ``/root/reference``.

  * on a GPU box the reference runs as its authors run it: real ``torch.autocast('cuda', bf16)``, FlashAttention-2
    (``flash_attn_func``, picked by kandinsky/models/nn.py:9-23), eager (``TORCHDYNAMO_DISABLE=1``) or compiled;
  * on CPU (the ``--impl reference`` arm and ``cpu_baseline``) ``nn.FA`` is replaced by an SDPA wrapper with the same
    [B, S, H, D] contract and the caller wraps the call in ``torch.autocast('cpu', bf16)``.
"""
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("K5_REFERENCE", "/root/reference")

# what the DiT path needs (SURVEY.md section 8a + Appendix D): no text encoders, no VAE (needs diffusers), no CLI
FILES = [
    "kandinsky/generation_utils.py", "kandinsky/magcache_utils.py", "kandinsky/models/dit.py", "kandinsky/models/nn.py",
    "kandinsky/models/utils.py", "kandinsky/models/parallelize.py", "configs/config_5s_nocfg.yaml",
    "configs/config_5s_sft.yaml", "configs/config_5s_distil.yaml", "configs/config_10s_sft.yaml", "LICENSE",
]


def populate(force=False):
    """Copy the reference's files (verbatim) into baseline/_ref/.  No-op when the reference tree is absent."""
    if not os.path.isdir(os.path.join(SOURCE, "kandinsky")):
        return False
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(REF_DIR, rel)
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return True


def available():
    return os.path.exists(os.path.join(REF_DIR, "kandinsky", "models", "dit.py"))


_mods = None


def import_reference(cpu_attention=False, keep_registered=False):
    """Returns {'utils','nn','dit','generation_utils'} modules of the reference.  cpu_attention: replace nn.FA by an
    SDPA wrapper (flash_attn has no CPU kernels).  keep_registered: leave the reference's modules in sys.modules under
    their own names (`kandinsky.models.nn` ...) - TorchDynamo re-imports a traced function's module by name, so the
    compiled run needs it; the drop-in mirror of the same name cannot be imported in that process afterwards."""
    global _mods
    import torch
    import torch.nn.functional as F

    if _mods is None:
        if not available():
            raise ImportError("baseline/_ref is empty: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                              "/root/reference exists")
        if not torch.cuda.is_available():
            torch.cuda.get_device_capability = lambda *a, **k: (10, 0)      # nn.py:9 asks at import time
        root = os.path.join(REF_DIR, "kandinsky")
        saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "kandinsky" or k.startswith("kandinsky.")}
        for k in saved:
            del sys.modules[k]
        try:
            for name, path in (("kandinsky", root), ("kandinsky.models", os.path.join(root, "models"))):
                m = types.ModuleType(name)
                m.__path__ = [path]
                sys.modules[name] = m
            mods = {}
            for name in ("kandinsky.models.utils", "kandinsky.models.nn", "kandinsky.models.dit",
                         "kandinsky.generation_utils"):
                mods[name.split(".")[-1]] = importlib.import_module(name)
        finally:
            # hand the name `kandinsky` back to whoever had it (the drop-in mirror lives under the same name)
            if not keep_registered:
                for k in [k for k in sys.modules if k == "kandinsky" or k.startswith("kandinsky.")]:
                    del sys.modules[k]
                for k, v in saved.items():
                    sys.modules[k] = v
        _mods = mods
    if cpu_attention:
        def fa(q, k, v):       # flash_attn_func contract: [B,S,H,D] in / out, non-causal, scale d^-0.5
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
            return o.transpose(1, 2)

        _mods["nn"].FA = fa
    return _mods


def build_model(cfg, state_dict, device):
    """The reference's DiffusionTransformer3D with a given state dict (strict load, as kandinsky/utils.py:89-116)."""
    mods = import_reference(cpu_attention=(str(device) == "cpu"))
    model = mods["dit"].get_dit(dict(cfg))
    res = model.load_state_dict(state_dict, assign=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.eval().to(device)
