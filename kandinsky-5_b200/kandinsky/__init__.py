"""B200-native drop-in for the `kandinsky` package of ai-forever/Kandinsky-5 (DiT denoising hot path).

    from kandinsky import get_T2V_pipeline          # same entry point as the reference (kandinsky/__init__.py:1)
"""
from .utils import get_T2V_pipeline  # noqa: F401
