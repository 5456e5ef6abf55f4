"""gpt.classify_gpt.classify_layer — same head as classify.py with the float16 kwarg
(mirrors gpt/classify_gpt.py:6-55)."""
from classify import classify_layer  # noqa: F401
