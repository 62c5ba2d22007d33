from .align_ctypes import *  # noqa: F401,F403  (mirrors ambivert/align/__init__.py:1)
