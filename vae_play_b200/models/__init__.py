from . import blocks, networks  # noqa: F401
