from .chunk_decode import chunk_decode_3d  # noqa: F401
