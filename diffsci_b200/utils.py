"""diffsci/utils.py:5-11."""
from __future__ import annotations


def get_minibatch_sizes(n: int, b: int) -> list[int]:
    """Split n items into chunks of b (last chunk holds the remainder)."""
    full, rem = divmod(n, b)
    return [b] * full + ([rem] if rem else [])
