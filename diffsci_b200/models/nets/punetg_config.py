"""PUNetGConfig -- same constructor signature / attribute names as the reference
(diffsci/models/nets/punetg_config.py:8-38) so existing scripts and YAML files keep working."""
from __future__ import annotations

import inspect
import pathlib
from typing import Any

import yaml


class PUNetGConfig:
    def __init__(self, input_channels: int = 1, output_channels: int = 1, dimension: int = 2,
                 model_channels: int = 64, channel_expansion: list[int] = [2, 4],
                 number_resnet_downward_block: int = 2, number_resnet_upward_block: int = 2,
                 number_resnet_attn_block: int = 2, number_resnet_before_attn_block: int = 2,
                 number_resnet_after_attn_block: int = 2, kernel_size: int = 3, in_out_kernel_size: int = 3,
                 in_embedding: bool = False, time_projection_scale: float = 30.0,
                 input_projection_scale: float = 1.0, transition_scale_factor: int = 2,
                 transition_kernel_size: int = 3, dropout: float = 0.0, cond_dropout: float = 0.0,
                 cond_drop: float = 0.0, cond_drop_learnable: bool = True, first_resblock_norm: str = "GroupLN",
                 second_resblock_norm: str = "GroupRMS", affine_norm: bool = True,
                 convolution_type: str = "default", num_groups: int = 1, attn_residual: bool = False,
                 attn_type: str = "default", bias: bool = True):
        for name, value in list(locals().items()):
            if name != "self":
                setattr(self, name, value)

    _FIELDS = None

    @classmethod
    def field_names(cls) -> list[str]:
        if cls._FIELDS is None:
            cls._FIELDS = [n for n in inspect.signature(cls.__init__).parameters if n != "self"]
        return cls._FIELDS

    @property
    def extended_channel_expansion(self) -> list[int]:
        return [1] + list(self.channel_expansion)

    @property
    def magnitude_preserving(self) -> bool:
        return self.convolution_type == "mp"

    def export_description(self) -> dict[str, Any]:
        return {n: getattr(self, n) for n in self.field_names()}

    @classmethod
    def from_description(cls, description: dict) -> "PUNetGConfig":
        return cls(**description)

    @classmethod
    def from_config_file(cls, config_file: pathlib.Path | str) -> "PUNetGConfig":
        with open(config_file, "r") as f:
            return cls.from_description(yaml.safe_load(f))
