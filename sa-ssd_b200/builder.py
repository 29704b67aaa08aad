"""Registry — mirror of mmdet/models/builder.py:13-56: ``type`` strings are resolved as
attributes of the sub-package of the matching role, so configs/car_cfg.py builds unchanged."""
from torch import nn

from . import backbones
from . import necks, single_stage_heads
from .config import obj_from_dict


def _build_module(cfg, parent=None, default_args=None):
    return cfg if isinstance(cfg, nn.Module) else obj_from_dict(cfg, parent, default_args)


def build(cfg, parent=None, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[_build_module(c, parent, default_args) for c in cfg])
    return _build_module(cfg, parent, default_args)


def build_backbone(cfg):
    return build(cfg, backbones)


def build_neck(cfg):
    return build(cfg, necks)


def build_single_stage_head(cfg):
    return build(cfg, single_stage_heads)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    from . import detectors
    return build(cfg, detectors, dict(train_cfg=train_cfg, test_cfg=test_cfg))
