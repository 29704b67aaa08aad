"""Minimal stand-in for the two mmcv entry points the reference uses on the
inference path: ``mmcv.Config.fromfile`` (tools/test.py:128) and
``mmcv.runner.obj_from_dict`` (mmdet/models/builder.py:1,13-16).

The reference configs are plain Python files whose module-level names become
attribute-accessible nested dicts; objects are built from dicts carrying a
``type`` key that is looked up as an attribute of a Python (sub)package."""
import os
import types


class ConfigDict(dict):
    """dict with attribute access, nested dicts converted on the way in."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, list):
            return [ConfigDict._wrap(x) for x in v]
        if isinstance(v, tuple):
            return tuple(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


class Config:
    """``Config.fromfile(path)`` executes a config ``.py`` and exposes its
    public module-level names (tools/test.py:128 usage: ``cfg.model``,
    ``cfg.test_cfg``, ``cfg.data.val``)."""

    def __init__(self, cfg_dict, filename=None):
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        if not filename.endswith(".py"):
            raise IOError("Only py type config files are supported")
        ns = {"__file__": filename, "__name__": "_sassd_config_"}
        with open(filename) as f:
            exec(compile(f.read(), filename, "exec"), ns)
        cfg = {k: v for k, v in ns.items()
               if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType))}
        return Config(cfg, filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)


def obj_from_dict(info, parent=None, default_args=None):
    """Build ``getattr(parent, info['type'])(**rest)`` — same contract as the
    mmcv helper the reference registry relies on (builder.py:13-16): ``type``
    may be a string (looked up on ``parent``) or a class; ``default_args`` only
    fill keys that are absent."""
    if not (isinstance(info, dict) and "type" in info):
        raise TypeError("info must be a dict containing the key 'type'")
    if default_args is not None and not isinstance(default_args, dict):
        raise TypeError("default_args must be a dict or None")
    args = dict(info)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        if parent is not None:
            obj_type = getattr(parent, obj_type)
        else:
            raise KeyError("cannot resolve type %r without a parent package" % obj_type)
    elif not isinstance(obj_type, type):
        raise TypeError("type must be a str or a class, got %s" % type(obj_type))
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    return obj_type(**args)
