"""Pure-Python stand-ins for the third-party packages the reference imports at module load but that are not
in this image: accelerate, fvcore, iopath, termcolor, terminaltables, omegaconf, yapf,
pycocotools, albumentations.  Only what executes on the model / criterion path is functional (loggers, the
DictConfig base class); everything else is an inert placeholder so that ``import util.misc`` etc. succeed.

Bench / test infrastructure for running the UNMODIFIED reference (``baseline/_ref``): never imported by the
product package.
"""
from __future__ import annotations

import importlib.util
import logging
import sys
import types


class _Inert:
    """Placeholder class: constructible, callable, attribute access returns another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Inert,), {})
        setattr(self, name, cls)
        return cls


def _module(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    mod = _StubModule(name)
    mod.__path__ = []  # behave as a package so that submodule imports resolve through sys.modules
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_module(parent), child, mod)
    return mod


class _MultiProcessAdapter(logging.LoggerAdapter):
    """accelerate.logging.get_logger returns an adapter whose methods take main_process_only / in_order."""

    def log(self, level, msg, *args, **kwargs):
        kwargs.pop("main_process_only", None)
        kwargs.pop("in_order", None)
        if self.isEnabledFor(level):
            msg, kwargs = self.process(msg, kwargs)
            self.logger.log(level, msg, *args, **kwargs)


def _get_logger(name=None, log_level=None):
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper() if isinstance(log_level, str) else log_level)
    return _MultiProcessAdapter(logger, {})


class _Meta:
    object_type = dict


class DictConfig(dict):
    """omegaconf.DictConfig as far as util/lazy_load.py and the backbones use it: a dict with attribute access,
    built as ``DictConfig(content=..., flags=...)``, with ``_metadata.object_type`` and deep copies."""

    def __init__(self, content=None, flags=None, **kw):
        super().__init__(content or {}, **kw)

    _metadata = _Meta()

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        import copy

        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    def __init__(self, content=None, flags=None):
        super().__init__(content or [])


class OmegaConf:
    @staticmethod
    def to_object(cfg):
        return cfg


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install() -> list:
    """Registers stand-ins for the packages that are really absent; returns their names."""
    done = []
    if _missing("accelerate"):
        acc = _module("accelerate")
        _module("accelerate.logging").get_logger = _get_logger
        _module("accelerate.utils")
        acc.__version__ = "0.0-stub"
        done.append("accelerate")
    if _missing("fvcore"):
        _module("fvcore.common.file_io")
        _module("fvcore.nn")
        done.append("fvcore")
    if _missing("iopath"):
        _module("iopath.common.file_io")
        done.append("iopath")
    if _missing("termcolor"):
        _module("termcolor").colored = lambda s, *a, **k: s
        done.append("termcolor")
    if _missing("terminaltables"):
        _module("terminaltables")
        done.append("terminaltables")
    if _missing("omegaconf"):
        om = _module("omegaconf")
        om.DictConfig, om.ListConfig, om.OmegaConf = DictConfig, ListConfig, OmegaConf
        done.append("omegaconf")
    if _missing("yapf"):
        _module("yapf.yapflib.yapf_api").FormatCode = lambda code, **k: (code, False)
        done.append("yapf")
    if _missing("pycocotools"):
        _module("pycocotools.mask")
        _module("pycocotools.coco")
        _module("pycocotools.cocoeval")
        done.append("pycocotools")
    if _missing("albumentations"):
        _module("albumentations")
        done.append("albumentations")
    return done
