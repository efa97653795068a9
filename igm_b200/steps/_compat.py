"""Binding to the reference's Step/Config machinery.

When the real ``igm`` package is importable (a production IGM install) the
drop-in subclasses ``igm.core.Step`` and is driven by ``Step.run``
(igm/core/step.py:226-322) unchanged.  In this image ``igm`` cannot be imported
(alabtools / h5py / ipyparallel are absent), so a duck-typed stand-in with the
same phase order (setup -> map(task) -> reduce -> cleanup, skip()) and the same
``Config.get/set`` path semantics (igm/core/config.py:98-127) is used; it keeps
no sqlite restart log.
"""
from __future__ import annotations

import logging
import os
from functools import partial
from hashlib import md5

logger = logging.getLogger("IGM")


def make_absolute_path(path, basedir="."):
    """igm/utils/files.py:118-122."""
    if os.path.isabs(path):
        return path
    return os.path.abspath(os.path.join(basedir, path))


_RAISE = object()


class Config(dict):
    """Slash-path get/set on a nested dict (igm/core/config.py:98-127), without
    the schema defaults of the reference (callers pass explicit defaults)."""

    def __init__(self, cfg=None):
        super().__init__()
        if cfg is not None:
            import copy
            self.update(copy.deepcopy(dict(cfg)))
        self.setdefault("runtime", {})
        self.setdefault("parameters", {})
        self["parameters"].setdefault("workdir", os.getcwd())
        self["parameters"].setdefault("tmp_dir", "tmp")
        for k in self.get("restraints", {}):
            self["runtime"].setdefault(k, {})

    def get(self, keypath, default=_RAISE):
        d = self
        try:
            for p in keypath.split("/"):
                d = d[p]
        except (KeyError, TypeError):
            if default is not _RAISE:
                return default
            raise KeyError("{} does not exist".format(keypath))
        return d

    def set(self, keypath, val):
        parts = keypath.split("/")
        d = self
        for p in parts[:-1]:
            if p not in d:
                d[p] = dict()
            d = d[p]
        d[parts[-1]] = val
        return val


class _SerialController:
    def map(self, fn, args):
        return [fn(a) for a in args]


class ShimStep(object):
    """Phase order and bookkeeping of igm.core.Step (igm/core/step.py:161-322)
    minus the StepDB restart log."""

    def __init__(self, cfg):
        self.controller = _SerialController()
        self.cfg = cfg
        self.tmp_extensions = []
        self.tmp_dir = make_absolute_path(self.cfg["parameters"].get("tmp_dir", "tmp/"),
                                          cfg["parameters"]["workdir"])
        self.keep_temporary_files = True
        os.makedirs(self.tmp_dir, exist_ok=True)
        if "current_iteration_name" not in cfg["runtime"]:
            self.cfg["runtime"]["current_iteration_name"] = self.name()
        if cfg["runtime"].get("step_no") is None:
            self.cfg["runtime"]["step_no"] = -1
        self.cfg["runtime"]["step_no"] += 1
        self.uid = md5("{:s}:{:d}".format(self.name(), self.cfg["runtime"]["step_no"]).encode()).hexdigest()
        self.cfg["runtime"]["step_hash"] = self.uid

    def setup(self):
        self.argument_list = []

    def before_map(self):
        return

    @staticmethod
    def task(struct_id, cfg, tmp_dir):
        pass

    def before_reduce(self):
        return

    def reduce(self):
        pass

    def cleanup(self):
        if not self.keep_temporary_files:
            for f in os.listdir(self.tmp_dir):
                if os.path.splitext(f)[1] in self.tmp_extensions:
                    os.remove(self.tmp_dir + "/" + f)

    def run(self):
        logger.info("%s - starting" % self.name())
        self.setup()
        serial_function = partial(self.__class__.task, cfg=self.cfg, tmp_dir=self.tmp_dir)
        self.before_map()
        self.controller.map(serial_function, self.argument_list)
        self.before_reduce()
        self.reduce()
        self.cleanup()
        self.cfg["runtime"].pop("step_hash", None)
        logger.info("%s - completed" % self.name())

    def name(self):
        return self.__class__.__name__

    def skip(self):
        return None


def _resolve_step_base():
    try:
        from igm.core import Step as RefStep   # real IGM install
        return RefStep, True
    except Exception:
        return ShimStep, False


Step, HAVE_REFERENCE_STEP = _resolve_step_base()
